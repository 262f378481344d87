"""host/hbt_output.h writes the rows of the .dat files with snprintf into one buffer per file; the reference streams
them (`output << std::scientific << std::setw(18) << std::setprecision(8) << v0 << "    " << v1 ... << std::endl`,
src/HBT_correlation.cpp:711-718, :760-774).  tests/tools/output_format_check.cpp formats 120 000 rows both ways (zeros,
signed zeros, denormals, huge exponents, inf, nan, random magnitudes over 60 decades) and compares the characters."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rows_equal_the_reference_stream_expression(tmp_path):
    exe = tmp_path / "output_format_check"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "hadronic_afterburner_toolkit_b200", "host"),
                    os.path.join(ROOT, "tests", "tools", "output_format_check.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("identical 0 of")
