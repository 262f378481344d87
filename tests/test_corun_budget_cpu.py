"""The two kernels of a whole batch must FIT next to each other on one SM — otherwise the co-run of launch_split_pair
(hbt_b200.cu) silently becomes one kernel after the other.  Registers and static shared memory of the two kernels are
read from the ptxas log the build writes (csrc/ptxas.log) and checked against the co-run split compiled into the
library: corun_same x same-event kernel + corun_mixed x v4 mixed-event kernel within 64 K registers and the 228 KB of
shared memory of an sm_100 SM (1 KB per resident CTA is reserved by the system)."""
import os
import re

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "hadronic_afterburner_toolkit_b200", "csrc")
SM_REGISTERS = 65536
SM_SHARED = 233472       # 228 KB
CTA_RESERVED = 1024


def kernel_resources(log, mangled_prefix):
    m = re.search(r"Compiling entry function '(" + re.escape(mangled_prefix) + r"[^']*)' for 'sm_100a'.*?Used (\d+) registers.*?(\d+) bytes smem", log, re.S)
    assert m, f"{mangled_prefix} not in ptxas.log"
    return int(m.group(2)), int(m.group(3))


def test_the_two_kernels_of_a_batch_fit_one_sm_together():
    path = os.path.join(CSRC, "ptxas.log")
    if not os.path.exists(path):
        pytest.skip("csrc/ptxas.log not there (library not built here)")
    log = open(path).read()
    regs_s, smem_s = kernel_resources(log, "_Z12hbt_pairs_v3ILb0ELb0ELb0EE")
    regs_m, smem_m = kernel_resources(log, "_Z18hbt_pairs_v4_mixed")
    src = open(os.path.join(CSRC, "hbt_b200.cu")).read()
    m = re.search(r"int corun_same = (\d+), corun_mixed = (\d+);", src)
    assert m
    n_s, n_m = int(m.group(1)), int(m.group(2))

    def regs_per_warp(r):  # allocated in units of 256 registers per warp
        return (r * 32 + 255) // 256 * 256

    assert n_s * regs_per_warp(regs_s) + n_m * regs_per_warp(regs_m) <= SM_REGISTERS
    assert n_s * (smem_s + CTA_RESERVED) + n_m * (smem_m + CTA_RESERVED) <= SM_SHARED
    # each kernel alone: the resident warps DESIGN.md quotes (20 same-event, 24 mixed-event)
    assert min(SM_REGISTERS // regs_per_warp(regs_s), SM_SHARED // (smem_s + CTA_RESERVED)) >= 20
    assert min(SM_REGISTERS // regs_per_warp(regs_m), SM_SHARED // (smem_m + CTA_RESERVED)) >= 24
