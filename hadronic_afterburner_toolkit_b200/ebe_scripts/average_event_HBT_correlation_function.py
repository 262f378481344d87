#!/usr/bin/env python3
"""Event averaging of the HBT correlation-function files.

Mirrors `/root/reference/ebe_scripts/average_event_HBT_correlation_function.py:18-48`: every
`<working_folder>/UrQMD*/UrQMD_results/HBT*` table (the 5-column files written by
`output_correlation_function`, `/root/reference/src/HBT_correlation.cpp:726-783`: q_out, q_side,
q_long, numerator sum of cos(q.x), pair-ratio-scaled denominator; 2 columns with ecoOutput=1) is
summed element-wise over the event folders, divided by the number of folders and written under
the same file name with `%.10e` fields separated by two blanks.

    python -m hadronic_afterburner_toolkit_b200.ebe_scripts.average_event_HBT_correlation_function \\
        working_folder results_folder
"""
import glob
import os
import sys
from typing import Dict, List

import numpy as np

EVENT_FOLDER_PATTERN = "UrQMD*"
RESULTS_FOLDER = "UrQMD_results"
FILE_PATTERN = "HBT*"


def average_tables(tables: List[np.ndarray]) -> np.ndarray:
    """Arithmetic mean of equally shaped tables, accumulated in file order like the reference
    (sum first, one division at the end: `:36-43`)."""
    if not tables:
        raise ValueError("no tables to average")
    total = np.zeros_like(np.asarray(tables[0], dtype=np.float64))
    for t in tables:
        t = np.asarray(t, dtype=np.float64)
        if t.shape != total.shape:
            raise ValueError(f"table shapes differ: {t.shape} vs {total.shape}")
        total += t
    return total / len(tables)


def average_event_folders(working_folder: str, avg_folder: str, verbose: bool = False) -> Dict[str, np.ndarray]:
    """Average every HBT file over the event folders of `working_folder`; returns
    {file name: averaged table} and writes each table into `avg_folder`."""
    working_folder, avg_folder = os.path.abspath(working_folder), os.path.abspath(avg_folder)
    folders = glob.glob(os.path.join(working_folder, EVENT_FOLDER_PATTERN))
    if not folders:
        raise FileNotFoundError(f"no {EVENT_FOLDER_PATTERN} folders under {working_folder}")
    os.makedirs(avg_folder, exist_ok=True)
    # the list of files is taken from the first folder, as the reference does (:27-30)
    names = [os.path.basename(f) for f in glob.glob(os.path.join(folders[0], RESULTS_FOLDER, FILE_PATTERN))]
    out = {}
    for name in names:
        tables = []
        for folder in folders:
            fn = os.path.join(folder, RESULTS_FOLDER, name)
            if verbose:
                print(f"processing {fn} ...")
            tables.append(np.loadtxt(fn))
        out[name] = average_tables(tables)
        np.savetxt(os.path.join(avg_folder, name), out[name], fmt="%.10e", delimiter="  ")
    return out


def main(argv=None) -> int:
    argv = sys.argv if argv is None else argv
    if len(argv) < 3:
        print(f"Usage: {argv[0]} working_folder results_folder")
        return 1
    average_event_folders(argv[1], argv[2], verbose=True)
    print("Analysis is done.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
