#!/usr/bin/env python3
"""3-D Gaussian fit of the HBT radii to a correlation-function table.

Mirrors `/root/reference/ebe_scripts/fit_HBT_radii.py:13-100`: the model is the Bertsch-Pratt
Gaussian of arXiv:1403.4972,

    C(q) = lambda * exp(-[(R_o q_o)^2 + (R_s q_s)^2 + (R_l q_l)^2 + 2 q_o q_s R_os^2 + 2 q_o q_l R_ol^2] / hbarc^2),

fitted with `scipy.optimize.curve_fit` (start values 1, 5, 5, 5, 0.1, 0.1; `absolute_sigma=True`)
to the bins with 0 < |q| < q_cut for q_cut in 0.05 ... 0.15 GeV.  One output row per q_cut:
q_cut, then (value, error) for lambda, R_out, R_side, R_long, R_os, R_ol [fm].

The reference reads 8-column tables out of an HDF5 database (columns 0-2 the bin's q, 6 the
correlation function, 7 its error: the layout `average_event_HBT_correlation_function_h5.py:215-218`
produces) and stores `HBT_radii_KT_*.dat` next to them; `fit_database` does exactly that when
`h5py` is importable.  `correlation_table` builds the same 8 columns from the 5-column files the
analysis binary writes today (`/root/reference/src/HBT_correlation.cpp:726-783`), so that a radius
fit can be run directly on them:

    python -m hadronic_afterburner_toolkit_b200.ebe_scripts.fit_HBT_radii database.h5
    python -m hadronic_afterburner_toolkit_b200.ebe_scripts.fit_HBT_radii HBT_correlation_function_KT_0.15_0.25.dat ...
"""
import os
import sys
from typing import Optional, Sequence

import numpy as np
from scipy.optimize import curve_fit

HBARC = 0.19733          # GeV fm, the script's own constant (:10)
ERR_FLOOR = 1e-15        # added to the error column (:11, :63)
Q_CUT_MAX_LIST = (0.05, 0.075, 0.1, 0.125, 0.15)
Q_CUT_MIN = 0.0
KT_CUT_LIST = ("0_0.2", "0.2_0.4", "0.4_0.6", "0.6_0.8")
START_VALUES = (1.0, 5.0, 5.0, 5.0, 0.1, 0.1)
HEADER = ("# q_cut[GeV]  lambda  lambda_err  R_out[fm]  R_out_err[fm]  "
          "R_side[fm]  R_side_err[fm]  R_long [fm]  R_long_err[fm]  "
          "R_os[fm]  R_os_err[fm]  R_ol[fm]  R_ol_err[fm]")


def gaussian_3d(q_arr, lambda_, R_out, R_side, R_long, R_os, R_ol):
    """The fit function, flattened to 1-D (`:13-29`); radii in fm, q in GeV."""
    q_out, q_side, q_long = q_arr
    ro, rs, rl, ros, rol = (R_out / HBARC, R_side / HBARC, R_long / HBARC, R_os / HBARC, R_ol / HBARC)
    expo = ((ro * q_out) ** 2 + (rs * q_side) ** 2 + (rl * q_long) ** 2
            + 2.0 * q_out * q_side * ros ** 2.0 + 2.0 * q_out * q_long * rol ** 2.0)
    return np.ravel(lambda_ * np.exp(-expo))


def grid_points(n_rows: int) -> int:
    """Points per q axis of a table with n_rows = nq^3 rows (`:58` takes int(n^(1/3)) + 1, which
    relies on the cube root of a perfect cube rounding down; this is the exact inverse)."""
    nq = int(round(n_rows ** (1.0 / 3.0)))
    if nq ** 3 != n_rows:
        raise ValueError(f"{n_rows} rows is not a cubic q grid")
    return nq


def fit_table(HBT_data, q_cut_max_list: Sequence[float] = Q_CUT_MAX_LIST, corr_col: int = 6, err_col: int = 7,
              mask: Optional[np.ndarray] = None, return_rsquared: bool = False):
    """Fit one (K_T) table; returns an array [len(q_cut_max_list)][13] (`:56-93`).  `mask`
    (optional, one flag per row) removes bins from every fit, e.g. bins without pairs."""
    HBT_data = np.nan_to_num(np.asarray(HBT_data, dtype=np.float64))
    nq = grid_points(HBT_data.shape[0])
    shape = (nq, nq, nq)
    q_out, q_side, q_long = (HBT_data[:, c].reshape(shape) for c in (0, 1, 2))
    corr = HBT_data[:, corr_col].reshape(shape)
    corr_err = HBT_data[:, err_col].reshape(shape) + ERR_FLOOR
    q_abs = np.sqrt(q_out ** 2.0 + q_side ** 2.0 + q_long ** 2.0)
    keep = np.ones(shape, dtype=bool) if mask is None else np.asarray(mask, dtype=bool).reshape(shape)
    rows, r2 = [], []
    for q_cut in q_cut_max_list:
        idx = (q_abs > Q_CUT_MIN) & (q_abs < q_cut) & keep
        q_arr = [q_out[idx], q_side[idx], q_long[idx]]
        params, cov = curve_fit(gaussian_3d, q_arr, np.ravel(corr[idx]), p0=list(START_VALUES),
                                sigma=np.ravel(corr_err[idx]), absolute_sigma=True)
        errors = np.sqrt(np.diag(cov))
        residual = corr[idx] - gaussian_3d(q_arr, *params).reshape(corr[idx].shape)
        r2.append(1.0 - np.var(residual) / np.var(corr[idx]))  # goodness of fit (:81-83)
        row = [q_cut]
        for value, err in zip(params, errors):
            row += [value, err]
        rows.append(row)
    out = np.array(rows)
    return (out, np.array(r2)) if return_rsquared else out


def correlation_table(table5, rel_floor: float = 0.0):
    """8-column table (the layout `fit_table` indexes) from the 5-column file the analysis writes
    today: columns 0-2 the bin's mean q, 3 the numerator sum of cos(q.x), 4 the denominator scaled
    by the pair ratio.  Column 6 = numerator / denominator (0 where the denominator is empty),
    column 7 = 1/sqrt(denominator) as a Poisson estimate of its statistical error (the reference
    gets its errors from event-to-event fluctuations, which a single file does not carry).
    Also returns the mask of bins with a non-empty denominator."""
    t = np.asarray(table5, dtype=np.float64)
    if t.ndim != 2 or t.shape[1] != 5:
        raise ValueError("expected the 5-column HBT_correlation_function_KT_*.dat layout")
    out = np.zeros((t.shape[0], 8))
    out[:, 0:3] = t[:, 0:3]
    out[:, 3] = t[:, 3]
    out[:, 4] = t[:, 4]
    ok = t[:, 4] > 0.0
    out[ok, 6] = t[ok, 3] / t[ok, 4]
    out[ok, 7] = np.maximum(1.0 / np.sqrt(t[ok, 4]), rel_floor)
    return out, ok


def fit_dat_file(path: str, q_cut_max_list: Sequence[float] = Q_CUT_MAX_LIST, out_path: Optional[str] = None):
    """Radii from one HBT_correlation_function_KT_*.dat (plain or .gz); writes HBT_radii_KT_*.dat
    beside it (same rows and header as the reference's dataset)."""
    table, ok = correlation_table(np.loadtxt(path))
    radii = fit_table(table, q_cut_max_list, mask=ok)
    if out_path is None:
        base = os.path.basename(path).replace("HBT_correlation_function", "HBT_radii")
        if base.endswith(".gz"):
            base = base[:-3]
        out_path = os.path.join(os.path.dirname(os.path.abspath(path)), base)
    np.savetxt(out_path, radii, fmt="%.10e", delimiter="  ", header=HEADER[2:])
    return radii


def fit_database(datafile_name: str, KT_cut_list: Sequence[str] = KT_CUT_LIST, verbose: bool = True) -> None:
    """The reference's own flow (`:45-100`): for every event group of the HDF5 database and every
    K_T cut, replace the dataset HBT_radii_KT_<cut>.dat by a fresh fit of
    HBT_correlation_function_KT_<cut>.dat."""
    import h5py  # not a dependency of anything else in this package

    database = h5py.File(datafile_name)
    for event in list(database.keys()):
        group = database[event]
        for KT_cut in KT_cut_list:
            filename = f"HBT_correlation_function_KT_{KT_cut}.dat"
            radii_name = f"HBT_radii_KT_{KT_cut}.dat"
            if radii_name in group.keys():
                del group[radii_name]
            if verbose:
                print(f"Analyzing {event}: {filename} ...")
            radii = fit_table(group.get(filename))
            if verbose:
                k = list(Q_CUT_MAX_LIST).index(0.1)
                print(f"R_out = {radii[k, 3]} fm, R_side = {radii[k, 5]} fm, R_long = {radii[k, 7]} fm")
            dataset = group.create_dataset(radii_name, data=radii, compression="gzip", compression_opts=9)
            dataset.attrs.create("header", np.bytes_(HEADER))
    database.close()


def main(argv=None) -> int:
    argv = sys.argv if argv is None else argv
    if len(argv) < 2:
        print(f"Usage: {argv[0]} database_file.h5 | HBT_correlation_function_KT_*.dat ...")
        return 0
    for name in argv[1:]:
        if name.endswith(".h5") or name.endswith(".hdf5"):
            fit_database(name)
        else:
            radii = fit_dat_file(name)
            k = list(Q_CUT_MAX_LIST).index(0.1)
            print(f"{name}: R_out = {radii[k, 3]:.4f} fm, R_side = {radii[k, 5]:.4f} fm, R_long = {radii[k, 7]:.4f} fm")
    return 0


if __name__ == "__main__":
    sys.exit(main())
