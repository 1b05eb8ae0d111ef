"""Regenerate the reference-derived fixtures under tests/golden/ (run in the build container,
where /root/reference is mounted; the GPU box only ever sees the committed .npz files).

Only *data columns* of the reference's sample files are extracted (no source code):

* exphydro_01013500.npz — first 10 000 rows of `/root/reference/data/exphydro/01013500.csv`
  (`tmean(C)`, `dayl(day)`, `prcp(mm/day)`, `flow(mm)`): the run behind the published table
  and NSE/PBIAS (`docs/src/get_start_en.md:288-301,379-387,453-468`).
* gr4j_sample.npz — `/root/reference/data/gr4j/sample.csv` columns `prec, pet, pr, q9, q1`
  (+ `prob, rout, qsim` for plausibility only): third-party GR4J simulation whose `pr → q9/q1`
  columns pin the unit-hydrograph semantics.
* m50_01013500.npz — first 2 000 rows of `/root/reference/data/m50/01013500.csv`
  (`Prcp, Temp, Lday, SnowWater, SoilWater`): forcing for the M50 test of
  `test/base/run_single_lumped.jl:192-203`.
* hbv_sample.npz — `/root/reference/data/hbv_edu/hbv_sample.csv` (`prec, temp, pet` and, for plausibility only, the
  third-party simulation's `qsim, snow, soil, s1, s2`; all 3652 rows).
"""
import csv
import os

import numpy as np

REF = "/root/reference/data"
OUT = os.path.dirname(os.path.abspath(__file__))


def read(path, cols, nrows=None):
    with open(path, newline="") as f:
        rd = csv.DictReader(f)
        rows = []
        for i, r in enumerate(rd):
            if nrows is not None and i >= nrows:
                break
            rows.append([float(r[c]) if r[c] not in ("", "NA", "missing") else 0.0 for c in cols])
    a = np.asarray(rows, dtype=np.float64)
    return {c: a[:, i] for i, c in enumerate(cols)}


def main():
    d = read(f"{REF}/exphydro/01013500.csv", ["tmean(C)", "dayl(day)", "prcp(mm/day)", "flow(mm)"], 10000)
    np.savez_compressed(f"{OUT}/exphydro_01013500.npz", temp=d["tmean(C)"], lday=d["dayl(day)"],
                        prcp=d["prcp(mm/day)"], flow=d["flow(mm)"])
    d = read(f"{REF}/gr4j/sample.csv", ["prec", "pet", "pr", "q9", "q1", "prob", "rout", "qsim"])
    np.savez_compressed(f"{OUT}/gr4j_sample.npz", **d)
    d = read(f"{REF}/m50/01013500.csv", ["Prcp", "Temp", "Lday", "SnowWater", "SoilWater"], 2000)
    np.savez_compressed(f"{OUT}/m50_01013500.npz", **d)
    d = read(f"{REF}/hbv_edu/hbv_sample.csv", ["prec", "temp", "pet", "qsim", "snow", "soil", "s1", "s2"])
    np.savez_compressed(f"{OUT}/hbv_sample.npz", **d)


if __name__ == "__main__":
    main()
