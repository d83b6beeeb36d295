#!/usr/bin/env python3
"""One-off check, not part of the pytest suite (minutes of host CPU, ~20 GB of host memory): the SHIPPED 5-level
ioniz_sphere_hires deck (80^3/52^3/96^3/176^3/320^3 = 39.8 M zones) for K steps, reference vs drop-in.

  reference : oracle/_ref/ioniz_smr_mpi/athena  -- the reference's MPI branch as one rank against oracle/mpi_stub
              (its serial build segfaults below level 1)
  drop-in   : build/dropin/ioniz_smr/athena     -- the unmodified driver linked against libathena_b200 (one B200)

Prints per-level max relative differences and the radiation iteration counts; exit code 0 only if every level is
bit-identical.  Usage (GPU box):  python tests/live_five_level_check.py [K=1]
"""
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from athena_b200.restart import read_restart  # noqa: E402


def run(exe, nlim, mpi):
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [exe, "-i", os.path.join(ROOT, "tests", "decks", "ioniz_sphere_hires.athinput"), f"time/nlim={nlim}"]
        t0 = time.time()
        r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
        wall = time.time() - t0
        if r.returncode != 0:
            raise SystemExit(f"{exe} failed:\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}")
        out = os.path.join(tmp, "id0") if mpi else tmp
        rst = sorted(f for f in os.listdir(out) if f.endswith(".rst"))
        d = read_restart(os.path.join(out, rst[-1]))
    rad = [int(a) for a in re.findall(r"Radiation done in (\d+) iterations", r.stderr)]
    return d, rad, wall


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    ref, rad_r, w_r = run(os.path.join(ROOT, "oracle", "_ref", "ioniz_smr_mpi", "athena"), k, True)
    b2, rad_b, w_b = run(os.path.join(ROOT, "build", "dropin", "ioniz_smr", "athena"), k, False)
    print(f"steps {k}: reference wall {w_r:.1f} s, drop-in wall {w_b:.1f} s (both include set-up and two restart dumps)")
    print("radiation iterations  reference:", rad_r)
    print("radiation iterations  drop-in  :", rad_b)
    print("time", ref["time"], b2["time"], "dt", ref["dt"], b2["dt"])
    ok = rad_r == rad_b and ref["time"] == b2["time"] and ref["dt"] == b2["dt"]
    for lev, (ga, gb) in enumerate(zip(ref["grids"], b2["grids"])):
        for v in ("d", "M1", "M2", "M3", "E", "s", "edgeflux"):
            a, b = ga[v], gb[v]
            same = np.array_equal(a, b, equal_nan=True)
            ok = ok and same
            with np.errstate(all="ignore"):
                rel = np.nanmax(np.abs(a - b) / np.maximum(np.abs(a), 1e-300)) if not same else 0.0
            print(f"level {lev} {v:8s} shape {a.shape} identical {same} max rel {rel:.3e} differing {(a != b).sum()}")
    print("FIVE_LEVEL_PARITY", "OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
