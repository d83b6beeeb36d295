#!/usr/bin/env python3
"""Fingerprint of the SHIPPED 5-level deck (tests/decks/ioniz_sphere_hires.athinput: 80^3/52^3/96^3/176^3/320^3,
39.8 M zones) after K steps of the reference's MPI branch run as one rank (oracle/_ref/ioniz_smr_mpi, oracle/mpi_stub).

The dump is 2.2 GB, so what is committed is tests/golden/hires5_k<K>.json: SHA-256 of the raw FP64 bytes of every
conserved variable and of EdgeFlux on every level, plus time, dt, the radiation iteration counts and a few sums
(to localise a mismatch).  A bit-identical run reproduces every hash.  ~10 min of one host core, ~6 GB of memory.

    python tests/golden/make_hires5_fingerprint.py [K=1] [existing_dump.rst]
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from athena_b200.restart import read_restart  # noqa: E402

VARS = ("d", "M1", "M2", "M3", "E", "s", "edgeflux")


def fingerprint(d, rad, vars_=VARS):
    from athena_b200.restart import fingerprint as fp
    return fp(d, rad, vars_=vars_)


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    if len(sys.argv) > 3:
        d = read_restart(sys.argv[2])
        rad = [int(a) for a in re.findall(r"Radiation done in (\d+) iterations", open(sys.argv[3]).read())]
    else:
        exe = os.path.join(ROOT, "oracle", "_ref", "ioniz_smr_mpi", "athena")
        with tempfile.TemporaryDirectory() as tmp:
            r = subprocess.run([exe, "-i", os.path.join(ROOT, "tests", "decks", "ioniz_sphere_hires.athinput"), f"time/nlim={k}"],
                               cwd=tmp, capture_output=True, text=True, check=True)
            rst = sorted(f for f in os.listdir(os.path.join(tmp, "id0")) if f.endswith(".rst"))
            d = read_restart(os.path.join(tmp, "id0", rst[-1]))
        rad = [int(a) for a in re.findall(r"Radiation done in (\d+) iterations", r.stderr)]
    fp = fingerprint(d, rad)
    path = os.path.join(HERE, f"hires5_k{k}.json")
    json.dump(fp, open(path, "w"), indent=1)
    print(path, "levels", len(fp["levels"]), "rad", rad)


if __name__ == "__main__":
    main()
