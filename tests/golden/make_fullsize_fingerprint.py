#!/usr/bin/env python3
"""SHA-256 fingerprints of the REFERENCE's dump at BASELINE.json's full sizes (the dumps are 0.9 GB each, so only their
fingerprints are committed; same scheme as make_hires5_fingerprint.py):

    ioniz256_k1  configs[2]: ioniz_sphere unigrid 256^3, 1 step   (oracle/_ref/ioniz_sphere, ~2 min of one core, 11 GB)
    blast256_k2  configs[1]: blast 256^3 periodic, 2 steps        (oracle/_ref/blast)
    ifront64_k30 configs[0]: ifront 64^3 as shipped, 30 steps     (oracle/_ref/ifront)
    ioniz64s_k2 / blast64s_k4 / ifront64s_k3: 32x32x64 grids for the multi-GPU parity gate inside bench.py
    ioniz96_k12 / ioniz80_k6 / ioniz1100_k4: ioniz_sphere 96^3 x 12 steps, 80x48x40 x 6 steps, 1100x12x10 x 4 steps for the sparse sub-cycle

A bit-identical run reproduces every hash (tests/test_gpu_longrun.py).   python tests/golden/make_fullsize_fingerprint.py
"""
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_hires5_fingerprint import fingerprint, read_restart  # noqa: E402

N256 = ["job/num_domains=1", "domain1/Nx1=256", "domain1/Nx2=256", "domain1/Nx3=256"]
CASES = {
    "ioniz256_k1": ("ioniz_sphere", "ioniz_sphere.athinput", 1, N256),
    "blast256_k2": ("blast", "blast.athinput", 2, N256),
    "ifront64_k30": ("ifront", "ifront.athinput", 30, ["job/num_domains=1"]),      # configs[0]: the deck as shipped, 64^3
    # the parity gate bench.py runs before every timed region, on 1/2/4/8 GPUs: Nx3 = 64 gives every one of 8 x3 slabs
    # 8 layers (two ghost widths); seconds of one core
    "ioniz64s_k2": ("ioniz_sphere", "ioniz_sphere.athinput", 2, ["job/num_domains=1", "domain1/Nx1=32", "domain1/Nx2=32", "domain1/Nx3=64"]),
    "blast64s_k4": ("blast", "blast.athinput", 4, ["job/num_domains=1", "domain1/Nx1=32", "domain1/Nx2=32", "domain1/Nx3=64", "problem/radius=0.25"]),
    "ifront64s_k3": ("ifront", "ifront.athinput", 3, ["job/num_domains=1", "domain1/Nx1=32", "domain1/Nx2=32", "domain1/Nx3=64"]),
    # the sparse sub-cycle (csrc/ion.cu): a box most of whose chunks are inert ambient gas, over enough steps and sub-cycle
    # iterations that a wrongly re-used chunk would show (Nx1 = 96: three chunks per ray; Nx1 = 80: a ragged last chunk)
    "ioniz96_k12": ("ioniz_sphere", "ioniz_sphere.athinput", 12, ["job/num_domains=1", "domain1/Nx1=96", "domain1/Nx2=96", "domain1/Nx3=96"]),
    "ioniz80_k6": ("ioniz_sphere", "ioniz_sphere.athinput", 6, ["job/num_domains=1", "domain1/Nx1=80", "domain1/Nx2=48", "domain1/Nx3=40"]),
    # more than 32 chunks per ray (k_ion_walk scans the records of a ray 32 at a time)
    "ioniz1100_k4": ("ioniz_sphere", "ioniz_sphere.athinput", 4, ["job/num_domains=1", "domain1/Nx1=1100", "domain1/Nx2=12", "domain1/Nx3=10"]),
}


def main():
    for name, (build, deck, nlim, OVERRIDES) in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        exe = os.path.join(ROOT, "oracle", "_ref", build, "athena")
        with tempfile.TemporaryDirectory() as tmp:
            r = subprocess.run([exe, "-i", os.path.join(ROOT, "tests", "decks", deck), f"time/nlim={nlim}"] + OVERRIDES,
                               cwd=tmp, capture_output=True, text=True, check=True)
            rst = sorted(f for f in os.listdir(tmp) if f.endswith(".rst"))
            d = read_restart(os.path.join(tmp, rst[-1]))
        rad = [int(a) for a in re.findall(r"Radiation done in (\d+) iterations", r.stderr)]
        fp = fingerprint(d, rad, vars_=("d", "M1", "M2", "M3", "E") + (("s", "edgeflux") if build != "blast" else ()))
        fp["deck"], fp["overrides"], fp["build"] = deck, OVERRIDES, build
        json.dump(fp, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
        print(name, "rad", rad, "time", fp["time"])


if __name__ == "__main__":
    main()
