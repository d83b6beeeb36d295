"""Reader for the reference's restart (.rst) dumps -- the only full-state FP64 dump format it has.

Layout (/root/reference/src/restart.c:463-760): an ASCII copy of the parameter file terminated by
`<par_end>`, then labelled raw arrays: `N_STEP` (int), `TIME`, `TIME_STEP` (double), and per Grid
`DENSITY`, `1-MOMENTUM`, `2-MOMENTUM`, `3-MOMENTUM`, `ENERGY` (active zones, k-j-i order, i fastest),
`EDGEFLUX` ((Nx3+1)(Nx2+1)(Nx1+1) doubles, ion builds only), `SCALAR n`.
Used for parity artefacts and by the host driver's restart hook.
"""
from __future__ import annotations

import numpy as np

from .athinput import ParFile


def _take_label(buf: bytes, pos: int, label: bytes) -> int:
    i = buf.find(label, pos)
    if i < 0:
        raise ValueError(f"restart file: label {label!r} not found")
    return i + len(label)


def read_restart(path: str, level_shapes=None) -> dict:
    """Returns dict(par=ParFile, nstep, time, dt, grids=[dict(d,M1,M2,M3,E,s?,edgeflux?)]).

    level_shapes: optional list of (Nx3,Nx2,Nx1) per Grid in file order; by default taken from the
    embedded parameter file's <domainN> blocks (single Grid per Domain, i.e. serial runs)."""
    buf = open(path, "rb").read()
    end = buf.find(b"<par_end>")
    par = ParFile.from_text(buf[:end].decode("latin-1"))
    pos = _take_label(buf, end, b"N_STEP\n")
    nstep = int(np.frombuffer(buf, dtype=np.int32, count=1, offset=pos)[0])
    pos = _take_label(buf, pos, b"\nTIME\n")
    time = float(np.frombuffer(buf, dtype=np.float64, count=1, offset=pos)[0])
    pos = _take_label(buf, pos + 8, b"\nTIME_STEP\n")
    dt = float(np.frombuffer(buf, dtype=np.float64, count=1, offset=pos)[0])
    pos += 8
    if level_shapes is None:
        nd = par.par_geti("job", "num_domains")
        level_shapes = []
        for n in range(1, nd + 1):
            b = f"domain{n}"
            level_shapes.append((par.par_geti(b, "Nx3"), par.par_geti(b, "Nx2"), par.par_geti(b, "Nx1")))
    grids = []
    for (n3, n2, n1) in level_shapes:
        g = {}
        ncell = n1 * n2 * n3
        for key, label in (("d", b"\nDENSITY\n"), ("M1", b"\n1-MOMENTUM\n"), ("M2", b"\n2-MOMENTUM\n"),
                           ("M3", b"\n3-MOMENTUM\n"), ("E", b"\nENERGY\n")):
            pos = _take_label(buf, pos, label)
            g[key] = np.frombuffer(buf, dtype=np.float64, count=ncell, offset=pos).reshape(n3, n2, n1).copy()
            pos += 8 * ncell
        if buf.startswith(b"\nEDGEFLUX\n", pos):
            pos += len(b"\nEDGEFLUX\n")
            ne = (n1 + 1) * (n2 + 1) * (n3 + 1)
            g["edgeflux"] = np.frombuffer(buf, dtype=np.float64, count=ne, offset=pos).reshape(n3 + 1, n2 + 1, n1 + 1).copy()
            pos += 8 * ne
        if buf.startswith(b"\nSCALAR 0\n", pos):
            pos += len(b"\nSCALAR 0\n")
            g["s"] = np.frombuffer(buf, dtype=np.float64, count=ncell, offset=pos).reshape(n3, n2, n1).copy()
            pos += 8 * ncell
        grids.append(g)
    return dict(par=par, nstep=nstep, time=time, dt=dt, grids=grids)


def fingerprint(d: dict, rad, vars_=("d", "M1", "M2", "M3", "E", "s", "edgeflux")) -> dict:
    """SHA-256 of the raw FP64 bytes of every variable of every Grid of a dump `d` (read_restart's layout), plus time, dt
    and the radiation iteration counts `rad`: what tests/golden/*.json commit of the REFERENCE's dumps that are too large
    to commit themselves.  A bit-identical run reproduces every hash (-0.0 is folded onto +0.0: equal values, equal bytes)."""
    import hashlib
    out = {"time": float(d["time"]).hex(), "dt": float(d["dt"]).hex(), "nstep": int(d["nstep"]), "rad": rad, "levels": []}
    for g in d["grids"]:
        lev = {}
        for v in vars_:
            a = np.ascontiguousarray(g[v], dtype=np.float64) + 0.0
            lev[v] = {"shape": list(a.shape), "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
                      "sum": float(np.nansum(a)), "absmax": float(np.nanmax(np.abs(a))), "nan": int(np.isnan(a).sum())}
        out["levels"].append(lev)
    return out
