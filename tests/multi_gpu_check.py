"""Run under torchrun on N GPUs: the x3-slab decomposed run of a golden case must equal the 1-GPU run of the
same engine BIT FOR BIT (hydro: no reduction is reordered; ion: the cross-rank reductions are min/max/integer
sum), hence inherits its parity with the reference.  Prints 'MULTI_GPU_OK <case>' per case on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cases import Case, VARS  # noqa: E402
from athena_b200.driver import Run  # noqa: E402


def fullsize(name, rank, world, local):
    """BASELINE's full 256^3 grid cut into x3 slabs over the ranks must hash to the REFERENCE's dump (SHA-256 fingerprints in
    tests/golden/<name>.json, generated from the reference binary) -- N-GPU parity at full size, not only against one GPU."""
    import importlib.util
    import json
    from cases import ROOT, DECKS
    from athena_b200.athinput import ParFile
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
    spec = importlib.util.spec_from_file_location("fp5", os.path.join(ROOT, "tests", "golden", "make_hires5_fingerprint.py"))
    fp5 = importlib.util.module_from_spec(spec); spec.loader.exec_module(fp5)
    par = ParFile.par_open(os.path.join(DECKS, gold["deck"]))
    par.par_cmdline(gold["overrides"])
    R = Run(par, gold["build"], device=local, dist=dist)
    R.run(gold["nstep"])
    mine = R.G.download_active()
    names = [v for v in gold["levels"][0] if v != "edgeflux"]
    loc = torch.from_numpy(np.stack([mine[v] for v in names])).cuda()
    parts = [torch.empty_like(loc) for _ in range(world)]
    dist.all_gather(parts, loc)
    ef_parts = None
    if "edgeflux" in gold["levels"][0]:
        ef = torch.from_numpy(R.G.download_edgeflux()).cuda()
        ef_parts = [torch.empty_like(ef) for _ in range(world)]
        dist.all_gather(ef_parts, ef)
    good = True
    if rank == 0:
        full = torch.cat(parts, dim=1).cpu().numpy()
        g = {v: full[n] for n, v in enumerate(names)}
        if ef_parts is not None:      # ray rows of every slab, the last edge row from the top slab
            g["edgeflux"] = torch.cat([e[:-1] for e in ef_parts[:-1]] + [ef_parts[-1]], dim=0).cpu().numpy()
        fp = fp5.fingerprint(dict(time=R.time, dt=R.dt, nstep=R.nstep, grids=[g]), [r[0] for r in R.rad_log], vars_=tuple(gold["levels"][0]))
        good = fp["rad"] == gold["rad"] and fp["time"] == gold["time"] and fp["dt"] == gold["dt"]
        bad = [v for v, ref in gold["levels"][0].items() if fp["levels"][0][v]["sha256"] != ref["sha256"]]
        good = good and not bad
        print(("MULTI_GPU_OK " if good else "MULTI_GPU_FAIL ") + f"{name} world={world} full-size fingerprint of the reference; differing: {bad}", flush=True)
    return good


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ok = True
    for name in sys.argv[1:]:
        if name.endswith("256_k1") or name.endswith("256_k2") or name in ("ioniz96_k12", "ioniz80_k6"):     # fingerprints of the reference's dump
            ok = fullsize(name, rank, world, local) and ok
            dist.barrier()
            continue
        case = Case(name)
        R = Run(case.par, case.problem, device=local, dist=dist, U0=case.U0, riemann=case.riemann)
        R.run(case.nsteps)
        mine = R.G.download_active()
        nv = 5 + case.nscal
        loc = torch.from_numpy(np.stack([mine[v] for v in VARS[:nv]])).cuda()
        parts = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(parts, loc)
        if rank == 0:
            full = torch.cat(parts, dim=1).cpu().numpy()
            single = Run(case.par, case.problem, device=local, dist=None, U0=case.U0, riemann=case.riemann)
            single.run(case.nsteps)
            ref = single.G.download_active()
            same = all(np.array_equal(full[n], ref[v], equal_nan=True) for n, v in enumerate(VARS[:nv]))
            same = same and R.time == single.time and R.dt == single.dt and R.rad_log == single.rad_log
            gold = all(np.allclose(full[n], case.U1[v], rtol=1e-9, atol=0) for n, v in enumerate(VARS[:nv]) if v in ("d", "E", "s"))
            print(("MULTI_GPU_OK " if (same and gold) else "MULTI_GPU_FAIL ") + f"{name} world={world} bit_equal_to_1gpu={same} golden={gold}", flush=True)
            ok = ok and same and gold
        dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
