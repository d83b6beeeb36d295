"""Run under torchrun on N GPUs: SMR with every level cut along x3 on common planes (driver.SmrRunSlabs) must equal
the reference's dumps (golden cases) and the one-GPU SmrRun bit for bit.  The shipped 2-level ioniz_sphere cut
(80^3 + 52^3, static potential + Userwork reset; problem() data generated per slab) is compared with the one-GPU run
to 1e-12: the potential is tabulated from slab-local MinX as in the reference's own decomposition.
Prints 'MULTI_GPU_SMR_OK <case>' per case on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cases import DECKS, VARS  # noqa: E402
from smr_cases import SmrCase  # noqa: E402
from athena_b200.athinput import ParFile  # noqa: E402
from athena_b200.driver import SmrRun, SmrRunSlabs  # noqa: E402


def compare(tag, R, single, nlev, rank, golden=None, rtol=0.0):
    ok = True
    msgs = []
    for n in range(nlev):
        full = R.gather_level(n)
        ef = R.gather_level(n, "edgeflux") if R.ion else None
        if rank != 0:
            continue
        ref = single.grids[n].download_active()
        for v in full:
            same = np.array_equal(full[v], ref[v], equal_nan=True) if rtol == 0 else np.allclose(full[v], ref[v], rtol=rtol, atol=0, equal_nan=True)
            if not same:
                ok = False; msgs.append(f"L{n}:{v} vs 1gpu")
            if golden is not None and not np.array_equal(full[v], golden.levels[n]["U1"][v], equal_nan=True):
                ok = False; msgs.append(f"L{n}:{v} vs golden")
        if ef is not None:
            ref_ef = single.grids[n].download_edgeflux()
            # the last edge row/column of a level is never written by a ray; compare the ray rows
            same = np.array_equal(ef[:-1], ref_ef[:-1]) if rtol == 0 else np.allclose(ef[:-1], ref_ef[:-1], rtol=rtol, atol=0)
            if not same:
                ok = False; msgs.append(f"L{n}:edgeflux vs 1gpu")
    if rank == 0:
        if rtol == 0:
            ok = ok and R.time == single.time and R.dt == single.dt and R.rad_log == single.rad_log
        else:
            ok = ok and abs(R.time - single.time) <= rtol * single.time and [r[:3] for r in R.rad_log] == [r[:3] for r in single.rad_log]
        if not ok:
            msgs.append(f"time {R.time!r} {single.time!r} dt {R.dt!r} {single.dt!r} rad {R.rad_log[:4]} {single.rad_log[:4]}")
        print(("MULTI_GPU_SMR_OK " if ok else "MULTI_GPU_SMR_FAIL ") + f"{tag} world={dist.get_world_size()} " + "; ".join(msgs), flush=True)
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ok = True
    for name in sys.argv[1:]:
        if name == "ioniz_hires2":
            par = ParFile.par_open(os.path.join(DECKS, "ioniz_sphere_hires.athinput"))
            par.par_cmdline(["job/num_domains=2"])
            R = SmrRunSlabs(par, "ioniz_sphere", None, dist, device=local)
            R.run(2)
            single = SmrRun(par, "ioniz_sphere", None, device=local) if rank == 0 else None
            if single is not None:
                single.run(2)
            ok = compare(name, R, single, 2, rank, rtol=1e-12) and ok
        else:
            case = SmrCase(name)
            R = SmrRunSlabs(case.par, case.build0, [lv["U0"] for lv in case.levels], dist, device=local)
            R.run(case.nsteps)
            single = None
            if rank == 0:
                single = SmrRun(case.par, case.build0, *[lv["U0"] for lv in case.levels], device=local)
                single.run(case.nsteps)
            ok = compare(name, R, single, case.nlevels, rank, golden=case) and ok
        dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
