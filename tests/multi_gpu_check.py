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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ok = True
    for name in sys.argv[1:]:
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
