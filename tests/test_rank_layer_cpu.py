"""The control plane of the C driver's multi-GPU build, on CPU: shim/mpi_b200 (the one-node rank layer that stands in for the
MPI this image lacks: ab2run + the ~16 MPI-1 calls the reference driver makes) and the start-up of the `--enable-mpi`
drop-in up to the point where it needs a GPU.  The data plane (NCCL inside the library) is covered by tests/test_gpu_multi.py."""
import os
import subprocess
import tempfile

import pytest

from cases import ROOT, DECKS

MPI_DIR = os.path.join(ROOT, "atmospheric-athena_b200", "shim", "mpi_b200")

PROGRAM = r"""
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
int main(int argc, char **argv)
{
  int r, n, i, sub = -1;
  static char big[3000000];                       /* larger than one shared-memory slot: exercises the chunked broadcast */
  double x, s, mn = -1.0, red = -1.0;
  int m, mx, ranks[2] = {1, 2};
  MPI_Group wg, g;
  MPI_Comm c;
  MPI_Init(&argc, &argv);
  MPI_Comm_rank(MPI_COMM_WORLD, &r); MPI_Comm_size(MPI_COMM_WORLD, &n);
  if (r == 0) for (i=0; i<3000000; i++) big[i] = (char)(i*7);
  MPI_Bcast(big, 3000000, MPI_CHAR, 0, MPI_COMM_WORLD);
  for (i=0; i<3000000; i++) if (big[i] != (char)(i*7)) { printf("rank %d: bcast corrupt at %d\n", r, i); return 1; }
  x = r + 1.5; MPI_Allreduce(&x, &s, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  m = r; MPI_Allreduce(&m, &mx, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
  MPI_Reduce(&x, &red, 1, MPI_DOUBLE, MPI_MIN, 2 % n, MPI_COMM_WORLD);
  MPI_Comm_group(MPI_COMM_WORLD, &wg); MPI_Group_incl(wg, 2, ranks, &g); MPI_Comm_create(MPI_COMM_WORLD, g, &c);   /* init_mesh.c:700-760 */
  if (c != MPI_COMM_NULL) { double y = r; MPI_Allreduce(&y, &mn, 1, MPI_DOUBLE, MPI_MIN, c); MPI_Comm_rank(c, &sub); }
  if (r == 0) { int v = 42; MPI_Send(&v, 1, MPI_INT, 1, 7, MPI_COMM_WORLD); }
  if (r == 1) { int v = 0; MPI_Status st; MPI_Recv(&v, 1, MPI_INT, 0, 7, MPI_COMM_WORLD, &st); if (v != 42 || st.MPI_TAG != 7) return 2; }
  MPI_Barrier(MPI_COMM_WORLD);
  printf("rank %d/%d sum %g max %d sub %d min %g red %g\n", r, n, s, mx, sub, mn, red);
  if (argc > 1 && r == atoi(argv[1])) MPI_Abort(MPI_COMM_WORLD, 5);
  MPI_Finalize();
  return 0;
}
"""


@pytest.fixture(scope="module")
def built():
    tmp = tempfile.mkdtemp(prefix="ab2mpi_")
    open(os.path.join(tmp, "t.c"), "w").write(PROGRAM)
    for cmd in (["gcc", "-O2", "-c", os.path.join(MPI_DIR, "mpi_b200.c"), "-o", os.path.join(tmp, "mpi.o")],
                ["gcc", "-O2", "-I", MPI_DIR, os.path.join(tmp, "t.c"), os.path.join(tmp, "mpi.o"), "-o", os.path.join(tmp, "t"), "-lrt"],
                ["gcc", "-O2", os.path.join(MPI_DIR, "ab2run.c"), os.path.join(tmp, "mpi.o"), "-o", os.path.join(tmp, "ab2run"), "-lrt"]):
        subprocess.run(cmd, check=True, capture_output=True)
    return tmp


def test_rank_layer_collectives_four_ranks(built):
    r = subprocess.run([os.path.join(built, "ab2run"), "-np", "4", os.path.join(built, "t")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = sorted(l for l in r.stdout.splitlines() if l.startswith("rank"))
    assert len(lines) == 4
    for rank, l in enumerate(lines):
        assert l.startswith(f"rank {rank}/4 sum 12 max 3 ")                  # 1.5+2.5+3.5+4.5
        sub = {1: 0, 2: 1}.get(rank, -1)
        assert f"sub {sub} min {1 if sub >= 0 else -1} " in l
        assert l.endswith("red 1.5" if rank == 2 else "red -1")


def test_rank_layer_single_process_and_abort(built):
    r = subprocess.run([os.path.join(built, "t")], capture_output=True, text=True, timeout=60)      # no launcher: rank 0 of 1
    assert r.returncode == 0 and "rank 0/1 sum 1.5 max 0" in r.stdout
    r = subprocess.run([os.path.join(built, "ab2run"), "-np", "3", os.path.join(built, "t"), "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 5, (r.returncode, r.stdout, r.stderr)          # MPI_Abort's code comes back, nothing hangs


@pytest.mark.parametrize("cfg,deck", [("blast_mpi", "blast.athinput"), ("ioniz_sphere_mpi", "ioniz_sphere.athinput")])
def test_mpi_dropin_control_plane_reaches_the_device_boundary(cfg, deck):
    """Two ranks of the UNMODIFIED driver (--enable-mpi) + shim: athinput broadcast, Domain decomposition into x3 slabs and
    problem() run on CPU; the first hot-path call then needs a GPU and, without one, dies loudly (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_multi.py")
    exe = os.path.join(ROOT, "build", "dropin", cfg, "athena")
    run = os.path.join(ROOT, "build", "dropin", "ab2run")
    if not (os.path.exists(exe) and os.path.exists(run)):
        pytest.skip("build/dropin not built")
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [run, "-np", "2", exe, "-i", os.path.join(DECKS, deck), "domain1/Nx1=16", "domain1/Nx2=16", "domain1/Nx3=16",
               "domain1/NGrid_x3=2", "time/nlim=1"]
        r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True, timeout=120)
    out = r.stdout + r.stderr
    assert r.returncode != 0
    assert "ID_Comm_world: 1" in out and "out of 2 1 1" in out, out[-2000:]          # init_mesh.c saw two x3 slabs
    assert "no CUDA device" in out, out[-2000:]
