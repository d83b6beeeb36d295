"""Multi-GPU parity (needs >= 2 GPUs on the box): see tests/multi_gpu_check.py."""
import os
import socket
import subprocess
import sys

import pytest

from cases import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_run_equals_one_gpu_run():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py"), "blast16_k6", "ifront32_k4", "ioniz24_k3",
           "ioniz256_k1", "blast256_k2"]        # the last two: BASELINE's full 256^3 over 2 slabs against the reference's fingerprint
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MULTI_GPU_OK") == 5, r.stdout


def test_two_gpu_smr_equals_one_gpu_and_reference():
    """every SMR level cut along x3 on common planes (driver.SmrRunSlabs): 2- and 3-level goldens, ion RT included"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cases = ["blastsmr_k4", "ifrontsmr_k3", "blastsmr3_k4", "ifrontsmr3_k3", "ioniz_hires2"]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_smr_check.py")] + cases
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MULTI_GPU_SMR_OK") == len(cases), r.stdout
