"""GPU suite, needs >= 2 devices (skipped on a 1-GPU box): real NCCL run of the multi-GPU data plane of the C ABI --
chs_comm_init, chs_integrate_batch_distributed (sharded ingest + in-place all-gather + ownership-sharded fused integration),
chs_update_meshes_distributed (dirty-set union, device-side ghost exchange, mesh gather) -- against the CPU oracle.
The log of a run on 2 and 8 B200s is kept under profiles/ (r02_multi_gpu_test_*.log)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_nccl_sharded_integration_and_meshing():
    world = max(w for w in (2, 4, 8) if w <= _gpus())
    port = 29700 + os.getpid() % 200
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_OK world=%d" % world in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("MULTI_GPU_CASE_OK") == 3
    print(out.stdout)
