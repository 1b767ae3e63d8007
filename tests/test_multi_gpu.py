"""GPU suite, needs >= 2 devices (skipped on a 1-GPU box): real NCCL run of the sharded path -- frame broadcast, ownership
sharded integration, distributed re-mesh with ghost-chunk exchange, mesh gather -- against the CPU oracle; then the fused
multi-frame path fed by sharded ingest + one all-gather."""
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
    world = min(_gpus(), 4)
    port = 29700 + os.getpid() % 200
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_OK world=%d" % world in out.stdout
    # second phase of the worker: fused multi-frame batches fed by sharded ingest + one all-gather (bench.py's N > 1 path)
    assert "MULTI_GPU_BATCH_OK world=%d" % world in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
