"""Multi-GPU path: element-wise partition + NCCL ghost exchange against one GPU doing the
whole mesh.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu() or _n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", [2, 4])
def test_two_rank_assembly_matches_single_gpu(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "mgpu_probe.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_PARALLELMAT_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_NATIVE_PARTITION_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    if world == 2:
        # the same with the streamed assembly forced on these small slabs: state chunks, element
        # ranges with the ghost-touching ones last, forward halo in between, residual chunks back
        env = dict(os.environ, A2DS_STREAM_CHUNKS="3", A2DS_STREAM_MIN_ELEMS="1")
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        assert "MGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
        assert "MGPU_NATIVE_PARTITION_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
