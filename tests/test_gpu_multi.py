"""Multi-GPU checks; skipped on a single-GPU box (the driver's `-m gpu` tier).  Run under torchrun by
tools/gpu_multi.sh / tools/peer_allreduce_check.py on 2-8 GPUs."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_allreduce_equals_nccl():
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        str(ROOT / "tools" / "peer_allreduce_check.py")], capture_output=True, text=True, timeout=600)
    assert "PEER_ALLREDUCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_adam_equals_replicated_adam():
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29534",
                        str(ROOT / "tools" / "sharded_adam_check.py"), "--P", "200000"], capture_output=True, text=True,
                       timeout=600)
    assert "SHARDED_ADAM_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_batch_step_on_gpus_equals_single_process_accumulation():
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29535",
                        str(ROOT / "tools" / "sharded_step_check.py")], capture_output=True, text=True, timeout=600)
    assert "SHARDED_STEP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
