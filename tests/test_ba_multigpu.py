"""Sharded global bundle adjustment on 2 GPUs equals the single-GPU solve (needs >= 2 GPUs; run with
`gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_global_ba_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    for attempt in range(2):          # (a rendezvous right after another torchrun job on the box can fail once: fresh port, once more)
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "ba_multi_check.py")],
                           capture_output=True, text=True, timeout=600)
        if r.returncode == 0 and "MULTI_OK" in r.stdout:
            break
    assert r.returncode == 0 and "MULTI_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
