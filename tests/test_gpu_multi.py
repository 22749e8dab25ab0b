"""Multi-GPU parity on real devices: spawns tests/multi_gpu_worker.py under torchrun with every GPU of
the box (at most 8); skipped on a single-GPU box.  The CPU-side logic of the same paths is covered by the world-size-2
gloo test in tests/test_host_cpu.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_knn_and_data_parallel_finetuning_on_all_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 8)
    env = dict(os.environ)
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    checks = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    # sharded kNN, own all-reduce, DataParallelStep x2, GraphedSaladStep x2 (NCCL / peer), GraphedFinetuneStep x2
    assert len(checks) == 8 and all(c.get("ok", c.get("all_ranks")) for c in checks), checks
