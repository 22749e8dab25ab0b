"""Kept for the command line of round 1: the checks now live in tests/multi_gpu_worker.py (run by tests/test_gpu_multi.py).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py"""
import os
import runpy

runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "multi_gpu_worker.py"),
               run_name="__main__")
