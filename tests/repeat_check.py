"""Single-GPU run of tests/run_dp_check.py (a one-rank NCCL group): the check's two passes over the same batch and state
must give the same gradient whether or not the bucketed all-reduce runs. Usage: python tests/repeat_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("WORLD_SIZE", "1")
os.environ.setdefault("RANK", "0")
os.environ.setdefault("LOCAL_RANK", "0")
torch.cuda.set_device(0)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % os.environ.get("PP_PORT", "29513"), rank=0, world_size=1)
import run_dp_check  # noqa: E402

run_dp_check.main()
