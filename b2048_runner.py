#!/usr/bin/env python
"""CLI of the batched trainer / evaluator (the reference's `python runner.py -conf cfg.json`, runner.py:854-893):

    python b2048_runner.py -conf cfg.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 b2048_runner.py -conf cfg.json
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import b2048  # noqa: E402
from b2048 import trainer  # noqa: E402

if __name__ == "__main__":
    trainer.main()
