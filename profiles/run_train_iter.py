"""REINFORCE iteration (BASELINE.json configs[2]: 65,536 boards, rollout to termination + update) on cuda:0.
Used for the per-kernel launch lists under profiles/ (run plain first, then under ncu)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b2048
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
r = b2048.bench_train_iter(torch.device("cuda", 0), boards=65536, iters=iters, precision=prec)
print(json.dumps(r))
