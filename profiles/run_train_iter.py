"""REINFORCE iteration (BASELINE.json configs[2]: 65,536 boards, rollout to termination + update) on cuda:0.
Used for the per-kernel launch lists under profiles/ (run plain first, then under ncu)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b2048
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
boards = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
critic = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
r = b2048.bench_train_iter(torch.device("cuda", 0), boards=boards, iters=iters, precision=prec, use_critic=critic)
print(json.dumps(r))
