"""Actor-critic iteration on the reference's documented one-hot 272-256-128-64 networks on cuda:0.
usage: run_onehot_iter.py [boards] [iters]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b2048
boards = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
r = b2048.bench_train_iter(torch.device("cuda", 0), boards=boards, iters=iters, precision="auto", use_critic=True, network="onehot")
print(json.dumps(r))
