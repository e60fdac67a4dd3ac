"""BASELINE.json configs[4]: sharded rollout sweep, 64 M boards in total over the ranks, 16-step episodes (max_steps = 16), one
policy-gradient update with a single gradient all-reduce.  The same leg runs inside bench.py at N >= 2 (`sharded_sweep`); this
driver runs it alone.  Launch under torchrun; with one process the per-GPU share of an 8-GPU run (8 M boards) is used."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, b2048
from b2048 import dist as bd

info = bd.init_distributed("nccl")
dev = torch.device("cuda", info.local_rank)
torch.cuda.set_device(dev)
total = int(sys.argv[1]) if len(sys.argv) > 1 else (64 << 20 if info.world_size > 1 else 8 << 20)
H = int(sys.argv[2]) if len(sys.argv) > 2 else 16
r = b2048.bench_sharded_sweep(dev, total_boards=total, info=info, horizon=H, iters=3)
if info.rank == 0:
    print(json.dumps(r))
if info.is_distributed:
    dist.destroy_process_group()
