import sys, os
sys.path.insert(0, os.getcwd())
import torch, b2048
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
boards = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
print(b2048.bench_rollout(dev, boards=boards, steps=steps, warmup=4, precision=1))
