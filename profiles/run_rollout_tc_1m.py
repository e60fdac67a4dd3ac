import sys, os
sys.path.insert(0, os.getcwd())
import torch, b2048
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
print(b2048.bench_rollout(dev, boards=1 << 20, steps=4, warmup=2, precision=1))
