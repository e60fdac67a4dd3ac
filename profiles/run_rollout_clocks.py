import sys, os
sys.path.insert(0, os.getcwd())
import torch, b2048
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
b2048.debug_set("tc_clocks", True)
print(b2048.bench_rollout(dev, boards=65536, steps=64, warmup=4, precision=1))
