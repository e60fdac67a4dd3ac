"""Rollout-to-termination time of 65,536 episodes vs the compaction interval (check_every)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, b2048
from b2048.rollout_bench import RUNNER_ENV
dev = torch.device("cuda", 0)
env = b2048.Batched2048Env(65536, b2048.Game2048EnvConfig(**RUNNER_ENV), device=dev, seed=0xB200)
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig(model_seed=0))
for ce in (8, 16, 24, 32, 48, 64, 32):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        ro = agent.rollout_many(env, precision=1, check_every=ce)
        e1.record(); torch.cuda.synchronize()
    print(ce, round(e0.elapsed_time(e1), 3), "ms  T =", ro.T, " live steps", int(ro.length.sum()))
