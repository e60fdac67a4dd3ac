"""Host-side enqueue time vs device time of rollout_many (is the launch loop host-bound?).
usage: run_rollout_host_time.py [boards] [onehot|default]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b2048
from b2048.rollout_bench import ONEHOT_ENV, RUNNER_ENV
boards = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
net = sys.argv[2] if len(sys.argv) > 2 else "onehot"
env = b2048.Batched2048Env(boards, b2048.Game2048EnvConfig(**(ONEHOT_ENV if net == "onehot" else RUNNER_ENV)), seed=1)
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 128, 64] if net == "onehot" else [256, 256], activation="ReLU",
                                                  init_distribution="HeNormal"), b2048.ReinforceAgentConfig(model_seed=0))
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    ro = agent.rollout_many(env, precision=1)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"iter {it}: T {ro.T} host enqueue {1e3 * (t1 - t0):.1f} ms, device {e0.elapsed_time(e1):.1f} ms, wall {1e3 * (t2 - t0):.1f} ms, "
          f"live steps {int(ro.length.sum())}")
