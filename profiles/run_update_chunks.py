"""Update time of the REINFORCE iteration (65,536 episodes, ~7.7 M stored steps) against the `chunk` argument of
update_from_rollout (samples per b2048_mlp_backward pipeline launch).  usage: run_update_chunks.py [boards] [critic 0/1]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b2048
from b2048.rollout_bench import RUNNER_ENV
boards = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
critic = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
dev = torch.device("cuda", 0)
env = b2048.Batched2048Env(boards, b2048.Game2048EnvConfig(**RUNNER_ENV), device=dev, seed=0xB200)
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig(gamma=0.99, learning_rate=1e-4, baseline_mode="batch", model_seed=0,
                                                        use_critic=critic, critic_learning_rate=5e-4))
ro = agent.rollout_many(env, precision="auto")
out = {"samples": int(ro.length.sum().item())}
for chunk in (1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    st = agent.save_state()
    ts = []
    for _ in range(4):
        agent.load_state(st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        info = agent.update_from_rollout(ro, chunk=chunk)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out[str(chunk)] = {"ms": min(ts[1:]), "grad_norm": info["actor_grad_norm"]}
print(json.dumps(out))
