"""BASELINE.json configs[4]: sharded rollout sweep, 64 M boards in total over the ranks (8 M per GPU on 8 GPUs), fixed
horizon (16 steps, reset-on-done), one policy-gradient update with the NCCL all-reduce.  Run under torchrun; with one
process the per-GPU share of an 8-GPU run (8 M boards) is used."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, b2048
from b2048 import dist as bd
from b2048.rollout_bench import RUNNER_ENV

info = bd.init_distributed("nccl")
dev = torch.device("cuda", info.local_rank)
torch.cuda.set_device(dev)
total = int(sys.argv[1]) if len(sys.argv) > 1 else (64 << 20 if info.world_size > 1 else 8 << 20)
H = int(sys.argv[2]) if len(sys.argv) > 2 else 16
env = bd.make_sharded_env(total, b2048.Game2048EnvConfig(**RUNNER_ENV), info, seed=0xB200, device=dev)
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig(gamma=0.99, learning_rate=1e-4, baseline_mode="batch", model_seed=0))
out = []
for it in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    if info.is_distributed:
        dist.barrier()
    e[0].record()
    ro = agent.rollout_many(env, horizon=H, precision=1, reset=(it == 0))
    e[1].record()
    upd = bd.sharded_update(agent, ro, info, total_episodes=total)
    e[2].record()
    torch.cuda.synchronize()
    t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], dtype=torch.float64, device=dev)
    bd.allreduce_max_(t)
    out.append((float(t[0]), float(t[1]), upd["actor_grad_norm"]))
if info.rank == 0:
    ro_ms, up_ms, gn = out[-1]
    print(json.dumps({"config": "BASELINE.json configs[4]", "n_gpus": info.world_size, "boards_total": total,
                      "boards_per_gpu": env.num_envs, "horizon": H, "rollout_ms": ro_ms, "update_ms": up_ms,
                      "rollout_steps_per_s": total * H / (ro_ms * 1e-3), "samples_per_s_update": total * H / (up_ms * 1e-3),
                      "actor_grad_norm": gn, "iters": out}))
if info.is_distributed:
    dist.destroy_process_group()
