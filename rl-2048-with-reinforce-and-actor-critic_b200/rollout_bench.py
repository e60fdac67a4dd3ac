"""Secondary benchmark leg: policy-rollout steps/s (BASELINE.json configs[2] shape: MLP 16-256-256-4 policy,
65,536 boards, one policy_step + one step_many launch per rollout step)."""
from __future__ import annotations

import numpy as np
import torch

from .MLP import MLPConfig
from .batched_env import Batched2048Env, Game2048EnvConfig
from .reinforce_agent import ReinforceAgent, ReinforceAgentConfig

RUNNER_ENV = dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5,
                  bonus_mode="off", max_steps=1024)
FLOPS_PER_STEP = 2 * (16 * 256 + 256 * 256 + 256 * 4)


def bench_rollout(dev, boards: int = 65536, steps: int = 64, warmup: int = 8, precision: int = 0, gid0: int = 0):
    env = Batched2048Env(boards, Game2048EnvConfig(**RUNNER_ENV), device=dev, seed=0xB200, gid0=gid0)
    agent = ReinforceAgent(env, MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                           ReinforceAgentConfig(gamma=0.99, learning_rate=1e-4, baseline_mode="batch", model_seed=0))
    env.reset_many()
    act = torch.zeros(boards, dtype=torch.uint8, device=dev)

    def one():
        agent.policy_step(env.board, env.flags, act, env.seed, env.gid0, env.t + 1, precision=precision)
        env.step_many(act, auto_reset=True)

    for _ in range(warmup):
        one()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rate = boards * steps / (ms * 1e-3)
    return {"metric": "policy-rollout steps/s", "value": rate, "unit": "rollout-steps/s", "boards": boards, "steps": steps,
            "ms_per_step": ms / steps, "mlp": "16-256-256-4 ReLU", "precision": "fp32 CUDA cores" if precision == 0 else "bf16 tcgen05",
            "flops_per_step": FLOPS_PER_STEP, "achieved_tflops": rate * FLOPS_PER_STEP / 1e12,
            "gpu_launches": 2 * steps}
