"""Small pass over the tensor-core kernels for compute-sanitizer (memcheck): fused rollout, TC update (actor and
critic), symmetry kernel.  Sizes are the smallest that take the tensor-core paths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, b2048
torch.cuda.set_device(0)
kw = dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5, max_steps=24)
env = b2048.Batched2048Env(4096 + 37, b2048.Game2048EnvConfig(**kw), seed=3)
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig(use_critic=True, optimizer="adam", baseline_mode="batch_norm", augmentation=False))
ro = agent.rollout_many(env, precision=1)
info = agent.update_from_rollout(ro, precision=1)
ro2 = agent.rollout_many(env, horizon=8, precision=1)
from b2048 import symmetry
ro3 = symmetry.augment_rollout(ro2)
torch.cuda.synchronize()
print("memcheck case OK", ro.T, info["actor_grad_norm"], info["critic_grad_norm"], ro3.B)
