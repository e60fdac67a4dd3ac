"""Small pass over the tensor-core kernels for compute-sanitizer (memcheck): fused rollout, the default (split-fp16) update in
its chunked form (actor and critic, value forward) and as the persistent pipeline, the single-bf16 opt-in, the symmetry
kernel.  Sizes are the smallest that take each path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, b2048
torch.cuda.set_device(0)
kw = dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5, max_steps=24)
env = b2048.Batched2048Env(4096 + 37, b2048.Game2048EnvConfig(**kw), seed=3)
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig(use_critic=True, optimizer="adam", baseline_mode="batch_norm", augmentation=False))
ro = agent.rollout_many(env, precision=1)
info = agent.update_from_rollout(ro)                    # float32-grade tensor-core path; n < 4 x 148 tiles: chunked kernels
mode = info["precision"]
info1 = agent.update_from_rollout(ro, precision=1)      # single-bf16 opt-in
ro2 = agent.rollout_many(env, horizon=8, precision=1)
from b2048 import symmetry
ro3 = symmetry.augment_rollout(ro2)
# the persistent pipeline: >= 4 x 148 tiles of 128 samples, more tiles than ring slots would need n > 36,864: take ~80 K samples
env2 = b2048.Batched2048Env(4096, b2048.Game2048EnvConfig(**kw), seed=4)
agent2 = b2048.ReinforceAgent(env2, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                              b2048.ReinforceAgentConfig(baseline_mode="batch"))
ro4 = agent2.rollout_many(env2, precision=1)
info2 = agent2.update_from_rollout(ro4)
torch.cuda.synchronize()
print("memcheck case OK", ro.T, mode, info["actor_grad_norm"], info["critic_grad_norm"], info1["actor_grad_norm"], ro3.B,
      int(ro4.length.sum()), info2["actor_grad_norm"])
# reward shaping on the fast paths (fused rollout kernel, shaped variant) and the shared-trunk agent's two descriptor views
kw_s = dict(kw, empty_tile_reward=0.05, merge_reward=0.3, bonus_mode="log2", bonus_scale=2.0, endgame_penalty=-7.5)
env3 = b2048.Batched2048Env(4096 + 5, b2048.Game2048EnvConfig(**kw_s), seed=5)
agent3 = b2048.SharedTrunkActorCritic(env3, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                      b2048.ReinforceAgentConfig(baseline_mode="batch_norm", optimizer="adam"), gae_lambda=0.9)
ro5 = agent3.rollout_many(env3, precision=1)
info3 = agent3.update_from_rollout(ro5)
env4 = b2048.Batched2048Env(40000, b2048.Game2048EnvConfig(**kw_s), seed=6)
env4.reset_many()
for _ in range(3):
    env4.step_many(action_mode="random_legal", auto_reset=True)
env4.step_many_n(5, action_mode="random_legal", auto_reset=True)
torch.cuda.synchronize()
print("memcheck case 2 OK", ro5.T, info3["actor_grad_norm"])
