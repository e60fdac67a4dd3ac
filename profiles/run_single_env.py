import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, b2048
env = b2048.Game2048Env(b2048.Game2048EnvConfig(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5))
obs, info = env.reset(seed=1)
n = 0; t0 = time.perf_counter()
while time.perf_counter() - t0 < 3.0:
    legal = np.flatnonzero(obs["action_mask"])
    obs, r, term, trunc, info = env.step(int(legal[0]) if len(legal) else 0)
    n += 1
    if term or trunc:
        obs, info = env.reset(seed=n)
print("drop-in Game2048Env.step:", n / (time.perf_counter() - t0), "steps/s")
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"), b2048.ReinforceAgentConfig())
t0 = time.perf_counter(); steps = 0
for ep in range(8):
    tr = agent.run_episode(ep, ep + 100)
    steps += len(tr["actions"])
print("drop-in run_episode:", steps / (time.perf_counter() - t0), "steps/s")
