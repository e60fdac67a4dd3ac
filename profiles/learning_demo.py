"""Learning demonstration on the GPU engine: REINFORCE (Adam) with 8192 parallel episodes per batch."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import torch, b2048
from b2048 import trainer
cfg = trainer.merge_config({
    "agent": {"optimizer": "adam", "learning_rate": 1e-3, "baseline_mode": "batch_norm", "gamma": 0.99},
    "train": {"batch_size": 8192, "num_batches": int(sys.argv[1]) if len(sys.argv) > 1 else 120, "out_dir": None},
    "eval": {"num_episodes": 8192}})
t0 = time.time()
rows = trainer.training(cfg)
dt = time.time() - t0
first = sum(r["avg_reward"] for r in rows[:5]) / 5
last = sum(r["avg_reward"] for r in rows[-5:]) / 5
steps = sum(r["steps"] for r in rows)
print(json.dumps({"first5_avg_reward": first, "last5_avg_reward": last, "batches": len(rows), "episode_steps": steps,
                  "seconds": dt, "episode_steps_per_s": steps / dt}))
