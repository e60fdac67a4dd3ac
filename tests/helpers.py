"""Shared helpers for the parity tests (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

ENV_CONFIGS = {
    "runner_default": dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5,
                           bonus_mode="off", max_steps=1024),
    "dataclass_default": dict(),
    "shaped_raw": dict(obs_mode="raw", reward_mode="sum", base_reward_scale=0.25, empty_tile_reward=0.05,
                       merge_reward=0.3, bonus_mode="raw", bonus_scale=0.125, step_reward=-0.01,
                       endgame_penalty=-7.5, max_steps=60),
    "onehot_log2bonus": dict(obs_mode="onehot", reward_mode="log2", base_reward_scale=1.0, empty_tile_reward=0.05,
                             bonus_mode="log2", bonus_scale=2.0, max_steps=None),
    "mask_off": dict(obs_mode="log2", obs_log2_scale=1.0, reward_mode="sum", use_action_mask=False,
                     invalid_action_penalty=-2.5, step_reward=0.125, max_steps=200),
}
# Game2048EnvConfig defaults (reference src/env.py:19-40)
ENV_DEFAULTS = dict(size=4, obs_mode="raw", obs_log2_scale=1.0, reward_mode="sum", base_reward_scale=1.0,
                    empty_tile_reward=0.0, merge_reward=0.0, bonus_mode="off", bonus_scale=1.0, step_reward=0.0,
                    endgame_penalty=0.0, use_action_mask=True, invalid_action_penalty=-1.0, max_steps=1024)


def full_env_kwargs(name):
    kw = dict(ENV_DEFAULTS)
    kw.update(ENV_CONFIGS[name])
    return kw


def random_boards(rng, n):
    e = rng.integers(0, 16, (n, 16)) * (rng.random((n, 16)) < rng.random((n, 1)))
    q = n // 4
    e[:q] = rng.integers(0, 4, (q, 16))
    e[q:2 * q] = rng.integers(1, 5, (q, 16))
    boards = np.zeros(n, np.uint64)
    for k in range(16):
        boards |= e[:, k].astype(np.uint64) << np.uint64(4 * k)
    return boards


_hc = None


def host_check_lib():
    """g++ build of the product's host+device headers (bit tricks + step body) for CPU-side checking."""
    global _hc
    if _hc is None:
        d = os.path.join(ROOT, "tests", "host_check")
        so = os.path.join(d, "libhostcheck.so")
        src = os.path.join(d, "host_check.cpp")
        hdr = os.path.join(ROOT, "rl-2048-with-reinforce-and-actor-critic_b200", "csrc")
        deps = [src] + [os.path.join(hdr, f) for f in ("b2048_device.cuh", "b2048_step.cuh", "b2048_step_fast.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src])
        _hc = C.CDLL(so)
    return _hc


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def load_update_fixture(tag):
    """update_*.npz -> (meta, actor0, critic0, list of updates) ; each update has per-episode lists."""
    import json
    g = np.load(os.path.join(GOLDEN, f"update_{tag}.npz"))
    meta = json.loads(str(g["meta"]))
    L = int(g["n_layers"])

    def params(prefix):
        if f"{prefix}/W0" not in g:
            return None
        return {"W": [g[f"{prefix}/W{i}"] for i in range(L)], "b": [g[f"{prefix}/b{i}"] for i in range(L)]}

    updates = []
    for u in range(int(g["n_updates"])):
        lens = g[f"u{u}/lens"]
        offs = np.concatenate([[0], np.cumsum(lens)])
        eps = []
        for i in range(len(lens)):
            s = slice(offs[i], offs[i + 1])
            eps.append(dict(boards=g[f"u{u}/boards"][s], masks=g[f"u{u}/masks"][s], actions=g[f"u{u}/actions"][s],
                            rewards=g[f"u{u}/rewards"][s]))
        updates.append(dict(episodes=eps, lens=lens, total_reward=g[f"u{u}/total_reward"], grad_norms=g[f"u{u}/grad_norms"],
                            adv=g[f"u{u}/adv"], actor=params(f"u{u}/actor"), critic=params(f"u{u}/critic")))
    return meta, params("actor0"), params("critic0"), updates


def rank_weights(total_rewards, conf):
    """reference _compute_episode_rank_weights (src/reinforce_agent.py:681-716), restated for the tests."""
    n = len(total_rewards)
    if not conf:
        return np.ones(n, np.float32)
    conf = np.asarray(conf, np.float32)
    order = np.argsort(total_rewards)
    w = np.zeros(n, np.float32)
    for rank, idx in enumerate(order):
        b = min(int((rank + 0.5) / n * len(conf)), len(conf) - 1)
        w[idx] = conf[b]
    m = w.mean()
    return w / m if m > 1e-8 else w


UPDATE_TAGS = ["reinforce_sgd", "reinforce_norm_sigmoid", "actor_critic_adam", "actor_critic_huber_sgd"]


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))
