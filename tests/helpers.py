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
        deps = [src] + [os.path.join(hdr, f) for f in ("b2048_device.cuh", "b2048_step.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src])
        _hc = C.CDLL(so)
    return _hc


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
