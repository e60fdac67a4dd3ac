"""The Python port used as the CPU baseline must follow the reference too: replay the golden episodes
(spawns taken from the oracle's Philox log) through oracle/pyport.py and compare with the reference outputs."""
import os

import numpy as np
import pytest

import oracle
from oracle.pyport import PyEnv
from oracle.ref_shim import ReplayRng
from helpers import ENV_CONFIGS, GOLDEN, full_env_kwargs


@pytest.mark.parametrize("name", ["runner_default", "shaped_raw", "mask_off"])
def test_pyport_replays_golden(name):
    g = np.load(os.path.join(GOLDEN, "episodes.npz"))
    seed, gid0 = int(g["seed"]), int(g["gid0"])
    board = g[f"{name}/board"]
    T, n = board.shape
    n = min(n, 6)
    kw = full_env_kwargs(name); kw.pop("size")
    mask_on = kw["use_action_mask"]
    okw = dict(kw)
    if okw["max_steps"] is None:
        okw["max_steps"] = 0
    cfg = oracle.make_cfg(action_mode="random_legal" if mask_on else "random_any", **okw)
    st, rlog = oracle.reset_many(n, seed, gid0, 0, with_log=True)
    envs = []
    for i in range(n):
        e = PyEnv(**kw)
        rr = ReplayRng()
        rr.push(rlog[i, 0], rlog[i, 1]); rr.push(rlog[i, 2], rlog[i, 3])
        e.t = 0; e.max_tile_seen = 4
        e.game.board[:] = 0; e.game.score = 0; e.game.rng = rr
        e.game.spawn(); e.game.spawn()
        assert oracle.pack_board(e.game.board) == int(g[f"{name}/board0"][i])
        envs.append((e, rr))
    alive = np.ones(n, bool)
    for t in range(1, T + 1):
        o = oracle.step_many(st, cfg, seed, gid0, t, with_log=True)
        for i in range(n):
            if not alive[i]:
                continue
            e, rr = envs[i]
            if o["spawn_log"][i, 0] >= 0:
                rr.push(o["spawn_log"][i, 0], o["spawn_log"][i, 1])
            obs, r, done, trunc = e.step(int(g[f"{name}/action"][t - 1][i]))
            assert oracle.pack_board(e.game.board) == int(board[t - 1][i])
            assert r == g[f"{name}/reward"][t - 1][i]
            fl = int(g[f"{name}/flags"][t - 1][i])
            assert done == bool(fl & 0x20) and trunc == bool(fl & 0x40)
            if mask_on:
                assert sum(int(v) << q for q, v in enumerate(obs["action_mask"])) == (fl & 0xF)
            if t <= g[f"{name}/obs"].shape[0]:
                ob = obs["board"] if mask_on else obs
                assert (np.asarray(ob, np.float32).reshape(-1) == g[f"{name}/obs"][t - 1][i]).all()
        alive = g[f"{name}/alive"][t - 1][:n]


def test_policy_rollout_port_runs():
    """The CPU rollout baseline (run_episode restated per environment) steps and resets."""
    from oracle import pyport
    n, dt = pyport.time_policy_rollout_steps(0.3, seed=3)
    assert n >= 32 and dt > 0
