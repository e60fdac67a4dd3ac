#!/usr/bin/env python
"""Generates tests/golden/priority.npz from the LIVE reference (build container only):

    python tests/golden/gen_priority_golden.py

The reference's scripted baselines ``action_gen_1`` (up, right, down, left) and ``action_gen_2`` (up, right, left,
down) from tools/simple_action_gen.py:16-33 drive the unmodified ``Game2048Env`` (spawns replayed from the Philox
stream as in gen_golden.py).  The ACTIONS come from the reference's functions; the oracle in B2048_ACT_PRIORITY
mode is asserted to choose the same action and produce the same board / reward / flags at every step."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.ref_shim import REFERENCE_ROOT, ReplayRng, load_reference  # noqa: E402

ref = load_reference()
Game2048Env, Game2048EnvConfig = ref.env.Game2048Env, ref.env.Game2048EnvConfig
spec = importlib.util.spec_from_file_location("simple_action_gen", os.path.join(REFERENCE_ROOT, "tools", "simple_action_gen.py"))
sag = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sag)

KW = dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="sum", base_reward_scale=1.0, max_steps=None)
POLICIES = {"urdl": (sag.action_gen_1, (0, 1, 2, 3)), "urld": (sag.action_gen_2, (0, 1, 3, 2))}


def main(n_boards=12, seed=0x2048, gid0=77, T_cap=6000):
    out = {"seed": np.uint64(seed), "gid0": np.uint64(gid0)}
    for name, (fn, prio) in POLICIES.items():
        cfg = oracle.make_cfg(reward_mode="sum", obs_mode="none", max_steps=0, action_mode="priority", action_priority=prio)
        st, rlog = oracle.reset_many(n_boards, seed, gid0, 0, with_log=True)
        envs = []
        for i in range(n_boards):
            env = Game2048Env(Game2048EnvConfig(**KW))
            rr = ReplayRng()
            env.game._set_seed = lambda seed=None: None
            env.game._rng = rr
            rr.push(rlog[i, 0], rlog[i, 1]); rr.push(rlog[i, 2], rlog[i, 3])
            obs, _ = env.reset(seed=1)
            assert oracle.pack_board(env.game.board) == int(st["board"][i])
            envs.append([env, rr, obs])
        out[f"{name}/board0"] = st["board"].copy(); out[f"{name}/flags0"] = st["flags"].copy()
        alive = np.ones(n_boards, bool)
        rec = dict(board=[], action=[], reward=[], flags=[], score=[], alive=[])
        t = 0
        while alive.any() and t < T_cap:
            t += 1
            o = oracle.step_many(st, cfg, seed, gid0, t, with_log=True)
            rb = np.zeros(n_boards, np.uint64); ra = np.zeros(n_boards, np.uint8); rw = np.zeros(n_boards, np.float64)
            rf = np.zeros(n_boards, np.uint8); rs = np.zeros(n_boards, np.uint32)
            for i in range(n_boards):
                if not alive[i]:
                    continue
                env, rr, obs = envs[i]
                a = fn(obs, obs["action_mask"])                       # the REFERENCE picks the action
                assert a == int(o["action"][i]), (name, i, t, a, int(o["action"][i]))
                k, four = o["spawn_log"][i, 0], o["spawn_log"][i, 1]
                if k >= 0:
                    rr.push(k, four)
                obs, rew, term, trunc, info = env.step(a)
                envs[i][2] = obs
                fl = sum(int(v) << q for q, v in enumerate(obs["action_mask"]))
                fl |= (oracle.F_CHANGED if k >= 0 else 0) | (oracle.F_DONE if term else 0) | (oracle.F_TRUNC if trunc else 0)
                rb[i] = oracle.pack_board(env.game.board); ra[i] = a; rw[i] = rew; rf[i] = fl; rs[i] = env.game.score
                assert rb[i] == st["board"][i] and rew == o["reward64"][i] and fl == o["flags"][i] and rs[i] == st["score"][i]
                if term or trunc:
                    alive[i] = False
            for key, val in (("board", rb), ("action", ra), ("reward", rw), ("flags", rf), ("score", rs), ("alive", alive.copy())):
                rec[key].append(val)
        assert not alive.any(), "raise T_cap"
        for key in rec:
            out[f"{name}/{key}"] = np.stack(rec[key])
        print(f"priority[{name}]: T={t} final scores {np.stack(rec['score']).max(0).tolist()}")
    np.savez_compressed(os.path.join(HERE, "priority.npz"), **out)


if __name__ == "__main__":
    main()
