#!/usr/bin/env python
"""Generates the committed golden fixtures from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py

The reference is imported unmodified (oracle/ref_shim.py registers a minimal
``gymnasium`` stand-in).  Spawns are *replayed* into the reference: the CPU
oracle derives (cell, value) choices from the Philox stream defined in
include/b2048.h and a ReplayRng feeds exactly those into
``Game2048._spawn`` (game2048.py:108-118).  Every fixture stores the inputs and
the REFERENCE's outputs; at generation time the oracle is asserted equal.

Fixtures written next to this file:
  row_lut.npz     all 65,536 rows through Game2048._row_move_left
  moves.npz       random boards x 4 actions through _move / get_action_mask / _is_done
  episodes.npz    replayed episodes through Game2048Env.step for several configs
  autoreset.npz   replayed stepping with reset-on-done
  seeded.json     seeded known answers of the reference
  mlp.npz         forward_logits / logits_to_probs / compute_returns / _compute_advantages
  update_*.npz    ReinforceAgent.update_batch before/after (REINFORCE sgd, actor-critic adam)
"""
from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.ref_shim import ReplayRng, load_reference  # noqa: E402

ref = load_reference()
Game2048 = ref.game2048.Game2048
Game2048Env = ref.env.Game2048Env
Game2048EnvConfig = ref.env.Game2048EnvConfig

ENV_CONFIGS = {
    # runner.py:116-131 defaults
    "runner_default": dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5,
                           bonus_mode="off", max_steps=1024),
    # dataclass defaults env.py:19-40
    "dataclass_default": dict(),
    # every shaped term switched on, raw bonus, short horizon (truncation)
    "shaped_raw": dict(obs_mode="raw", reward_mode="sum", base_reward_scale=0.25, empty_tile_reward=0.05,
                       merge_reward=0.3, bonus_mode="raw", bonus_scale=0.125, step_reward=-0.01,
                       endgame_penalty=-7.5, max_steps=60),
    # runner.py:10-63 docstring example flavour: onehot, log2 bonus, no step limit
    "onehot_log2bonus": dict(obs_mode="onehot", reward_mode="log2", base_reward_scale=1.0, empty_tile_reward=0.05,
                             bonus_mode="log2", bonus_scale=2.0, max_steps=None),
    # mask off: illegal moves allowed, invalid_action_penalty path (env.py:206-207)
    "mask_off": dict(obs_mode="log2", obs_log2_scale=1.0, reward_mode="sum", use_action_mask=False,
                     invalid_action_penalty=-2.5, step_reward=0.125, max_steps=200),
}


def to_oracle_cfg(kw, action_mode="buffer", auto_reset=False, obs=True):
    c = Game2048EnvConfig(**kw)
    return oracle.make_cfg(reward_mode=c.reward_mode, bonus_mode=c.bonus_mode, obs_mode=c.obs_mode if obs else "none",
                           use_action_mask=c.use_action_mask, max_steps=c.max_steps, action_mode=action_mode,
                           auto_reset=auto_reset, base_reward_scale=c.base_reward_scale,
                           empty_tile_reward=c.empty_tile_reward, merge_reward=c.merge_reward,
                           bonus_scale=c.bonus_scale, step_reward=c.step_reward, endgame_penalty=c.endgame_penalty,
                           invalid_action_penalty=c.invalid_action_penalty, obs_log2_scale=c.obs_log2_scale)


def make_replay_env(kw):
    env = Game2048Env(Game2048EnvConfig(**kw))
    rr = ReplayRng()
    env.game._set_seed = lambda seed=None: None     # keep the replay rng across reset (instance-level patch)
    env.game._rng = rr
    return env, rr


def flat_obs(obs):
    b = obs["board"] if isinstance(obs, dict) else obs
    return np.asarray(b, np.float32).reshape(-1)


# --------------------------------------------------------------------------- rows

def gen_row_lut():
    g = Game2048()
    out = np.zeros(65536, np.uint16)
    mrg = np.zeros(65536, np.uint8)
    sc = np.zeros(65536, np.int32)
    for r in range(65536):
        e = [(r >> (4 * c)) & 0xF for c in range(4)]
        row = np.array([0 if x == 0 else 1 << x for x in e], dtype=np.int64)
        g._new_merged = []
        nr = g._row_move_left(row)
        o = 0
        for c, v in enumerate(nr):
            ex = 0 if v == 0 else int(v).bit_length() - 1
            o |= min(ex, 15) << (4 * c)            # 65536 saturates to nibble 15 (overflow is flagged separately)
        mb = 0
        for k, v in enumerate(g._new_merged):
            ex = int(v).bit_length() - 1
            mb |= (1 if ex == 16 else ex) << (4 * k)
        out[r], mrg[r], sc[r] = o, mb, sum(g._new_merged)
    o2, m2, s2 = oracle.row_lut()
    assert (o2 == out).all() and (m2 == mrg).all() and (s2 == sc).all()
    np.savez_compressed(os.path.join(HERE, "row_lut.npz"), left=out, merge=mrg, score=sc)
    print("row_lut: changed", int((out != np.arange(65536)).sum()), "score sum", int(sc.sum()))


# --------------------------------------------------------------------------- moves

def random_boards(rng, n):
    """Packed boards with a spread of fill levels / tile heights (no 15s so the reference stays in domain;
    a second batch contains 15s but never two adjacent equal 15s in a line)."""
    boards = np.zeros(n, np.uint64)
    for i in range(n):
        kind = i % 4
        if kind == 0:
            e = rng.integers(0, 4, 16)
        elif kind == 1:
            e = rng.integers(0, 12, 16) * (rng.random(16) < 0.7)
        elif kind == 2:
            e = rng.integers(1, 6, 16)              # full boards, many merges / dead ends
        else:
            e = rng.integers(0, 15, 16) * (rng.random(16) < rng.random())
        b = 0
        for k in range(16):
            b |= int(e[k]) << (4 * k)
        boards[i] = b
    return boards


def gen_moves(n=3000, seed=20481):
    rng = np.random.default_rng(seed)
    boards = random_boards(rng, n)
    # explicit edge cases
    edge = [0x0, 0x1, 0x1111111111111111, 0x2121121221211212, 0x1234123412341234, 0xEEEE000000000000,
            0x0000000000001111, 0x1000100010001000, 0xFEDCBA9876543210, 0xE0E0E0E0D0D0D0D0]
    boards[: len(edge)] = np.array(edge, dtype=np.uint64)
    res = np.zeros((n, 4), np.uint64)
    msum = np.zeros((n, 4), np.int32)
    nmerge = np.zeros((n, 4), np.int8)
    mmax = np.zeros((n, 4), np.int32)
    changed = np.zeros((n, 4), np.uint8)
    mask = np.zeros(n, np.uint8)
    done = np.zeros(n, np.uint8)
    g = Game2048()
    for i in range(n):
        tiles = oracle.unpack_board(boards[i])
        g.board = tiles.copy()
        m = g.get_action_mask()
        mask[i] = sum(v << a for a, v in enumerate(m))
        done[i] = g._is_done()
        for a in range(4):
            g.board = tiles.copy()
            g._new_merged = []
            ch = g._move(a)
            res[i, a] = oracle.pack_board(g.board)
            msum[i, a] = sum(g._new_merged)
            nmerge[i, a] = len(g._new_merged)
            mmax[i, a] = max(g._new_merged, default=0)
            changed[i, a] = ch
    for a in range(4):
        o, ms, mi, fl = oracle.move_many(boards, np.full(n, a, np.uint8))
        assert (o == res[:, a]).all() and (ms == msum[:, a]).all()
        assert (((fl & oracle.F_CHANGED) != 0) == (changed[:, a] != 0)).all()
        assert not (fl & oracle.F_OVERFLOW).any()
    om, od = oracle.mask_done(boards)
    assert (om == mask).all() and (od == done).all()
    np.savez_compressed(os.path.join(HERE, "moves.npz"), boards=boards, result=res, merge_sum=msum, n_merge=nmerge,
                        merge_max=mmax, changed=changed, mask=mask, done=done)
    print("moves:", n, "boards; legal-any", int((mask != 0).sum()), "done", int(done.sum()))


# --------------------------------------------------------------------------- episodes

def gen_episodes(n_boards=24, seed=0xB200, gid0=1000):
    out = {}
    for name, kw in ENV_CONFIGS.items():
        mask_on = Game2048EnvConfig(**kw).use_action_mask
        cfg = to_oracle_cfg(kw, action_mode="random_legal" if mask_on else "random_any")
        st, rlog = oracle.reset_many(n_boards, seed, gid0, 0, with_log=True)
        envs = []
        for i in range(n_boards):
            env, rr = make_replay_env(kw)
            rr.push(rlog[i, 0], rlog[i, 1])
            rr.push(rlog[i, 2], rlog[i, 3])
            obs, info = env.reset(seed=123)
            assert oracle.pack_board(env.game.board) == int(st["board"][i]), name
            envs.append((env, rr, obs))
        alive = np.ones(n_boards, bool)
        T_cap = 700
        rec = dict(board=[], action=[], reward=[], flags=[], score=[], max_tile=[], step=[], obs=[], alive=[])
        board0 = st["board"].copy()
        flags0 = st["flags"].copy()
        obs_w = 272 if kw.get("obs_mode") == "onehot" else 16
        t = 0
        while alive.any() and t < T_cap:
            t += 1
            o = oracle.step_many(st, cfg, seed, gid0, t, want_obs=True, with_log=True)
            r_board = np.zeros(n_boards, np.uint64)
            r_reward = np.zeros(n_boards, np.float64)
            r_flags = np.zeros(n_boards, np.uint8)
            r_score = np.zeros(n_boards, np.uint32)
            r_maxt = np.zeros(n_boards, np.int32)
            r_step = np.zeros(n_boards, np.uint32)
            r_obs = np.zeros((n_boards, obs_w), np.float32)
            for i in range(n_boards):
                if not alive[i]:
                    continue
                env, rr, _ = envs[i]
                a = int(o["action"][i])
                k, four = o["spawn_log"][i, 0], o["spawn_log"][i, 1]
                if k >= 0:
                    rr.push(k, four)
                obs, rew, term, trunc, info = env.step(a)
                assert not rr.queue, (name, i, t)
                fl = 0
                if isinstance(obs, dict):
                    fl |= sum(int(v) << q for q, v in enumerate(obs["action_mask"]))
                else:
                    fl |= sum(int(v) << q for q, v in enumerate(env.game.get_action_mask()))
                # is_changed is not returned by env.step; a spawn happens iff the move changed the board
                # (game2048.py:56-57; a changed move always frees a cell), so k >= 0 <=> changed
                assert (k >= 0) or info["invalid_action"] or term
                fl |= (oracle.F_CHANGED if k >= 0 else 0) | (oracle.F_DONE if term else 0) | (oracle.F_TRUNC if trunc else 0)
                r_board[i] = oracle.pack_board(env.game.board)
                r_reward[i] = rew
                r_flags[i] = fl
                r_score[i] = env.game.score
                r_maxt[i] = env.max_tile_seen
                r_step[i] = env._step_count
                r_obs[i] = flat_obs(obs)
                # oracle == reference, bit for bit
                assert r_board[i] == st["board"][i], (name, i, t)
                assert rew == o["reward64"][i], (name, i, t, rew, o["reward64"][i])
                assert fl == o["flags"][i], (name, i, t, fl, o["flags"][i])
                assert r_score[i] == st["score"][i] and r_step[i] == st["step"][i]
                assert r_maxt[i] == (1 << int(st["max_exp"][i]))
                assert (r_obs[i] == o["obs"][i]).all()
                assert info["merged"] is not None and sum(info["merged"]) == o["merge_sum"][i]
                if term or trunc:
                    alive[i] = False
            rec["board"].append(r_board); rec["action"].append(o["action"].copy()); rec["reward"].append(r_reward)
            rec["flags"].append(r_flags); rec["score"].append(r_score); rec["max_tile"].append(r_maxt)
            rec["step"].append(r_step); rec["obs"].append(r_obs if t <= 12 else None)
            rec["alive"].append(alive.copy())
        T = t
        out[f"{name}/board0"] = board0
        out[f"{name}/flags0"] = flags0
        for k in ("board", "action", "reward", "flags", "score", "max_tile", "step", "alive"):
            out[f"{name}/{k}"] = np.stack(rec[k])
        out[f"{name}/obs"] = np.stack([x for x in rec["obs"] if x is not None])
        lens = np.stack(rec["alive"]).sum(0) + 1
        print(f"episodes[{name}]: T={T} mean_len={lens.mean():.1f} max_score={int(np.stack(rec['score']).max())}")
    out["seed"] = np.uint64(seed)
    out["gid0"] = np.uint64(gid0)
    np.savez_compressed(os.path.join(HERE, "episodes.npz"), **out)


def gen_autoreset(n_boards=16, T=400, seed=77, gid0=5):
    kw = dict(ENV_CONFIGS["runner_default"])
    kw["max_steps"] = 90
    cfg = to_oracle_cfg(kw, action_mode="random_legal", auto_reset=True)
    st, rlog = oracle.reset_many(n_boards, seed, gid0, 0, with_log=True)
    envs = []
    for i in range(n_boards):
        env, rr = make_replay_env(kw)
        rr.push(rlog[i, 0], rlog[i, 1]); rr.push(rlog[i, 2], rlog[i, 3])
        env.reset(seed=1)
        envs.append((env, rr))
    boards, rewards, flags, actions = [], [], [], []
    n_resets = 0
    for t in range(1, T + 1):
        o = oracle.step_many(st, cfg, seed, gid0, t, with_log=True)
        rb = np.zeros(n_boards, np.uint64); rw = np.zeros(n_boards, np.float64); rf = np.zeros(n_boards, np.uint8)
        for i in range(n_boards):
            env, rr = envs[i]
            lg = o["spawn_log"][i]
            if lg[0] >= 0:
                rr.push(lg[0], lg[1])
            obs, rew, term, trunc, info = env.step(int(o["action"][i]))
            fl = (oracle.F_CHANGED if lg[0] >= 0 else 0) | (oracle.F_DONE if term else 0) | (oracle.F_TRUNC if trunc else 0)
            if term or trunc:
                rr.push(lg[2], lg[3]); rr.push(lg[4], lg[5])
                obs, info = env.reset(seed=1)
                n_resets += 1
                assert env.game.score == st["score"][i] == 0 and env._step_count == st["step"][i] == 0
            fl |= sum(int(v) << q for q, v in enumerate(obs["action_mask"]))
            rb[i] = oracle.pack_board(env.game.board); rw[i] = rew; rf[i] = fl
            assert rb[i] == st["board"][i] and rew == o["reward64"][i] and fl == o["flags"][i], (t, i)
        boards.append(rb); rewards.append(rw); flags.append(rf); actions.append(o["action"].copy())
    np.savez_compressed(os.path.join(HERE, "autoreset.npz"), seed=np.uint64(seed), gid0=np.uint64(gid0),
                        max_steps=np.int32(90), board=np.stack(boards), reward=np.stack(rewards),
                        flags=np.stack(flags), action=np.stack(actions))
    print("autoreset: resets", n_resets)


# --------------------------------------------------------------------------- seeded facts

def gen_seeded():
    facts = {}
    g = Game2048()
    facts["reset_seed"] = {str(s): g.reset(seed=s) for s in (0, 1, 2)}
    g.reset(seed=1)
    steps = []
    for a in (0, 1, 2, 3):
        ch, state, merged, done = g.step(a)
        steps.append(dict(action=a, changed=bool(ch), state=state, merged=merged, done=bool(done)))
    facts["reset1_steps"] = steps
    facts["reset1_score"] = int(g.score)
    runner = ref.runner
    it3, it7 = runner.make_fixed_seed_iter(3), runner.make_fixed_seed_iter(7)
    facts["seed_iter_3"] = [next(it3) for _ in range(3)]
    facts["seed_iter_7"] = [next(it7) for _ in range(3)]
    # BASELINE.json config 1: runner defaults, seeds 3 / 7, first 8 training episodes
    env = Game2048Env(Game2048EnvConfig(**runner.DEFAULT_ENV_KWARGS))
    agent = ref.agent.ReinforceAgent(env, ref.MLP.MLPConfig(**runner.DEFAULT_MLP_KWARGS),
                                     ref.agent.ReinforceAgentConfig(**runner.DEFAULT_AGENT_KWARGS))
    it_e, it_p = runner.make_fixed_seed_iter(3), runner.make_fixed_seed_iter(7)
    eps = []
    for _ in range(8):
        tr = agent.run_episode(next(it_e), next(it_p))
        eps.append([len(tr["actions"]), float(tr["total_reward"]), int(tr["max_tile"])])
    facts["runner_default_first8"] = eps
    facts["actor_shapes"] = [list(W.shape) for W in agent.params["W"]]
    with open(os.path.join(HERE, "seeded.json"), "w") as f:
        json.dump(facts, f, indent=1)
    print("seeded:", eps)


# --------------------------------------------------------------------------- MLP / learner

def traj_to_arrays(traj, kw):
    """Packed boards / masks / actions / rewards of one reference trajectory."""
    T = len(traj["actions"])
    boards = np.zeros(T, np.uint64)
    masks = np.zeros(T, np.uint8)
    for t, o in enumerate(traj["obs"]):
        b = np.asarray(o["board"])
        if kw.get("obs_mode") == "onehot":
            e = b.reshape(16, 17).argmax(1)
        elif kw.get("obs_mode") == "log2":
            e = np.rint(b.reshape(16) / np.float32(kw.get("obs_log2_scale", 1.0))).astype(int)
        else:
            e = np.array([0 if v == 0 else int(v).bit_length() - 1 for v in b.reshape(16).astype(np.int64)])
        boards[t] = sum(int(x) << (4 * k) for k, x in enumerate(e))
        masks[t] = sum(int(v) << q for q, v in enumerate(o["action_mask"]))
    return boards, masks, np.array(traj["actions"], np.uint8), np.array(traj["rewards"], np.float64)


def gen_mlp():
    runner = ref.runner
    out = {}
    for tag, env_kw, mlp_kw in (
        ("default", dict(runner.DEFAULT_ENV_KWARGS), dict(runner.DEFAULT_MLP_KWARGS)),
        ("onehot", dict(obs_mode="onehot", reward_mode="log2", max_steps=None),
         dict(hidden_sizes=[256, 128, 64], activation="ReLU", init_distribution="HeNormal")),
        ("sigmoid", dict(obs_mode="log2", obs_log2_scale=1.0), dict(hidden_sizes=[48], activation="Sigmoid",
                                                                     init_distribution="XavierNormal")),
    ):
        env = Game2048Env(Game2048EnvConfig(**env_kw))
        agent = ref.agent.ReinforceAgent(env, ref.MLP.MLPConfig(**mlp_kw), ref.agent.ReinforceAgentConfig(model_seed=0))
        tr = agent.run_episode(11, 13)
        boards, masks, actions, rewards = traj_to_arrays(tr, env_kw)
        X = np.stack([flat_obs(o) for o in tr["obs"]])
        M = np.stack([o["action_mask"] for o in tr["obs"]])
        logits, acts, pre = ref.MLP.forward_logits(agent.params, X, mlp_kw["activation"])
        probs = ref.MLP.logits_to_probs(logits, M)
        # single-sample path must agree with the batched one the kernels are compared to
        l1, _, _ = ref.MLP.forward_logits(agent.params, X[0], mlp_kw["activation"])
        assert np.allclose(l1, logits[0], rtol=1e-5, atol=1e-6)
        out[f"{tag}/boards"] = boards
        out[f"{tag}/masks"] = masks
        out[f"{tag}/logits"] = logits.astype(np.float32)
        out[f"{tag}/probs"] = probs.astype(np.float32)
        greedy = np.array([int(np.argmax(p * m)) for p, m in zip(probs, M)], np.uint8)
        out[f"{tag}/greedy"] = greedy
        for i, (W, b) in enumerate(zip(agent.params["W"], agent.params["b"])):
            out[f"{tag}/W{i}"] = W.astype(np.float32)
            out[f"{tag}/b{i}"] = b.astype(np.float32)
        out[f"{tag}/n_layers"] = np.int32(len(agent.params["W"]))
        out[f"{tag}/obs_scale"] = np.float32(env_kw.get("obs_log2_scale", 1.0))
        print(f"mlp[{tag}]: T={len(boards)} in={X.shape[1]}")
    # compute_returns / _compute_advantages
    env = Game2048Env(Game2048EnvConfig(**runner.DEFAULT_ENV_KWARGS))
    rng = np.random.default_rng(5)
    lens = [1, 2, 7, 33, 150, 64]
    rew = [list((rng.integers(0, 9, L) * 0.5 - (rng.random(L) < 0.1) * 3.0).astype(float)) for L in lens]
    for gamma in (0.99, 1.0, 0.5):
        for mode in ("off", "each", "batch", "batch_norm"):
            ag = ref.agent.ReinforceAgent(env, ref.MLP.MLPConfig(hidden_sizes=[8]),
                                          ref.agent.ReinforceAgentConfig(gamma=gamma, baseline_mode=mode))
            rets = [ag.compute_returns(r) for r in rew]
            w = np.array([1.0, 0.5, 2.0, 1.0, 0.25, 1.25], np.float32)
            adv = ag._compute_advantages(rets, w)
            out[f"ret/{gamma}/{mode}/returns"] = np.concatenate(rets)
            out[f"ret/{gamma}/{mode}/adv"] = np.concatenate(adv).astype(np.float32)
    out["ret/lens"] = np.array(lens, np.int32)
    out["ret/rewards"] = np.concatenate([np.array(r) for r in rew])
    out["ret/weights"] = np.array([1.0, 0.5, 2.0, 1.0, 0.25, 1.25], np.float32)
    np.savez_compressed(os.path.join(HERE, "mlp.npz"), **out)


def gen_update(tag, env_kw, mlp_kw, agent_kw, n_eps=6, n_updates=2):
    env = Game2048Env(Game2048EnvConfig(**env_kw))
    agent = ref.agent.ReinforceAgent(env, ref.MLP.MLPConfig(**mlp_kw), ref.agent.ReinforceAgentConfig(**agent_kw))
    out = {}

    def dump_params(prefix, params):
        for i, (W, b) in enumerate(zip(params["W"], params["b"])):
            out[f"{prefix}/W{i}"] = np.array(W, np.float32)
            out[f"{prefix}/b{i}"] = np.array(b, np.float32)

    dump_params("actor0", agent.params)
    if agent.critic_params is not None:
        dump_params("critic0", agent.critic_params)
    it_e, it_p = ref.runner.make_fixed_seed_iter(3), ref.runner.make_fixed_seed_iter(7)
    norms, advs = [], []
    orig_clip = agent.clip_grads_global_norm
    orig_adv = agent._compute_advantages

    def clip_spy(gW, gb):
        v = orig_clip(gW, gb)
        norms.append(float(v))
        return v

    def adv_spy(returns_list, w):
        a = orig_adv(returns_list, w)
        advs.append(np.concatenate(a).astype(np.float32))
        return a

    agent.clip_grads_global_norm = clip_spy
    agent._compute_advantages = adv_spy
    for u in range(n_updates):
        trajs = [agent.run_episode(next(it_e), next(it_p)) for _ in range(n_eps)]
        lens = np.array([len(t["actions"]) for t in trajs], np.int32)
        arrs = [traj_to_arrays(t, env_kw) for t in trajs]
        out[f"u{u}/lens"] = lens
        out[f"u{u}/boards"] = np.concatenate([a[0] for a in arrs])
        out[f"u{u}/masks"] = np.concatenate([a[1] for a in arrs])
        out[f"u{u}/actions"] = np.concatenate([a[2] for a in arrs])
        out[f"u{u}/rewards"] = np.concatenate([a[3] for a in arrs])
        out[f"u{u}/total_reward"] = np.array([t["total_reward"] for t in trajs], np.float64)
        n0 = len(norms)
        agent.update_batch(trajs)
        out[f"u{u}/grad_norms"] = np.array(norms[n0:], np.float64)     # [actor] or [actor, critic]
        out[f"u{u}/adv"] = advs[-1]
        dump_params(f"u{u}/actor", agent.params)
        if agent.critic_params is not None:
            dump_params(f"u{u}/critic", agent.critic_params)
    out["n_layers"] = np.int32(len(agent.params["W"]))
    out["n_updates"] = np.int32(n_updates)
    out["meta"] = np.array(json.dumps(dict(env=env_kw, mlp=mlp_kw, agent=agent_kw)))
    np.savez_compressed(os.path.join(HERE, f"update_{tag}.npz"), **out)
    print(f"update[{tag}]: lens {out['u0/lens'].tolist()} norms {out['u0/grad_norms'].tolist()}")


def main():
    gen_row_lut()
    gen_moves()
    gen_episodes()
    gen_autoreset()
    gen_seeded()
    gen_mlp()
    runner = ref.runner
    env_d = dict(runner.DEFAULT_ENV_KWARGS)
    gen_update("reinforce_sgd", env_d, dict(hidden_sizes=[64, 48], activation="ReLU", init_distribution="HeNormal"),
               dict(gamma=0.99, learning_rate=1e-2, baseline_mode="batch", optimizer="sgd", max_grad_norm=1.0))
    gen_update("reinforce_norm_sigmoid", dict(env_d, obs_log2_scale=0.25),
               dict(hidden_sizes=[40], activation="Sigmoid", init_distribution="XavierUniform"),
               dict(gamma=0.9, learning_rate=5e-2, baseline_mode="batch_norm", optimizer="adam", max_grad_norm=0.5,
                    reward_rank_weights=[2.0, 1.0]))
    gen_update("actor_critic_adam", dict(obs_mode="onehot", reward_mode="log2", max_steps=None, empty_tile_reward=0.05),
               dict(hidden_sizes=[48, 32], activation="ReLU", init_distribution="HeNormal"),
               dict(gamma=0.99, learning_rate=0.01, baseline_mode="batch", optimizer="adam", use_critic=True,
                    critic_learning_rate=5e-4, max_grad_norm=1.0))
    gen_update("actor_critic_huber_sgd", env_d,
               dict(hidden_sizes=[32], activation="Sigmoid", init_distribution="XavierNormal"),
               dict(gamma=0.95, learning_rate=0.02, baseline_mode="each", optimizer="sgd", use_critic=True,
                    critic_learning_rate=1e-2, critic_loss_type="huber", huber_delta=0.5, max_grad_norm=10.0))


if __name__ == "__main__":
    main()
