"""TEST / BASELINE INFRASTRUCTURE ONLY — per-environment Python + NumPy port of the reference's
env path, used solely to time "what the reference does on a CPU" on the GPU box (where
/root/reference does not exist).  bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm run it.

It restates the reference's algorithm at the reference's own granularity — one environment, one
step at a time, int64 4x4 ndarray of raw tile values, rotate -> slide rows left -> rotate back,
``np.argwhere`` spawn with a NumPy ``Generator`` (PCG64), legal mask by previewing all four moves,
float64 shaped reward — so its cost profile is the reference's (minus the reference's eager debug
string formatting, game2048.py:69 / reinforce_agent.py:147-149, which is logging, not algorithm; the
port is therefore a slightly *faster* baseline than the reference itself).

Reference lines followed: src/game2048.py:26-34 (reset), :40-70 (step), :95-99/:189-237 (mask),
:108-118 (spawn), :120-165 (move), :172-187 (done); src/env.py:131-150 (obs), :197-261 (reward),
:264-302 (step).  Checked against the golden episodes in tests/test_pyport.py.
"""
from __future__ import annotations

import time

import numpy as np


class PyGame:
    def __init__(self):
        self.board = np.zeros((4, 4), dtype=np.int64)
        self.score = 0
        self.steps = 0
        self.rng = np.random.default_rng()
        self.merged: list[int] = []

    def reset(self, seed=None):
        self.rng = np.random.default_rng(seed)
        self.board = np.zeros((4, 4), dtype=np.int64)
        self.score = 0
        self.steps = 0
        self.spawn()
        self.spawn()

    def spawn(self):
        free = np.argwhere(self.board == 0)
        if free.size == 0:
            return
        r, c = free[self.rng.integers(len(free))]
        self.board[r, c] = 2 if self.rng.random() < 0.9 else 4

    @staticmethod
    def slide_row(row, merged):
        tiles = row[row != 0]
        out = np.zeros_like(row)
        i = w = 0
        n = len(tiles)
        while i < n:
            if i + 1 < n and tiles[i] == tiles[i + 1]:
                v = int(tiles[i]) * 2
                out[w] = v
                if merged is not None:
                    merged.append(v)
                i += 2
            else:
                out[w] = int(tiles[i])
                i += 1
            w += 1
        return out

    @staticmethod
    def slide_board(board, action, merged):
        k = 3 - action
        b = np.rot90(board, k=-k) if k % 4 else board
        nb = np.zeros_like(b)
        changed = False
        for r in range(4):
            new = PyGame.slide_row(b[r], merged)
            nb[r] = new
            if not np.array_equal(new, b[r]):
                changed = True
        back = (4 - k) % 4
        if back:
            nb = np.rot90(nb, k=-back)
        return nb, changed

    def is_done(self):
        b = self.board
        if (b == 0).any():
            return False
        for r in range(4):
            for c in range(4):
                v = b[r, c]
                if (r + 1 < 4 and b[r + 1, c] == v) or (c + 1 < 4 and b[r, c + 1] == v):
                    return False
        return True

    def mask(self):
        state = self.board.tolist()
        out = []
        for a in range(4):
            _, ch = PyGame.slide_board(np.array(state, dtype=int), a, None)
            out.append(1 if ch else 0)
        return out

    def step(self, action):
        self.steps += 1
        self.merged = []
        self.board, changed = PyGame.slide_board(self.board, action, self.merged)
        done = self.is_done()
        self.score += sum(self.merged)
        if changed:
            self.spawn()
            done = self.is_done()
        return changed, list(self.merged), done


class PyEnv:
    """Game2048Env restated (dict obs with mask, float64 reward)."""

    def __init__(self, **kw):
        d = dict(obs_mode="raw", obs_log2_scale=1.0, reward_mode="sum", base_reward_scale=1.0, empty_tile_reward=0.0,
                 merge_reward=0.0, bonus_mode="off", bonus_scale=1.0, step_reward=0.0, endgame_penalty=0.0,
                 use_action_mask=True, invalid_action_penalty=-1.0, max_steps=1024)
        d.update(kw)
        d.pop("size", None)
        self.c = d
        self.game = PyGame()
        self.t = 0
        self.max_tile_seen = 4

    def obs(self):
        c = self.c
        b = self.game.board.astype(np.float32)
        if c["obs_mode"] == "log2":
            nzm = b > 0
            b[nzm] = np.log2(b[nzm])
            b *= c["obs_log2_scale"]
        elif c["obs_mode"] == "onehot":
            e = np.zeros(b.shape, dtype=np.int32)
            nzm = b > 0
            e[nzm] = np.log2(b[nzm]).astype(np.int32)
            b = np.eye(17, dtype=np.float32)[e]
        if c["use_action_mask"]:
            return {"board": b, "action_mask": np.array(self.game.mask(), dtype=np.int8)}
        return b

    def reset(self, seed=None):
        self.t = 0
        self.max_tile_seen = 4
        self.game.reset(seed)
        return self.obs()

    def reward(self, merged, done, invalid):
        c = self.c
        if not c["use_action_mask"] and invalid:
            return c["invalid_action_penalty"]
        if c["reward_mode"] == "sum":
            r = float(sum(merged))
        else:
            r = 0.0
            for v in merged:
                r += float(np.log2(v))
        r *= c["base_reward_scale"]
        if c["empty_tile_reward"] != 0.0:
            r += c["empty_tile_reward"] * float(np.sum(self.game.board == 0))
        if c["merge_reward"] != 0.0:
            r += c["merge_reward"] * float(len(merged))
        mx = max(merged, default=0)
        if mx >= 8 and mx > self.max_tile_seen:
            bonus = 0.0
            if c["bonus_mode"] == "raw":
                bonus = float(mx)
            elif c["bonus_mode"] == "log2":
                bonus = float(np.log2(mx))
            self.max_tile_seen = mx
            r += bonus * c["bonus_scale"]
        r += c["step_reward"]
        if done and c["endgame_penalty"] != 0.0:
            r += c["endgame_penalty"]
        return r

    def step(self, action):
        self.t += 1
        changed, merged, done = self.game.step(action)
        invalid = (not changed) and (not done)
        r = self.reward(merged, done, invalid)
        trunc = self.c["max_steps"] is not None and self.t >= self.c["max_steps"] and not done
        return self.obs(), r, done, trunc


RUNNER_DEFAULT_ENV = dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5,
                          bonus_mode="off", max_steps=1024)


def time_random_legal_steps(seconds: float, seed: int = 0, env_kw: dict | None = None) -> tuple[int, float]:
    """Random-legal stepping with reset-on-done for ~`seconds`; returns (env steps, elapsed seconds)."""
    env = PyEnv(**(env_kw or RUNNER_DEFAULT_ENV))
    rng = np.random.default_rng(seed)
    obs = env.reset(seed)
    n = 0
    t0 = time.perf_counter()
    while True:
        legal = np.flatnonzero(obs["action_mask"])
        a = int(legal[rng.integers(len(legal))]) if len(legal) else 0
        obs, r, done, trunc = env.step(a)
        n += 1
        if done or trunc:
            obs = env.reset(seed + n)
        if (n & 63) == 0 and time.perf_counter() - t0 >= seconds:
            break
    return n, time.perf_counter() - t0


def _worker(args):
    seconds, seed = args
    return time_random_legal_steps(seconds, seed)


def time_multiprocess(seconds: float, procs: int) -> tuple[float, int]:
    """Sum of steps/s over `procs` independent worker processes (the reference itself is single-threaded)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(seconds, 1000 + i) for i in range(procs)])
    rate = sum(n / dt for n, dt in res)
    return rate, sum(n for n, _ in res)


def time_multiprocess_series(seconds: float, procs: int, count: int) -> list[tuple[float, int]]:
    """`count` consecutive samples of time_multiprocess on ONE worker pool (no per-sample fork cost)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    out = []
    with ctx.Pool(procs) as pool:
        for k in range(count):
            res = pool.map(_worker, [(seconds, 1000 + 131 * k + i) for i in range(procs)])
            out.append((sum(n / dt for n, dt in res), sum(n for n, _ in res)))
    return out


# ---------------------------------------------------------------------------------------------- policy rollout
def time_policy_rollout_steps(seconds: float, seed: int = 0) -> tuple[int, float]:
    """ReinforceAgent.run_episode restated at the reference's granularity (src/reinforce_agent.py:126-252,
    src/MLP.py:22-196): per step encode_observation -> three float32 `a @ W + b` products with ReLU -> masked
    max-subtracted softmax -> rng.choice(4, p=probs) -> env.step, trajectory lists appended, on one environment with the
    runner-default 16-256-256-4 network.  (Without the reference's eager debug strings and ANSI renders, which are
    logging: this port is therefore a slightly faster baseline than the reference itself.)"""
    rng = np.random.default_rng(seed)
    sizes = [16, 256, 256, 4]
    W = [(rng.normal(size=(i, o)) * np.sqrt(2.0 / i)).astype(np.float32) for i, o in zip(sizes[:-1], sizes[1:])]
    b = [np.zeros(o, np.float32) for o in sizes[1:]]
    env = PyEnv(**RUNNER_DEFAULT_ENV)
    obs = env.reset(seed)
    n = 0
    obs_list, act_list, rew_list = [], [], []
    t0 = time.perf_counter()
    while True:
        x = obs["board"].astype(np.float32).flatten()
        mask = obs["action_mask"]
        a = x
        for l in range(3):
            z = a @ W[l] + b[l]
            a = np.maximum(z, 0.0) if l < 2 else z
        logits = np.where(mask.astype(bool), a, -1e9)
        e = np.exp(logits - np.max(logits))
        probs = e / np.sum(e)
        action = int(rng.choice(4, p=probs))
        nobs, r, done, trunc = env.step(action)
        obs_list.append(obs); act_list.append(action); rew_list.append(float(r))
        obs = nobs
        n += 1
        if done or trunc:
            obs = env.reset(seed + n)
            obs_list, act_list, rew_list = [], [], []
        if (n & 31) == 0 and time.perf_counter() - t0 >= seconds:
            break
    return n, time.perf_counter() - t0


def _rollout_worker(args):
    seconds, seed = args
    return time_policy_rollout_steps(seconds, seed)


def time_rollout_multiprocess(seconds: float, procs: int) -> tuple[float, int]:
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        res = pool.map(_rollout_worker, [(seconds, 5000 + i) for i in range(procs)])
    return sum(n / dt for n, dt in res), sum(n for n, _ in res)
