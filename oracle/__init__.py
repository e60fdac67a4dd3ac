"""TEST INFRASTRUCTURE ONLY — CPU oracle for the b2048 hot path.

``oracle/b2048_oracle.c`` restates the reference's game rules / env step in
plain C (cell by cell, no LUT); ``oracle/learner.py`` restates the MLP /
REINFORCE / actor-critic arithmetic in NumPy; ``oracle/pyport.py`` is a
per-environment pure-Python port used only as the CPU baseline timing.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this package.  The product package never does.

Parity pin: the reference ships no golden vectors, so the oracle is pinned
against the live reference imported from /root/reference (build container
only) — see ``tests/golden/gen_golden.py`` and ``tests/test_oracle_vs_golden.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libb2048_oracle.so")


class EnvCfg(C.Structure):
    """ctypes mirror of b2048_env_cfg (include/b2048.h)."""

    _fields_ = [
        ("reward_mode", C.c_int32), ("bonus_mode", C.c_int32), ("obs_mode", C.c_int32),
        ("use_action_mask", C.c_int32), ("max_steps", C.c_int32), ("action_mode", C.c_int32),
        ("auto_reset", C.c_int32), ("action_priority", C.c_int32),
        ("base_reward_scale", C.c_double), ("empty_tile_reward", C.c_double), ("merge_reward", C.c_double),
        ("bonus_scale", C.c_double), ("step_reward", C.c_double), ("endgame_penalty", C.c_double),
        ("invalid_action_penalty", C.c_double), ("obs_log2_scale", C.c_float), ("reserved_f", C.c_float),
    ]


REWARD = {"sum": 0, "log2": 1}
BONUS = {"off": 0, "raw": 1, "log2": 2}
OBS = {"none": 0, "raw": 1, "log2": 2, "onehot": 3}
ACT = {"buffer": 0, "random_legal": 1, "random_any": 2, "priority": 3}

F_MASK, F_CHANGED, F_DONE, F_TRUNC, F_OVERFLOW = 0x0F, 0x10, 0x20, 0x40, 0x80


def make_cfg(reward_mode="sum", bonus_mode="off", obs_mode="none", use_action_mask=True, max_steps=1024,
             action_mode="buffer", auto_reset=False, base_reward_scale=1.0, empty_tile_reward=0.0,
             merge_reward=0.0, bonus_scale=1.0, step_reward=0.0, endgame_penalty=0.0,
             invalid_action_penalty=-1.0, obs_log2_scale=1.0, action_priority=(0, 1, 2, 3)) -> EnvCfg:
    prio = sum(int(a) << (4 * k) for k, a in enumerate(action_priority)) if action_mode == "priority" else 0
    return EnvCfg(REWARD[reward_mode], BONUS[bonus_mode], OBS[obs_mode], int(bool(use_action_mask)),
                  int(max_steps) if max_steps else 0, ACT[action_mode], int(bool(auto_reset)), prio,
                  float(base_reward_scale), float(empty_tile_reward), float(merge_reward), float(bonus_scale),
                  float(step_reward), float(endgame_penalty), float(invalid_action_penalty),
                  float(obs_log2_scale), 0.0)


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("b2048_oracle.c", "oracle_bench.c")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_bench_steps.restype = C.c_double
    return _lib


def _p(a, ct=C.c_void_p):
    return None if a is None else a.ctypes.data_as(ct)


# ------------------------------------------------------------------ packing helpers

def pack_board(tiles) -> int:
    """4x4 raw tile values (Game2048.board) -> packed uint64."""
    t = np.asarray(tiles, dtype=np.int64).reshape(16)
    b = 0
    for i, v in enumerate(t):
        e = 0 if v == 0 else int(v).bit_length() - 1
        assert v == 0 or (1 << e) == v, v
        assert e <= 15
        b |= e << (4 * i)
    return b


def unpack_board(b: int) -> np.ndarray:
    b = int(b)
    e = np.array([(b >> (4 * i)) & 0xF for i in range(16)], dtype=np.int64)
    return np.where(e > 0, np.left_shift(1, e), 0).astype(np.int64).reshape(4, 4)


# ------------------------------------------------------------------ C entry points

def row_move_left(row: int):
    o, m, s = C.c_uint16(), C.c_uint8(), C.c_int32()
    lib().orc_row_move_left(C.c_uint16(row), C.byref(o), C.byref(m), C.byref(s))
    return o.value, m.value, s.value


def row_lut():
    out = np.zeros(65536, np.uint16)
    mrg = np.zeros(65536, np.uint8)
    sc = np.zeros(65536, np.int32)
    L = lib()
    o, m, s = C.c_uint16(), C.c_uint8(), C.c_int32()
    for r in range(65536):
        L.orc_row_move_left(C.c_uint16(r), C.byref(o), C.byref(m), C.byref(s))
        out[r], mrg[r], sc[r] = o.value, m.value, s.value
    return out, mrg, sc


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def reset_many(n, seed, gid0=0, t=0, with_log=False):
    board = np.zeros(n, np.uint64)
    score = np.zeros(n, np.uint32)
    step = np.zeros(n, np.uint32)
    max_exp = np.zeros(n, np.uint8)
    flags = np.zeros(n, np.uint8)
    log = np.zeros((n, 4), np.int32) if with_log else None
    lib().orc_reset_many(_p(board), _p(score), _p(step), _p(max_exp), _p(flags), _p(log), C.c_int64(n),
                         C.c_uint64(seed), C.c_uint64(gid0), C.c_uint32(t))
    st = dict(board=board, score=score, step=step, max_exp=max_exp, flags=flags)
    return (st, log) if with_log else st


def step_many(state, cfg: EnvCfg, seed, gid0, t, action=None, want_obs=False, with_log=False, use_state=True):
    """In-place step of ``state`` (dict from reset_many). Returns dict of outputs."""
    n = state["board"].shape[0]
    merge_sum = np.zeros(n, np.int32)
    reward = np.zeros(n, np.float32)
    reward64 = np.zeros(n, np.float64)
    flags = np.zeros(n, np.uint8)
    action_out = np.zeros(n, np.uint8)
    obs = None
    if want_obs and cfg.obs_mode:
        obs = np.zeros((n, 272 if cfg.obs_mode == 3 else 16), np.float32)
    log = np.zeros((n, 6), np.int32) if with_log else None
    if action is not None:
        action = np.ascontiguousarray(action, dtype=np.uint8)
    lib().orc_step_many(_p(state["board"]), _p(state["board"]),
                        _p(state["score"]) if use_state else None,
                        _p(state["step"]) if use_state else None,
                        _p(state["max_exp"]) if use_state else None,
                        _p(action), _p(action_out), None, C.byref(cfg), _p(merge_sum), _p(reward), _p(reward64),
                        _p(flags), _p(obs), _p(log), C.c_int64(n), C.c_uint64(seed), C.c_uint64(gid0),
                        C.c_uint32(t))
    state["flags"] = flags
    return dict(merge_sum=merge_sum, reward=reward, reward64=reward64, flags=flags, action=action_out, obs=obs,
                spawn_log=log)


def move_many(board, action):
    board = np.ascontiguousarray(board, dtype=np.uint64)
    action = np.ascontiguousarray(action, dtype=np.uint8)
    n = board.shape[0]
    out = np.zeros(n, np.uint64)
    merge_sum = np.zeros(n, np.int32)
    merge_info = np.zeros((n, 4), np.uint8)
    flags = np.zeros(n, np.uint8)
    lib().orc_move_many(_p(board), _p(out), _p(action), _p(merge_sum), _p(merge_info), _p(flags), C.c_int64(n))
    return out, merge_sum, merge_info, flags


def mask_done(board):
    board = np.ascontiguousarray(board, dtype=np.uint64)
    n = board.shape[0]
    mask = np.zeros(n, np.uint8)
    done = np.zeros(n, np.uint8)
    lib().orc_mask_done(_p(board), _p(mask), _p(done), C.c_int64(n))
    return mask, done


def encode_obs(board, obs_mode: str, scale=1.0):
    board = np.ascontiguousarray(board, dtype=np.uint64)
    n = board.shape[0]
    obs = np.zeros((n, 272 if obs_mode == "onehot" else 16), np.float32)
    lib().orc_encode_obs(_p(board), _p(obs), C.c_int32(OBS[obs_mode]), C.c_float(scale), C.c_int64(n))
    return obs


def reverse_scan(x, length, c):
    x = np.ascontiguousarray(x, dtype=np.float32)
    T, B = x.shape
    y = np.zeros_like(x)
    ln = None if length is None else np.ascontiguousarray(length, dtype=np.int32)
    lib().orc_reverse_scan(_p(x), _p(y), _p(ln), C.c_double(c), C.c_int32(T), C.c_int64(B))
    return y


def bench_steps(n, n_steps, n_threads, cfg: EnvCfg, seed=0xB200):
    st = reset_many(n, seed)
    reward = np.zeros(n, np.float32)
    flags = np.zeros(n, np.uint8)
    secs = lib().orc_bench_steps(_p(st["board"]), _p(st["score"]), _p(st["step"]), _p(st["max_exp"]), _p(reward),
                                 _p(flags), C.byref(cfg), C.c_int64(n), C.c_uint64(seed), C.c_uint64(0),
                                 C.c_uint32(1), C.c_int(n_steps), C.c_int(n_threads))
    return secs
