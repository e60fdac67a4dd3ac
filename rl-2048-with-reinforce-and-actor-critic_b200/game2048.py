"""Drop-in for the reference's ``src/game2048.py`` — one 4x4 game whose rules run on the GPU.

Same public surface as the reference class (src/game2048.py:11-99): ``reset(seed) -> state``,
``step(action) -> (is_changed, state, new_merged, is_done)``, ``state``, ``board`` (int64 [4,4] of raw
tile values), ``score``, ``step_count``, ``render()``, ``get_action_mask()``, ``_rng``.

The board lives in HBM as one packed uint64; ``step`` launches the same kernels the batched engine
uses (``b2048_move_many`` for the slide/merge preview, ``b2048_step_many`` for the committed step).
Spawn randomness stays where the reference keeps it — a NumPy ``Generator`` seeded by ``reset(seed)``
(src/game2048.py:102-118): the wrapper draws ``rng.integers(n_empty)`` / ``rng.random()`` exactly like
the reference and hands the choice to the kernel (``spawn_replay``), so the same seed yields the same
game as the reference, move for move.
"""
from __future__ import annotations

import ctypes as C
import logging
import secrets

import numpy as np
import torch

from . import _lib
from .batched_env import Batched2048Env, Game2048EnvConfig, _ptr, _stream

Action = int  # 0: up, 1: right, 2: down, 3: left


def unpack_tiles(packed: int) -> np.ndarray:
    e = np.array([(int(packed) >> (4 * i)) & 0xF for i in range(16)], dtype=np.int64)
    return np.where(e > 0, np.left_shift(np.int64(1), e), 0).astype(np.int64).reshape(4, 4)


def pack_tiles(tiles) -> int:
    b = 0
    for i, v in enumerate(np.asarray(tiles, dtype=np.int64).reshape(16)):
        v = int(v)
        if v == 0:
            continue
        e = v.bit_length() - 1
        if (1 << e) != v or e > 15:
            raise ValueError(f"tile {v} is not a power of two <= 32768 (4-bit exponent domain)")
        b |= e << (4 * i)
    return b


class Game2048:
    def __init__(self, size: int = 4, device: str | torch.device = "cuda", _env: Batched2048Env | None = None):
        if size != 4:
            raise ValueError("only size=4 boards are supported by the packed 4-bit-exponent engine")
        self.size = size
        self._benv = _env or Batched2048Env(1, Game2048EnvConfig(max_steps=None), device=device)
        self._dev = self._benv.device
        self.step_count: int = 0
        self.score: int = 0
        self._rng: np.random.Generator = np.random.default_rng()
        self._new_merged: list[int] = []
        self._logger = logging.getLogger(__name__ + ".Game2048")
        self._packed: int = 0
        # The single-env drop-in pays one host round trip per kernel call, so everything a call returns lives in ONE
        # device buffer read back with ONE copy (the Batched2048Env tensors of this env are views into it), and
        # everything a call takes (action, spawn replay) goes up with one copy from a pinned host buffer.
        #   _pack: board i64 @0 | reward f64 @8 | score i32 @16 | step i32 @20 | reward f32 @24 | max_exp u8 @28 |
        #          flags u8 @29 | observation f32[272] @32
        be = self._benv
        self._pack = torch.zeros(32 + 4 * 272, dtype=torch.uint8, device=self._dev)
        be.board = self._pack[0:8].view(torch.int64)
        self._r64 = self._pack[8:16].view(torch.float64)
        be.score = self._pack[16:20].view(torch.int32)
        be.step_count = self._pack[20:24].view(torch.int32)
        be.reward = self._pack[24:28].view(torch.float32)
        be.max_exp = self._pack[28:29]
        be.flags = self._pack[29:30]
        be.max_exp.fill_(2)
        self._obs_dev = self._pack[32:].view(torch.float32)
        self._host = np.zeros(32 + 4 * 272, dtype=np.uint8)
        #   _pre: preview outputs  moved board i64 @0 | merge info u8[4] @8 | flags u8 @12
        self._pre = torch.zeros(16, dtype=torch.uint8, device=self._dev)
        self._tmp_board = self._pre[0:8].view(torch.int64)
        self._tmp_info = self._pre[8:12]
        self._tmp_flags = self._pre[12:13]
        #   _ctl: action u8 @0 | spawn replay u8[2] @1
        self._ctl = torch.zeros(3, dtype=torch.uint8, device=self._dev)
        self._ctl_host = torch.zeros(3, dtype=torch.uint8).pin_memory()
        self._act = self._ctl[0:1]
        self._replay = self._ctl[1:3]

    # ------------------------------------------------------------------ state views
    @property
    def board(self) -> np.ndarray:
        return unpack_tiles(self._packed)

    @board.setter
    def board(self, tiles) -> None:
        self._set_packed(pack_tiles(tiles))

    @property
    def state(self) -> list[list[int]]:
        return self.board.tolist()

    def _set_packed(self, packed: int) -> None:
        self._packed = int(packed)
        self._benv.board.copy_(torch.tensor([np.uint64(self._packed).astype(np.int64)], dtype=torch.int64))

    def _set_seed(self, seed: int | None = None) -> None:
        if seed is None:
            seed = secrets.randbits(64)
        self._rng = np.random.default_rng(seed)

    def _draw_spawn(self, n_empty: int) -> int:
        """The two draws of Game2048._spawn (src/game2048.py:113, :117), encoded for the kernel."""
        idx = int(self._rng.integers(n_empty))
        four = 0 if self._rng.random() < 0.9 else 1
        return 0x80 | (four << 4) | idx

    # ------------------------------------------------------------------ API
    def reset(self, seed: int | None = None) -> list[list[int]]:
        self._set_seed(seed)
        self.step_count = 0
        self.score = 0
        self._ctl_host[1] = self._draw_spawn(16)
        self._ctl_host[2] = self._draw_spawn(15)
        self._ctl.copy_(self._ctl_host, non_blocking=True)
        self._benv.reset_many(spawn_replay=self._replay)
        self._sync_from_device()
        return self.state

    def _sync_from_device(self) -> None:
        """ONE device -> host copy of everything the last call produced."""
        self._host = self._pack.cpu().numpy()
        self._packed = int(self._host[0:8].view(np.uint64)[0])
        self._flags = int(self._host[29])

    # host views of the last synchronised state (Game2048Env reads them instead of touching the device again)
    def _host_reward64(self) -> float:
        return float(self._host[8:16].view(np.float64)[0])

    def _host_score(self) -> int:
        return int(self._host[16:20].view(np.int32)[0])

    def _host_max_exp(self) -> int:
        return int(self._host[28])

    def _host_obs(self, width: int) -> np.ndarray:
        return self._host[32: 32 + 4 * width].view(np.float32).copy()

    def _preview(self, action: int):
        """Game2048._move on a copy: (moved packed board, merged exponents per line, flags)."""
        lib = self._benv._lib
        self._ctl_host[0] = int(action)
        self._ctl.copy_(self._ctl_host, non_blocking=True)
        with torch.cuda.device(self._dev):
            _lib.check(lib.b2048_move_many(self._benv._h, _ptr(self._benv.board), _ptr(self._tmp_board), _ptr(self._act),
                                           None, _ptr(self._tmp_info), _ptr(self._tmp_flags), 1, _stream()),
                       "b2048_move_many")
        pre = self._pre.cpu().numpy()                       # one copy: moved board, merge info, flags
        moved = int(pre[0:8].view(np.uint64)[0])
        info = pre[8:12].copy()
        flags = int(pre[12])
        return moved, info, flags

    def step(self, action: Action) -> tuple[bool, list[list[int]], list[int], bool]:
        if action not in (0, 1, 2, 3):
            raise ValueError("invalid action")              # src/game2048.py:44-45
        self.step_count += 1
        moved, info, pflags = self._preview(action)
        is_changed = bool(pflags & _lib.F_CHANGED)
        merged = []
        for byte in info:                                     # one byte per line: two merged exponents
            for nib in (int(byte) & 0xF, int(byte) >> 4):
                if nib:
                    merged.append(1 << (16 if nib == 1 else nib))
        self._new_merged = merged
        replay = 0
        if is_changed:
            n_empty = sum(1 for i in range(16) if not (moved >> (4 * i)) & 0xF)
            replay = self._draw_spawn(n_empty)
        self._ctl_host[1] = replay
        self._commit(action)
        self.score += sum(merged)
        is_done = bool(self._flags & _lib.F_DONE)
        return is_changed, self.state, list(merged), is_done

    def _commit(self, action: int) -> None:
        """Committed step through the fused kernel (raw rules: no reward shaping, no truncation)."""
        self._ctl_host[0] = int(action)
        self._ctl.copy_(self._ctl_host, non_blocking=True)
        self._benv.step_many(self._act, spawn_replay=self._replay[:1])
        self._sync_from_device()

    def render(self) -> str:
        state = self.state
        width = max(4, max((len(str(x)) for row in state for x in row), default=1))
        sep = "+" + "+".join(["-" * width] * self.size) + "+"
        lines = [sep]
        for row in state:
            lines.append("|" + "|".join((f"{x}" if x else " ").rjust(width) for x in row) + "|")
            lines.append(sep)
        return "\n".join(lines)

    def get_action_mask(self) -> list[int]:
        return [(self._flags >> a) & 1 for a in range(4)]

    def _is_done(self) -> bool:
        return bool(self._flags & _lib.F_DONE)
