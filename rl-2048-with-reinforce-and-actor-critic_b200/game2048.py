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
        self._replay = torch.zeros(2, dtype=torch.uint8, device=self._dev)
        self._act = torch.zeros(1, dtype=torch.uint8, device=self._dev)
        self._tmp_board = torch.zeros(1, dtype=torch.int64, device=self._dev)
        self._tmp_info = torch.zeros(4, dtype=torch.uint8, device=self._dev)
        self._tmp_flags = torch.zeros(1, dtype=torch.uint8, device=self._dev)

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
        r = np.array([self._draw_spawn(16), self._draw_spawn(15)], dtype=np.uint8)
        self._replay.copy_(torch.from_numpy(r))
        self._benv.reset_many(spawn_replay=self._replay)
        self._sync_from_device()
        return self.state

    def _sync_from_device(self) -> None:
        self._packed = int(self._benv.board.cpu().numpy().view(np.uint64)[0])
        self._flags = int(self._benv.flags.cpu()[0])

    def _preview(self, action: int):
        """Game2048._move on a copy: (moved packed board, merged exponents per line, flags)."""
        lib = self._benv._lib
        self._act.fill_(int(action))
        with torch.cuda.device(self._dev):
            _lib.check(lib.b2048_move_many(self._benv._h, _ptr(self._benv.board), _ptr(self._tmp_board), _ptr(self._act),
                                           None, _ptr(self._tmp_info), _ptr(self._tmp_flags), 1, _stream()),
                       "b2048_move_many")
        moved = int(self._tmp_board.cpu().numpy().view(np.uint64)[0])
        info = self._tmp_info.cpu().numpy()
        flags = int(self._tmp_flags.cpu()[0])
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
        self._replay[0] = replay
        self._commit(action)
        self.score += sum(merged)
        is_done = bool(self._flags & _lib.F_DONE)
        return is_changed, self.state, list(merged), is_done

    def _commit(self, action: int) -> None:
        """Committed step through the fused kernel (raw rules: no reward shaping, no truncation)."""
        self._act.fill_(int(action))
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
