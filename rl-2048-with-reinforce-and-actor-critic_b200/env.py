"""Drop-in for the reference's ``src/env.py``: ``Game2048Env.reset()/step()`` with the reference's
observation dict, reward shaping, termination / truncation and ``info`` keys — computed by the fused
CUDA step kernel on a single packed board (B = 1 view of the batched engine).

``gymnasium`` is optional: when it is importable the class derives from ``gym.Env`` and exposes real
``spaces``; otherwise light stand-ins with the same attributes (``.n``, ``.contains``, ``.shape``) are used.
"""
from __future__ import annotations

import logging
from typing import Any

import numpy as np
import torch

from . import _lib
from .batched_env import Batched2048Env, Game2048EnvConfig
from .game2048 import Action, Game2048

try:  # pragma: no cover - depends on the image
    import gymnasium as gym
    from gymnasium import spaces
    _EnvBase = gym.Env
except Exception:  # gymnasium is not installed in the build image
    class _Space:
        pass

    class _Discrete(_Space):
        def __init__(self, n):
            self.n = int(n)

        def contains(self, x) -> bool:
            try:
                return int(x) == x and 0 <= int(x) < self.n
            except Exception:
                return False

    class _Box(_Space):
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class _Dict(_Space):
        def __init__(self, d):
            self.spaces = dict(d)

    class spaces:  # noqa: N801 - mimics the gymnasium module name
        Space, Discrete, Box, Dict = _Space, _Discrete, _Box, _Dict

    class _EnvBase:
        metadata: dict = {}

        def reset(self, *, seed=None, options=None):
            return None


class Game2048Env(_EnvBase):
    metadata = {"render_modes": ["human", "ansi"]}

    def __init__(self, config: Game2048EnvConfig | None = None, device: str | torch.device = "cuda") -> None:
        super().__init__()
        self.config = config or Game2048EnvConfig()
        self._benv = Batched2048Env(1, self.config, device=device)          # validates modes like env.py:79-110
        self.game = Game2048(size=self.config.size, _env=self._benv)
        self._max_num: float = 16.0
        self._step_count: int = 0
        self.max_tile_seen: int = 4
        self._logger = logging.getLogger(__name__ + ".Game2048Env")
        if not self._logger.handlers:
            self._logger.addHandler(logging.NullHandler())
        self.action_space = spaces.Discrete(4)
        self.observation_space = self._build_observation_space()
        # observation and float64 reward live in the game's packed device buffer (one read-back per step)
        self._obs_buf = self.game._obs_dev[: self._benv.obs_width].view(1, self._benv.obs_width)
        self._r64 = self.game._r64

    @property
    def state(self) -> list[list[int]]:
        return self.game.state

    def _build_observation_space(self):
        size, mode = self.config.size, self.config.obs_mode
        if mode == "raw":
            board_space = spaces.Box(low=0, high=2 ** self._max_num, shape=(size, size), dtype=np.float32)
        elif mode == "log2":
            board_space = spaces.Box(low=0.0, high=self._max_num, shape=(size, size), dtype=np.float32)
        elif mode == "onehot":
            board_space = spaces.Box(low=0.0, high=1.0, shape=(size, size, int(self._max_num) + 1), dtype=np.float32)
        else:
            raise ValueError(f"Unsupported obs_mode: {mode}")
        if self.config.use_action_mask:
            return spaces.Dict({"board": board_space,
                                "action_mask": spaces.Box(low=0, high=1, shape=(self.action_space.n,), dtype=np.int8)})
        return board_space

    def _shape_obs(self, flat: np.ndarray):
        if self.config.obs_mode == "onehot":
            board = flat.reshape(4, 4, 17)
        else:
            board = flat.reshape(4, 4)
        if self.config.use_action_mask:
            return {"board": board, "action_mask": np.array(self.game.get_action_mask(), dtype=np.int8)}
        return board

    def _get_obs(self):
        return self._shape_obs(self._benv.encode_obs().cpu().numpy()[0])

    def _obs_from_host(self):
        return self._shape_obs(self.game._host_obs(self._benv.obs_width))

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        super().reset(seed=seed)
        self._step_count = 0
        self.max_tile_seen = 4
        state = self.game.reset(seed=seed)                     # also resets score / step / max_exp on the device
        return self._get_obs(), {"score": self.game.score, "raw_state": state}

    def step(self, action: Action):
        assert self.action_space.contains(action), f"Invalid action: {action}"    # env.py:265
        self._step_count += 1
        g = self.game
        g.step_count += 1
        moved, info, pflags = g._preview(action)
        is_changed = bool(pflags & _lib.F_CHANGED)
        merged = []
        for byte in info:
            for nib in (int(byte) & 0xF, int(byte) >> 4):
                if nib:
                    merged.append(1 << (16 if nib == 1 else nib))
        g._new_merged = merged
        replay = 0
        if is_changed:
            n_empty = sum(1 for i in range(16) if not (moved >> (4 * i)) & 0xF)
            replay = g._draw_spawn(n_empty)
        g._ctl_host[0] = int(action)
        g._ctl_host[1] = replay
        g._ctl.copy_(g._ctl_host, non_blocking=True)
        # the fused kernel: move, spawn, reward (float64, env.py:197-261), done / truncated, mask, observation
        self._benv.step_many(g._act, spawn_replay=g._replay[:1], reward64_out=self._r64, obs_out=self._obs_buf)
        g._sync_from_device()                                   # the step's only read-back
        g.score = g._host_score()
        flags = g._flags
        reward = g._host_reward64()
        self.max_tile_seen = 1 << g._host_max_exp()
        terminated = bool(flags & _lib.F_DONE)
        truncated = bool(flags & _lib.F_TRUNC)
        invalid_action = (not is_changed) and (not terminated)
        obs = self._obs_from_host()
        info_d = {"score": g.score, "raw_state": g.state, "merged": merged, "invalid_action": invalid_action,
                  "step_index": self._step_count}
        return obs, reward, terminated, truncated, info_d

    def render(self, mode: str = "human") -> str | None:
        text = self.game.render()
        if mode == "human":
            print(text)
            return None
        if mode == "ansi":
            return text
        raise NotImplementedError(f"Unsupported render mode: {mode}")

    @staticmethod
    def get_symmetries(obs, action: Action):
        """The 8 dihedral variants of (observation, action) in the reference's order (src/env.py:317-397):
        identity + three counter-clockwise quarter turns, then the same four for the left-right mirror."""
        if isinstance(obs, dict):
            board, mask = obs["board"], obs["action_mask"]
        else:
            board, mask = obs, None

        def emit(b, m, a):
            return ({"board": b, "action_mask": m}, a) if m is not None else (b, a)

        out = []
        for flipped in (False, True):
            b = np.fliplr(board.copy()) if flipped else board.copy()
            a = ({1: 3, 3: 1}.get(action, action)) if flipped else action
            m = None if mask is None else (mask[[0, 3, 2, 1]] if flipped else mask.copy())
            for _ in range(4):
                out.append(emit(b, m, a))
                b = np.rot90(b, k=1, axes=(0, 1))
                a = (a - 1) % 4
                m = None if m is None else np.roll(m, shift=-1)
        return out
