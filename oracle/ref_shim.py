"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference from /root/reference.

Used only in the build container (where /root/reference exists) by
``tests/golden/gen_golden.py`` and by the ``not gpu`` tests that pin the oracle
against the live reference.  Nothing here runs on the GPU box and nothing under
the product package imports it.

The reference's ``src/env.py:7-8`` imports ``gymnasium`` which is not installed
here and cannot be fetched (no network).  gymnasium only supplies the
``gym.Env`` base class, the ``spaces.*`` descriptors and the ``contains``
assert (``env.py:70-71``, ``:85-124``, ``:180``, ``:265``): none of the path's
arithmetic.  A minimal stand-in is registered in ``sys.modules`` before the
import so the reference code itself runs unchanged.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B2048_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "game2048.py"))


def _install_gymnasium_stub() -> None:
    if "gymnasium" in sys.modules:
        return
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Env:  # gymnasium.Env: only reset(seed=) bookkeeping is used
        metadata: dict = {}

        def reset(self, *, seed=None, options=None):
            self._np_random_seed = seed
            return None

    class Space:
        pass

    class Discrete(Space):
        def __init__(self, n):
            self.n = int(n)

        def contains(self, x) -> bool:
            try:
                xi = int(x)
            except Exception:
                return False
            return xi == x and 0 <= xi < self.n

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Dict(Space):
        def __init__(self, d):
            self.spaces = dict(d)

    gym.Env = Env
    spaces.Space, spaces.Discrete, spaces.Box, spaces.Dict = Space, Discrete, Box, Dict
    gym.spaces = spaces
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = spaces


def load_reference():
    """Return a namespace with the reference's modules (game2048, env, MLP, agent, runner)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_gymnasium_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    ns = types.SimpleNamespace()
    ns.game2048 = importlib.import_module("src.game2048")
    ns.env = importlib.import_module("src.env")
    ns.MLP = importlib.import_module("src.MLP")
    ns.agent = importlib.import_module("src.reinforce_agent")
    try:
        ns.runner = importlib.import_module("runner")
    except Exception:  # runner pulls optional deps; not needed for the path
        ns.runner = None
    return ns


class ReplayRng:
    """Stand-in for ``Game2048._rng`` that replays externally chosen spawns.

    ``Game2048._spawn`` (game2048.py:108-118) makes exactly two calls:
    ``rng.integers(len(empties))`` and ``rng.random()``.  Feed (k, is_four)
    pairs; ``integers`` returns k, ``random`` returns 1.0 for a 4, 0.0 for a 2.
    """

    def __init__(self):
        self.queue: list[tuple[int, int]] = []
        self._pending_val = 0

    def push(self, k: int, is_four: int) -> None:
        self.queue.append((int(k), int(is_four)))

    def integers(self, n):
        k, four = self.queue.pop(0)
        assert 0 <= k < int(n), (k, n)
        self._pending_val = four
        return k

    def random(self):
        return 1.0 if self._pending_val else 0.0
