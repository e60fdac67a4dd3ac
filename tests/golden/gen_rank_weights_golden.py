#!/usr/bin/env python
"""Generates tests/golden/rank_weights.npz from the LIVE reference (build container only): the unmodified
``ReinforceAgent._compute_episode_rank_weights`` (src/reinforce_agent.py:681-716) on tie-free and tied rewards."""
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402

ref = load_reference()
fn = ref.agent.ReinforceAgent._compute_episode_rank_weights


def main():
    rng = np.random.default_rng(7)
    out = {}
    for k, (n, conf, ties) in enumerate([(1000, [0.0, 0.5, 1.0, 2.5], False), (37, [1.0, 3.0], False), (5, [0.2, 0.3, 0.5, 1.0, 2.0, 4.0, 8.0], False),
                                         (2000, [0.0, 0.0, 1.0, 3.0], True), (64, [], False)]):
        r = rng.normal(size=n) * 100
        if ties:
            r = np.round(r / 25.0) * 12.5
        me = SimpleNamespace(agent_config=SimpleNamespace(reward_rank_weights=conf if conf else None))
        w = fn(me, [float(x) for x in r])
        out[f"case{k}/reward"] = r
        out[f"case{k}/conf"] = np.asarray(conf, np.float32)
        out[f"case{k}/weights"] = np.asarray(w, np.float32)
        out[f"case{k}/ties"] = np.int64(ties)
    out["n_cases"] = np.int64(5)
    np.savez_compressed(os.path.join(HERE, "rank_weights.npz"), **out)
    print("rank_weights.npz written")


if __name__ == "__main__":
    main()
