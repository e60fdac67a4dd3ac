#!/usr/bin/env python
"""Generates tests/golden/symmetries.npz from the LIVE reference (build container only):

    python tests/golden/gen_symmetry_golden.py

Random boards / legal masks / actions go through the unmodified ``Game2048Env.get_symmetries``
(src/env.py:317-397); the fixture stores the inputs and the reference's 8 variants, re-packed as
4-bit exponents / mask bits.  The oracle restatement (oracle/learner.py: symmetries) is asserted equal here."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import learner  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402
from helpers import random_boards  # noqa: E402

ref = load_reference()
Env = ref.env.Game2048Env


def unpack(b):
    return np.array([(int(b) >> (4 * i)) & 15 for i in range(16)], dtype=np.int64).reshape(4, 4)


def pack(c):
    return sum(int(x) << (4 * i) for i, x in enumerate(np.asarray(c).reshape(16)))


def main():
    rng = np.random.default_rng(2048)
    n = 512
    boards = random_boards(rng, n)
    masks = rng.integers(0, 16, n).astype(np.uint8)
    actions = rng.integers(0, 4, n).astype(np.uint8)
    out_b = np.zeros((8, n), np.uint64); out_m = np.zeros((8, n), np.uint8); out_a = np.zeros((8, n), np.uint8)
    for i in range(n):
        e = unpack(boards[i])
        tiles = np.where(e > 0, 1 << e, 0).astype(np.float32)                    # raw-mode observation
        mask = np.array([(masks[i] >> k) & 1 for k in range(4)], dtype=np.int8)
        syms = Env.get_symmetries({"board": tiles, "action_mask": mask}, int(actions[i]))
        assert len(syms) == 8
        for v, (obs, a) in enumerate(syms):
            t = np.asarray(obs["board"]).astype(np.int64)
            ex = np.where(t > 0, np.round(np.log2(np.maximum(t, 1))).astype(np.int64), 0)
            out_b[v, i] = pack(ex)
            out_m[v, i] = sum(int(x) << k for k, x in enumerate(obs["action_mask"]))
            out_a[v, i] = a
    ob, om, oa = learner.symmetries(boards, masks, actions)
    assert (ob == out_b).all() and (om == out_m).all() and (oa == out_a).all(), "oracle != reference"
    np.savez_compressed(os.path.join(HERE, "symmetries.npz"), boards=boards, masks=masks, actions=actions,
                        out_boards=out_b, out_masks=out_m, out_actions=out_a)
    print("symmetries.npz written; oracle == reference on", n, "cases")


if __name__ == "__main__":
    main()
