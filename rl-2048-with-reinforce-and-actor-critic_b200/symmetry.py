"""D4 symmetry augmentation on packed boards (reference src/env.py:317-397 ``get_symmetries`` and
src/reinforce_agent.py:773-808 ``_augment_trajectories``): the 8 (board, action, mask) variants are
nibble permutations of the packed board, a relabelling of the action and a permutation of the 4 mask
bits.  Pure data movement on the device (no arithmetic), expressed with torch indexing."""
from __future__ import annotations

import numpy as np
import torch


def _variant_tables():
    """cell permutation / action map / mask-bit permutation of the 8 variants, in the reference's order."""
    base = np.arange(16).reshape(4, 4)
    perms, amaps, mperms = [], [], []
    for flipped in (False, True):
        b = np.fliplr(base.copy()) if flipped else base.copy()
        amap = np.array([0, 3, 2, 1]) if flipped else np.arange(4)        # action a -> amap[a]
        mperm = np.array([0, 3, 2, 1]) if flipped else np.arange(4)       # new_mask[i] = mask[mperm[i]]
        for _ in range(4):
            perms.append(b.reshape(16).copy())       # new_cells[i] = cells[perm[i]]
            amaps.append(amap.copy())
            mperms.append(mperm.copy())
            b = np.rot90(b, k=1)
            amap = (amap - 1) % 4
            mperm = np.roll(mperm, -1)
    return perms, amaps, mperms


_PERMS, _AMAPS, _MPERMS = _variant_tables()


def transform_boards(boards: torch.Tensor, variant: int) -> torch.Tensor:
    shifts = torch.arange(16, device=boards.device, dtype=torch.int64) * 4
    cells = (boards.unsqueeze(-1) >> shifts) & 15
    perm = torch.as_tensor(_PERMS[variant], device=boards.device)
    return (cells[..., perm] << shifts).sum(-1)


def transform_actions(actions: torch.Tensor, variant: int) -> torch.Tensor:
    amap = torch.as_tensor(_AMAPS[variant], device=actions.device, dtype=torch.uint8)
    return amap[actions.long()]


def transform_flags(flags: torch.Tensor, variant: int) -> torch.Tensor:
    mperm = _MPERMS[variant]
    out = flags & 0xF0
    for i in range(4):
        out = out | (((flags >> int(mperm[i])) & 1) << i)
    return out


def augment_rollout(ro):
    """Rollout with 8x the episodes (every dihedral variant), weights repeated, n_traj = 8 B."""
    from .reinforce_agent import Rollout
    T = ro.T
    boards = torch.cat([transform_boards(ro.boards, v) for v in range(8)], dim=1)
    flags = torch.cat([transform_flags(ro.flags, v) for v in range(8)], dim=1)
    actions = torch.cat([transform_actions(ro.actions, v) for v in range(8)], dim=1)
    rewards = ro.rewards.repeat(1, 8)
    length = ro.length.repeat(8)
    w = None if ro.ep_weight is None else ro.ep_weight.repeat(8)
    return Rollout(boards.contiguous(), flags.contiguous(), actions.contiguous(), rewards.contiguous(), length, T, w,
                   8 * ro.B)
