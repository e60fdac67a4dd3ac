"""TEST INFRASTRUCTURE ONLY — NumPy restatement of the reference's MLP / REINFORCE / actor-critic
arithmetic (batched over samples instead of per time step; mathematically the same sums).

Pinned against fixtures produced by the live reference (tests/golden/mlp.npz, update_*.npz; see
tests/test_oracle_learner.py).  Reference lines followed (paths relative to the reference repo):
  forward        src/MLP.py:159-196        probs   src/MLP.py:139-156
  returns        src/reinforce_agent.py:255-273
  advantages     src/reinforce_agent.py:276-325, :864-881
  actor grads    src/reinforce_agent.py:502-555, :328-354, :639-678
  critic / TD    src/reinforce_agent.py:403-498, :884-910
  clip / SGD / Adam  src/reinforce_agent.py:835-861, :565-575, :719-770
"""
from __future__ import annotations

import numpy as np


def encode(boards: np.ndarray, obs_mode: str, scale: float = 1.0) -> np.ndarray:
    """packed uint64 boards -> float32 [n, 16] (raw/log2) or [n, 272] (onehot) like env.py:131-150 + MLP.py:41."""
    b = np.asarray(boards, dtype=np.uint64)
    e = np.stack([((b >> np.uint64(4 * i)) & np.uint64(15)).astype(np.int64) for i in range(16)], axis=1)
    if obs_mode == "onehot":
        return np.eye(17, dtype=np.float32)[e].reshape(len(b), 272)
    if obs_mode == "raw":
        return np.where(e > 0, np.left_shift(1, e), 0).astype(np.float32)
    return e.astype(np.float32) * np.float32(scale)


def act(z, mode):
    return np.maximum(z, 0.0) if mode == "ReLU" else 1.0 / (1.0 + np.exp(-z))


def act_grad_from_z(z, mode):
    if mode == "ReLU":
        return (z > 0).astype(np.float32)
    s = 1.0 / (1.0 + np.exp(-z))
    return s * (1.0 - s)


def forward(params, X, mode):
    a = X.astype(np.float32)
    acts, pres = [a], []
    L = len(params["W"])
    for i in range(L):
        z = a @ params["W"][i] + params["b"][i]
        pres.append(z)
        a = act(z, mode) if i < L - 1 else z
        acts.append(a)
    return a, acts, pres


def probs_from_logits(logits, mask_bits=None):
    lg = logits.astype(np.float32)
    if mask_bits is not None:
        m = np.stack([(mask_bits >> a) & 1 for a in range(4)], axis=1).astype(bool)
        lg = np.where(m, lg, np.float32(-1e9))
    mx = lg.max(axis=-1, keepdims=True)
    e = np.exp(lg - mx)
    return e / e.sum(axis=-1, keepdims=True)


def returns(rewards, gamma):
    G, out = 0.0, np.zeros(len(rewards), np.float32)
    for t in reversed(range(len(rewards))):
        G = float(rewards[t]) + gamma * G
        out[t] = G
    return out


def advantages(values_list, weights, mode):
    if mode == "off":
        return [v.astype(np.float32) for v in values_list]
    if mode == "each":
        return [(v - float(v.mean())).astype(np.float32) for v in values_list]
    allv = np.concatenate(values_list)
    allw = np.concatenate([np.full(len(v), w, dtype=np.float32) for v, w in zip(values_list, weights)])
    sw = np.sum(allw)
    if sw < 1e-8:
        mean, std = 0.0, 1.0
    else:
        mean = np.sum(allv * allw) / sw
        std = np.sqrt(np.sum(allw * (allv - mean) ** 2) / sw)
    if mode == "batch":
        return [(v - mean).astype(np.float32) for v in values_list]
    std = max(std, 1e-8)
    return [((v - mean) / std).astype(np.float32) for v in values_list]


def backprop(params, acts, pres, dlogits, mode):
    """Sum over samples of the per-sample outer products: dW_l = A_l^T D_l, db_l = sum D_l."""
    L = len(params["W"])
    gW, gb = [None] * L, [None] * L
    delta = dlogits.astype(np.float32)
    for l in reversed(range(L)):
        gW[l] = acts[l].T @ delta
        gb[l] = delta.sum(0)
        if l > 0:
            delta = (delta @ params["W"][l].T) * act_grad_from_z(pres[l - 1], mode)
    return gW, gb


def clip(gW, gb, max_norm):
    norm = np.sqrt(sum(float(np.linalg.norm(g)) ** 2 for g in gW) + sum(float(np.linalg.norm(g)) ** 2 for g in gb))
    coef = max_norm / max(norm, 1e-8)
    if coef < 1.0:
        gW = [g * coef for g in gW]
        gb = [g * coef for g in gb]
    return gW, gb, norm


class Learner:
    """update_batch restated on flat episode lists: boards/masks/actions/rewards per episode."""

    def __init__(self, actor, critic, *, activation, obs_mode, obs_scale, gamma, lr, baseline, optimizer="sgd",
                 beta1=0.9, beta2=0.999, use_critic=False, critic_lr=1e-3, max_grad_norm=1.0, critic_loss="mse",
                 huber_delta=1.0):
        self.actor = {"W": [w.copy() for w in actor["W"]], "b": [b.copy() for b in actor["b"]]}
        self.critic = None if critic is None else {"W": [w.copy() for w in critic["W"]], "b": [b.copy() for b in critic["b"]]}
        self.k = dict(activation=activation, obs_mode=obs_mode, obs_scale=obs_scale, gamma=gamma, lr=lr, baseline=baseline,
                      optimizer=optimizer, beta1=beta1, beta2=beta2, use_critic=use_critic, critic_lr=critic_lr,
                      max_grad_norm=max_grad_norm, critic_loss=critic_loss, huber_delta=huber_delta)
        self.adam = {}
        self.t = {"actor": 0, "critic": 0}

    def _step(self, name, params, gW, gb, lr, sign):
        k = self.k
        if k["optimizer"] == "sgd":
            for l in range(len(gW)):
                params["W"][l] = params["W"][l] + np.float32(sign * lr) * gW[l]
                params["b"][l] = params["b"][l] + np.float32(sign * lr) * gb[l]
            return
        st = self.adam.setdefault(name, {"mW": [np.zeros_like(w) for w in gW], "vW": [np.zeros_like(w) for w in gW],
                                         "mb": [np.zeros_like(b) for b in gb], "vb": [np.zeros_like(b) for b in gb]})
        self.t[name] += 1
        t, b1, b2 = self.t[name], k["beta1"], k["beta2"]
        for l in range(len(gW)):
            for key_m, key_v, g, pk in (("mW", "vW", gW[l], "W"), ("mb", "vb", gb[l], "b")):
                st[key_m][l] = b1 * st[key_m][l] + (1.0 - b1) * g
                st[key_v][l] = b2 * st[key_v][l] + (1.0 - b2) * (g * g)
                mh = st[key_m][l] / (1.0 - b1 ** t)
                vh = st[key_v][l] / (1.0 - b2 ** t)
                params[pk][l] = (params[pk][l] + sign * lr * mh / (np.sqrt(vh) + 1e-8)).astype(np.float32)

    def update(self, episodes, weights=None):
        """episodes: list of dicts with boards (uint64), masks (uint8), actions, rewards (float)."""
        k = self.k
        n_traj = len(episodes)
        weights = np.ones(n_traj, np.float32) if weights is None else np.asarray(weights, np.float32)
        X = [encode(e["boards"], k["obs_mode"], k["obs_scale"]) for e in episodes]
        out = {}
        if k["use_critic"]:
            td_list, gWc, gbc = [], None, None
            for e, x, w in zip(episodes, X, weights):
                T = len(x)
                v, acts, pres = forward(self.critic, x, k["activation"])
                v = v.reshape(-1)
                v_next = np.concatenate([v[1:], v[-1:]])
                md = np.ones(T, np.float32); md[-1] = 0.0
                target = np.asarray(e["rewards"], np.float32) + k["gamma"] * v_next * md
                td = target - v
                td_list.append(td.astype(np.float32))
                diff = v - target
                g = diff if k["critic_loss"] == "mse" else np.where(np.abs(diff) <= k["huber_delta"], diff,
                                                                     k["huber_delta"] * np.sign(diff))
                g = (g.astype(np.float32) * np.float32(float(w) / (T * n_traj))).reshape(-1, 1)
                a, b = backprop(self.critic, acts, pres, g, k["activation"])
                gWc = a if gWc is None else [p + q for p, q in zip(gWc, a)]
                gbc = b if gbc is None else [p + q for p, q in zip(gbc, b)]
            adv = advantages(td_list, weights, k["baseline"])
            out["td"] = td_list
        else:
            rets = [returns(e["rewards"], k["gamma"]) for e in episodes]
            adv = advantages(rets, weights, k["baseline"])
            out["returns"] = rets
        gW = gb = None
        for e, x, a_e, w in zip(episodes, X, adv, weights):
            T = len(x)
            logits, acts, pres = forward(self.actor, x, k["activation"])
            p = probs_from_logits(logits, np.asarray(e["masks"]))
            onehot = np.eye(4, dtype=np.float32)[np.asarray(e["actions"], np.int64)]
            d = (a_e.reshape(-1, 1) * (onehot - p.astype(np.float32))) * np.float32(float(w) / (T * n_traj))
            a, b = backprop(self.actor, acts, pres, d, k["activation"])
            gW = a if gW is None else [p_ + q for p_, q in zip(gW, a)]
            gb = b if gb is None else [p_ + q for p_, q in zip(gb, b)]
        gW, gb, norm_a = clip(gW, gb, k["max_grad_norm"])
        out["adv"] = adv
        out["actor_grad_norm"] = norm_a
        if k["use_critic"]:
            gWc, gbc, norm_c = clip(gWc, gbc, k["max_grad_norm"])
            out["critic_grad_norm"] = norm_c
        self._step("actor", self.actor, gW, gb, k["lr"], +1.0)
        if k["use_critic"]:
            self._step("critic", self.critic, gWc, gbc, k["critic_lr"], -1.0)
        return out


# ---------------------------------------------------------------------------------------------- bf16 emulation
# The tensor-core gradient path (csrc/b2048_learn_tc.cu) rounds operands to bfloat16 and accumulates in float32.
# ReLU units whose pre-activation is within that rounding error of zero switch on / off relative to the float32
# arithmetic, which shows up as a few-percent difference in a zero-mean gradient sum; to check the KERNELS (layouts,
# indexing, masks) tightly, the tests compare them with this restatement of the same roundings.
def bf16_round(x):
    """round-to-nearest-even float32 -> bfloat16 -> float32"""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def backprop_bf16(params, X, mask_bits, actions, coef, head_mode):
    """Forward + backward of a 3-layer ReLU MLP with the tensor-core path's roundings.  Returns
    (gW, gb, stages) where stages holds the bf16 intermediates A1, H1, H2, d3 (fp32), DL2, DL1."""
    W = [bf16_round(w) for w in params["W"]]
    b1, b2, b3 = bf16_round(params["b"][0]), bf16_round(params["b"][1]), params["b"][2].astype(np.float32)
    A1 = bf16_round(X)
    z1 = A1 @ W[0] + b1
    H1 = bf16_round(np.maximum(z1, 0))
    z2 = H1 @ W[1] + b2
    H2 = bf16_round(np.maximum(z2, 0))
    out = H2 @ W[2] + b3
    if head_mode == 0:
        p = probs_from_logits(out, mask_bits)
        d3 = (coef[:, None] * (np.eye(4, dtype=np.float32)[actions] - p)).astype(np.float32)
    else:
        d3 = coef[:, None].astype(np.float32)
    d3b = bf16_round(d3)
    DL2 = bf16_round((d3b @ W[2].T) * (z2 > 0))
    DL1 = bf16_round((DL2 @ W[1].T) * (z1 > 0))
    gW = [A1.T @ DL1, H1.T @ DL2, H2.T @ d3b]
    gb = [DL1.sum(0), DL2.sum(0), d3.sum(0)]
    return gW, gb, dict(A1=A1, H1=H1, H2=H2, d3=d3, DL2=DL2, DL1=DL1)


# ---------------------------------------------------------------------------------------------- D4 symmetries
def symmetries(boards, masks, actions):
    """Game2048Env.get_symmetries (src/env.py:317-397) restated on packed boards, 4-bit masks and actions.
    Returns (boards [8, n] uint64, masks [8, n] uint8, actions [8, n] uint8) in the reference's variant order:
    identity + three np.rot90(k=1) turns (action (a-1)%4, mask np.roll(-1)), then the same four starting from
    np.fliplr (actions 1 <-> 3, mask [0,3,2,1])."""
    b = np.asarray(boards, dtype=np.uint64)
    n = len(b)
    cells = np.stack([((b >> np.uint64(4 * i)) & np.uint64(15)).astype(np.int64) for i in range(16)], axis=1).reshape(n, 4, 4)
    mk = np.stack([(np.asarray(masks) >> k) & 1 for k in range(4)], axis=1).astype(np.int64)
    ac = np.asarray(actions).astype(np.int64)
    ob, om, oa = [], [], []

    def pack(c):
        flat = c.reshape(n, 16)
        out = np.zeros(n, np.uint64)
        for i in range(16):
            out |= flat[:, i].astype(np.uint64) << np.uint64(4 * i)
        return out

    for flipped in (False, True):
        cb = cells[:, :, ::-1].copy() if flipped else cells.copy()                  # np.fliplr per board
        ca = np.where(ac == 1, 3, np.where(ac == 3, 1, ac)) if flipped else ac.copy()
        cm = mk[:, [0, 3, 2, 1]].copy() if flipped else mk.copy()
        for _ in range(4):
            ob.append(pack(cb))
            om.append(sum(cm[:, k] << k for k in range(4)).astype(np.uint8))
            oa.append(ca.astype(np.uint8))
            cb = np.rot90(cb, k=1, axes=(1, 2)).copy()
            ca = (ca - 1) % 4
            cm = np.roll(cm, shift=-1, axis=1)
    return np.stack(ob), np.stack(om), np.stack(oa)
