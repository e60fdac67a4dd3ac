"""Drop-in for the reference's ``src/MLP.py``: same names, signatures and error behaviour; the
arithmetic (``z = a @ W + b``, Sigmoid / ReLU, masked softmax) runs in the library's CUDA kernels.

Parameters keep the reference's layout — ``{"W": [float32 (in, out)], "b": [float32 (out,)]}`` — and
the ``.npz`` interchange format (``n_layers, W_i, b_i``; src/MLP.py:97-126) is byte-compatible, so
weights move freely between the two implementations.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch

from . import _lib
from ._lib import ACTV, OBS, MlpDesc
from .batched_env import _ptr, _stream, get_handle


@dataclass
class MLPConfig:
    """Same fields / defaults as the reference's MLPConfig (src/MLP.py:10-19)."""
    hidden_sizes: list[int] = field(default_factory=list)
    activation: str = "Sigmoid"                 # "Sigmoid" / "ReLU"
    init_distribution: str = "XavierNormal"     # "XavierNormal" / "HeNormal" / "XavierUniform" / "Normal"
    last_init_normal: bool = True

    @property
    def num_layers(self) -> int:
        return len(self.hidden_sizes)


def encode_observation(obs) -> tuple[np.ndarray, np.ndarray | None]:
    """src/MLP.py:22-43: flatten the board to float32, pass the mask through."""
    if isinstance(obs, dict):
        board, action_mask = obs["board"], obs["action_mask"]
    else:
        board, action_mask = obs, None
    return np.asarray(board).astype(np.float32).flatten(), action_mask


def init_model_params(input_dim: int, hidden_sizes: list[int], output_dim: int, rng: np.random.Generator,
                      init_distribution: str = "normal", last_init_normal: bool = True) -> dict[str, Any]:
    """src/MLP.py:45-94.  Draws come from the caller's NumPy generator in the reference's order, so the
    same ``model_seed`` gives the same weights.  (The reference's ``last_init_normal`` branch never
    fires — its index test compares against ``len(layer_sizes) - 1`` — and its own default
    ``init_distribution="normal"`` raises ValueError; both behaviours are kept.)"""
    sizes = [input_dim] + list(hidden_sizes) + [output_dim]
    Ws, bs = [], []
    for fan_in, fan_out in zip(sizes[:-1], sizes[1:]):
        if init_distribution == "XavierNormal":
            W = rng.normal(0.0, np.sqrt(2.0 / (fan_in + fan_out)), size=(fan_in, fan_out)).astype(np.float32)
        elif init_distribution == "HeNormal":
            W = rng.normal(0.0, np.sqrt(2.0 / fan_in), size=(fan_in, fan_out)).astype(np.float32)
        elif init_distribution == "XavierUniform":
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            W = rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)
        elif init_distribution == "Normal":
            W = rng.standard_normal((fan_in, fan_out), dtype=np.float32) * 0.01
        else:
            raise ValueError(f"Unsupported init_distribution: {init_distribution}")
        Ws.append(W)
        bs.append(np.zeros((fan_out,), dtype=np.float32))
    return {"W": Ws, "b": bs}


def load_model_params(file_path: str | None = "params.npz") -> dict[str, Any]:
    data = np.load(file_path)
    n = int(data["n_layers"])
    return {"W": [data[f"W_{i}"] for i in range(n)], "b": [data[f"b_{i}"] for i in range(n)]}


def save_model_params(params: dict[str, Any], file_path: str | None = "params.npz") -> None:
    Ws, bs = list(params["W"]), list(params["b"])
    assert len(Ws) == len(bs), "W/b layer count mismatch"
    out = {"n_layers": np.array(len(Ws), dtype=np.int64)}
    for i, (W, b) in enumerate(zip(Ws, bs)):
        out[f"W_{i}"] = np.asarray(W)
        out[f"b_{i}"] = np.asarray(b)
    np.savez(file_path, **out)


# ---------------------------------------------------------------------------------------------- device side

class DeviceMLP:
    """Flat float32 parameter vector on the GPU (W_0, b_0, W_1, b_1, ...) plus the b2048_mlp_desc that
    points into it.  This is what every kernel consumes."""

    def __init__(self, params: dict[str, Any], activation: str, obs_mode: str = "log2", obs_log2_scale: float = 1.0,
                 device: torch.device | str = "cuda"):
        if activation not in ACTV:
            raise ValueError(f"Unsupported activation: {activation}")          # src/MLP.py:136
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.activation = activation
        self.obs_mode = obs_mode
        self.obs_log2_scale = float(obs_log2_scale)
        Ws = [np.ascontiguousarray(W, dtype=np.float32) for W in params["W"]]
        bs = [np.ascontiguousarray(b, dtype=np.float32) for b in params["b"]]
        assert len(Ws) == len(bs), "W/b layer count mismatch"
        if len(Ws) > _lib.B2048_MAX_LAYERS:
            raise ValueError(f"at most {_lib.B2048_MAX_LAYERS} layers supported")
        self.dims = [Ws[0].shape[0]] + [W.shape[1] for W in Ws]
        self.n_layers = len(Ws)
        flat = np.concatenate([np.concatenate([W.reshape(-1), b.reshape(-1)]) for W, b in zip(Ws, bs)])
        self.theta = torch.from_numpy(flat).to(self.device)
        self.n_params = int(flat.size)
        self._build_desc()

    def _build_desc(self):
        d = MlpDesc()
        d.n_layers = self.n_layers
        d.activation = ACTV[self.activation]
        d.obs_mode = OBS[self.obs_mode]
        d.obs_log2_scale = self.obs_log2_scale
        off = 0
        base = self.theta.data_ptr()
        self.offsets = []
        for l in range(self.n_layers):
            d.dims[l] = self.dims[l]
            nW = self.dims[l] * self.dims[l + 1]
            d.W[l] = base + 4 * off
            d.b[l] = base + 4 * (off + nW)
            self.offsets.append((off, off + nW, off + nW + self.dims[l + 1]))
            off += nW + self.dims[l + 1]
        d.dims[self.n_layers] = self.dims[self.n_layers]
        self.desc = d

    def to_params(self) -> dict[str, Any]:
        flat = self.theta.detach().cpu().numpy()
        Ws, bs = [], []
        for l, (o0, o1, o2) in enumerate(self.offsets):
            Ws.append(flat[o0:o1].reshape(self.dims[l], self.dims[l + 1]).copy())
            bs.append(flat[o1:o2].copy())
        return {"W": Ws, "b": bs}

    def load_params(self, params: dict[str, Any]) -> None:
        dims = [np.asarray(params["W"][0]).shape[0]] + [np.asarray(W).shape[1] for W in params["W"]]
        if dims != self.dims:
            raise ValueError(f"parameter shapes {dims} do not match the network {self.dims}")
        flat = np.concatenate([np.concatenate([np.asarray(W, np.float32).reshape(-1), np.asarray(b, np.float32).reshape(-1)])
                               for W, b in zip(params["W"], params["b"])])
        self.theta.copy_(torch.from_numpy(flat))


def _dense_forward(params, x: np.ndarray, activation_mode: str):
    if activation_mode not in ACTV:
        raise ValueError(f"Unsupported activation: {activation_mode}")
    single = x.ndim == 1
    xb = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    dev = torch.device("cuda", torch.cuda.current_device())
    net = DeviceMLP(params, activation_mode, "raw", 1.0, dev)
    n = xb.shape[0]
    xd = torch.from_numpy(xb).to(dev)
    L = net.n_layers
    acts = [torch.empty((n, net.dims[l + 1]), dtype=torch.float32, device=dev) for l in range(L)]
    pres = [torch.empty((n, net.dims[l + 1]), dtype=torch.float32, device=dev) for l in range(L)]
    act_ptrs = (C.c_void_p * (L + 1))(None, *[a.data_ptr() for a in acts])
    pre_ptrs = (C.c_void_p * L)(*[p.data_ptr() for p in pres])
    lib = _lib.load()
    _lib.check(lib.b2048_dense_forward(get_handle(dev), _ptr(xd), C.byref(net.desc), act_ptrs, pre_ptrs, n, _stream()),
               "b2048_dense_forward")
    acts_np = [xb] + [a.cpu().numpy() for a in acts]
    pres_np = [p.cpu().numpy() for p in pres]
    if single:
        acts_np = [a[0] for a in acts_np]
        pres_np = [p[0] for p in pres_np]
    return acts_np[-1], acts_np, pres_np


def forward_logits(params: dict[str, list[np.ndarray]], x: np.ndarray, activation_mode: str
                   ) -> tuple[np.ndarray, list[np.ndarray], list[np.ndarray]]:
    """src/MLP.py:159-196: returns (logits, activations [a_0..a_L], pre_activations [z_0..z_{L-1}]);
    accepts a single vector [in] or a batch [T, in].  Computed by b2048_dense_forward on the GPU."""
    assert len(params["W"]) == len(params["b"]), "W/b layer count mismatch"
    return _dense_forward(params, np.asarray(x), activation_mode)


def logits_to_probs(logits: np.ndarray, action_mask: np.ndarray | None = None) -> np.ndarray:
    """src/MLP.py:139-156 (masked fill -1e9, max-subtracted softmax) on the GPU via torch elementwise ops on
    the caller's host arrays.  The fused kernels (b2048_policy_step) never call this; it exists for API parity."""
    dev = torch.device("cuda", torch.cuda.current_device())
    lg = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32)).to(dev)
    if action_mask is not None:
        m = torch.from_numpy(np.ascontiguousarray(action_mask).astype(bool)).to(dev)
        lg = torch.where(m, lg, torch.full_like(lg, -1e9))
    mx = lg.max(dim=-1, keepdim=True).values
    e = torch.exp(lg - mx)
    return (e / e.sum(dim=-1, keepdim=True)).cpu().numpy()
