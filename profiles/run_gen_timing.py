"""Timing of the shape-generic tensor-core kernels (csrc/b2048_mlp_gen.cu) on the reference's documented one-hot
272-256-128-64 network: policy step (one fp16 MMA per product), float32-grade forward, and a full gradient pass
(gen_mlp_kernel in update mode + the gen_dw_kernel GEMMs), against the fp32 CUDA-core kernels.  CUDA events, after warm-up.
usage: run_gen_timing.py [n_samples] [iters] [skip_fp32]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import b2048
from b2048 import _lib
from helpers import random_boards

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
skip32 = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
hidden = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [256, 128, 64]
obs = sys.argv[5] if len(sys.argv) > 5 else "onehot"
lib = _lib.load()
env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(obs_mode=obs, obs_log2_scale=0.0625))
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=hidden, activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig(use_critic=False, model_seed=0))
rng = np.random.default_rng(0)
boards = torch.from_numpy(random_boards(rng, n).view(np.int64)).cuda()
flags = torch.full((n,), 15, dtype=torch.uint8, device="cuda")
acts = torch.from_numpy(rng.integers(0, 4, n).astype(np.uint8)).cuda()
coef = torch.from_numpy((rng.normal(size=n) * 1e-6).astype(np.float32)).cuda()
net = agent._actor
ptr = lambda t: C.c_void_p(t.data_ptr())
stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
act_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
res = {"n": n, "network": f"{obs} {net.dims}"}
flops = 2 * sum(a * b for a, b in zip(net.dims[:-1], net.dims[1:]))
for name, prec in (("policy_step_tc_ms", 1), ("policy_step_fp32_ms", 0)):
    if prec == 0 and skip32:
        continue
    res[name] = timed(lambda: agent.policy_step(boards, flags, act_out, seed=1, gid0=0, t=1, precision=prec))
for name, prec in (("forward_split_ms", 3), ("forward_fp32_ms", 0)):
    if prec == 0 and skip32:
        continue
    res[name] = timed(lambda: _lib.check(lib.b2048_mlp_forward(agent._h, ptr(boards), C.byref(net.desc), ptr(out), n, prec, stream()), "fwd"))
ws_floats = int(lib.b2048_backward_workspace_floats(C.byref(net.desc), n))
ws = torch.zeros(ws_floats, dtype=torch.float32, device="cuda")
for name, prec in (("backward_tc_ms", 3), ("backward_fp32_ms", 0)):
    if prec == 0 and skip32:
        continue

    def bw():
        net.grad.zero_()
        _lib.check(lib.b2048_mlp_backward(agent._h, ptr(boards), ptr(flags), ptr(acts), ptr(coef), C.byref(net.desc), ptr(net.grad), n, 0,
                                          ptr(ws), ws_floats, n, prec, stream()), "bwd")
    res[name] = timed(bw)
res["policy_tc_steps_per_s"] = n / (res["policy_step_tc_ms"] * 1e-3)
res["policy_tc_tflops"] = res["policy_tc_steps_per_s"] * flops / 1e12
res["backward_tc_samples_per_s"] = n / (res["backward_tc_ms"] * 1e-3)
print(json.dumps(res))
