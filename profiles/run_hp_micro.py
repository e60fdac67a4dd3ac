"""Micro-benchmark of the split-fp16 update kernels on random boards: b2048_mlp_forward precision 3 (fwd_hp_kernel without
image stores) and b2048_mlp_backward precision 3 / 1 (all kernels of one chunk).  Run plain, then under an ncu launch list."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, b2048
from b2048 import _lib
from helpers import random_boards
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.cuda.set_device(0)
lib = _lib.load()
if os.environ.get("NO_PIPE"):
    b2048.debug_set("no_update_pipe", True)
env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(obs_mode="log2", obs_log2_scale=0.0625))
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig())
rng = np.random.default_rng(0)
boards = torch.from_numpy(random_boards(rng, n).view(np.int64)).cuda()
flags = torch.full((n,), 0xF, dtype=torch.uint8, device="cuda")
acts = torch.from_numpy(rng.integers(0, 4, n).astype(np.uint8)).cuda()
coef = torch.from_numpy((rng.normal(size=n) * 1e-6).astype(np.float32)).cuda()
net = agent._actor
out = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
wsf = int(lib.b2048_backward_workspace_floats(C.byref(net.desc), n))
ws = torch.zeros(wsf, dtype=torch.float32, device="cuda")


def timed(fn, name):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per call ({n} samples)", flush=True)


timed(lambda: _lib.check(lib.b2048_mlp_forward(agent._h, p(boards), C.byref(net.desc), p(out), n, 3, st()), "fwd"), "mlp_forward precision 3 (no images)")
timed(lambda: _lib.check(lib.b2048_mlp_forward(agent._h, p(boards), C.byref(net.desc), p(out), n, 1, st()), "fwd"), "mlp_forward precision 1 (bf16)")
for prec in (3, 1):
    def bw():
        net.grad.zero_()
        _lib.check(lib.b2048_mlp_backward(agent._h, p(boards), p(flags), p(acts), p(coef), C.byref(net.desc), p(net.grad), n, 0, p(ws), wsf,
                                          n, prec, st()), "bwd")
    timed(bw, f"mlp_backward precision {prec}")
if len(sys.argv) > 3:       # in-kernel phase clocks of the forward kernel
    b2048.debug_set("tc_clocks", True)
    _lib.check(lib.b2048_mlp_forward(agent._h, p(boards), C.byref(net.desc), p(out), n, 3, st()), "fwd")
    _lib.check(lib.b2048_mlp_backward(agent._h, p(boards), p(flags), p(acts), p(coef), C.byref(net.desc), p(net.grad), n, 0, p(ws), wsf,
                                      n, 3, st()), "bwd")
    b2048.debug_set("tc_clocks", False)
