"""b2048_mlp_backward precision 3 on n random samples at several chunk sizes: do the activation images stay in L2?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, b2048
from b2048 import _lib
from helpers import random_boards
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 3
chunks = [int(c) for c in sys.argv[3:]] or [1 << 20, 1 << 18, 1 << 16, 37888, 18944]
torch.cuda.set_device(0)
lib = _lib.load()
env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(obs_mode="log2", obs_log2_scale=0.0625))
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig())
rng = np.random.default_rng(0)
boards = torch.from_numpy(random_boards(rng, n).view(np.int64)).cuda()
flags = torch.full((n,), 0xF, dtype=torch.uint8, device="cuda")
acts = torch.from_numpy(rng.integers(0, 4, n).astype(np.uint8)).cuda()
coef = torch.from_numpy((rng.normal(size=n) * 1e-6).astype(np.float32)).cuda()
net = agent._actor
p = lambda t: C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
ref = None
split = os.environ.get("PIPE_SPLIT")
if split:
    nb, n2, n13, r16 = [int(x) for x in split.split(",")]
    b2048.debug_set("pipe_split", nb | (n2 << 8) | (n13 << 16) | (r16 << 24))
for chunk in chunks:
    b2048.debug_set("no_update_pipe", chunk > 0 and os.environ.get("NO_PIPE") is not None)
    wsf = int(lib.b2048_backward_workspace_floats(C.byref(net.desc), chunk))
    ws = torch.zeros(wsf, dtype=torch.float32, device="cuda")
    def bw():
        net.grad.zero_()
        _lib.check(lib.b2048_mlp_backward(agent._h, p(boards), p(flags), p(acts), p(coef), C.byref(net.desc), p(net.grad), n, 0, p(ws), wsf,
                                          chunk, prec, st()), "bwd")
    bw(); torch.cuda.synchronize()
    g = net.grad.clone()
    if ref is None:
        ref = g
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bw(); bw(); e1.record(); torch.cuda.synchronize()
    print(f"chunk {chunk}: {e0.elapsed_time(e1) / 2 * 1e3 / (n / 2**20):.1f} us per 1M samples, grad rel diff vs first {float((g - ref).norm() / ref.norm()):.1e}", flush=True)
