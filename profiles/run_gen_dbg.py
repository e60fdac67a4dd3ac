"""In-kernel phase clocks of gen_mlp_kernel (b2048.debug_set("tc_clocks")): one policy step, one split forward, one update pass.
usage: run_gen_dbg.py [n_samples]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import b2048
from b2048 import _lib
from helpers import random_boards
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 8
lib = _lib.load()
env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(obs_mode="onehot"))
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 128, 64], activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig(model_seed=0))
rng = np.random.default_rng(0)
boards = torch.from_numpy(random_boards(rng, n).view(np.int64)).cuda()
flags = torch.full((n,), 15, dtype=torch.uint8, device="cuda")
acts = torch.from_numpy(rng.integers(0, 4, n).astype(np.uint8)).cuda()
coef = torch.from_numpy((rng.normal(size=n) * 1e-6).astype(np.float32)).cuda()
net = agent._actor
ptr = lambda t: C.c_void_p(t.data_ptr())
stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
act_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
ws_floats = int(lib.b2048_backward_workspace_floats(C.byref(net.desc), n))
ws = torch.zeros(ws_floats, dtype=torch.float32, device="cuda")
def run():
    agent.policy_step(boards, flags, act_out, seed=1, gid0=0, t=1, precision=1)
    _lib.check(lib.b2048_mlp_forward(agent._h, ptr(boards), C.byref(net.desc), ptr(out), n, 3, stream()), "fwd")
    net.grad.zero_()
    _lib.check(lib.b2048_mlp_backward(agent._h, ptr(boards), ptr(flags), ptr(acts), ptr(coef), C.byref(net.desc), ptr(net.grad), n, 0,
                                      ptr(ws), ws_floats, n, 3, stream()), "bwd")
    torch.cuda.synchronize()
run()
b2048.debug_set("tc_clocks", True)
print("==== policy step (single), split forward, update pass", file=sys.stderr)
run()
