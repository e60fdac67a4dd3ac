"""Debug driver of the persistent update pipeline: one b2048_mlp_backward precision 3 call with the progress watchdog on."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, b2048
from b2048 import _lib
from helpers import random_boards
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 700
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
wsf = int(lib.b2048_backward_workspace_floats(C.byref(net.desc), n))
ws = torch.zeros(wsf, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
b2048.debug_set("tc_clocks", True)
net.grad.zero_()
_lib.check(lib.b2048_mlp_backward(agent._h, p(boards), p(flags), p(acts), p(coef), C.byref(net.desc), p(net.grad), n, 0, p(ws), wsf, n, 3,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "bwd")
torch.cuda.synchronize()
print("pipeline finished; grad norm", float(net.grad.norm()))
