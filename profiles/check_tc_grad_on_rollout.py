"""Is the several-percent TC-vs-fp32 difference on ROLLOUT-derived gradients inherent to bf16 (then the bf16-rounding
oracle reproduces the TC gradient) or a kernel bug (then it does not)?"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, b2048
from oracle import learner
from b2048 import _lib
torch.cuda.set_device(0)
KW = dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5, max_steps=31)
B = 18944
env = b2048.Batched2048Env(B, b2048.Game2048EnvConfig(**KW), seed=123, gid0=5)
agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                             b2048.ReinforceAgentConfig(model_seed=11, baseline_mode="batch"))
ro = agent.rollout_many(env, precision=1)
T = ro.T
# per-sample inputs exactly as update_from_rollout builds them
info = agent.update_from_rollout(ro, precision=0)      # fills the coef buffer (and steps the params; restore below)
coef = agent._scratch["coef"][: T * B].clone()
tg = torch.arange(T, device="cuda").unsqueeze(1)
live = (tg < ro.length.unsqueeze(0)).reshape(-1)
boards = ro.boards[:T].reshape(-1)[live]; flags = ro.flags[:T].reshape(-1)[live]; acts = ro.actions[:T].reshape(-1)[live]
cf = coef[live]
n = int(boards.numel())
params = agent.params
lib = _lib.load()
def grads(prec):
    net = agent._actor
    net.grad.zero_()
    wsf = int(lib.b2048_backward_workspace_floats(C.byref(net.desc), n))
    ws = torch.zeros(wsf, dtype=torch.float32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.b2048_mlp_backward(agent._h, p(boards), p(flags), p(acts), p(cf), C.byref(net.desc), p(net.grad), n, 0, p(ws), wsf,
                                      n, prec, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "bwd")
    torch.cuda.synchronize()
    return net.grad.cpu().numpy().copy()
g_tc, g_32 = grads(1), grads(0)
X = learner.encode(boards.cpu().numpy().view(np.uint64), "log2", 0.0625)
gW, gb, _ = learner.backprop_bf16(params, X, flags.cpu().numpy() & 0xF, acts.cpu().numpy().astype(np.int64), cf.cpu().numpy(), 0)
g_or = np.concatenate([np.concatenate([w.reshape(-1), b.reshape(-1)]) for w, b in zip(gW, gb)])
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
print("samples", n, "| tc vs bf16-rounding oracle", rel(g_tc, g_or), "| tc vs fp32 kernels", rel(g_tc, g_32),
      "| bf16 oracle vs fp32 kernels", rel(g_or, g_32), "| norms", np.linalg.norm(g_tc), np.linalg.norm(g_32), np.linalg.norm(g_or))
