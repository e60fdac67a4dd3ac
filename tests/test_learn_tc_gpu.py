"""Single-bf16 tcgen05 gradient path (b2048_mlp_backward precision=1, an explicit OPT-IN since round 2: fb_tc_kernel +
atb_tc_kernel) against the NumPy restatement of the same bf16 roundings (tight: the kernels are exact) and against the
float32 arithmetic of the reference (loose: a bf16 forward flips ReLU units near zero, a 3-30 % error on cancelling
gradients — which is why the DEFAULT tensor-core mode is the split-fp16 path of tests/test_learn_hp_gpu.py, held to 1e-2).

The first test also un-swizzles the intermediate bf16 images the two kernels exchange (H1, H2, DL2, DL1, A1^T,
d3^T) so that a layout bug is localised to one stage."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import learner  # noqa: E402
from helpers import random_boards, rel_err  # noqa: E402


@pytest.fixture(scope="module")
def b2048():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import b2048 as m
    return m


def dev64(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def bf16_bytes_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def unswizzle_act(raw, n):
    """activation image bytes -> float32 [n, 256].  Layout: [tile64][slab 4][row 64][128 B], chunk ^= row & 7."""
    t = raw.reshape(-1, 4, 64, 8, 8)                     # tile, slab, row, physical chunk, element (uint16)
    rows = np.arange(64)
    out = np.empty_like(t)
    for c in range(8):                                   # logical chunk c of row r sits at physical chunk c ^ (r & 7)
        out[:, :, rows, c, :] = t[:, :, rows, c ^ (rows & 7), :]
    x = out.transpose(0, 2, 1, 3, 4).reshape(-1, 256)    # tile, row, slab, chunk, elem -> [sample, feature]
    return bf16_bytes_to_f32(x[:n])


def unswizzle_small(raw, n):
    """small K-major image bytes -> float32 [n, 16].  Layout: [tile64][row j 16][128 B = 64 samples], chunk ^= j & 7."""
    t = raw.reshape(-1, 16, 8, 8)                        # tile, j, physical chunk, element
    js = np.arange(16)
    out = np.empty_like(t)
    for c in range(8):
        out[:, js, c, :] = t[:, js, c ^ (js & 7), :]
    x = out.reshape(-1, 16, 64).transpose(0, 2, 1).reshape(-1, 16)
    return bf16_bytes_to_f32(x[:n])


def make_agent(b2048, n_out_critic=False, obs_mode="log2", scale=0.0625, seed=0, use_critic=False):
    env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(obs_mode=obs_mode, obs_log2_scale=scale))
    agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig(use_critic=use_critic))
    rng = np.random.default_rng(seed)
    p = b2048.init_model_params(16, [256, 256], 4, rng, "HeNormal")
    p["b"] = [rng.normal(size=b.shape).astype(np.float32) * 0.1 for b in p["b"]]
    agent.params = p
    return agent


def call_backward(b2048, agent, net, boards, masks, actions, coef, head_mode, precision, chunk):
    from b2048 import _lib
    lib = _lib.load()
    n = len(boards)
    bd = dev64(boards)
    fl = torch.from_numpy(masks).cuda() if masks is not None else None
    ac = torch.from_numpy(actions).cuda() if actions is not None else None
    cf = torch.from_numpy(coef).cuda()
    ws_floats = int(lib.b2048_backward_workspace_floats(C.byref(net.desc), min(chunk, n)))
    ws = torch.zeros(ws_floats, dtype=torch.float32, device="cuda")
    net.grad.zero_()
    ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    st = lib.b2048_mlp_backward(agent._h, ptr(bd), ptr(fl), ptr(ac), ptr(cf), C.byref(net.desc), ptr(net.grad), n, head_mode,
                                ptr(ws), ws_floats, min(chunk, n), precision, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(st, "b2048_mlp_backward")
    torch.cuda.synchronize()
    return net.grad.cpu().numpy().copy(), ws


def split_grads(flat, n_out):
    o, out = 0, []
    for (i, j) in ((16, 256), (256, 256), (256, n_out)):
        W = flat[o:o + i * j].reshape(i, j); o += i * j
        b = flat[o:o + j]; o += j
        out.append((W, b))
    return out


def oracle_grads(params, boards, masks, actions, coef, head_mode, obs_mode, scale):
    X = learner.encode(boards, obs_mode, scale)
    out, acts, pres = learner.forward(params, X, "ReLU")
    if head_mode == 0:
        p = learner.probs_from_logits(out, masks)
        onehot = np.eye(4, dtype=np.float32)[actions]
        d = coef[:, None] * (onehot - p)
    else:
        d = coef[:, None].astype(np.float32)
    gW, gb = learner.backprop(params, acts, pres, d, "ReLU")
    return gW, gb, acts, pres, d


def make_case(rng, n, zero_mean=True, scale=1e-3):
    boards = random_boards(rng, n)
    masks, _ = oracle.mask_done(boards)
    masks = masks.astype(np.uint8)
    masks[masks == 0] = 0xF
    actions = rng.integers(0, 4, n).astype(np.uint8)
    for _ in range(4):                                   # push every action onto a legal one
        bad = ((masks >> actions) & 1) == 0
        actions[bad] = (actions[bad] + 1) % 4
    coef = (rng.normal(size=n) * scale).astype(np.float32)
    if not zero_mean:
        coef = np.abs(coef)
    coef[rng.random(n) < 0.1] = 0.0                      # padded / dead slots carry coef 0
    return boards, masks, actions, coef


def test_tc_backward_stages_and_grads(b2048):
    """Every intermediate image and every gradient tensor against the bf16-rounding restatement (tight), then
    against the float32 restatement of the reference (loose: ReLU units within bf16 rounding of zero flip)."""
    from b2048 import _lib
    n = 128 * 70 + 37                                   # ragged last tile
    rng = np.random.default_rng(11)
    boards, masks, actions, coef = make_case(rng, n)
    agent = make_agent(b2048)
    net = agent._actor
    grads, ws = call_backward(b2048, agent, net, boards, masks, actions, coef, 0, 1, n)
    params = agent.params
    X = learner.encode(boards, "log2", 0.0625)
    gW, gb, st = learner.backprop_bf16(params, X, masks, actions, coef, 0)

    lay = (C.c_int64 * 7)()
    _lib.check(_lib.load().b2048_backward_tc_layout(n, lay), "layout")
    raw = ws.view(torch.uint8)
    base = (-raw.data_ptr()) % 1024
    np_pad = (n + 127) // 128 * 128

    def img(off, nbytes):
        return raw[base + off: base + off + nbytes].cpu().numpy().view(np.uint16)

    act_bytes, small_bytes = np_pad // 64 * 32768, np_pad // 64 * 2048
    errs = {}
    errs["A1^T"] = rel_err(unswizzle_small(img(lay[4], small_bytes), n), st["A1"])
    errs["H1"] = rel_err(unswizzle_act(img(lay[0], act_bytes), n), st["H1"])
    errs["H2"] = rel_err(unswizzle_act(img(lay[1], act_bytes), n), st["H2"])
    d3t = unswizzle_small(img(lay[5], small_bytes), n)
    errs["d3^T"] = rel_err(d3t[:, :4], st["d3"])
    errs["DL2"] = rel_err(unswizzle_act(img(lay[2], act_bytes), n), st["DL2"])
    errs["DL1"] = rel_err(unswizzle_act(img(lay[3], act_bytes), n), st["DL1"])
    got = split_grads(grads, 4)
    for l, (W, b) in enumerate(got):
        errs[f"dW{l}"] = rel_err(W, gW[l])
        errs[f"db{l}"] = rel_err(b, gb[l])
    print("tc backward stage errors vs bf16 restatement:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert (d3t[:, 4:] == 0).all()
    tol = {"A1^T": 1e-6, "H1": 2e-3, "H2": 3e-3, "d3^T": 1e-2, "DL2": 1e-2, "DL1": 1e-2}
    for k, v in errs.items():
        assert v < tol.get(k, 1e-2), (k, v, errs)
    # float32 restatement of the reference (zero-mean coefficients: the gradient is a noise-like sum, so the ReLU
    # sign flips of bf16 show up at the few-percent level)
    gW32, gb32, _, _, _ = oracle_grads(params, boards, masks, actions, coef, 0, "log2", 0.0625)
    for l, (W, b) in enumerate(got):
        assert rel_err(W, gW32[l]) < 6e-2, (f"dW{l} vs fp32", rel_err(W, gW32[l]))
        assert rel_err(b, gb32[l]) < 6e-2, (f"db{l} vs fp32", rel_err(b, gb32[l]))


@pytest.mark.parametrize("head_mode,n,chunk,zero_mean", [(0, 50000, 16384, False), (1, 20000, 1 << 20, False),
                                                         (0, 4096, 4096, False), (0, 30000, 1 << 20, True)])
def test_tc_backward_vs_fp32_path(b2048, head_mode, n, chunk, zero_mean):
    """Tensor-core path vs the fp32 CUDA-core path through the same entry point.  1e-2 (north_star's bf16 bar) on a
    coherent gradient; 6e-2 on a zero-mean one (see above)."""
    rng = np.random.default_rng(5 + head_mode)
    boards, masks, actions, coef = make_case(rng, n, zero_mean=zero_mean, scale=1e-4)
    tol = 6e-2 if zero_mean else 1e-2
    agent = make_agent(b2048, use_critic=(head_mode == 1), seed=3)
    if head_mode == 1:
        p = agent.critic_params
        p["b"] = [rng.normal(size=b.shape).astype(np.float32) * 0.1 for b in p["b"]]
        agent.critic_params = p
        net, n_out = agent._critic, 1
    else:
        net, n_out = agent._actor, 4
    args = (boards, masks if head_mode == 0 else None, actions if head_mode == 0 else None, coef, head_mode)
    g_tc, _ = call_backward(b2048, agent, net, *args, 1, chunk)
    g_32, _ = call_backward(b2048, agent, net, *args, 0, chunk)
    a, b = split_grads(g_tc, n_out), split_grads(g_32, n_out)
    errs = {}
    for l in range(3):
        errs[f"dW{l}"] = rel_err(a[l][0], b[l][0])
        errs[f"db{l}"] = rel_err(a[l][1], b[l][1])
    print("tc vs fp32 gradient errors:", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < tol, (k, v, errs)


def test_tc_backward_refuses_other_shapes(b2048):
    from b2048 import _lib
    env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(obs_mode="log2"))
    agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[64, 32], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig())
    n = 5000
    rng = np.random.default_rng(0)
    boards = random_boards(rng, n)
    with pytest.raises(_lib.B2048Error):
        call_backward(b2048, agent, agent._actor, boards, None, np.zeros(n, np.uint8), np.ones(n, np.float32), 0, 1, n)
    # precision 2 (auto) falls back to the fp32 kernels for the same call
    g, _ = call_backward(b2048, agent, agent._actor, boards, None, np.zeros(n, np.uint8), np.ones(n, np.float32), 0, 2, n)
    assert np.isfinite(g).all() and np.abs(g).sum() > 0


def test_update_from_rollout_tc_matches_fp32(b2048):
    """Whole update (returns scan -> advantages -> gradients -> clip -> SGD) with the tensor-core gradient path vs
    fp32: same gradient norm to 1e-2, parameter step within the zero-mean-gradient bound."""
    from helpers import full_env_kwargs
    n, seed = 8192, 21
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 48
    outs = []
    for prec in (0, 1):
        benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=seed, gid0=0)
        agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                     b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode="batch", learning_rate=1e-2))
        agent.params = b2048.init_model_params(16, [256, 256], 4, np.random.default_rng(1), "HeNormal")
        before = agent._actor.theta.cpu().numpy().copy()
        ro = agent.rollout_many(benv, precision=0)
        info = agent.update_from_rollout(ro, precision=prec)
        outs.append((agent._actor.theta.cpu().numpy() - before, info["actor_grad_norm"]))
    (d0, g0), (d1, g1) = outs
    print("update tc vs fp32: grad norms", g0, g1, "step rel err", rel_err(d1, d0))
    assert abs(g0 - g1) / g0 < 2e-2, (g0, g1)
    assert rel_err(d1, d0) < 6e-2, rel_err(d1, d0)


def test_tc_value_forward_vs_fp32(b2048):
    """b2048_mlp_forward precision=1 (critic V(s) on tcgen05) vs the fp32 kernel: 1e-2 relative."""
    from b2048 import _lib
    lib = _lib.load()
    n = 20000 + 11
    rng = np.random.default_rng(9)
    boards = random_boards(rng, n)
    agent = make_agent(b2048, use_critic=True, seed=4)
    p = agent.critic_params
    p["b"] = [rng.normal(size=b.shape).astype(np.float32) * 0.1 for b in p["b"]]
    agent.critic_params = p
    bd = dev64(boards)
    outs = []
    for prec in (0, 1):
        out = torch.zeros(n, dtype=torch.float32, device="cuda")
        agent._values(bd, out, prec)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    ref, _, _ = learner.forward(agent.critic_params, learner.encode(boards, "log2", 0.0625), "ReLU")
    assert rel_err(outs[0], ref[:, 0]) < 1e-3
    assert rel_err(outs[1], ref[:, 0]) < 1e-2, rel_err(outs[1], ref[:, 0])


def test_actor_critic_update_tc_matches_fp32(b2048):
    """Actor-critic update (critic forward -> TD errors -> critic grads -> baseline-processed TD advantages -> actor
    grads -> Adam on both networks) with every GEMM on tensor cores vs the fp32 kernels."""
    from helpers import full_env_kwargs
    n, seed = 8192, 33
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 40
    outs = []
    for prec in (0, 1):
        benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=seed, gid0=0)
        agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                     b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode="batch_norm", learning_rate=1e-3,
                                                                critic_learning_rate=5e-4, use_critic=True, optimizer="adam",
                                                                model_seed=5))
        a0 = agent._actor.theta.cpu().numpy().copy(); c0 = agent._critic.theta.cpu().numpy().copy()
        ro = agent.rollout_many(benv, precision=0)
        info = agent.update_from_rollout(ro, precision=prec)
        outs.append((agent._actor.theta.cpu().numpy() - a0, agent._critic.theta.cpu().numpy() - c0,
                     info["actor_grad_norm"], info["critic_grad_norm"], info["td"].cpu().numpy().copy()))
    (da0, dc0, ga0, gc0, td0), (da1, dc1, ga1, gc1, td1) = outs
    print("actor-critic tc vs fp32: grad norms", (ga0, ga1), (gc0, gc1), "td rel err", rel_err(td1, td0))
    assert rel_err(td1, td0) < 1e-2
    assert abs(ga0 - ga1) / ga0 < 3e-2 and abs(gc0 - gc1) / gc0 < 3e-2
    # Adam normalises every coordinate to ~lr, so the parameter step is compared by direction
    cos_a = float(np.dot(da0, da1) / (np.linalg.norm(da0) * np.linalg.norm(da1)))
    cos_c = float(np.dot(dc0, dc1) / (np.linalg.norm(dc0) * np.linalg.norm(dc1)))
    assert cos_a > 0.9 and cos_c > 0.9, (cos_a, cos_c)


def test_tc_gradient_on_rollout_matches_bf16_oracle(b2048):
    """On a ROLLOUT-derived policy gradient (advantage-weighted, heavy cancellation) the tensor-core kernels agree with the
    bf16-rounding restatement to 2e-3 while bf16 and float32 themselves differ by several percent: that gap is the
    arithmetic (ReLU units within bf16 rounding of zero), not the kernels."""
    from b2048 import _lib
    from helpers import full_env_kwargs
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 24
    B = 8192
    env = b2048.Batched2048Env(B, b2048.Game2048EnvConfig(**kw), seed=123, gid0=5)
    agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig(model_seed=11, baseline_mode="batch"))
    ro = agent.rollout_many(env, precision=1)
    T = ro.T
    agent.update_from_rollout(ro, precision=0)                      # fills the per-slot coefficient buffer
    coef = agent._scratch["coef"][: T * B].clone()
    live = (torch.arange(T, device="cuda").unsqueeze(1) < ro.length.unsqueeze(0)).reshape(-1)
    boards = ro.boards[:T].reshape(-1)[live].cpu().numpy().view(np.uint64)
    flags = ro.flags[:T].reshape(-1)[live].cpu().numpy()
    acts = ro.actions[:T].reshape(-1)[live].cpu().numpy()
    cf = coef[live].cpu().numpy()
    n = len(boards)
    assert n >= 4096
    g_tc, _ = call_backward(b2048, agent, agent._actor, boards, flags, acts, cf, 0, 1, n)
    g_32, _ = call_backward(b2048, agent, agent._actor, boards, flags, acts, cf, 0, 0, n)
    X = learner.encode(boards, "log2", 0.0625)
    gW, gb, _ = learner.backprop_bf16(agent.params, X, flags & 0xF, acts.astype(np.int64), cf, 0)
    g_or = np.concatenate([np.concatenate([w.reshape(-1), b.reshape(-1)]) for w, b in zip(gW, gb)])
    e_or, e_32 = rel_err(g_tc, g_or), rel_err(g_tc, g_32)
    print(f"rollout gradient, {n} samples: tc vs bf16 oracle {e_or:.2e}, tc vs fp32 {e_32:.2e}")
    assert e_or < 2e-3, e_or
    assert e_32 < 0.15, e_32
