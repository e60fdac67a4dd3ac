"""Float32-grade tensor-core gradient path (b2048_mlp_backward precision 3, what "auto" selects: fwd_hp_kernel with
split-fp16 operands + bwd_tc_kernel + atb_tc_kernel on fp16 images) against the float32 restatement of the reference's
forward_logits / _backpropagation (oracle/learner.py; src/MLP.py:159-196, src/reinforce_agent.py:502-555, :639-678),
against float64, and against the fp32 CUDA-core kernels.  Bar: 1e-2 relative per gradient tensor (north_star's
tensor-core tolerance) on coherent AND zero-mean / rollout-derived (heavily cancelling) gradients; the forward pass itself
is held to 1e-5."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import learner  # noqa: E402
from helpers import full_env_kwargs, random_boards, rel_err  # noqa: E402
from test_learn_tc_gpu import call_backward, dev64, make_agent, make_case, oracle_grads, split_grads  # noqa: E402

HP = 3


@pytest.fixture(scope="module")
def b2048():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import b2048 as m
    return m


def forward64(params, X):
    a = X.astype(np.float64)
    L = len(params["W"])
    pres = []
    for i in range(L):
        z = a @ params["W"][i].astype(np.float64) + params["b"][i].astype(np.float64)
        pres.append(z)
        a = np.maximum(z, 0.0) if i < L - 1 else z
    return a, pres


@pytest.mark.parametrize("n,head", [(128 * 70 + 37, "actor"), (4096, "critic"), (200000 + 5, "actor")])
def test_hp_forward_vs_float64(b2048, n, head):
    """b2048_mlp_forward precision 3: head outputs within 1e-5 of the float64 forward (split-fp16 operands carry 22 bits;
    single bf16 is at 4e-3, single fp16 at 5e-4)."""
    import ctypes as C
    from b2048 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(17)
    boards = random_boards(rng, n)
    agent = make_agent(b2048, use_critic=(head == "critic"), seed=6)
    if head == "critic":
        p = agent.critic_params
        p["b"] = [rng.normal(size=b.shape).astype(np.float32) * 0.1 for b in p["b"]]
        agent.critic_params = p
        net, params = agent._critic, agent.critic_params
    else:
        net, params = agent._actor, agent.params
    n_out = net.dims[-1]
    out = torch.zeros((n, n_out), dtype=torch.float32, device="cuda")
    bd = dev64(boards)
    _lib.check(lib.b2048_mlp_forward(agent._h, C.c_void_p(bd.data_ptr()), C.byref(net.desc), C.c_void_p(out.data_ptr()), n, HP,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "b2048_mlp_forward")
    torch.cuda.synchronize()
    ref, _ = forward64(params, learner.encode(boards, "log2", 0.0625))
    err = rel_err(out.cpu().numpy(), ref)
    print(f"split-fp16 forward vs float64 ({head}, n = {n}): {err:.2e}")
    assert err < 1e-5, err


@pytest.mark.parametrize("head_mode,n,chunk,zero_mean", [(0, 50000, 16384, False), (1, 20000, 1 << 20, False),
                                                         (0, 4096, 4096, False), (0, 30000, 1 << 20, True),
                                                         (0, 128 * 70 + 37, 1 << 20, True), (1, 33000, 8192, True)])
def test_hp_backward_vs_fp32(b2048, head_mode, n, chunk, zero_mean):
    """Every gradient tensor within 1e-2 of the float32 restatement of the reference AND of the fp32 kernels — on
    coherent and on zero-mean (noise-like, heavily cancelling) coefficient sets, one chunk and several, ragged tiles."""
    rng = np.random.default_rng(5 + head_mode)
    boards, masks, actions, coef = make_case(rng, n, zero_mean=zero_mean, scale=1e-4)
    agent = make_agent(b2048, use_critic=(head_mode == 1), seed=3)
    if head_mode == 1:
        p = agent.critic_params
        p["b"] = [rng.normal(size=b.shape).astype(np.float32) * 0.1 for b in p["b"]]
        agent.critic_params = p
        net, n_out, params = agent._critic, 1, agent.critic_params
    else:
        net, n_out, params = agent._actor, 4, agent.params
    args = (boards, masks if head_mode == 0 else None, actions if head_mode == 0 else None, coef, head_mode)
    g_hp, _ = call_backward(b2048, agent, net, *args, HP, chunk)
    g_32, _ = call_backward(b2048, agent, net, *args, 0, chunk)
    gW, gb, _, _, _ = oracle_grads(params, boards, masks if head_mode == 0 else None, actions, coef, head_mode, "log2", 0.0625)
    a, b = split_grads(g_hp, n_out), split_grads(g_32, n_out)
    errs = {}
    for l in range(3):
        errs[f"dW{l} vs fp32 kernels"] = rel_err(a[l][0], b[l][0])
        errs[f"db{l} vs fp32 kernels"] = rel_err(a[l][1], b[l][1])
        errs[f"dW{l} vs oracle"] = rel_err(a[l][0], gW[l])
        errs[f"db{l} vs oracle"] = rel_err(a[l][1], gb[l])
    print("split-fp16 gradient errors:", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < 1e-2, (k, v, errs)


def test_hp_gradient_on_rollout_within_1e2_of_fp32(b2048):
    """A ROLLOUT-derived policy gradient (advantage-weighted, zero-mean: the case single-bf16 arithmetic misses by
    5-30 %): the default tensor-core mode is within 1e-2 of the fp32 kernels and of the float64 gradient."""
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
    g_hp, _ = call_backward(b2048, agent, agent._actor, boards, flags, acts, cf, 0, HP, n)
    g_bf, _ = call_backward(b2048, agent, agent._actor, boards, flags, acts, cf, 0, 1, n)
    g_32, _ = call_backward(b2048, agent, agent._actor, boards, flags, acts, cf, 0, 0, n)
    # float64 gradient of the same batch
    params = agent.params
    X = learner.encode(boards, "log2", 0.0625)
    out, pres = forward64(params, X)
    p = learner.probs_from_logits(out.astype(np.float32), flags & 0xF).astype(np.float64)
    d = cf[:, None].astype(np.float64) * (np.eye(4)[acts.astype(np.int64)] - p)
    h1, h2 = np.maximum(pres[0], 0), np.maximum(pres[1], 0)
    dl2 = (d @ params["W"][2].astype(np.float64).T) * (pres[1] > 0)
    dl1 = (dl2 @ params["W"][1].astype(np.float64).T) * (pres[0] > 0)
    g_64 = np.concatenate([np.concatenate([w.reshape(-1), b.reshape(-1)]) for w, b in
                           ((X.astype(np.float64).T @ dl1, dl1.sum(0)), (h1.T @ dl2, dl2.sum(0)), (h2.T @ d, d.sum(0)))])
    e_32, e_64, e_bf, e_3264 = rel_err(g_hp, g_32), rel_err(g_hp, g_64), rel_err(g_bf, g_32), rel_err(g_32, g_64)
    print(f"rollout gradient, {n} samples: split-fp16 vs fp32 kernels {e_32:.2e}, vs float64 {e_64:.2e}; "
          f"single bf16 vs fp32 {e_bf:.2e}; fp32 kernels vs float64 {e_3264:.2e}")
    assert e_32 < 1e-2 and e_64 < 1e-2, (e_32, e_64)


def test_update_from_rollout_auto_matches_fp32(b2048):
    """Whole update (returns scan -> advantages -> gradients -> clip -> SGD) with precision "auto" (tensor cores) vs the
    fp32 kernels: gradient norm and parameter step within 1e-2."""
    n, seed = 8192, 21
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 48
    outs = []
    for prec in (0, "auto"):
        benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=seed, gid0=0)
        agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                     b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode="batch", learning_rate=1e-2))
        agent.params = b2048.init_model_params(16, [256, 256], 4, np.random.default_rng(1), "HeNormal")
        before = agent._actor.theta.cpu().numpy().copy()
        ro = agent.rollout_many(benv, precision=0)
        info = agent.update_from_rollout(ro, precision=prec)
        outs.append((agent._actor.theta.cpu().numpy() - before, info["actor_grad_norm"], info["precision"]))
    (d0, g0, m0), (d1, g1, m1) = outs
    print(f"update {m1} vs {m0}: grad norms {g0} {g1}, step rel err {rel_err(d1, d0):.2e}")
    assert m1.startswith("fp16 split tcgen05") and m0.startswith("fp32")
    assert abs(g0 - g1) / g0 < 1e-2, (g0, g1)
    assert rel_err(d1, d0) < 1e-2, rel_err(d1, d0)


def test_actor_critic_update_auto_matches_fp32(b2048):
    """Actor-critic update (critic forward -> TD errors -> critic grads -> baseline-processed TD advantages -> actor grads
    -> Adam on both networks) with every GEMM on the tensor cores in the default mode vs the fp32 kernels: TD errors,
    both gradient norms within 1e-2 (the value forward is float32-grade too, so the TD differences do not amplify it)."""
    n, seed = 8192, 33
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 40
    outs = []
    for prec in (0, "auto"):
        benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=seed, gid0=0)
        agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                     b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode="batch_norm", learning_rate=1e-3,
                                                                critic_learning_rate=5e-4, use_critic=True, optimizer="sgd",
                                                                model_seed=5))
        a0 = agent._actor.theta.cpu().numpy().copy(); c0 = agent._critic.theta.cpu().numpy().copy()
        ro = agent.rollout_many(benv, precision=0)
        info = agent.update_from_rollout(ro, precision=prec)
        outs.append((agent._actor.theta.cpu().numpy() - a0, agent._critic.theta.cpu().numpy() - c0,
                     info["actor_grad_norm"], info["critic_grad_norm"], info["td"].cpu().numpy().copy()))
    (da0, dc0, ga0, gc0, td0), (da1, dc1, ga1, gc1, td1) = outs
    print("actor-critic auto vs fp32: grad norms", (ga0, ga1), (gc0, gc1), "td rel err", rel_err(td1, td0),
          "steps", rel_err(da1, da0), rel_err(dc1, dc0))
    assert rel_err(td1, td0) < 1e-4
    assert abs(ga0 - ga1) / ga0 < 1e-2 and abs(gc0 - gc1) / gc0 < 1e-2
    assert rel_err(da1, da0) < 1e-2 and rel_err(dc1, dc0) < 1e-2


def test_hp_refuses_raw_observations_and_auto_falls_back(b2048):
    """fp16 activations could overflow on raw tile values: precision 3 is refused loudly, "auto" runs the fp32 kernels."""
    from b2048 import _lib
    agent = make_agent(b2048, obs_mode="raw", scale=1.0)
    n = 5000
    rng = np.random.default_rng(0)
    boards, masks, actions, coef = make_case(rng, n)
    with pytest.raises(_lib.B2048Error):
        call_backward(b2048, agent, agent._actor, boards, masks, actions, coef, 0, HP, n)
    g2, _ = call_backward(b2048, agent, agent._actor, boards, masks, actions, coef, 0, 2, n)
    g0, _ = call_backward(b2048, agent, agent._actor, boards, masks, actions, coef, 0, 0, n)
    assert rel_err(g2, g0) < 1e-5          # same fp32 kernels (atomic summation order differs run to run)


@pytest.mark.parametrize("head_mode,n", [(0, 128 * 700 + 37), (1, 128 * 600), (0, 300000)])
def test_hp_update_pipeline_matches_chunked_kernels_and_oracle(b2048, head_mode, n):
    """Large batches run as ONE persistent launch (update_pipe_kernel: forward / backward / dW roles exchanging tiles through
    an L2-resident ring of 288 slots with acquire / release flags; more tiles than slots, so slots are recycled).  Same
    gradient as the chunked kernel sequence (fp32 atomic order only) and within 1e-2 of the float32 restatement."""
    rng = np.random.default_rng(50 + head_mode)
    boards, masks, actions, coef = make_case(rng, n, zero_mean=True, scale=1e-5)
    agent = make_agent(b2048, use_critic=(head_mode == 1), seed=8)
    if head_mode == 1:
        p = agent.critic_params
        p["b"] = [rng.normal(size=b.shape).astype(np.float32) * 0.1 for b in p["b"]]
        agent.critic_params = p
        net, n_out, params = agent._critic, 1, agent.critic_params
    else:
        net, n_out, params = agent._actor, 4, agent.params
    args = (boards, masks if head_mode == 0 else None, actions if head_mode == 0 else None, coef, head_mode)
    g_pipe, _ = call_backward(b2048, agent, net, *args, HP, 1 << 20)
    g_pipe2, _ = call_backward(b2048, agent, net, *args, HP, 1 << 20)          # run-to-run: atomic summation order only
    b2048.debug_set("no_update_pipe", True)
    try:
        g_chunk, _ = call_backward(b2048, agent, net, *args, HP, 65536)
    finally:
        b2048.debug_set("no_update_pipe", False)
    gW, gb, _, _, _ = oracle_grads(params, boards, masks if head_mode == 0 else None, actions, coef, head_mode, "log2", 0.0625)
    g_or = np.concatenate([np.concatenate([w.reshape(-1), b.reshape(-1)]) for w, b in zip(gW, gb)])
    e_rr, e_ck, e_or = rel_err(g_pipe2, g_pipe), rel_err(g_pipe, g_chunk), rel_err(g_pipe, g_or)
    print(f"update pipeline, {n} samples: run-to-run {e_rr:.1e}, vs chunked kernels {e_ck:.1e}, vs float32 oracle {e_or:.1e}")
    assert e_rr < 1e-5 and e_ck < 1e-4, (e_rr, e_ck)
    a, b = split_grads(g_pipe, n_out), split_grads(g_or, n_out)
    for l in range(3):
        assert rel_err(a[l][0], b[l][0]) < 1e-2 and rel_err(a[l][1], b[l][1]) < 1e-2, (l, rel_err(a[l][0], b[l][0]), rel_err(a[l][1], b[l][1]))


def test_auto_update_never_applies_a_non_finite_fp16_gradient(b2048):
    """A network whose hidden activations exceed the fp16 range (first-layer weights scaled by 1e6) makes the split-fp16
    path produce non-finite values; the default update detects that and redoes the pass on the fp32 kernels."""
    n = 8192
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 12
    benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=9, gid0=0)
    agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode="batch", learning_rate=1e-3))
    ro = agent.rollout_many(benv, precision=0)
    p = agent.params
    p["W"][0] = (p["W"][0] * 1e6).astype(np.float32)
    p["W"][1] = (p["W"][1] * 1e-6).astype(np.float32)
    agent.params = p
    th0 = agent._actor.theta.clone()
    info = agent.update_from_rollout(ro, precision="auto")
    assert "fell back to fp32" in info["precision"], info["precision"]
    d_auto = (agent._actor.theta - th0).cpu().numpy()
    agent._actor.theta.copy_(th0)
    info0 = agent.update_from_rollout(ro, precision=0)
    d_32 = (agent._actor.theta - th0).cpu().numpy()
    assert np.isfinite(d_auto).all() and np.isfinite(info["actor_grad_norm"])
    assert rel_err(d_auto, d_32) < 1e-4


def test_full_size_actor_critic_update_auto_vs_fp32(b2048):
    """BASELINE.json configs[3] size (262,144 boards, actor + separate critic, TD(0), Adam): the default tensor-core update
    (two pipeline launches + the float32-grade value forward) against the fp32 kernels on the same rollout — TD errors to 1e-4,
    both unclipped gradients to 1e-2."""
    n = 262144
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 40
    benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=77, gid0=0)
    agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode="batch_norm", learning_rate=1e-4,
                                                            critic_learning_rate=5e-4, use_critic=True, optimizer="adam",
                                                            model_seed=2, max_grad_norm=1e30))
    ro = agent.rollout_many(benv, precision="auto")
    st = agent.save_state()
    outs = []
    for prec in ("auto", 0):
        agent.load_state(st)
        info = agent.update_from_rollout(ro, precision=prec)
        outs.append((agent._actor.grad.clone(), agent._critic.grad.clone(), info["td"].clone(), info["precision"]))
    (ga1, gc1, td1, m1), (ga0, gc0, td0, m0) = outs
    ea = float((ga1 - ga0).norm() / ga0.norm()); ec = float((gc1 - gc0).norm() / gc0.norm())
    et = float((td1 - td0).norm() / td0.norm())
    print(f"configs[3] size, {int(ro.length.sum())} samples, {m1}: actor grad {ea:.2e}, critic grad {ec:.2e}, td {et:.2e} vs fp32 kernels")
    assert m1.startswith("fp16 split tcgen05") and "fell back" not in m1
    assert et < 1e-4 and ea < 1e-2 and ec < 1e-2, (et, ea, ec)


def test_sharded_sweep_leg_runs_on_one_gpu(b2048):
    """The BASELINE.json configs[4] bench leg (16-step episodes via max_steps = 16, one update) at a small size."""
    r = b2048.bench_sharded_sweep(torch.device("cuda", 0), total_boards=1 << 16, horizon=16, iters=1)
    assert r["horizon"] == 16 and r["samples"] > 0.9 * 16 * (1 << 16) and np.isfinite(r["actor_grad_norm"])
    assert r["update_precision"].startswith("fp16 split tcgen05")


class _Recorded(Exception):
    pass


@pytest.mark.parametrize("prec,use_critic,baseline", [(0, False, "batch"), (0, True, "batch_norm"), ("auto", False, "batch_norm"),
                                                      ("auto", True, "batch")])
def test_one_message_exchange_equals_single_process_update(b2048, prec, use_critic, baseline):
    """exchange="one_message" (north_star: one all-reduce for the policy gradient): two emulated ranks, each
    holding half of the episodes, exchange exactly ONE buffer [g_A | g_B | critic gradient | baseline sums] per update, and the
    applied update equals the one a single process holding every episode applies with the default exchange
    (reinforce_agent.py:303-322, :502-555: g = (g_A - mean g_B) / std by linearity of the backward pass)."""
    n, seed = 8192, 77
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 40
    mlp = b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal")
    acfg = b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode=baseline, learning_rate=1e-2, critic_learning_rate=1e-3,
                                      use_critic=use_critic, model_seed=5)
    cfg = b2048.Game2048EnvConfig(**kw)
    env1 = b2048.Batched2048Env(n, cfg, seed=seed, gid0=0)
    single = b2048.ReinforceAgent(env1, mlp, acfg)
    th0 = single._actor.theta.clone()
    c0 = single._critic.theta.clone() if use_critic else None
    u1 = single.update_from_rollout(single.rollout_many(env1, precision=0), precision=prec)
    ranks = []
    for lo, hi in ((0, n // 2 - 100), (n // 2 - 100, n)):
        env = b2048.Batched2048Env(hi - lo, cfg, seed=seed, gid0=lo)
        agent = b2048.ReinforceAgent(env, mlp, acfg)
        ro = agent.rollout_many(env, precision=0)
        ro.n_traj = n
        ranks.append((agent, ro))
    sent = []

    def record(t):
        sent.append(t.clone())
        raise _Recorded()

    for agent, ro in ranks:                      # phase 1: what every rank would send (parameters are not touched yet)
        with pytest.raises(_Recorded):
            agent.update_from_rollout(ro, allreduce=record, precision=prec, exchange="one_message")
    assert len(sent) == 2 and sent[0].shape == sent[1].shape and sent[0].dtype == torch.float64
    total = sent[0] + sent[1]
    for agent, ro in ranks:                      # phase 2: the same update with the summed message delivered
        calls = []

        def deliver(t):
            calls.append(int(t.numel()))
            t.copy_(total)

        upd = agent.update_from_rollout(ro, allreduce=deliver, precision=prec, exchange="one_message")
        assert len(calls) == 1, calls
        d_ref, d_got = single._actor.theta - th0, agent._actor.theta - th0
        rel = float((d_got - d_ref).norm() / d_ref.norm())
        gn = abs(upd["actor_grad_norm"] - u1["actor_grad_norm"]) / u1["actor_grad_norm"]
        # fp32 kernels: only the summation order differs.  Tensor cores: g_A and mean g_B are each formed with the fp16 backward's
        # rounding noise and then cancel (|g_A| ~ 3 |g|), so the bar is north_star's tensor-core tolerance, not the 2e-3 the
        # default exchange reaches
        tol = 1e-4 if prec == 0 else 1e-2
        print(f"one-message exchange ({upd['precision']}, {baseline}, critic={use_critic}): 1 all-reduce of {calls[0]} float64; "
              f"update rel err {rel:.2e}, grad-norm rel err {gn:.2e}")
        assert rel < tol and gn < tol, (rel, gn)
        if use_critic:
            dc_ref, dc_got = single._critic.theta - c0, agent._critic.theta - c0
            relc = float((dc_got - dc_ref).norm() / dc_ref.norm())
            assert relc < tol, relc
    assert torch.equal(ranks[0][0]._actor.theta, ranks[1][0]._actor.theta)
