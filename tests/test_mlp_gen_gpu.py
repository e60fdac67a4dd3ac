"""Shape-generic tensor-core MLP kernels (csrc/b2048_mlp_gen.cu: gen_mlp_kernel / gen_dw_kernel) — the tcgen05 path of every
network other than the runner-default 16-256-256-4, in particular the reference's documented one-hot [256, 128, 64]
configuration (runner.py:27-47).  Checked against the float32 / float64 restatement of the reference's
forward_logits / _backpropagation (oracle/learner.py; src/MLP.py:159-196, src/reinforce_agent.py:502-555, :639-678), against
the fp32 CUDA-core kernels and against the reference's own outputs (tests/golden/mlp.npz, learner.npz).
Bars: split-fp16 forward 1e-5 (float32 grade), single-fp16 policy logits 1e-2, every gradient tensor 1e-2."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import learner  # noqa: E402
from helpers import GOLDEN, full_env_kwargs, random_boards, rel_err  # noqa: E402
from test_learn_tc_gpu import call_backward, dev64, make_case  # noqa: E402

HP = 3
SHAPES = {  # name -> (obs_mode, hidden sizes)
    "onehot_256_128_64": ("onehot", [256, 128, 64]),      # the reference's documented configuration
    "log2_128": ("log2", [128]),                          # one hidden layer
    "log2_64_192_64_128": ("log2", [64, 192, 64, 128]),   # four hidden layers, every slab count
    "onehot_256_256": ("onehot", [256, 256]),
    "sigmoid_log2_128_64": ("log2", [128, 64], "Sigmoid"),          # the dataclass-default activation (MLP.py:13)
    "sigmoid_onehot_256_128": ("onehot", [256, 128], "Sigmoid"),
}


def shape_of(name):
    t = SHAPES[name]
    return t[0], t[1], (t[2] if len(t) > 2 else "ReLU")


@pytest.fixture(scope="module")
def b2048():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import b2048 as m
    return m


def make_gen_agent(b2048, obs_mode, hidden, seed=0, use_critic=False, actv="ReLU", **agent_kw):
    env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(obs_mode=obs_mode, obs_log2_scale=0.0625))
    agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=list(hidden), activation=actv, init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig(use_critic=use_critic, **agent_kw))
    rng = np.random.default_rng(seed)
    kin = 272 if obs_mode == "onehot" else 16
    p = b2048.init_model_params(kin, list(hidden), 4, rng, "HeNormal")
    p["b"] = [rng.normal(size=b.shape).astype(np.float32) * 0.1 for b in p["b"]]
    agent.params = p
    if use_critic:
        pc = b2048.init_model_params(kin, list(hidden), 1, rng, "HeNormal")
        pc["b"] = [rng.normal(size=b.shape).astype(np.float32) * 0.1 for b in pc["b"]]
        agent.critic_params = pc
    return agent


def forward64(params, X, actv="ReLU"):
    a = X.astype(np.float64)
    L = len(params["W"])
    for i in range(L):
        z = a @ params["W"][i].astype(np.float64) + params["b"][i].astype(np.float64)
        a = (np.maximum(z, 0.0) if actv == "ReLU" else 1.0 / (1.0 + np.exp(-z))) if i < L - 1 else z
    return a


def split_flat(flat, dims):
    o, out = 0, []
    for i, j in zip(dims[:-1], dims[1:]):
        W = flat[o:o + i * j].reshape(i, j); o += i * j
        b = flat[o:o + j]; o += j
        out.append((W, b))
    return out


def oracle_grads(params, boards, masks, actions, coef, head_mode, obs_mode, actv="ReLU"):
    X = learner.encode(boards, obs_mode, 0.0625)
    out, acts, pres = learner.forward(params, X, actv)
    if head_mode == 0:
        p = learner.probs_from_logits(out, masks)
        d = coef[:, None] * (np.eye(4, dtype=np.float32)[actions] - p)
    else:
        d = coef[:, None].astype(np.float32)
    return learner.backprop(params, acts, pres, d, actv)


def mlp_forward(agent, net, boards, precision):
    from b2048 import _lib
    lib = _lib.load()
    n = len(boards)
    out = torch.zeros((n, net.dims[-1]), dtype=torch.float32, device="cuda")
    bd = dev64(boards)
    _lib.check(lib.b2048_mlp_forward(agent._h, C.c_void_p(bd.data_ptr()), C.byref(net.desc), C.c_void_p(out.data_ptr()), n, precision,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "b2048_mlp_forward")
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("shape", list(SHAPES))
@pytest.mark.parametrize("n", [4096, 128 * 150 + 37])
def test_gen_forward_split_vs_float64(b2048, shape, n):
    """b2048_mlp_forward precision 3 on the generic kernel: head outputs (actor logits and the critic's V) within 1e-5 of the
    float64 forward; ragged last tile; more tiles than CTAs."""
    obs_mode, hidden, actv = shape_of(shape)
    rng = np.random.default_rng(17)
    boards = random_boards(rng, n)
    agent = make_gen_agent(b2048, obs_mode, hidden, seed=6, use_critic=True, actv=actv)
    X = learner.encode(boards, obs_mode, 0.0625)
    for net, params in ((agent._actor, agent.params), (agent._critic, agent.critic_params)):
        got = mlp_forward(agent, net, boards, HP)
        ref = forward64(params, X, actv)
        err = rel_err(got, ref)
        print(f"{shape} n = {n} n_out = {net.dims[-1]}: split-fp16 forward vs float64 {err:.2e}")
        assert err < 1e-5, (shape, err)
        got32 = mlp_forward(agent, net, boards, 0)
        assert rel_err(got, got32) < 1e-5


@pytest.mark.parametrize("shape", list(SHAPES))
def test_gen_policy_step_vs_fp32(b2048, shape):
    """b2048_policy_step precision 1 on the generic kernel (one fp16 MMA per product): logits / probabilities within 1e-2 of
    the fp32 kernel; sampled actions are legal and agree with the fp32 kernel's wherever the uniform is not within the
    probability difference of a CDF edge; greedy actions agree wherever the top-2 gap exceeds the logit error."""
    obs_mode, hidden, actv = shape_of(shape)
    rng = np.random.default_rng(3)
    n = 128 * 40 + 5
    boards = random_boards(rng, n)
    masks, done = oracle.mask_done(boards)
    masks = np.where(masks == 0, 0xF, masks).astype(np.uint8)
    agent = make_gen_agent(b2048, obs_mode, hidden, seed=2, actv=actv)
    bd, fl = dev64(boards), torch.from_numpy(masks).cuda()
    res = {}
    for prec in (0, 1):
        act = torch.zeros(n, dtype=torch.uint8, device="cuda")
        pr = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        lg = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        agent.policy_step(bd, fl, act, seed=99, gid0=7, t=3, probs_out=pr, logits_out=lg, precision=prec)
        gr = torch.zeros(n, dtype=torch.uint8, device="cuda")
        agent.policy_step(bd, fl, gr, seed=99, gid0=7, t=3, greedy=True, precision=prec)
        torch.cuda.synchronize()
        res[prec] = (act.cpu().numpy(), pr.cpu().numpy(), lg.cpu().numpy(), gr.cpu().numpy())
    a0, p0, l0, g0 = res[0]
    a1, p1, l1, g1 = res[1]
    e_l, e_p = rel_err(l1, l0), float(np.abs(p1 - p0).max())
    print(f"{shape}: fp16 tcgen05 policy vs fp32: logits {e_l:.2e}, max |dp| {e_p:.2e}, sampled agree {np.mean(a0 == a1):.4f}, "
          f"greedy agree {np.mean(g0 == g1):.4f}")
    assert e_l < 1e-2 and e_p < 1e-2
    assert np.all((masks >> a1) & 1), "sampled an illegal action"
    assert np.all((masks >> g1) & 1), "greedy picked an illegal action"
    assert np.mean(a0 == a1) > 0.99 and np.mean(g0 == g1) > 0.99
    # against the reference restatement
    X = learner.encode(boards, obs_mode, 0.0625)
    out, _, _ = learner.forward(agent.params, X, actv)
    assert rel_err(l1, out) < 1e-2


@pytest.mark.parametrize("shape,head_mode,n,chunk,zero_mean", [
    ("onehot_256_128_64", 0, 50000, 16384, False), ("onehot_256_128_64", 1, 20000, 1 << 20, False),
    ("onehot_256_128_64", 0, 128 * 170 + 37, 1 << 20, True), ("onehot_256_128_64", 1, 33000, 8192, True),
    ("log2_128", 0, 30000, 1 << 20, True), ("log2_64_192_64_128", 0, 4096, 4096, False),
    ("log2_64_192_64_128", 1, 128 * 160 + 1, 1 << 20, True), ("onehot_256_256", 0, 25000, 1 << 20, True),
    ("sigmoid_log2_128_64", 0, 30000, 8192, True), ("sigmoid_log2_128_64", 1, 128 * 150 + 3, 1 << 20, False),
    ("sigmoid_onehot_256_128", 0, 20000, 1 << 20, True)])
def test_gen_backward_vs_fp32(b2048, shape, head_mode, n, chunk, zero_mean):
    """b2048_mlp_backward precision 3 on the generic kernels: every gradient tensor within 1e-2 of the float32 restatement of
    the reference AND of the fp32 kernels — coherent and zero-mean (heavily cancelling) coefficients, one chunk and several,
    ragged tiles, more tiles than CTAs, policy and value heads."""
    obs_mode, hidden, actv = shape_of(shape)
    rng = np.random.default_rng(5 + head_mode)
    boards, masks, actions, coef = make_case(rng, n, zero_mean=zero_mean, scale=1e-4)
    agent = make_gen_agent(b2048, obs_mode, hidden, seed=3, use_critic=(head_mode == 1), actv=actv)
    net, params = (agent._critic, agent.critic_params) if head_mode == 1 else (agent._actor, agent.params)
    args = (boards, masks if head_mode == 0 else None, actions if head_mode == 0 else None, coef, head_mode)
    g_hp, _ = call_backward(b2048, agent, net, *args, HP, chunk)
    g_32, _ = call_backward(b2048, agent, net, *args, 0, chunk)
    gW, gb = oracle_grads(params, boards, masks if head_mode == 0 else None, actions, coef, head_mode, obs_mode, actv)
    a, b = split_flat(g_hp, net.dims), split_flat(g_32, net.dims)
    errs = {}
    for l in range(len(net.dims) - 1):
        errs[f"dW{l} vs fp32 kernels"] = rel_err(a[l][0], b[l][0])
        errs[f"db{l} vs fp32 kernels"] = rel_err(a[l][1], b[l][1])
        errs[f"dW{l} vs oracle"] = rel_err(a[l][0], gW[l])
        errs[f"db{l} vs oracle"] = rel_err(a[l][1], gb[l])
    print(f"{shape} head {head_mode} n {n}: generic tcgen05 gradient errors:", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < 1e-2, (k, v, errs)


def test_gen_matches_reference_mlp_fixture(b2048):
    """The reference's own forward_logits / logits_to_probs outputs for its documented one-hot 272-256-128-64-4 ReLU network
    (tests/golden/mlp.npz, tag 'onehot': produced by the imported reference), reproduced by the tensor-core kernels; the
    fixture's 72 boards are tiled to a tensor-core batch."""
    import os
    g = np.load(os.path.join(GOLDEN, "mlp.npz"))
    tag = "onehot"
    L = int(g[f"{tag}/n_layers"])
    params = {"W": [g[f"{tag}/W{i}"] for i in range(L)], "b": [g[f"{tag}/b{i}"] for i in range(L)]}
    hidden = [w.shape[1] for w in params["W"][:-1]]
    assert params["W"][0].shape[0] == 272 and hidden == [256, 128, 64]
    agent = make_gen_agent(b2048, "onehot", hidden, seed=1)
    agent.params = params
    boards, masks = g[f"{tag}/boards"].astype(np.uint64), g[f"{tag}/masks"]
    nb = len(boards)
    reps = (4096 + nb - 1) // nb
    big, bigm = np.tile(boards, reps), np.tile(masks, reps)
    ref = g[f"{tag}/logits"]
    got = mlp_forward(agent, agent._actor, big, HP)
    for r in (0, reps - 1):
        assert rel_err(got[r * nb:(r + 1) * nb], ref) < 1e-3          # north_star: 1e-3 at float32 grade
    n = len(big)
    bd, fl = dev64(big), torch.from_numpy(bigm).cuda()
    act = torch.zeros(n, dtype=torch.uint8, device="cuda")
    pr = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    lg = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    agent.policy_step(bd, fl, act, 1, 0, 1, greedy=True, probs_out=pr, logits_out=lg, precision=1)
    torch.cuda.synchronize()
    assert rel_err(lg.cpu().numpy()[:nb], ref) < 1e-2                  # north_star: 1e-2 on the reduced-precision path
    assert np.abs(pr.cpu().numpy()[:nb] - g[f"{tag}/probs"]).max() < 1e-2
    top2 = np.sort(g[f"{tag}/probs"] * np.stack([(masks >> a) & 1 for a in range(4)], 1), 1)
    clear = (top2[:, -1] - top2[:, -2]) > 2e-2
    assert (act.cpu().numpy()[:nb][clear] == g[f"{tag}/greedy"][clear]).all()


def test_gen_actor_critic_update_matches_fp32(b2048):
    """SURVEY §8d config 4 as the reference documents it (one-hot 272-256-128-64 actor and critic, Adam, critic lr 5e-4):
    rollout with the tensor-core policy, then update_from_rollout on the default (tensor-core) path vs the fp32 kernels —
    parameter steps and gradient norms within 1e-2."""
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 20; kw["obs_mode"] = "onehot"
    B = 8192
    deltas = {}
    for prec in ("auto", 0):
        env = b2048.Batched2048Env(B, b2048.Game2048EnvConfig(**kw), seed=21, gid0=3)
        agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 128, 64], activation="ReLU", init_distribution="HeNormal"),
                                     b2048.ReinforceAgentConfig(model_seed=4, use_critic=True, optimizer="adam", learning_rate=0.01,
                                                                critic_learning_rate=5e-4, baseline_mode="batch_norm"))
        assert agent.tc_supported()
        ro = agent.rollout_many(env, precision=1)
        th_a, th_c = agent._actor.theta.clone(), agent._critic.theta.clone()
        info = agent.update_from_rollout(ro, precision=prec)
        deltas[prec] = ((agent._actor.theta - th_a).cpu().numpy(), (agent._critic.theta - th_c).cpu().numpy(),
                        info["actor_grad_norm"], info["critic_grad_norm"], info["precision"], ro.boards.cpu().numpy())
    assert np.array_equal(deltas["auto"][5], deltas[0][5]), "the two rollouts differ"
    assert "tcgen05" in deltas["auto"][4] and "fp32" in deltas[0][4], (deltas["auto"][4], deltas[0][4])
    ga, gc = deltas["auto"][2] / deltas[0][2], deltas["auto"][3] / deltas[0][3]
    print(f"one-hot actor-critic update: grad-norm ratios actor {ga:.5f} critic {gc:.5f}; precision = {deltas['auto'][4]}")
    assert abs(ga - 1) < 1e-2 and abs(gc - 1) < 1e-2
    # Adam's first step is lr * sign(g) (|g| >> eps): compare the steps where the gradient is not at the noise floor
    for k in (0, 1):
        a, b = deltas["auto"][k], deltas[0][k]
        assert np.mean(np.sign(a) == np.sign(b)) > 0.97


def _gen_rollout(b2048, n, max_steps, compact, greedy=False, hidden=(256, 128, 64), obs_mode="onehot", seed=4242, gid0=17):
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = max_steps; kw["obs_mode"] = obs_mode
    b2048.debug_set("no_compact_rollout", not compact)
    try:
        benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=seed, gid0=gid0)
        agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=list(hidden), activation="ReLU", init_distribution="HeNormal"),
                                     b2048.ReinforceAgentConfig(model_seed=3))
        assert agent.tc_supported() and not agent._fused_shape()
        ro = agent.rollout_many(benv, greedy=greedy, precision=1, check_every=16)
        torch.cuda.synchronize()
    finally:
        b2048.debug_set("no_compact_rollout", False)
    return kw, benv, ro


def test_gen_rollout_compact_equals_plain_loop(b2048):
    """Run-to-termination rollout of the one-hot network: the policy kernel visiting only the live boards of every chunk
    (slot_map, device-side count) produces bit for bit the rollout of the loop that evaluates every board at every step."""
    outs = []
    for compact in (False, True):
        kw, benv, ro = _gen_rollout(b2048, 20000, 70, compact)
        outs.append((ro.T, ro.length.cpu().numpy(), ro.actions.cpu().numpy(), ro.rewards.cpu().numpy(), ro.boards.cpu().numpy(),
                     benv.score.cpu().numpy(), benv.board.cpu().numpy()))
    (Ta, La, Aa, Ra, Ba, Sa, Fa), (Tb, Lb, Ab, Rb, Bb, Sb, Fb) = outs
    assert (La == Lb).all() and (Sa == Sb).all() and (Fa == Fb).all()
    T = min(Ta, Tb)
    live = np.arange(T)[:, None] < La[None, :]
    assert (Aa[:T][live] == Ab[:T][live]).all() and (Ra[:T][live] == Rb[:T][live]).all() and (Ba[:T][live] == Bb[:T][live]).all()


@pytest.mark.parametrize("n,max_steps,hidden,obs_mode", [(32768, 120, (256, 128, 64), "onehot"), (8192, 40, (128,), "log2")])
def test_gen_rollout_replays_in_oracle(b2048, n, max_steps, hidden, obs_mode):
    """Rollout with the shape-generic tcgen05 policy kernel + the step kernel against the CPU oracle: the recorded actions
    replayed through oracle.step_many reproduce every live board, reward and flags byte, the lengths and the final counters."""
    seed, gid0 = 4242, 17
    kw, benv, ro = _gen_rollout(b2048, n, max_steps, True, hidden=hidden, obs_mode=obs_mode, seed=seed, gid0=gid0)
    T = ro.T
    boards = ro.boards.cpu().numpy().view(np.uint64); flags = ro.flags.cpu().numpy()
    actions = ro.actions.cpu().numpy(); rewards = ro.rewards.cpu().numpy(); length = ro.length.cpu().numpy()
    okw = dict(kw); okw.pop("size")
    cfg = oracle.make_cfg(action_mode="buffer", auto_reset=False, **okw)
    st = oracle.reset_many(n, seed, gid0, 0)
    assert (st["board"] == boards[0]).all()
    assert length.min() >= 1 and length.max() == T <= max_steps
    for t in range(T):
        live = length > t
        prev = {k: st[k].copy() for k in ("board", "score", "step", "max_exp")}
        o = oracle.step_many(st, cfg, seed, gid0, t + 1, action=actions[t])
        assert (st["board"][live] == boards[t + 1][live]).all(), t
        assert (o["reward"][live] == rewards[t][live]).all(), t
        assert (o["flags"][live] == flags[t + 1][live]).all(), t
        dead = ~live
        for k in prev:
            st[k][dead] = prev[k][dead]
    assert (benv.board.cpu().numpy().view(np.uint64) == st["board"]).all()
    assert (benv.score.cpu().numpy() == st["score"]).all()
    assert (benv.step_count.cpu().numpy() == st["step"]).all()
    m = flags[:T] & 0xF
    livem = np.arange(T)[:, None] < length[None, :]
    assert ((((m >> actions) & 1) == 1) | (m == 0))[livem].all()
