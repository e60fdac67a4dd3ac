"""GPU parity tests of the policy / learner kernels and of the reference-API drop-ins.
Floating point: 1e-3 relative (north_star's fp32 bar) against outputs of the live reference stored in
tests/golden/ and against the NumPy oracle; integer / index results (greedy actions, boards) exact."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import learner  # noqa: E402
from helpers import GOLDEN, UPDATE_TAGS, full_env_kwargs, load_update_fixture, random_boards, rank_weights, rel_err  # noqa: E402

TOL = 1e-3


@pytest.fixture(scope="module")
def b2048():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import b2048 as m
    return m


def dev64(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def make_agent(b2048, env_kw, mlp_kw, agent_kw, actor=None, critic=None):
    env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(**env_kw))
    agent = b2048.ReinforceAgent(env, b2048.MLPConfig(**mlp_kw), b2048.ReinforceAgentConfig(**agent_kw))
    if actor is not None:
        agent.params = actor
    if critic is not None:
        agent.critic_params = critic
    return agent


NETS = [("default", dict(obs_mode="log2", obs_log2_scale=0.0625), dict(hidden_sizes=[256, 256], activation="ReLU")),
        ("onehot", dict(obs_mode="onehot"), dict(hidden_sizes=[256, 128, 64], activation="ReLU")),
        ("sigmoid", dict(obs_mode="log2", obs_log2_scale=1.0), dict(hidden_sizes=[48], activation="Sigmoid"))]


@pytest.mark.parametrize("tag,env_kw,mlp_kw", NETS)
def test_policy_step_matches_reference(b2048, tag, env_kw, mlp_kw):
    g = np.load(os.path.join(GOLDEN, "mlp.npz"))
    L = int(g[f"{tag}/n_layers"])
    params = {"W": [g[f"{tag}/W{i}"] for i in range(L)], "b": [g[f"{tag}/b{i}"] for i in range(L)]}
    agent = make_agent(b2048, env_kw, mlp_kw, {}, actor=params)
    boards, masks = g[f"{tag}/boards"], g[f"{tag}/masks"]
    n = len(boards)
    bd, fl = dev64(boards), torch.from_numpy(masks).cuda()
    act = torch.zeros(n, dtype=torch.uint8, device="cuda")
    probs = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    logits = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    agent.policy_step(bd, fl, act, 1, 0, 1, greedy=True, probs_out=probs, logits_out=logits)
    assert rel_err(logits.cpu().numpy(), g[f"{tag}/logits"]) < TOL
    assert np.abs(probs.cpu().numpy() - g[f"{tag}/probs"]).max() < TOL
    assert (act.cpu().numpy() == g[f"{tag}/greedy"]).all()
    # sampled actions: always legal, and the empirical distribution of one board follows its probabilities
    agent.policy_step(bd, fl, act, 7, 0, 3, greedy=False)
    a = act.cpu().numpy()
    assert (((masks >> a) & 1) == 1).all()
    rep = 20000
    b1 = dev64(np.full(rep, boards[5], np.uint64))
    f1 = torch.full((rep,), int(masks[5]), dtype=torch.uint8, device="cuda")
    a1 = torch.zeros(rep, dtype=torch.uint8, device="cuda")
    agent.policy_step(b1, f1, a1, 99, 0, 5, greedy=False)
    freq = np.bincount(a1.cpu().numpy(), minlength=4) / rep
    assert np.abs(freq - g[f"{tag}/probs"][5]).max() < 0.02


def test_policy_step_large_batch_vs_oracle(b2048):
    rng = np.random.default_rng(3)
    n = 70001
    boards = random_boards(rng, n)
    masks, _ = oracle.mask_done(boards)
    params = b2048.init_model_params(16, [256, 256], 4, np.random.default_rng(0), "HeNormal")
    agent = make_agent(b2048, dict(obs_mode="log2", obs_log2_scale=0.0625), dict(hidden_sizes=[256, 256], activation="ReLU"),
                       {}, actor=params)
    logits = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    probs = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    act = torch.zeros(n, dtype=torch.uint8, device="cuda")
    agent.policy_step(dev64(boards), torch.from_numpy(masks).cuda(), act, 5, 0, 1, greedy=True, probs_out=probs,
                      logits_out=logits)
    X = learner.encode(boards, "log2", 0.0625)
    ref_logits, _, _ = learner.forward(params, X, "ReLU")
    assert rel_err(logits.cpu().numpy(), ref_logits) < TOL
    ref_p = learner.probs_from_logits(ref_logits, masks)
    assert np.abs(probs.cpu().numpy() - ref_p).max() < TOL
    legal = masks != 0
    a = act.cpu().numpy()
    assert (((masks[legal] >> a[legal]) & 1) == 1).all()


def test_forward_logits_dropin(b2048):
    g = np.load(os.path.join(GOLDEN, "mlp.npz"))
    for tag, obs_mode, actv in (("default", "log2", "ReLU"), ("onehot", "onehot", "ReLU"), ("sigmoid", "log2", "Sigmoid")):
        L = int(g[f"{tag}/n_layers"])
        params = {"W": [g[f"{tag}/W{i}"] for i in range(L)], "b": [g[f"{tag}/b{i}"] for i in range(L)]}
        X = learner.encode(g[f"{tag}/boards"], obs_mode, float(g[f"{tag}/obs_scale"]))
        logits, acts, pres = b2048.forward_logits(params, X, actv)
        assert rel_err(logits, g[f"{tag}/logits"]) < TOL
        rl, racts, rpres = learner.forward(params, X, actv)
        assert len(acts) == L + 1 and len(pres) == L
        for a, r in zip(acts, racts):
            assert rel_err(a, r) < TOL
        for a, r in zip(pres, rpres):
            assert rel_err(a, r) < TOL
        l1, a1, p1 = b2048.forward_logits(params, X[0], actv)
        assert l1.shape == (4,) and rel_err(l1, g[f"{tag}/logits"][0]) < TOL
        m = np.stack([(g[f"{tag}/masks"] >> a) & 1 for a in range(4)], 1).astype(np.int8)
        p = b2048.logits_to_probs(logits, m)
        assert np.abs(p - g[f"{tag}/probs"]).max() < TOL
    with pytest.raises(ValueError):
        b2048.forward_logits(params, X, "Tanh")
    with pytest.raises(ValueError):
        b2048.init_model_params(16, [8], 4, np.random.default_rng(0))       # reference default "normal" raises


def test_reverse_scan_and_advantages(b2048):
    import ctypes as C
    g = np.load(os.path.join(GOLDEN, "mlp.npz"))
    lens = g["ret/lens"]; rew = g["ret/rewards"]; w = g["ret/weights"]
    T, B = int(lens.max()), len(lens)
    x = np.zeros((T, B), np.float32)
    offs = np.concatenate([[0], np.cumsum(lens)])
    for b in range(B):
        x[: lens[b], b] = rew[offs[b]:offs[b + 1]]
    lib = b2048._lib.load()
    h = b2048.get_handle(torch.device("cuda", 0))
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    xd, ld, wd = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(w).cuda()
    for gamma in (0.99, 1.0, 0.5):
        ref = g[f"ret/{gamma}/off/returns"]
        y = torch.zeros_like(xd)
        b2048._lib.check(lib.b2048_reverse_scan_f64(p(xd), p(y), p(ld), gamma, T, B, None))
        got = np.concatenate([y.cpu().numpy()[: lens[b], b] for b in range(B)])
        assert (got == ref).all()                                   # float64 recurrence: bit-exact
        y2 = torch.zeros_like(xd)
        b2048._lib.check(lib.b2048_reverse_scan(p(xd), p(y2), p(ld), gamma, T, B, None))   # B < 2048: warp-shuffle scan
        got2 = np.concatenate([y2.cpu().numpy()[: lens[b], b] for b in range(B)])
        assert np.abs(got2 - ref).max() <= TOL * np.abs(ref).max()
        assert float(y.cpu().numpy()[lens[0]:, 0].sum()) == 0.0
        for mode_i, mode in enumerate(("off", "each", "batch", "batch_norm")):
            adv = torch.zeros_like(xd); coef = torch.zeros_like(xd)
            stats = torch.zeros(4, dtype=torch.float64, device="cuda"); em = torch.zeros(B, device="cuda")
            b2048._lib.check(lib.b2048_advantages(h, p(y), p(ld), p(wd), mode_i, float(B), T, B, p(adv), p(coef), p(stats),
                                                  0, p(em), None))
            refa = g[f"ret/{gamma}/{mode}/adv"]
            gota = np.concatenate([adv.cpu().numpy()[: lens[b], b] for b in range(B)])
            assert np.abs(gota - refa).max() <= TOL * max(1.0, np.abs(refa).max()), (gamma, mode)
            c = coef.cpu().numpy()
            for b in range(B):
                assert np.allclose(c[: lens[b], b], adv.cpu().numpy()[: lens[b], b] * w[b] / (lens[b] * B), rtol=1e-5, atol=1e-9)
    # large batch: the per-board float64 path inside b2048_reverse_scan
    rng = np.random.default_rng(0)
    T2, B2 = 300, 5000
    x2 = rng.normal(size=(T2, B2)).astype(np.float32)
    l2 = rng.integers(1, T2 + 1, B2).astype(np.int32)
    y3 = torch.zeros((T2, B2), device="cuda")
    b2048._lib.check(lib.b2048_reverse_scan(p(torch.from_numpy(x2).cuda()), p(y3), p(torch.from_numpy(l2).cuda()), 0.97, T2,
                                            B2, None))
    ref3 = oracle.reverse_scan(x2, l2, float(np.float32(0.97)))
    assert (y3.cpu().numpy() == ref3).all()


def episodes_to_rollout(b2048, eps):
    B = len(eps)
    lens = np.array([len(e["actions"]) for e in eps], np.int32)
    T = int(lens.max())
    boards = np.zeros((T + 1, B), np.uint64); flags = np.zeros((T + 1, B), np.uint8)
    actions = np.zeros((T, B), np.uint8); rewards = np.zeros((T, B), np.float32)
    for b, e in enumerate(eps):
        boards[: lens[b], b] = e["boards"]; flags[: lens[b], b] = e["masks"]
        actions[: lens[b], b] = e["actions"]; rewards[: lens[b], b] = e["rewards"].astype(np.float32)
    return b2048.Rollout(dev64(boards), torch.from_numpy(flags).cuda(), torch.from_numpy(actions).cuda(),
                         torch.from_numpy(rewards).cuda(), torch.from_numpy(lens).cuda(), T)


@pytest.mark.parametrize("tag", UPDATE_TAGS)
def test_update_matches_reference(b2048, tag):
    """ReinforceAgent.update_batch of the live reference (grad norms, advantages, parameters after the update)."""
    meta, actor0, critic0, updates = load_update_fixture(tag)
    agent = make_agent(b2048, meta["env"], meta["mlp"], meta["agent"], actor=actor0, critic=critic0)
    prev = actor0
    for u in updates:
        ro = episodes_to_rollout(b2048, u["episodes"])
        w = rank_weights(u["total_reward"], meta["agent"].get("reward_rank_weights"))
        ro.ep_weight = torch.from_numpy(w).cuda()
        info = agent.update_from_rollout(ro, chunk=4096)
        lens = u["lens"]
        adv = info["advantages"].cpu().numpy()
        got = np.concatenate([adv[: lens[b], b] for b in range(len(lens))])
        assert rel_err(got, u["adv"]) < TOL
        assert abs(info["actor_grad_norm"] - u["grad_norms"][0]) < TOL * u["grad_norms"][0]
        if critic0 is not None:
            assert abs(info["critic_grad_norm"] - u["grad_norms"][1]) < TOL * u["grad_norms"][1]
        new = agent.params
        for l in range(len(actor0["W"])):
            dref = u["actor"]["W"][l] - prev["W"][l]
            assert rel_err(new["W"][l] - prev["W"][l], dref) < TOL, (tag, l)       # the update itself
            assert rel_err(new["W"][l], u["actor"]["W"][l]) < 1e-4
            assert rel_err(new["b"][l] - prev["b"][l], u["actor"]["b"][l] - prev["b"][l]) < TOL
        if critic0 is not None:
            newc = agent.critic_params
            for l in range(len(critic0["W"])):
                assert rel_err(newc["W"][l], u["critic"]["W"][l]) < 1e-4
            agent.critic_params = u["critic"]
        prev = u["actor"]
        agent.params = u["actor"]


def test_update_large_batch_vs_oracle(b2048):
    """A few thousand ragged episodes through the chunked backward path (several chunks, partial tiles)."""
    rng = np.random.default_rng(1)
    B = 300
    lens = rng.integers(1, 90, B)
    eps = []
    for b in range(B):
        boards = random_boards(rng, int(lens[b]))
        masks, _ = oracle.mask_done(boards)
        masks = np.where(masks == 0, 0xF, masks).astype(np.uint8)
        acts = np.array([rng.choice([a for a in range(4) if (m >> a) & 1]) for m in masks], np.uint8)
        eps.append(dict(boards=boards, masks=masks, actions=acts, rewards=rng.integers(0, 6, int(lens[b])) * 0.5))
    actor = b2048.init_model_params(16, [128, 64], 4, np.random.default_rng(5), "HeNormal")
    critic = b2048.init_model_params(16, [128, 64], 1, np.random.default_rng(6), "HeNormal")
    kw = dict(gamma=0.97, learning_rate=3e-3, baseline_mode="batch_norm", optimizer="adam", use_critic=True,
              critic_learning_rate=1e-3, max_grad_norm=0.7)
    L = learner.Learner(actor, critic, activation="ReLU", obs_mode="log2", obs_scale=0.25, gamma=0.97, lr=3e-3,
                        baseline="batch_norm", optimizer="adam", use_critic=True, critic_lr=1e-3, max_grad_norm=0.7)
    out = L.update(eps)
    # fp32 kernels at the float32 bar; "auto" = the shape-generic tensor-core kernels for this 16-128-64 network (13 K samples)
    # at north_star's tensor-core bar
    for prec, tol in ((0, TOL), ("auto", 1e-2)):
        agent = make_agent(b2048, dict(obs_mode="log2", obs_log2_scale=0.25), dict(hidden_sizes=[128, 64], activation="ReLU"), kw,
                           actor={k: [x.copy() for x in v] for k, v in actor.items()},
                           critic={k: [x.copy() for x in v] for k, v in critic.items()})
        info = agent.update_from_rollout(episodes_to_rollout(b2048, eps), chunk=5000, precision=prec)
        assert ("tcgen05" in info["precision"]) == (prec == "auto"), info["precision"]
        assert abs(info["actor_grad_norm"] - out["actor_grad_norm"]) < tol * out["actor_grad_norm"]
        assert abs(info["critic_grad_norm"] - out["critic_grad_norm"]) < tol * out["critic_grad_norm"]
        new, newc = agent.params, agent.critic_params
        for l in range(3):
            da, ra = new["W"][l] - actor["W"][l], L.actor["W"][l] - actor["W"][l]
            dc, rc = newc["W"][l] - critic["W"][l], L.critic["W"][l] - critic["W"][l]
            if prec == 0:
                assert rel_err(da, ra) < tol and rel_err(dc, rc) < tol
            else:
                # Adam's first step is lr * g / (|g| + eps) ~ lr * sign(g): a gradient element below the tensor-core path's
                # 5e-4-class error may change sign, so the steps are compared by sign agreement (the gradients themselves are
                # held to 1e-2 in tests/test_mlp_gen_gpu.py)
                assert np.mean(np.sign(da) == np.sign(ra)) > 0.995 and np.mean(np.sign(dc) == np.sign(rc)) > 0.995


def fixed_seed_iter(base_seed):
    """runner.make_fixed_seed_iter (reference runner.py:244-261) restated."""
    rng = np.random.default_rng(base_seed)
    while True:
        yield int(rng.integers(low=0, high=np.iinfo(np.int64).max, dtype=np.int64))


def test_game2048_dropin_seeded_facts(b2048):
    with open(os.path.join(GOLDEN, "seeded.json")) as f:
        facts = json.load(f)
    g = b2048.Game2048()
    for s in ("0", "1", "2"):
        assert g.reset(seed=int(s)) == facts["reset_seed"][s]
    g.reset(seed=1)
    for st in facts["reset1_steps"]:
        ch, state, merged, done = g.step(st["action"])
        assert (ch, state, merged, done) == (st["changed"], st["state"], st["merged"], st["done"])
    assert g.score == facts["reset1_score"]
    assert g.get_action_mask() == [int(x) for x in np.array(g.get_action_mask())]
    with pytest.raises(ValueError):
        g.step(4)
    assert "+----" in g.render()
    it = fixed_seed_iter(3)
    assert [next(it) for _ in range(3)] == facts["seed_iter_3"]


def test_reference_config1_first_episodes(b2048):
    """BASELINE.json configs[0]: runner defaults, seeds 3 / 7 -> the reference's own (T, total_reward, max_tile)."""
    with open(os.path.join(GOLDEN, "seeded.json")) as f:
        facts = json.load(f)
    env_kw = full_env_kwargs("runner_default")
    env = b2048.Game2048Env(b2048.Game2048EnvConfig(**env_kw))
    agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig(gamma=0.99, learning_rate=1e-4, baseline_mode="batch", model_seed=0))
    assert [list(W.shape) for W in agent.params["W"]] == facts["actor_shapes"]
    it_e, it_p = fixed_seed_iter(3), fixed_seed_iter(7)
    trajs = []
    for k in range(3):
        tr = agent.run_episode(next(it_e), next(it_p))
        trajs.append(tr)
        assert [len(tr["actions"]), float(tr["total_reward"]), int(tr["max_tile"])] == facts["runner_default_first8"][k]
        assert len(tr["states"]) == len(tr["obs"]) == len(tr["rewards"])
        assert tr["obs"][0]["board"].dtype == np.float32 and tr["obs"][0]["action_mask"].dtype == np.int8
    r = agent.compute_returns(trajs[2]["rewards"])
    assert (r == learner.returns(trajs[2]["rewards"], 0.99)).all()
    before = agent.params
    agent.update_batch(trajs)
    after = agent.params
    assert any(np.abs(a - b).max() > 0 for a, b in zip(after["W"], before["W"]))
    # same update through the NumPy oracle of the reference's update_batch
    eps = []
    for tr in trajs:
        bm = [agent._obs_to_packed(o) for o in tr["obs"]]
        eps.append(dict(boards=np.array([x[0] for x in bm], np.uint64), masks=np.array([x[1] for x in bm], np.uint8),
                        actions=np.array(tr["actions"], np.uint8), rewards=np.array(tr["rewards"])))
    L = learner.Learner(before, None, activation="ReLU", obs_mode="log2", obs_scale=0.0625, gamma=0.99, lr=1e-4,
                        baseline="batch")
    out = L.update(eps)
    assert abs(agent.last_update_info["actor_grad_norm"] - out["actor_grad_norm"]) < TOL * out["actor_grad_norm"]
    for l in range(3):
        # lr = 1e-4: the step is a few float32 ulps of the weights, so compare weights at ulp level
        assert np.abs(after["W"][l] - L.actor["W"][l]).max() < 2e-7


def test_env_dropin_step_semantics(b2048):
    env = b2048.Game2048Env(b2048.Game2048EnvConfig(**full_env_kwargs("shaped_raw")))
    obs, info = env.reset(seed=5)
    assert set(info) == {"score", "raw_state"} and obs["board"].shape == (4, 4)
    with pytest.raises(AssertionError):
        env.step(7)
    total = 0.0
    for t in range(60):
        legal = np.flatnonzero(obs["action_mask"])
        obs, r, term, trunc, info = env.step(int(legal[0]) if len(legal) else 0)
        assert set(info) == {"score", "raw_state", "merged", "invalid_action", "step_index"}
        assert isinstance(r, float) and info["step_index"] == t + 1
        total += r
        if term or trunc:
            break
    assert trunc or term
    assert env.render("ansi").count("\n") == 8
    env1 = b2048.Game2048Env(b2048.Game2048EnvConfig(obs_mode="onehot"))
    o, _ = env1.reset(seed=0)
    assert o["board"].shape == (4, 4, 17) and float(o["board"].sum()) == 16.0
    syms = b2048.Game2048Env.get_symmetries(o, 1)
    assert len(syms) == 8 and syms[1][1] == 0 and syms[4][1] == 3


def test_rollout_many_replays_in_oracle(b2048):
    """Run-to-termination rollout of a batch: recorded actions replayed through the CPU oracle reproduce every
    board / reward / flag bit-exactly; finished boards are frozen, lengths match."""
    n, seed = 3000, 77
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 150
    benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=seed, gid0=11)
    agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[64, 32], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode="batch", learning_rate=1e-3))
    ro = agent.rollout_many(benv)
    T = ro.T
    boards = ro.boards.cpu().numpy().view(np.uint64); flags = ro.flags.cpu().numpy()
    actions = ro.actions.cpu().numpy(); rewards = ro.rewards.cpu().numpy(); length = ro.length.cpu().numpy()
    assert length.min() >= 1 and length.max() == T <= 150
    okw = dict(kw); okw.pop("size")
    cfg = oracle.make_cfg(action_mode="buffer", **okw)
    st = oracle.reset_many(n, seed, 11, 0)
    assert (st["board"] == boards[0]).all()
    for t in range(T):
        live = length > t
        o = oracle.step_many(st, cfg, seed, 11, t + 1, action=actions[t])
        assert (st["board"][live] == boards[t + 1][live]).all()
        assert (o["reward"][live] == rewards[t][live]).all()
        assert (o["flags"][live] == flags[t + 1][live]).all()
        ended = length == t + 1
        assert ((flags[t + 1][ended] & 0x60) != 0).all() and ((flags[t + 1][live & ~ended] & 0x60) == 0).all()
        dead = ~live
        assert (rewards[t][dead] == 0).all() and (boards[t + 1][dead] == boards[t][dead]).all()
        # keep the oracle's frozen boards in sync with the device's
        st["board"][dead] = boards[t + 1][dead]
    # the actions were legal wherever a legal move existed
    m = flags[:T] & 0xF
    tt = np.arange(T)[:, None]
    livem = tt < length[None, :]
    assert ((((m >> actions) & 1) == 1) | (m == 0))[livem].all()
    info = agent.update_from_rollout(ro)
    assert np.isfinite(info["actor_grad_norm"]) and info["actor_grad_norm"] > 0


def test_rollout_fixed_horizon_replays_in_oracle(b2048):
    """Fixed-horizon rollout with reset-on-done through the C rollout loop (b2048_rollout_many)."""
    n, seed, H = 33000, 5, 40
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 25
    benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=seed, gid0=3)
    agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[32], activation="Sigmoid", init_distribution="XavierNormal"),
                                 b2048.ReinforceAgentConfig())
    ro = agent.rollout_many(benv, horizon=H)
    assert ro.T == H and int(ro.length.min()) == H
    boards = ro.boards.cpu().numpy().view(np.uint64); flags = ro.flags.cpu().numpy()
    actions = ro.actions.cpu().numpy(); rewards = ro.rewards.cpu().numpy()
    okw = dict(kw); okw.pop("size")
    cfg = oracle.make_cfg(action_mode="buffer", auto_reset=True, **okw)
    st = oracle.reset_many(n, seed, 3, 0)
    for t in range(H):
        o = oracle.step_many(st, cfg, seed, 3, t + 1, action=actions[t])
        assert (st["board"] == boards[t + 1]).all() and (o["reward"] == rewards[t]).all() and (o["flags"] == flags[t + 1]).all()
    assert (benv.score.cpu().numpy() == st["score"]).all() and benv.t == H
    assert ((flags[1:] & 0x60) != 0).sum() >= n           # truncation at 25 steps resets every board at least once


def test_trainer_smoke(b2048, tmp_path):
    from b2048 import trainer
    cfg = trainer.merge_config({"mlp": {"hidden_sizes": [32, 32]},
                                "agent": {"optimizer": "adam", "learning_rate": 1e-3, "baseline_mode": "batch_norm"},
                                "train": {"batch_size": 512, "num_batches": 34, "out_dir": str(tmp_path)},
                                "eval": {"num_episodes": 256}})
    rows = trainer.training(cfg)
    assert len(rows) == 34 and all(np.isfinite(r["avg_reward"]) for r in rows)
    files = os.listdir(tmp_path)
    assert "config.json" in files and "training_stats.csv" in files and "final.npz" in files
    lines = open(os.path.join(tmp_path, "training_stats.csv")).read().strip().splitlines()
    assert lines[0] == "batch,avg_reward,max_reward,min_reward,max_tile_counts" and len(lines) == 35
    cfg["eval"]["model_path"] = os.path.join(tmp_path, "final.npz")
    res = trainer.evaluation(cfg)
    assert res["avg_reward"] > 0 and len(res["max_tile_counts"]) == 9 and sum(res["max_tile_counts"]) <= 256
    assert [int(l.split(",")[0]) for l in lines[1:]] == list(range(1, 35))            # 1-based batches like runner.py:610
    # resume from the full checkpoint (actor + Adam moments + counters) and continue the batch numbering
    assert "final_checkpoint.npz" in files
    cfg2 = trainer.merge_config({"mlp": {"hidden_sizes": [32, 32]},
                                 "agent": {"optimizer": "adam", "learning_rate": 1e-3, "baseline_mode": "batch_norm"},
                                 "train": {"batch_size": 512, "num_batches": 37, "start_batch": 34, "out_dir": str(tmp_path),
                                           "resume": os.path.join(tmp_path, "final_checkpoint.npz")}})
    more = trainer.training(cfg2)
    assert [r["batch"] for r in more] == [35, 36, 37] and all(np.isfinite(r["avg_reward"]) for r in more)
    lines = open(os.path.join(tmp_path, "training_stats.csv")).read().strip().splitlines()
    assert len(lines) == 38 and lines[0].startswith("batch,")                          # resumed rows are appended


def test_checkpoint_round_trip_and_rank_weights(b2048, tmp_path):
    """save_checkpoint / load_checkpoint restore everything an update mutates (actor, critic, Adam moments, step
    counters): an agent resumed from the checkpoint takes exactly the step the original takes.  The update runs with
    reward_rank_weights (global reward ranks on the device, reinforce_agent.py:681-716)."""
    from helpers import full_env_kwargs
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 30

    def make():
        env = b2048.Batched2048Env(2048, b2048.Game2048EnvConfig(**kw), seed=9)
        agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[64, 32], activation="ReLU", init_distribution="HeNormal"),
                                     b2048.ReinforceAgentConfig(use_critic=True, optimizer="adam", baseline_mode="batch_norm",
                                                                reward_rank_weights=[0.0, 0.5, 1.0, 2.5], model_seed=4))
        return env, agent

    env, a = make()
    a.update_from_rollout(a.rollout_many(env))
    path = str(tmp_path / "ck.npz")
    a.save_checkpoint(path)
    env2, b = make()
    b.load_checkpoint(path)
    assert torch.equal(a._actor.theta, b._actor.theta) and torch.equal(a._critic.adam_v, b._critic.adam_v)
    env.seed = env2.seed = 123                       # the same second rollout + update from both
    env2.t = env.t
    before = a._actor.theta.clone()
    ra, rb = a.rollout_many(env), b.rollout_many(env2)
    assert torch.equal(ra.actions, rb.actions) and torch.equal(ra.length, rb.length)
    a.update_from_rollout(ra)
    b.update_from_rollout(rb)
    # (the gradient atomics make the float sum order run-dependent: compare the steps, not the bits)
    da, db = a._actor.theta - before, b._actor.theta - before
    assert float((da - db).norm() / da.norm()) < 1e-3
    assert float((a._critic.theta - b._critic.theta).norm() / a._critic.theta.norm()) < 1e-5
    assert a._adam_t == b._adam_t == 2


def test_select_action_returns_forward_cache(b2048):
    """select_action's 3rd / 4th return values are forward_logits' cached activations / pre-activations
    (reference src/reinforce_agent.py:139-143, :192; src/MLP.py:159-196), here for every stored network shape."""
    g = np.load(os.path.join(GOLDEN, "mlp.npz"))
    for tag, env_kw, mlp_kw in NETS:
        L = int(g[f"{tag}/n_layers"])
        params = {"W": [g[f"{tag}/W{i}"] for i in range(L)], "b": [g[f"{tag}/b{i}"] for i in range(L)]}
        env = b2048.Game2048Env(b2048.Game2048EnvConfig(**env_kw))
        agent = b2048.ReinforceAgent(env, b2048.MLPConfig(**mlp_kw), b2048.ReinforceAgentConfig())
        agent.params = params
        obs, _ = env.reset(seed=5)
        for _ in range(3):
            action, probs, acts, pres = agent.select_action(obs, np.random.default_rng(1))
            x, mask = b2048.encode_observation(obs)
            rl, racts, rpres = learner.forward(params, x[None, :], mlp_kw["activation"])
            assert len(acts) == L + 1 and len(pres) == L
            for a, r in zip(acts, racts):
                assert a.shape == r[0].shape and rel_err(a, r[0]) < TOL
            for a, r in zip(pres, rpres):
                assert a.shape == r[0].shape and rel_err(a, r[0]) < TOL
            mbits = sum(int(v) << k for k, v in enumerate(mask))
            assert np.abs(probs - learner.probs_from_logits(rl, np.array([mbits]))[0]).max() < TOL
            assert mask[action] == 1
            obs, *_ = env.step(action)
        a2, p2, acts2, pres2 = agent.select_action(obs, np.random.default_rng(1), return_cache=False)
        assert acts2 == [] and pres2 == []
