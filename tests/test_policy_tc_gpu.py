"""bf16 tcgen05 policy kernel (precision=1) against the fp32 reference arithmetic: 1e-2 relative (north_star's
bf16 bar), sampled / greedy actions legal, distribution consistent with the fp32 probabilities."""
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


@pytest.mark.parametrize("obs_mode,scale,n", [("log2", 0.0625, 4096), ("log2", 1.0, 70001), ("raw", 1.0, 12345)])
def test_tc_policy_vs_fp32(b2048, obs_mode, scale, n):
    rng = np.random.default_rng(3)
    boards = random_boards(rng, n)
    if obs_mode == "raw":        # keep raw tile values moderate so fp32 and bf16 are comparable
        boards &= np.uint64(0x7777777777777777)
    masks, _ = oracle.mask_done(boards)
    params = b2048.init_model_params(16, [256, 256], 4, np.random.default_rng(0), "HeNormal")
    params["b"] = [rng.normal(size=b.shape).astype(np.float32) * 0.1 for b in params["b"]]
    env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(obs_mode=obs_mode, obs_log2_scale=scale))
    agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU"), b2048.ReinforceAgentConfig())
    agent.params = params
    assert agent.tc_supported()
    bd, fl = dev64(boards), torch.from_numpy(masks).cuda()
    logits = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    probs = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    act = torch.full((n,), 9, dtype=torch.uint8, device="cuda")
    agent.policy_step(bd, fl, act, 5, 0, 1, greedy=True, probs_out=probs, logits_out=logits, precision=1)
    torch.cuda.synchronize()
    X = learner.encode(boards, obs_mode, scale)
    ref_logits, _, _ = learner.forward(params, X, "ReLU")
    err = rel_err(logits.cpu().numpy(), ref_logits)
    assert err < 1e-2, err
    ref_p = learner.probs_from_logits(ref_logits, masks)
    p = probs.cpu().numpy()
    assert np.abs(p.sum(1) - 1).max() < 1e-5
    tight = obs_mode != "raw"      # raw tile values give logits of magnitude ~1e2: a 1e-2 relative logit error moves probs
    if tight:
        assert np.abs(p - ref_p).max() < 5e-2
    a = act.cpu().numpy()
    legal = masks != 0
    assert (a < 4).all() and (((masks[legal] >> a[legal]) & 1) == 1).all()
    # greedy agrees with the fp32 argmax wherever the fp32 margin is not tiny
    m = np.stack([(masks >> k) & 1 for k in range(4)], 1)
    q = ref_p * m
    top2 = np.sort(q, 1)[:, -2:]
    clear = legal & ((top2[:, 1] - top2[:, 0]) > 0.05)
    assert (a[clear] == q[clear].argmax(1)).mean() > (0.999 if tight else 0.97)
    # sampling: legal, and fp32 / bf16 paths draw from (nearly) the same distribution with the same uniforms
    a1 = torch.zeros(n, dtype=torch.uint8, device="cuda")
    a0 = torch.zeros(n, dtype=torch.uint8, device="cuda")
    agent.policy_step(bd, fl, a1, 7, 0, 3, greedy=False, precision=1)
    agent.policy_step(bd, fl, a0, 7, 0, 3, greedy=False, precision=0)
    a1, a0 = a1.cpu().numpy(), a0.cpu().numpy()
    assert (((masks[legal] >> a1[legal]) & 1) == 1).all()
    assert (a1 == a0).mean() > (0.97 if tight else 0.9)


def test_tc_policy_unsupported_shape_is_loud(b2048):
    env = b2048.Batched2048Env(1, b2048.Game2048EnvConfig(obs_mode="log2"))
    # (hidden sizes that are not multiples of 64, or Sigmoid: outside the hand-specialised AND the shape-generic tcgen05 kernels)
    agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[64, 48], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig())
    assert not agent.tc_supported()
    bd = torch.zeros(4096, dtype=torch.int64, device="cuda")
    act = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    with pytest.raises(b2048.B2048Error):
        agent.policy_step(bd, None, act, 0, 0, 0, precision=1)
    agent.policy_step(bd, None, act, 0, 0, 0, precision="auto")     # falls back to the fp32 kernel by design


def test_tc_rollout_statistics(b2048):
    """Rollouts driven by the tensor-core policy behave like the fp32 ones (same mean return within noise)."""
    n = 8192
    cfg = b2048.Game2048EnvConfig(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5)
    out = {}
    for prec in (0, 1):
        env = b2048.Batched2048Env(n, cfg, seed=11)
        agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                     b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode="batch"))
        ro = agent.rollout_many(env, precision=prec)
        out[prec] = (float(ro.total_reward().mean()), float(ro.length.float().mean()))
    assert abs(out[0][0] - out[1][0]) / out[0][0] < 0.05 and abs(out[0][1] - out[1][1]) / out[0][1] < 0.05


@pytest.mark.parametrize("n,horizon,max_steps", [(4096, 24, 1024), (33000, None, 40), (33000, 50, 30), (150000, 12, 1024)])
def test_fused_rollout_kernel_equals_two_kernel_loop(b2048, n, horizon, max_steps):
    """policy_tc_kernel<rollout> (policy on tcgen05 + env step in ONE persistent launch for the whole horizon) must
    reproduce the policy-kernel / step-kernel loop bit for bit: same boards, flags, actions, rewards, episode
    lengths and counters.  Covers one tile per CTA (no prefetch), a mix of one and two, and many tiles per CTA;
    run-to-termination (frozen episodes) and fixed horizon with reset-on-done."""
    import os
    from helpers import full_env_kwargs
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = max_steps
    outs = []
    for fused in (False, True):
        b2048.debug_set("no_fused_rollout", not fused)
        try:
            benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=77, gid0=5)
            agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                         b2048.ReinforceAgentConfig())
            agent.params = b2048.init_model_params(16, [256, 256], 4, np.random.default_rng(2), "HeNormal")
            ro = agent.rollout_many(benv, horizon=horizon, precision=1)
            torch.cuda.synchronize()
            T = ro.T
            outs.append(dict(T=T, boards=ro.boards[:T + 1].cpu().numpy(), flags=ro.flags[:T + 1].cpu().numpy(),
                             actions=ro.actions[:T].cpu().numpy(), rewards=ro.rewards[:T].cpu().numpy(),
                             length=ro.length.cpu().numpy(), score=benv.score.cpu().numpy(), step=benv.step_count.cpu().numpy(),
                             max_exp=benv.max_exp.cpu().numpy(), final_board=benv.board.cpu().numpy(),
                             final_flags=benv.flags.cpu().numpy()))
        finally:
            b2048.debug_set("no_fused_rollout", False)
    a, b = outs
    assert a["T"] == b["T"]
    T, L = a["T"], a["length"]
    assert (a["length"] == b["length"]).all()
    # Run-to-termination on the fused kernel plays only the live boards of every chunk (slot_map), so slices beyond an
    # episode's end are not written there: compare the record where it is defined (t < length; boards / flags t <= length)
    live = np.arange(T)[:, None] < L[None, :]
    live1 = np.arange(T + 1)[:, None] <= L[None, :]
    for k in ("actions", "rewards"):
        assert (a[k][live] == b[k][live]).all(), (k, int((a[k][live] != b[k][live]).sum()))
    for k in ("boards", "flags"):
        assert (a[k][live1] == b[k][live1]).all(), (k, int((a[k][live1] != b[k][live1]).sum()))
    for k in ("score", "step", "max_exp", "final_board", "final_flags"):
        assert (a[k] == b[k]).all(), (k, int((a[k] != b[k]).sum()))
    if horizon is not None:                                   # fixed horizon: every slot is live, the whole record matches
        for k in ("boards", "flags", "actions", "rewards"):
            assert (a[k] == b[k]).all(), k
    assert a["rewards"][live].sum() > 0


def test_fused_rollout_greedy_equals_two_kernel_loop(b2048):
    """Greedy (evaluation) rollouts: fused persistent kernel == policy-kernel / step-kernel loop, raw observations."""
    import os
    from helpers import full_env_kwargs
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 60; kw["obs_mode"] = "raw"; kw["obs_log2_scale"] = 1.0
    outs = []
    for fused in (False, True):
        b2048.debug_set("no_fused_rollout", not fused)
        try:
            benv = b2048.Batched2048Env(20000, b2048.Game2048EnvConfig(**kw), seed=5, gid0=0)
            agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                         b2048.ReinforceAgentConfig())
            p = b2048.init_model_params(16, [256, 256], 4, np.random.default_rng(8), "HeNormal")
            p["W"][0] = p["W"][0] * 0.05                 # raw tile values are large: keep the logits moderate
            agent.params = p
            ro = agent.rollout_many(benv, greedy=True, precision=1)
            torch.cuda.synchronize()
            outs.append((ro.T, ro.length.cpu().numpy(), ro.actions.cpu().numpy(), ro.rewards.cpu().numpy(),
                         benv.score.cpu().numpy(), benv.board.cpu().numpy()))
        finally:
            b2048.debug_set("no_fused_rollout", False)
    (Ta, La, Aa, Ra, Sa, Ba), (Tb, Lb, Ab, Rb, Sb, Bb) = outs
    assert (La == Lb).all() and (Sa == Sb).all() and (Ba == Bb).all()
    T = min(Ta, Tb)
    live = np.arange(T)[:, None] < La[None, :]
    assert (Aa[:T][live] == Ab[:T][live]).all() and (Ra[:T][live] == Rb[:T][live]).all()


@pytest.mark.parametrize("n,horizon,max_steps,env_name", [(65536, None, 150, "runner_default"), (262144, None, 40, "runner_default"),
                                                          (65536, 48, 30, "runner_default"), (262144, 20, 1024, "runner_default"),
                                                          (65536, None, 120, "shaped_log2"), (65536, 40, 25, "shaped_log2")])
def test_fused_tc_rollout_replays_in_oracle(b2048, n, horizon, max_steps, env_name):
    """The fused persistent tcgen05 rollout kernel (policy_tc_kernel<rollout>, 16-256-256-4) against the CPU ORACLE
    directly (not against the repo's other path): the recorded actions replayed through oracle.step_many reproduce
    every live board, reward and flags byte, the episode lengths, and the final score / step / max-tile counters —
    run to termination (live-board compaction between chunks) and fixed horizon with reset-on-done."""
    from helpers import full_env_kwargs
    seed, gid0 = 4242, 17
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = max_steps
    if env_name == "shaped_log2":     # every reward-shaping term of env.py:226-259 inside the fused kernel (step_fast_rnd)
        kw.update(empty_tile_reward=0.05, merge_reward=0.3, bonus_mode="log2", bonus_scale=2.0, step_reward=-0.01, endgame_penalty=-7.5)
    benv = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=seed, gid0=gid0)
    agent = b2048.ReinforceAgent(benv, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig(model_seed=3))
    assert agent.tc_supported()
    ro = agent.rollout_many(benv, horizon=horizon, precision=1)
    torch.cuda.synchronize()
    T = ro.T
    boards = ro.boards.cpu().numpy().view(np.uint64); flags = ro.flags.cpu().numpy()
    actions = ro.actions.cpu().numpy(); rewards = ro.rewards.cpu().numpy(); length = ro.length.cpu().numpy()
    okw = dict(kw); okw.pop("size")
    fixed = horizon is not None
    cfg = oracle.make_cfg(action_mode="buffer", auto_reset=fixed, **okw)
    st = oracle.reset_many(n, seed, gid0, 0)
    assert (st["board"] == boards[0]).all() and ((st["flags"] & 0xF) == (flags[0] & 0xF)).all()
    if fixed:
        assert T == horizon and (length == T).all()
    else:
        assert length.min() >= 1 and length.max() == T <= max_steps
    frozen = {k: st[k].copy() for k in ("score", "step", "max_exp")}
    for t in range(T):
        live = length > t
        prev = {k: st[k].copy() for k in ("board", "score", "step", "max_exp")}
        o = oracle.step_many(st, cfg, seed, gid0, t + 1, action=actions[t])
        assert (st["board"][live] == boards[t + 1][live]).all(), t
        assert (o["reward"][live] == rewards[t][live]).all(), t
        assert (o["flags"][live] == flags[t + 1][live]).all(), t
        ended = length == t + 1
        if not fixed:
            assert ((flags[t + 1][ended] & 0x60) != 0).all() and ((flags[t + 1][live & ~ended] & 0x60) == 0).all()
            # finished episodes are frozen on the device: keep the oracle's copy of them untouched too
            dead = ~live
            for k in prev:
                st[k][dead] = prev[k][dead]
    # final env state of every board (counters included) as the oracle left it
    assert (benv.board.cpu().numpy().view(np.uint64) == st["board"]).all()
    assert (benv.score.cpu().numpy() == st["score"]).all()
    assert (benv.step_count.cpu().numpy() == st["step"]).all()
    assert (benv.max_exp.cpu().numpy() == st["max_exp"]).all()
    # sampled actions are legal wherever a legal move exists
    m = flags[:T] & 0xF
    livem = np.arange(T)[:, None] < length[None, :]
    assert ((((m >> actions) & 1) == 1) | (m == 0))[livem].all()
