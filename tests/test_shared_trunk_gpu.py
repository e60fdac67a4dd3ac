"""Shared-trunk actor-critic (BASELINE.json configs[3]'s wording; an addition to the reference's separate critic —
b2048/shared_trunk.py): the merged gradient against a plain PyTorch float32 autograd restatement of the same objective, the
lambda advantage scan against a torch loop, lambda = 0 against the reference-equivalent TD(0) advantages of the separate-critic
agent, and a short learning run."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import full_env_kwargs, rel_err  # noqa: E402


@pytest.fixture(scope="module")
def b2048():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import b2048 as m
    return m


def make(b2048, n, seed, hidden=(256, 256), lam=0.0, vc=0.5, baseline="batch_norm", max_steps=24, optimizer="sgd", obs="log2"):
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = max_steps
    if obs == "onehot":
        kw.update(obs_mode="onehot", obs_log2_scale=1.0)
    env = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=seed, gid0=0)
    agent = b2048.SharedTrunkActorCritic(
        env, b2048.MLPConfig(hidden_sizes=list(hidden), activation="ReLU", init_distribution="HeNormal"),
        b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode=baseline, learning_rate=1e-3, optimizer=optimizer, model_seed=5,
                                   max_grad_norm=1e9),
        value_coef=vc, gae_lambda=lam)
    return env, agent


def torch_objective_grad(agent, ro, coef, gcoef, vc, scale, onehot):
    """autograd gradient of J = sum coef log pi(a|s) - vc sum gcoef V(s) over the live samples, float32, in the layout of
    the shared vector [trunk | W_pi b_pi | W_v b_v]."""
    T, B = ro.T, ro.B
    live = (torch.arange(T, device="cuda").unsqueeze(1) < ro.length.unsqueeze(0)).reshape(-1)
    boards = ro.boards[:T].reshape(-1)[live]
    flags = ro.flags[:T].reshape(-1)[live].long()
    acts = ro.actions[:T].reshape(-1)[live].long()
    cf, gc = coef.reshape(-1)[live], gcoef.reshape(-1)[live]
    e = torch.stack([(boards >> (4 * i)) & 15 for i in range(16)], 1)
    x = torch.nn.functional.one_hot(e, 17).float().reshape(-1, 272) if onehot else e.float() * scale
    p = agent.params
    vh = agent.value_head
    Ws = [torch.tensor(W, device="cuda", requires_grad=True) for W in p["W"]]
    bs = [torch.tensor(b, device="cuda", requires_grad=True) for b in p["b"]]
    Wv = torch.tensor(vh["W"], device="cuda", requires_grad=True)
    bv = torch.tensor(vh["b"], device="cuda", requires_grad=True)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        h = x
        for W, b in zip(Ws[:-1], bs[:-1]):
            h = torch.relu(h @ W + b)
        logits = h @ Ws[-1] + bs[-1]
        V = (h @ Wv + bv).reshape(-1)
        legal = ((flags.unsqueeze(1) >> torch.arange(4, device="cuda")) & 1).bool()
        logp = torch.log_softmax(logits.masked_fill(~legal, float("-inf")), 1).gather(1, acts.unsqueeze(1)).reshape(-1)
        J = (cf * logp).sum() - vc * (gc * V).sum()
        J.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    parts = []
    for W, b in zip(Ws, bs):
        parts += [W.grad.reshape(-1), b.grad.reshape(-1)]
    parts += [Wv.grad.reshape(-1), bv.grad.reshape(-1)]
    return torch.cat(parts), V.detach(), live


@pytest.mark.parametrize("prec,hidden,obs,tol", [(0, (256, 256), "log2", 2e-4), ("auto", (256, 256), "log2", 1e-2),
                                                 (0, (64, 32), "log2", 2e-4), ("auto", (256, 128, 64), "onehot", 1e-2)])
def test_shared_trunk_gradient_vs_torch_autograd(b2048, prec, hidden, obs, tol):
    n = 4096
    env, agent = make(b2048, n, seed=11, hidden=hidden, vc=0.7, obs=obs)
    ro = agent.rollout_many(env, precision=0)
    th0 = agent._shared_net.theta.clone()
    info = agent.update_from_rollout(ro, precision=prec)
    T, B = ro.T, ro.B
    coef = agent._scratch["coef"][: T * B].clone()
    gcoef = agent._scratch["gcoef"][: T * B].clone()
    got = agent._shared_net.grad.clone()
    assert got.numel() == agent._actor.n_params + hidden[-1] + 1
    agent._shared_net.theta.copy_(th0)                       # the torch restatement differentiates at the pre-update parameters
    ref, V, live = torch_objective_grad(agent, ro, coef, gcoef, 0.7, 0.0625, obs == "onehot")
    nt, na = agent._n_trunk, agent._n_policy
    errs = {"trunk": rel_err(got[:nt].cpu().numpy(), ref[:nt].cpu().numpy()),
            "policy head": rel_err(got[nt:na].cpu().numpy(), ref[nt:na].cpu().numpy()),
            "value head": rel_err(got[na:].cpu().numpy(), ref[na:].cpu().numpy())}
    print(f"shared trunk {list(hidden)} {obs} ({info['precision']}): gradient vs torch autograd fp32 {errs}")
    assert max(errs.values()) < tol, errs
    # the values the TD errors were built from are the value view's forward
    td = info["td"].reshape(-1)[live]
    r = ro.rewards[:T].reshape(-1)[live]
    Vg = torch.zeros(T * B, device="cuda"); Vg[live] = V
    t_idx = torch.arange(T, device="cuda").unsqueeze(1).expand(T, B).reshape(-1)
    nxt_live = (t_idx + 1) < ro.length.repeat(T)
    Vn = torch.where(nxt_live, torch.roll(Vg, -B), torch.zeros_like(Vg))[live]
    assert rel_err(td.cpu().numpy(), (r + 0.99 * Vn - V).cpu().numpy()) < (1e-4 if prec == 0 else 1e-3)
    # SGD step without clipping: theta += lr * g
    agent._shared_net.theta.copy_(th0)
    agent.update_from_rollout(ro, precision=prec)
    step = (agent._shared_net.theta - th0).cpu().numpy()
    # (theta' - theta) carries the float32 rounding of theta ~ 0.1 against steps ~ 1e-6
    assert rel_err(step, 1e-3 * agent._shared_net.grad.cpu().numpy()) < 5e-3


def test_gae_lambda_scan_and_lambda0_equals_td0(b2048):
    n = 4096
    env, agent = make(b2048, n, seed=12, lam=0.9, baseline="off", max_steps=32)
    ro = agent.rollout_many(env, precision=0)
    info = agent.update_from_rollout(ro, precision=0)
    T, B = ro.T, ro.B
    td = info["td"].double().cpu().numpy().reshape(T, B)
    L = ro.length.cpu().numpy()
    A = np.zeros((T, B))
    run = np.zeros(B)
    for t in range(T - 1, -1, -1):
        livet = t < L
        run = np.where(livet, td[t] + 0.99 * 0.9 * run, 0.0)
        A[t] = run
    got = info["advantages"].cpu().numpy().reshape(T, B)
    assert rel_err(got, A.astype(np.float32)) < 1e-6
    # lambda = 0: the advantages are the TD errors themselves (the reference's actor-critic, reinforce_agent.py:495-498)
    env0, a0 = make(b2048, n, seed=12, lam=0.0, baseline="off", max_steps=32)
    r0 = a0.rollout_many(env0, precision=0)
    i0 = a0.update_from_rollout(r0, precision=0)
    assert torch.equal(i0["advantages"], i0["td"])


def test_shared_trunk_rollout_is_the_policy_view_and_state_round_trips(b2048, tmp_path):
    """Rollouts run on the policy view with the fused tensor-core kernel; checkpoints restore the whole shared vector."""
    n = 8192
    env, agent = make(b2048, n, seed=13, optimizer="adam", max_steps=40)
    assert agent.tc_supported() and agent._fused_shape()
    ro = agent.rollout_many(env, precision="auto")
    # a plain ReinforceAgent holding the same policy parameters plays the same episodes
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 40
    env2 = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=13, gid0=0)
    plain = b2048.ReinforceAgent(env2, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                 b2048.ReinforceAgentConfig(gamma=0.99, model_seed=5))
    plain.params = agent.params
    ro2 = plain.rollout_many(env2, precision="auto")
    assert torch.equal(ro.length, ro2.length) and torch.equal(ro.rewards[: ro.T], ro2.rewards[: ro2.T])
    agent.update_from_rollout(ro)
    path = str(tmp_path / "shared.npz")
    agent.save_checkpoint(path)
    st = agent.save_state()
    ro = agent.rollout_many(env, precision="auto")
    agent.update_from_rollout(ro)
    after = agent._shared_net.theta.clone()
    agent.load_state(st)
    env_b, agent_b = make(b2048, n, seed=13, optimizer="adam", max_steps=40)
    agent_b.load_checkpoint(path)
    assert torch.equal(agent_b._shared_net.theta, agent._shared_net.theta) and agent_b._adam_t == agent._adam_t == 1
    assert not torch.equal(after, agent._shared_net.theta)
    cp = agent.critic_params
    assert [w.shape for w in cp["W"]] == [(16, 256), (256, 256), (256, 1)]


def test_shared_trunk_learns(b2048):
    """A few dozen Adam updates with the default (tensor-core) update raise the average return (the TD error is printed only:
    its scale grows with the episode lengths of the improving policy)."""
    n = 8192
    kw = full_env_kwargs("runner_default"); kw["max_steps"] = 1024
    env = b2048.Batched2048Env(n, b2048.Game2048EnvConfig(**kw), seed=99, gid0=0)
    agent = b2048.SharedTrunkActorCritic(
        env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
        b2048.ReinforceAgentConfig(gamma=0.99, baseline_mode="batch_norm", learning_rate=1e-3, optimizer="adam", model_seed=1),
        value_coef=0.5, gae_lambda=0.9)
    first = last = None
    for it in range(40):
        ro = agent.rollout_many(env, precision="auto")
        ret = float(ro.total_reward().mean())
        info = agent.update_from_rollout(ro)
        live = torch.arange(ro.T, device="cuda").unsqueeze(1) < ro.length.unsqueeze(0)
        td_rms = float((info["td"][: ro.T] ** 2)[live].mean().sqrt())
        if it == 0:
            first = (ret, td_rms)
        last = (ret, td_rms)
    print(f"shared-trunk actor-critic, 40 Adam updates of {n} episodes: mean return {first[0]:.1f} -> {last[0]:.1f}, "
          f"TD rms {first[1]:.3f} -> {last[1]:.3f}")
    assert last[0] > 1.3 * first[0]
