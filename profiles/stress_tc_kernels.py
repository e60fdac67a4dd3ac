"""Stress run of the barrier-heavy tensor-core kernels: many shapes of the fused persistent rollout kernel against the
two-kernel loop (bit-exact on the live region), and the default tensor-core update (split-fp16 forward; one persistent
pipeline launch for large batches) against the fp32 kernels at the 1e-2 bar, repeated with different seeds.  A protocol race would show up as a mismatch or a hang (run it under `timeout`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, b2048

torch.cuda.set_device(0)
KW = dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)


def _samples(agent, ro):
    T, B = ro.T, ro.B
    lv = (torch.arange(T, device="cuda").unsqueeze(1) < ro.length.unsqueeze(0)).reshape(-1)
    cf = agent._scratch["coef"][: T * B][lv]
    return ro.boards[:T].reshape(-1)[lv], ro.flags[:T].reshape(-1)[lv], ro.actions[:T].reshape(-1)[lv].long(), cf


def grad_of(agent, ro, th0, prec):
    """flat actor gradient (before clipping) of update_from_rollout in the given precision, parameters restored"""
    agent._actor.theta.copy_(th0)
    agent.agent_config.max_grad_norm, keep = 1e30, agent.agent_config.max_grad_norm
    agent.update_from_rollout(ro, precision=prec)
    agent.agent_config.max_grad_norm = keep
    g = agent._actor.grad.clone().double()
    agent._actor.theta.copy_(th0)
    return g


def grad_float64(agent, ro, th0):
    """the same gradient in float64 (torch on the device, chunked): sum_s coef_s d log pi(a_s | s_s) / d theta"""
    agent._actor.theta.copy_(th0)
    agent.update_from_rollout(ro, precision=0)          # fills the coefficient buffer
    agent._actor.theta.copy_(th0)
    bd, fl, ac, cf = _samples(agent, ro)
    P = agent.params
    W = [torch.from_numpy(w).cuda().double() for w in P["W"]]
    b = [torch.from_numpy(x).cuda().double() for x in P["b"]]
    sh = torch.arange(16, device="cuda", dtype=torch.int64) * 4
    gW = [torch.zeros_like(w) for w in W]; gb = [torch.zeros_like(x) for x in b]
    for c0 in range(0, bd.numel(), 1 << 20):
        sl = slice(c0, c0 + (1 << 20))
        X = (((bd[sl].unsqueeze(1) >> sh) & 15).float() * 0.0625).double()     # log2 observations x 0.0625 (exact)
        z1 = X @ W[0] + b[0]; h1 = z1.clamp_min(0)
        z2 = h1 @ W[1] + b[1]; h2 = z2.clamp_min(0)
        lg = (h2 @ W[2] + b[2]).float()                                        # softmax in float32 like the reference
        legal = ((fl[sl].long().unsqueeze(1) >> torch.arange(4, device="cuda")) & 1).bool()
        lg = torch.where(legal, lg, torch.full_like(lg, -1e9))
        p = torch.softmax(lg, dim=1).double()
        d3 = cf[sl].double().unsqueeze(1) * (torch.nn.functional.one_hot(ac[sl], 4).double() - p)
        d2 = (d3 @ W[2].T) * (z2 > 0)
        d1 = (d2 @ W[1].T) * (z1 > 0)
        for l, (a, d) in enumerate(((X, d1), (h1, d2), (h2, d3))):
            gW[l] += a.T @ d; gb[l] += d.sum(0)
    return torch.cat([torch.cat([w.reshape(-1), x]) for w, x in zip(gW, gb)])


t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 120.0)
n_cases = 0
while time.time() < t_end:
    B = int(rng.choice([4096, 4097, 8192, 18944, 19000, 33000, 65536, 100001, 150000, 300000]))
    horizon = None if rng.random() < 0.5 else int(rng.integers(3, 70))
    max_steps = int(rng.integers(10, 90))
    greedy = bool(rng.random() < 0.3)
    seed = int(rng.integers(1, 1 << 30))
    outs = []
    for fused in (False, True):
        b2048.debug_set("no_fused_rollout", not fused)
        env = b2048.Batched2048Env(B, b2048.Game2048EnvConfig(max_steps=max_steps, **KW), seed=seed, gid0=seed % 1000)
        agent = b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                                     b2048.ReinforceAgentConfig(model_seed=seed % 97))
        ro = agent.rollout_many(env, horizon=horizon, greedy=greedy, precision=1, check_every=int(rng.integers(5, 60)) if fused else 32)
        torch.cuda.synchronize()
        outs.append((ro.T, ro.length.clone(), ro.actions.clone(), ro.rewards.clone(), ro.boards.clone(), env.score.clone(),
                     env.board.clone(), agent, ro))
    b2048.debug_set("no_fused_rollout", False)
    (Ta, La, Aa, Ra, Ba, Sa, Fa, _, _), (Tb, Lb, Ab, Rb, Bb, Sb, Fb, agent, ro) = outs
    assert torch.equal(La, Lb) and torch.equal(Sa, Sb) and torch.equal(Fa, Fb), ("state", B, horizon, max_steps)
    T = min(Ta, Tb)
    live = torch.arange(T, device="cuda").unsqueeze(1) < La.unsqueeze(0)
    live1 = torch.arange(T + 1, device="cuda").unsqueeze(1) <= La.unsqueeze(0)
    assert torch.equal(Aa[:T][live], Ab[:T][live]) and torch.equal(Ra[:T][live], Rb[:T][live]), ("record", B, horizon, max_steps)
    assert torch.equal(Ba[: T + 1][live1], Bb[: T + 1][live1]), ("boards", B, horizon, max_steps)
    if horizon is not None:     # reset-on-done lanes hold several episodes: not an update batch (update_from_rollout refuses)
        n_cases += 1
        print(f"case {n_cases}: B={B} horizon={horizon} max_steps={max_steps} greedy={greedy} T={T} rollout OK", flush=True)
        continue
    # default tensor-core update (float32-grade forward, the persistent pipeline for large batches) vs the fp32 kernels on the
    # fused rollout: the unclipped GRADIENT within the 1e-2 parity bar on every case (round 1's single-bf16 path was at
    # 0.1-0.3 here).  (The parameter step theta' - theta is not compared: at lr 1e-3 it is ~1e-6 per element against
    # theta ~ 0.1, so its own float32 representation carries ~0.5 % of noise.)
    th0 = agent._actor.theta.clone()
    ghp = grad_of(agent, ro, th0, "auto")
    g32 = grad_of(agent, ro, th0, 0)
    rel = float((ghp - g32).norm() / g32.norm())
    note = ""
    if rel > 3e-3:
        # Weak, heavily cancelling gradient: both against the float64 gradient of the same samples (torch, on the device)
        g64 = grad_float64(agent, ro, th0)
        e32, ehp = float((g32 - g64).norm() / g64.norm()), float((ghp - g64).norm() / g64.norm())
        note = f" [vs float64: fp32 kernels {e32:.1e}, tensor cores {ehp:.1e}]"
    assert np.isfinite(rel) and rel < 1e-2, ("gradient vs fp32 kernels", B, max_steps, rel, note)
    n_cases += 1
    print(f"case {n_cases}: B={B} horizon={horizon} max_steps={max_steps} greedy={greedy} T={T} gradient rel {rel:.4f}{note} OK", flush=True)
print("stress OK:", n_cases, "cases")
