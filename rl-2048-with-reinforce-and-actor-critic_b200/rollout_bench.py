"""Secondary benchmark leg: policy-rollout steps/s (BASELINE.json configs[2] shape: MLP 16-256-256-4 policy,
65,536 boards, one policy_step + one step_many launch per rollout step)."""
from __future__ import annotations

import numpy as np
import torch

from .MLP import MLPConfig
from .batched_env import Batched2048Env, Game2048EnvConfig
from .reinforce_agent import ReinforceAgent, ReinforceAgentConfig

RUNNER_ENV = dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5,
                  bonus_mode="off", max_steps=1024)
FLOPS_PER_STEP = 2 * (16 * 256 + 256 * 256 + 256 * 4)


def bench_rollout(dev, boards: int = 65536, steps: int = 64, warmup: int = 8, precision=0, gid0: int = 0, network: str = "default"):
    """network="onehot": the reference's documented one-hot 272-256-128-64-4 policy (runner.py:27-47) — on the tensor cores
    that is the shape-generic policy kernel + the step kernel, two launches per step."""
    onehot = network == "onehot"
    env = Batched2048Env(boards, Game2048EnvConfig(**(ONEHOT_ENV if onehot else RUNNER_ENV)), device=dev, seed=0xB200, gid0=gid0)
    hidden = [256, 128, 64] if onehot else [256, 256]
    agent = ReinforceAgent(env, MLPConfig(hidden_sizes=hidden, activation="ReLU", init_distribution="HeNormal"),
                           ReinforceAgentConfig(gamma=0.99, learning_rate=1e-4, baseline_mode="batch", model_seed=0))
    env.reset_many()
    agent.rollout_many(env, horizon=steps, precision=precision, reset=False)   # warm-up; also sizes the rollout buffers
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    agent.rollout_many(env, horizon=steps, precision=precision, reset=False)     # fixed horizon, reset-on-done
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rate = boards * steps / (ms * 1e-3)
    dims = agent._actor.dims
    flops = 2 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    fused = agent._fused_shape()
    return {"metric": "policy-rollout steps/s", "value": rate, "unit": "rollout-steps/s", "boards": boards, "steps": steps,
            "ms_per_step": ms / steps, "mlp": "-".join(str(d) for d in dims) + " ReLU" + (" (one-hot observations)" if onehot else ""),
            "precision": "fp32 CUDA cores" if precision == 0 else ("bf16 tcgen05 (TMEM accumulators)" if fused else
                                                                   "fp16 tcgen05, shape-generic policy kernel (TMEM accumulators)"),
            "flops_per_step": flops, "achieved_tflops": rate * flops / 1e12,
            # fp32 / generic shapes: policy kernel + step kernel per step; fused tcgen05: image prep + ONE persistent policy+env
            # kernel per 256-step chunk (policy_tc_kernel<rollout>)
            "gpu_launches": 2 * steps if (precision == 0 or not fused) else 2 * ((steps + 255) // 256)}


# the reference's documented configuration (runner.py:10-63): one-hot observations,
# hidden [256, 128, 64], Adam lr 0.01, critic lr 5e-4, empty-tile reward 0.05
ONEHOT_ENV = dict(obs_mode="onehot", obs_log2_scale=1.0, reward_mode="log2", base_reward_scale=1.0, bonus_mode="off",
                  empty_tile_reward=0.05, max_steps=1024)


def bench_train_iter(dev, boards: int = 65536, info=None, iters: int = 2, precision="auto", use_critic: bool = False,
                     update_precisions=("auto",), network: str = "default", shared_trunk: bool = False, gae_lambda: float = 0.95):
    """BASELINE.json configs[2]: REINFORCE rollout (to termination, max_steps 1024) + one update (gamma 0.99,
    baseline 'batch', SGD lr 1e-4, clip 1.0) on `boards` episodes per GPU; gradients all-reduced over ranks.
    use_critic=True is configs[3]: actor + separate critic (reference semantics, reinforce_agent.py:403-498), TD(0)
    advantages, Adam, both networks 16-256-256-{4,1} ReLU.
    shared_trunk=True is configs[3]'s literal wording: ONE 16-256-256 trunk with a policy head and a value head
    (shared_trunk.py), advantages from the lambda scan of the TD errors, one Adam step on the whole vector.
    update_precisions: the update of the LAST iteration is also timed (on a restored copy of the parameters) in these
    other modes of update_from_rollout, e.g. ("auto", 1) reports the default mode and the single-bf16 opt-in."""
    import sys
    from . import dist as bd
    info = info or bd.DistInfo()
    onehot = network == "onehot"
    env = bd.make_sharded_env(boards * info.world_size, Game2048EnvConfig(**(ONEHOT_ENV if onehot else RUNNER_ENV)), info,
                              seed=0xB200, device=dev)
    if onehot:      # runner.py:27-47
        acfg = ReinforceAgentConfig(gamma=0.99, learning_rate=0.01, critic_learning_rate=5e-4, baseline_mode="batch",
                                    use_critic=True, optimizer="adam", model_seed=0)
    else:
        acfg = (ReinforceAgentConfig(gamma=0.99, learning_rate=1e-4, critic_learning_rate=5e-4, baseline_mode="batch_norm",
                                     use_critic=True, optimizer="adam", model_seed=0) if use_critic else
                ReinforceAgentConfig(gamma=0.99, learning_rate=1e-4, baseline_mode="batch", model_seed=0))
    hidden = [256, 128, 64] if onehot else [256, 256]
    mlp_cfg = MLPConfig(hidden_sizes=hidden, activation="ReLU", init_distribution="HeNormal")
    if shared_trunk:
        from .shared_trunk import SharedTrunkActorCritic
        agent = SharedTrunkActorCritic(env, mlp_cfg, acfg, value_coef=0.5, gae_lambda=gae_lambda)
    else:
        agent = ReinforceAgent(env, mlp_cfg, acfg)
    use_critic = use_critic or onehot
    allreduce = bd.allreduce_sum_ if info.is_distributed else None
    out, alt = [], {}
    for it in range(iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        e[0].record()
        ro = agent.rollout_many(env, precision=precision)
        e[1].record()
        ro.n_traj = boards * info.world_size
        if it == iters - 1:
            saved = agent.save_state()
        upd = agent.update_from_rollout(ro, allreduce=allreduce, precision=update_precisions[0])
        e[2].record()
        torch.cuda.synchronize()
        live_steps = int(ro.length.sum().item())
        t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), float(live_steps)], dtype=torch.float64, device=dev)
        tm = t.clone()
        bd.allreduce_max_(tm)
        ts = t.clone()
        bd.allreduce_sum_(ts)
        out.append({"rollout_ms": float(tm[0]), "update_ms": float(tm[1]), "episode_steps": int(ts[2].item()),
                    "T": ro.T, "mean_len": float(ro.length.float().mean().item()),
                    "mean_return": float(ro.total_reward().mean().item()), "actor_grad_norm": upd["actor_grad_norm"]})
        if it == iters - 1:
            after = agent.save_state()
            for prec in update_precisions[1:]:
                agent.load_state(saved)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a0.record()
                u2 = agent.update_from_rollout(ro, allreduce=allreduce, precision=prec)
                a1.record()
                torch.cuda.synchronize()
                ta = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
                bd.allreduce_max_(ta)
                alt[str(prec)] = {"update_ms": float(ta[0]), "actor_grad_norm": u2["actor_grad_norm"]}
            agent.load_state(after)
    print("[bench_train_iter] " + ("actor-critic" if use_critic else "REINFORCE") + " iterations: " + repr(out), file=sys.stderr)
    last = out[-1]
    tot_ms = last["rollout_ms"] + last["update_ms"]
    res = {"metric": ("actor-critic" if use_critic else "REINFORCE") + " iteration (rollout to termination + update)",
           "boards_per_gpu": boards,
           "network": ("272-256-128-64-4 ReLU actor + 272-256-128-64-1 critic, one-hot observations (runner.py:27-47)" if onehot else
                       f"shared 16-256-256 ReLU trunk + policy head (4) + value head (1), GAE lambda {gae_lambda}" if shared_trunk else
                       "16-256-256-4 ReLU actor" + (" + 16-256-256-1 critic" if use_critic else "")),
           "rollout_precision": ("bf16 tcgen05 (fused persistent kernel)" if agent._fused_shape() else "fp16 tcgen05 (shape-generic policy kernel + step kernel)") if (agent.tc_supported() and precision != 0) else "fp32 CUDA cores",
           "episode_steps_per_s": last["episode_steps"] / (tot_ms * 1e-3), "rollout_ms": last["rollout_ms"],
           "update_ms": last["update_ms"], "update_precision": agent.last_update_info.get("precision"),
           "T": last["T"], "mean_len": last["mean_len"], "mean_return": last["mean_return"],
           "actor_grad_norm": last["actor_grad_norm"], "episode_steps": last["episode_steps"],
           "exchange": "1 all-reduce of the flat fp32 gradient buffer [actor | critic] + 4 float64 baseline sums per update"}
    for k, v in alt.items():
        res["update_ms_precision_" + k] = v["update_ms"]
    return res


def bench_sharded_sweep(dev, total_boards: int = 64 << 20, info=None, horizon: int = 16, iters: int = 2):
    """BASELINE.json configs[4]: sharded rollout sweep, `total_boards` boards in total over the ranks (32 M / 16 M / 8 M per
    GPU at 2 / 4 / 8 GPUs), a fixed 16-step horizon expressed the reference's way — Game2048EnvConfig.max_steps = 16, so
    every episode is truncated after 16 steps (env.py:279-286) and each lane holds exactly one episode — then ONE
    REINFORCE update with a single gradient all-reduce over NVLink (tensor-core rollout and update kernels)."""
    import sys
    from . import dist as bd
    info = info or bd.DistInfo()
    kw = dict(RUNNER_ENV, max_steps=horizon)
    env = bd.make_sharded_env(total_boards, Game2048EnvConfig(**kw), info, seed=0xB200, device=dev)
    agent = ReinforceAgent(env, MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"),
                           ReinforceAgentConfig(gamma=0.99, learning_rate=1e-4, baseline_mode="batch", model_seed=0))
    allreduce = bd.allreduce_sum_ if info.is_distributed else None
    out = []
    for it in range(iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        if info.is_distributed:
            torch.distributed.barrier()
        e[0].record()
        ro = agent.rollout_many(env, precision="auto", check_every=horizon)
        e[1].record()
        ro.n_traj = total_boards
        upd = agent.update_from_rollout(ro, allreduce=allreduce)
        e[2].record()
        torch.cuda.synchronize()
        t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], dtype=torch.float64, device=dev)
        bd.allreduce_max_(t)
        s = torch.tensor([float(ro.length.sum().item())], dtype=torch.float64, device=dev)
        bd.allreduce_sum_(s)
        out.append({"rollout_ms": float(t[0]), "update_ms": float(t[1]), "samples": int(s.item()),
                    "actor_grad_norm": upd["actor_grad_norm"]})
    print("[bench_sharded_sweep] iterations: " + repr(out), file=sys.stderr)
    last = out[-1]
    return {"metric": "BASELINE.json configs[4]: sharded rollout sweep + one all-reduced update", "boards_total": total_boards,
            "boards_per_gpu": env.num_envs, "n_gpus": info.world_size, "horizon": horizon,
            "horizon_semantics": "max_steps = 16 truncation (one episode per lane, run-to-termination rollout)",
            "rollout_ms": last["rollout_ms"], "update_ms": last["update_ms"], "samples": last["samples"],
            "rollout_steps_per_s": last["samples"] / (last["rollout_ms"] * 1e-3),
            "update_samples_per_s": last["samples"] / (last["update_ms"] * 1e-3),
            "update_precision": agent.last_update_info.get("precision"), "actor_grad_norm": last["actor_grad_norm"],
            "exchange": "1 all-reduce of the flat fp32 gradient (71,172 floats) + 4 float64 baseline sums"}


def bench_env_trained_boards(dev, boards: int = 1 << 20, train_batches: int = 100, harvest_steps: int = 256, steps: int = 200,
                             gid0: int = 0, seed: int = 0xB200):
    """north_star: env throughput "on synthetic random-policy and trained-policy boards" (input 2b).
    A 16-256-256-4 policy is trained for `train_batches` REINFORCE batches of 8,192 episodes (Adam, a few seconds), then
    `boards` boards are played `harvest_steps` greedy steps with it (fuller boards, higher tiles than random play) and
    the fused env step is timed on THAT state distribution: the harvested boards are restored before every timed launch
    and the L2 is flushed, exactly like the headline measurement."""
    acfg = ReinforceAgentConfig(gamma=0.99, learning_rate=1e-3, baseline_mode="batch_norm", optimizer="adam", model_seed=0)
    tenv = Batched2048Env(8192, Game2048EnvConfig(**RUNNER_ENV), device=dev, seed=seed + 1, gid0=gid0)
    agent = ReinforceAgent(tenv, MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal"), acfg)
    first = last = 0.0
    for b in range(train_batches):
        tenv.seed = (seed + 0x9E3779B97F4A7C15 * (b + 1)) & (2**64 - 1)
        ro = agent.rollout_many(tenv, precision="auto")
        avg = float(ro.total_reward().mean())
        first = avg if b == 0 else first
        last = avg
        agent.update_from_rollout(ro)
    env = Batched2048Env(boards, Game2048EnvConfig(**RUNNER_ENV), device=dev, seed=seed, gid0=gid0)
    env.reset_many()
    for _ in range(harvest_steps):                       # the random-policy reference distribution of the headline number
        env.step_many(action_mode="random_legal", auto_reset=True)
    rnd_board = env.board.clone()
    env.reset_many()
    agent.rollout_many(env, horizon=harvest_steps, greedy=True, precision="auto", reset=False)
    saved = [t.clone() for t in (env.board, env.flags, env.score, env.step_count, env.max_exp)]

    def stats(bd):
        shifts = torch.arange(16, device=dev, dtype=torch.int64) * 4
        cells = (bd.unsqueeze(-1) >> shifts) & 15
        return {"mean_max_exponent": float(cells.max(dim=1).values.float().mean()),
                "mean_occupied_cells": float((cells > 0).sum(dim=1).float().mean())}

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = 0.0
    for k in range(steps + 3):
        for dst, src in zip((env.board, env.flags, env.score, env.step_count, env.max_exp), saved):
            dst.copy_(src)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.step_many(action_mode="random_legal", auto_reset=True)
        e1.record()
        e1.synchronize()
        if k >= 3:
            ms += e0.elapsed_time(e1)
    return {"metric": "env-steps/s on trained-policy boards", "value": boards * steps / (ms * 1e-3), "unit": "env-steps/s",
            "boards": boards, "steps": steps, "ms_per_step": ms / steps,
            "policy": f"16-256-256-4 ReLU, {train_batches} REINFORCE batches of 8192 episodes (avg return {first:.0f} -> {last:.0f}), "
                      f"{harvest_steps} greedy steps",
            "trained_boards": stats(saved[0]), "random_policy_boards": stats(rnd_board), "gpu_launches": steps}
