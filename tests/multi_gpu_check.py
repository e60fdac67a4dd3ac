"""Multi-GPU check, launched by hand on the GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

1. sharding invariance on real GPUs: the union of the ranks' shards equals a single-GPU run bit for bit;
2. the NCCL-all-reduced update equals the update of one process holding every episode (fp32 sum order only).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2048  # noqa: E402
from b2048 import dist as bd  # noqa: E402

ENV = dict(obs_mode="log2", obs_log2_scale=0.0625, reward_mode="log2", base_reward_scale=0.5, max_steps=200)


def make_agent(env, baseline, hidden=(64, 64)):
    return b2048.ReinforceAgent(env, b2048.MLPConfig(hidden_sizes=list(hidden), activation="ReLU", init_distribution="HeNormal"),
                                b2048.ReinforceAgentConfig(gamma=0.99, learning_rate=1e-2, baseline_mode=baseline, model_seed=3))


def main():
    info = bd.init_distributed("nccl")
    dev = torch.device("cuda", info.local_rank)
    total, seed = 40000 + 3, 4242
    cfg = b2048.Game2048EnvConfig(**ENV)
    # ---- 1. env streams
    env = bd.make_sharded_env(total, cfg, info, seed=seed)
    env.reset_many()
    for _ in range(50):
        env.step_many(action_mode="random_legal", auto_reset=True)
    lo, hi = bd.shard_range(total, info.rank, info.world_size)
    allb = torch.zeros(total, dtype=torch.int64, device=dev)
    allb[lo:hi] = env.board
    bd.allreduce_sum_(allb)
    if info.rank == 0:
        full = b2048.Batched2048Env(total, cfg, device=dev, seed=seed)
        full.reset_many()
        for _ in range(50):
            full.step_many(action_mode="random_legal", auto_reset=True)
        assert torch.equal(full.board, allb), "sharded env differs from the single-GPU run"
        print("env sharding invariance OK")
    # ---- 2. update
    for baseline in ("batch", "batch_norm", "off"):
        env = bd.make_sharded_env(total, cfg, info, seed=seed + 1)
        agent = make_agent(env, baseline)
        ro = agent.rollout_many(env)
        upd = bd.sharded_update(agent, ro, info, precision=0)     # fp32 kernels on both sides: the sharding / collective logic is under test
        theta = agent._actor.theta.clone()
        chk = theta.clone()
        dist.broadcast(chk, src=0)
        assert torch.equal(chk, theta), "ranks diverged after the all-reduced update"
        if info.rank == 0:
            env1 = b2048.Batched2048Env(total, cfg, device=dev, seed=seed + 1)
            a1 = make_agent(env1, baseline)
            r1 = a1.rollout_many(env1)
            u1 = a1.update_from_rollout(r1, precision=0)
            d_ref = (a1._actor.theta - torch.from_numpy(np.concatenate(
                [np.concatenate([W.reshape(-1), b]) for W, b in zip(*[make_agent(env1, baseline).params[k] for k in ("W", "b")])])).to(dev))
            d_got = theta - (a1._actor.theta - d_ref)
            rel = float((d_got - d_ref).norm() / d_ref.norm())
            gn = abs(upd["actor_grad_norm"] - u1["actor_grad_norm"]) / u1["actor_grad_norm"]
            print(f"baseline={baseline}: update rel err {rel:.2e}, grad-norm rel err {gn:.2e}")
            assert rel < 1e-3 and gn < 1e-3
    # ---- 3. tensor-core paths (16-256-256-4): the fused rollout kernel is sharding-invariant bit for bit (Philox is
    #         keyed on the global board id), and the all-reduced tcgen05 update equals the single-process one up to
    #         the fp32 summation order of the gradient atomics
    env = bd.make_sharded_env(total, cfg, info, seed=seed + 2)
    agent = make_agent(env, "batch", hidden=(256, 256))
    ro = agent.rollout_many(env, precision=1)
    lo, hi = bd.shard_range(total, info.rank, info.world_size)
    T_all = torch.tensor([ro.T], device=dev)
    dist.all_reduce(T_all, op=dist.ReduceOp.MAX)
    Tm = int(T_all.item())
    acts = torch.zeros((Tm, total), dtype=torch.int32, device=dev)
    lens = torch.zeros(total, dtype=torch.int32, device=dev)
    acts[: ro.T, lo:hi] = ro.actions.int() * (torch.arange(ro.T, device=dev).unsqueeze(1) < ro.length.unsqueeze(0))
    lens[lo:hi] = ro.length
    bd.allreduce_sum_(acts)
    bd.allreduce_sum_(lens)
    upd = bd.sharded_update(agent, ro, info)
    theta = agent._actor.theta.clone()
    if info.rank == 0:
        env1 = b2048.Batched2048Env(total, cfg, device=dev, seed=seed + 2)
        a1 = make_agent(env1, "batch", hidden=(256, 256))
        th0 = a1._actor.theta.clone()
        r1 = a1.rollout_many(env1, precision=1)
        assert torch.equal(r1.length, lens), "fused rollout: episode lengths differ between sharded and single-GPU runs"
        live = torch.arange(r1.T, device=dev).unsqueeze(1) < r1.length.unsqueeze(0)
        assert torch.equal((r1.actions.int() * live)[:Tm], acts[: r1.T]), "fused rollout: actions differ"
        u1 = a1.update_from_rollout(r1)      # same mode as the sharded update: "auto" = the float32-grade tensor-core path
        d_ref, d_got = a1._actor.theta - th0, theta - th0
        rel = float((d_got - d_ref).norm() / d_ref.norm())
        gn = abs(upd["actor_grad_norm"] - u1["actor_grad_norm"]) / u1["actor_grad_norm"]
        print(f"tensor-core path ({upd['precision']}): fused rollout sharding invariance OK; update rel err {rel:.2e}, grad-norm rel err {gn:.2e}")
        assert rel < 1e-3 and gn < 1e-3
    # ---- 4. x8 symmetry augmentation and the actor-critic update under sharding: every rank applies exactly the update
    #         of one process holding all episodes (n_traj stays the GLOBAL episode count x 8), and an update issues one
    #         gradient all-reduce ([actor | critic] in one flat message) plus, for the batch baselines, the 4-double
    #         statistics message that has to precede the gradient
    calls = []
    real_allreduce = bd.allreduce_sum_

    def counting_allreduce(t):
        calls.append(int(t.numel()))
        return real_allreduce(t)

    #         exchange="one_message" (g = (g_A - mean g_B) / std formed after the exchange) issues exactly one
    #         all-reduce per update whatever the baseline
    ac_kw = dict(use_critic=True, baseline_mode="batch_norm", optimizer="adam", critic_learning_rate=1e-3)
    for name, kw, expect_calls, exchange in (
            ("augmentation", dict(augmentation=True, baseline_mode="batch"), 2, "default"),
            ("actor-critic", ac_kw, 2, "default"),
            ("actor-critic, baseline off", dict(use_critic=True, baseline_mode="off"), 1, "default"),
            ("REINFORCE batch baseline, one-message exchange", dict(baseline_mode="batch"), 1, "one_message"),
            ("REINFORCE batch_norm + augmentation, one-message exchange", dict(augmentation=True, baseline_mode="batch_norm"), 1,
             "one_message"),
            ("actor-critic, one-message exchange", ac_kw, 1, "one_message")):
        acfg = b2048.ReinforceAgentConfig(gamma=0.99, learning_rate=1e-2, model_seed=3, **kw)
        mlp = b2048.MLPConfig(hidden_sizes=[64, 64], activation="ReLU", init_distribution="HeNormal")
        n_ep = 6000 + 1
        env = bd.make_sharded_env(n_ep, cfg, info, seed=seed + 3)
        agent = b2048.ReinforceAgent(env, mlp, acfg)
        th0 = agent._actor.theta.clone()
        ro = agent.rollout_many(env)
        ro.n_traj = n_ep
        calls.clear()
        upd = agent.update_from_rollout(ro, allreduce=counting_allreduce, precision=0, exchange=exchange)
        assert len(calls) == expect_calls, (name, calls)
        theta = agent._actor.theta.clone()
        ctheta = None if agent._critic is None else agent._critic.theta.clone()
        if info.rank == 0:
            env1 = b2048.Batched2048Env(n_ep, cfg, device=dev, seed=seed + 3)
            a1 = b2048.ReinforceAgent(env1, mlp, acfg)
            c0 = None if a1._critic is None else a1._critic.theta.clone()
            u1 = a1.update_from_rollout(a1.rollout_many(env1), precision=0)
            rel = float(((theta - th0) - (a1._actor.theta - th0)).norm() / (a1._actor.theta - th0).norm())
            msg = f"{name}: {len(calls)} all-reduce(s) per update {calls}; actor update rel err {rel:.2e}"
            assert rel < 1e-3, msg
            if ctheta is not None:
                relc = float(((ctheta - c0) - (a1._critic.theta - c0)).norm() / (a1._critic.theta - c0).norm())
                msg += f", critic {relc:.2e}"
                assert relc < 1e-3, msg
            print(msg)
    # ---- 5. the shape-generic tensor-core kernels (the reference's documented one-hot 272-256-128-64 network, actor + critic):
    #         the rollout (generic policy kernel on the live boards + step kernel) is sharding-invariant bit for bit, and the
    #         all-reduced tensor-core update equals the single-process one (fp16 loss scales are per rank: 1e-3)
    cfg_oh = b2048.Game2048EnvConfig(**dict(ENV, obs_mode="onehot", obs_log2_scale=1.0, max_steps=60))
    acfg = b2048.ReinforceAgentConfig(gamma=0.99, learning_rate=1e-2, critic_learning_rate=1e-2, model_seed=3, use_critic=True,
                                      baseline_mode="batch")
    mlp = b2048.MLPConfig(hidden_sizes=[256, 128, 64], activation="ReLU", init_distribution="HeNormal")
    env = bd.make_sharded_env(total, cfg_oh, info, seed=seed + 5)
    agent = b2048.ReinforceAgent(env, mlp, acfg)
    th0, c0 = agent._actor.theta.clone(), agent._critic.theta.clone()
    ro = agent.rollout_many(env, precision=1)
    lo, hi = bd.shard_range(total, info.rank, info.world_size)
    T_all = torch.tensor([ro.T], device=dev)
    dist.all_reduce(T_all, op=dist.ReduceOp.MAX)
    Tm = int(T_all.item())
    acts = torch.zeros((Tm, total), dtype=torch.int32, device=dev)
    lens = torch.zeros(total, dtype=torch.int32, device=dev)
    acts[: ro.T, lo:hi] = ro.actions.int() * (torch.arange(ro.T, device=dev).unsqueeze(1) < ro.length.unsqueeze(0))
    lens[lo:hi] = ro.length
    bd.allreduce_sum_(acts)
    bd.allreduce_sum_(lens)
    upd = bd.sharded_update(agent, ro, info)
    theta, ctheta = agent._actor.theta.clone(), agent._critic.theta.clone()
    if info.rank == 0:
        env1 = b2048.Batched2048Env(total, cfg_oh, device=dev, seed=seed + 5)
        a1 = b2048.ReinforceAgent(env1, mlp, acfg)
        r1 = a1.rollout_many(env1, precision=1)
        assert torch.equal(r1.length, lens), "generic rollout: episode lengths differ between sharded and single-GPU runs"
        live = torch.arange(r1.T, device=dev).unsqueeze(1) < r1.length.unsqueeze(0)
        assert torch.equal((r1.actions.int() * live)[:Tm], acts[: r1.T]), "generic rollout: actions differ"
        u1 = a1.update_from_rollout(r1)
        assert "shape-generic" in upd["precision"] and "shape-generic" in u1["precision"], (upd["precision"], u1["precision"])
        rel = float(((theta - th0) - (a1._actor.theta - th0)).norm() / (a1._actor.theta - th0).norm())
        relc = float(((ctheta - c0) - (a1._critic.theta - c0)).norm() / (a1._critic.theta - c0).norm())
        print(f"shape-generic tensor-core kernels, one-hot 272-256-128-64 actor + critic: rollout sharding invariance OK; "
              f"update rel err actor {rel:.2e} critic {relc:.2e}")
        assert rel < 1e-3 and relc < 1e-3
    # ---- 6. shared-trunk actor-critic (shared_trunk.py) under sharding: ONE flat gradient (trunk | policy head | value head)
    #         is exchanged after the value-view gradient has been merged locally; lambda scan + batch_norm baseline
    acfg = b2048.ReinforceAgentConfig(gamma=0.99, learning_rate=1e-2, model_seed=3, baseline_mode="batch_norm", optimizer="adam")
    mlp = b2048.MLPConfig(hidden_sizes=[256, 256], activation="ReLU", init_distribution="HeNormal")
    n_ep = 12000 + 1
    env = bd.make_sharded_env(n_ep, cfg, info, seed=seed + 6)
    agent = b2048.SharedTrunkActorCritic(env, mlp, acfg, value_coef=0.5, gae_lambda=0.9)
    th0 = agent._shared_net.theta.clone()
    ro = agent.rollout_many(env, precision="auto")
    ro.n_traj = n_ep
    calls.clear()
    agent.update_from_rollout(ro, allreduce=counting_allreduce)
    theta = agent._shared_net.theta.clone()
    if info.rank == 0:
        env1 = b2048.Batched2048Env(n_ep, cfg, device=dev, seed=seed + 6)
        a1 = b2048.SharedTrunkActorCritic(env1, mlp, acfg, value_coef=0.5, gae_lambda=0.9)
        a1.update_from_rollout(a1.rollout_many(env1, precision="auto"))
        rel = float(((theta - th0) - (a1._shared_net.theta - th0)).norm() / (a1._shared_net.theta - th0).norm())
        print(f"shared-trunk actor-critic (lambda 0.9, tensor-core update): {len(calls)} all-reduces per update {calls}; "
              f"update rel err {rel:.2e}")
        assert len(calls) == 2 and calls[1] == agent._shared_net.n_params and rel < 2e-3
    dist.barrier()
    if info.rank == 0:
        print("multi-GPU check OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
