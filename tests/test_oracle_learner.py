"""Pins the NumPy learner oracle (oracle/learner.py) against outputs of the live reference
(tests/golden/mlp.npz, update_*.npz).  Tolerance: 1e-3 relative, the north_star's fp32 bar
(observed differences are ~1e-6: only float32 summation order differs)."""
import os

import numpy as np
import pytest

from oracle import learner
from helpers import GOLDEN, UPDATE_TAGS, load_update_fixture, rank_weights, rel_err

TOL = 1e-3


@pytest.mark.parametrize("tag,obs_mode,act", [("default", "log2", "ReLU"), ("onehot", "onehot", "ReLU"),
                                              ("sigmoid", "log2", "Sigmoid")])
def test_forward_and_probs(tag, obs_mode, act):
    g = np.load(os.path.join(GOLDEN, "mlp.npz"))
    L = int(g[f"{tag}/n_layers"])
    params = {"W": [g[f"{tag}/W{i}"] for i in range(L)], "b": [g[f"{tag}/b{i}"] for i in range(L)]}
    X = learner.encode(g[f"{tag}/boards"], obs_mode, float(g[f"{tag}/obs_scale"]))
    logits, _, _ = learner.forward(params, X, act)
    assert rel_err(logits, g[f"{tag}/logits"]) < TOL
    p = learner.probs_from_logits(logits, g[f"{tag}/masks"])
    assert np.abs(p - g[f"{tag}/probs"]).max() < TOL
    m = np.stack([(g[f"{tag}/masks"] >> a) & 1 for a in range(4)], 1)
    assert (np.argmax(p * m, 1) == g[f"{tag}/greedy"]).all()


def test_returns_and_advantages():
    g = np.load(os.path.join(GOLDEN, "mlp.npz"))
    lens = g["ret/lens"]; rew = g["ret/rewards"]; w = g["ret/weights"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    for gamma in (0.99, 1.0, 0.5):
        rets = [learner.returns(rew[offs[i]:offs[i + 1]], gamma) for i in range(len(lens))]
        assert (np.concatenate(rets) == g[f"ret/{gamma}/off/returns"]).all()
        for mode in ("off", "each", "batch", "batch_norm"):
            adv = np.concatenate(learner.advantages(rets, w, mode))
            ref = g[f"ret/{gamma}/{mode}/adv"]
            assert np.abs(adv - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("tag", UPDATE_TAGS)
def test_update_batch(tag):
    meta, actor0, critic0, updates = load_update_fixture(tag)
    a = meta["agent"]
    L = learner.Learner(actor0, critic0, activation=meta["mlp"]["activation"], obs_mode=meta["env"]["obs_mode"],
                        obs_scale=meta["env"].get("obs_log2_scale", 1.0), gamma=a["gamma"], lr=a["learning_rate"],
                        baseline=a["baseline_mode"], optimizer=a["optimizer"], use_critic=a.get("use_critic", False),
                        critic_lr=a.get("critic_learning_rate", 1e-3), max_grad_norm=a.get("max_grad_norm", 1.0),
                        critic_loss=a.get("critic_loss_type", "mse"), huber_delta=a.get("huber_delta", 1.0))
    prev_actor = actor0
    for u in updates:
        w = rank_weights(u["total_reward"], a.get("reward_rank_weights"))
        out = L.update(u["episodes"], w)
        assert rel_err(np.concatenate(out["adv"]), u["adv"]) < TOL
        assert abs(out["actor_grad_norm"] - u["grad_norms"][0]) < TOL * u["grad_norms"][0]
        if critic0 is not None:
            assert abs(out["critic_grad_norm"] - u["grad_norms"][1]) < TOL * u["grad_norms"][1]
        for l in range(len(actor0["W"])):
            # compare the parameter *change* (the update), not the parameters, so the bar is meaningful
            dref = u["actor"]["W"][l] - prev_actor["W"][l]
            dgot = L.actor["W"][l] - prev_actor["W"][l]
            assert rel_err(dgot, dref) < 2e-3, (tag, l)
            assert rel_err(L.actor["W"][l], u["actor"]["W"][l]) < 1e-5
        if critic0 is not None:
            for l in range(len(critic0["W"])):
                assert rel_err(L.critic["W"][l], u["critic"]["W"][l]) < 1e-5
        prev_actor = {"W": [x.copy() for x in u["actor"]["W"]], "b": [x.copy() for x in u["actor"]["b"]]}
        # follow the reference's parameters so errors do not compound across updates
        L.actor = {"W": [x.copy() for x in u["actor"]["W"]], "b": [x.copy() for x in u["actor"]["b"]]}
        if critic0 is not None:
            L.critic = {"W": [x.copy() for x in u["critic"]["W"]], "b": [x.copy() for x in u["critic"]["b"]]}
