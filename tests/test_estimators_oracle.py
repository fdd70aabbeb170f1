"""Oracle restatements of the remaining estimators / whitening / value loss / KL reward shaping
against golden vectors produced by executing the reference (tests/golden/make_golden.py estimators) and, when
/root/reference is present, against the live reference. CPU only."""
import numpy as np
import pytest
import torch

from oracle import grpo_oracle as O

TAGS = ("s", "m", "l")


def t(a):
    return torch.from_numpy(np.asarray(a))


def uid_of(g, tag):
    return np.array([str(u) for u in g[f"{tag}_uid"]], dtype=object)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("rname", ["sparse", "dense"])
def test_outcome_estimators_golden(golden, tag, rname):
    g = golden("estimators")
    rew, mask = t(g[f"{tag}_{rname}"]), t(g[f"{tag}_mask"])
    adv, ret = O.compute_rloo_outcome_advantage(rew.clone(), mask, uid_of(g, tag))
    assert adv is ret
    np.testing.assert_array_equal(adv.numpy(), g[f"{tag}_{rname}_rloo"])
    adv, ret = O.compute_remax_outcome_advantage(rew.clone(), t(g[f"{tag}_baselines"]), mask)
    assert adv is ret
    np.testing.assert_array_equal(adv.numpy(), g[f"{tag}_{rname}_remax"])


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("rname", ["sparse", "dense"])
def test_recurrent_estimators_golden(golden, tag, rname):
    g = golden("estimators")
    rew, mask, values = t(g[f"{tag}_{rname}"]), t(g[f"{tag}_mask"]), t(g[f"{tag}_values"])
    for gamma in (1.0, 0.97):
        adv, ret = O.compute_reinforce_plus_plus_outcome_advantage(rew.clone(), mask, gamma)
        np.testing.assert_array_equal(ret.numpy(), g[f"{tag}_{rname}_rpp_ret_{gamma}"])
        np.testing.assert_array_equal(adv.numpy(), g[f"{tag}_{rname}_rpp_adv_{gamma}"])
    for gamma, lam in ((1.0, 1.0), (0.99, 0.95)):
        adv, ret = O.compute_gae_advantage_return(rew.clone(), values, mask, gamma, lam)
        np.testing.assert_array_equal(ret.numpy(), g[f"{tag}_{rname}_gae_ret_{gamma}_{lam}"])
        np.testing.assert_array_equal(adv.numpy(), g[f"{tag}_{rname}_gae_adv_{gamma}_{lam}"])


@pytest.mark.parametrize("tag", TAGS)
def test_whiten_value_loss_kl_rewards_golden(golden, tag):
    g = golden("estimators")
    mask, values = t(g[f"{tag}_mask"]), t(g[f"{tag}_values"])
    np.testing.assert_array_equal(
        np.array([float(O.masked_var(values, mask)), float(O.masked_var(values, mask, unbiased=False))], dtype=np.float32),
        g[f"{tag}_var"])
    np.testing.assert_array_equal(O.masked_whiten(values, mask).numpy(), g[f"{tag}_whiten"])
    vp = t(g[f"{tag}_vpreds"]).requires_grad_(True)
    loss, frac = O.compute_value_loss(vp, t(g[f"{tag}_returns"]), values, mask, 0.5)
    loss.backward()
    np.testing.assert_array_equal(np.array([float(loss.detach()), float(frac)], dtype=np.float32), g[f"{tag}_vf"])
    np.testing.assert_array_equal(vp.grad.numpy(), g[f"{tag}_vf_grad"])
    old, ref, sparse = t(g[f"{tag}_old"]), t(g[f"{tag}_ref"]), t(g[f"{tag}_sparse"])
    for mode in O.KL_MODES:
        rewards, cur = O.kl_penalty_rewards(sparse, old, ref, mask, 0.05, mode)
        np.testing.assert_array_equal(rewards.numpy(), g[f"{tag}_klrew_{mode}"])
        assert np.float32(cur) == g[f"{tag}_klcur_{mode}"][0]
    rewards, cur = O.kl_penalty_rewards(sparse, None, None, mask, 0.05)
    assert torch.equal(rewards, sparse) and cur == 0.0
    np.testing.assert_array_equal(O.compute_rewards(sparse, old, ref, 0.05).numpy(), g[f"{tag}_compute_rewards"])


def test_degenerate_masks(golden):
    g = golden("estimators")
    x = torch.tensor([[1.0, 2.0, 4.0]])
    assert np.float32(O.masked_var(x, torch.tensor([[0, 1, 0]]))) == g["var_one"][0]
    assert float(O.masked_var(x, torch.zeros(1, 3))) == 0.0
    with pytest.raises(AssertionError):
        O.compute_rloo_outcome_advantage(torch.ones(3, 2), torch.ones(3, 2, dtype=torch.int64),
                                         np.array(["a", "a", "b"], dtype=object))


def test_kl_controllers_golden(golden):
    import spatialthinker_b200.core_algos as ca  # host-only arithmetic, no CUDA involved

    ctl = ca.AdaptiveKLController(init_kl_coef=0.01, target_kl=0.1, horizon=1000.0)
    trace = []
    for cur in (0.05, 0.2, 0.11, 0.0):
        ctl.update(current_kl=cur, n_steps=128)
        trace.append(ctl.kl_coef)
    np.testing.assert_array_equal(np.array(trace), golden("estimators")["adaptive_kl_trace"])
    fixed = ca.FixedKLController(0.02)
    fixed.update(1.0, 10)
    assert fixed.kl_coef == 0.02

    class Cfg:
        kl_type, kl_coef, kl_target, kl_horizon = "adaptive", 0.01, 0.1, 0.0

    with pytest.raises(AssertionError):
        ca.get_kl_controller(Cfg)
    Cfg.kl_type = "nope"
    with pytest.raises(ValueError):
        ca.get_kl_controller(Cfg)
    Cfg.kl_type = "fixed"
    assert isinstance(ca.get_kl_controller(Cfg), ca.FixedKLController)


def test_oracle_vs_live_reference(reference_modules):
    VF, ca = reference_modules
    g = torch.Generator().manual_seed(77)
    bsz, n, tl = 40, 8, 61
    lens = torch.randint(1, tl + 1, (bsz,), generator=g)
    mask = (torch.arange(tl)[None] < lens[:, None]).long()
    rew = torch.randn(bsz, tl, generator=g) * mask
    values = torch.randn(bsz, tl, generator=g)
    uid = np.repeat(np.array([f"p{i}" for i in range(bsz // n)], dtype=object), n)[torch.randperm(bsz, generator=g).numpy()]
    base = torch.rand(bsz, generator=g)
    pairs = [
        (O.compute_rloo_outcome_advantage(rew.clone(), mask, uid), ca.compute_rloo_outcome_advantage(rew.clone(), mask, uid)),
        (O.compute_remax_outcome_advantage(rew.clone(), base, mask), ca.compute_remax_outcome_advantage(rew.clone(), base, mask)),
        (O.compute_reinforce_plus_plus_outcome_advantage(rew.clone(), mask, 0.9),
         ca.compute_reinforce_plus_plus_outcome_advantage(rew.clone(), mask, 0.9)),
        (O.compute_gae_advantage_return(rew.clone(), values, mask, 0.98, 0.9),
         ca.compute_gae_advantage_return(rew.clone(), values, mask, 0.98, 0.9)),
        (O.compute_value_loss(values + 0.3, rew, values, mask, 0.2), ca.compute_value_loss(values + 0.3, rew, values, mask, 0.2)),
    ]
    for mine, theirs in pairs:
        for a, b in zip(mine, theirs):
            assert torch.equal(torch.as_tensor(a), torch.as_tensor(b))
    assert torch.equal(O.masked_whiten(values, mask), VF.masked_whiten(values, mask))
    assert torch.equal(O.masked_var(values, mask), VF.masked_var(values, mask))
