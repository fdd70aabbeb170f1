"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN FUNCTIONS.

Run in the build container only (needs the read-only checkout at /root/reference; it cannot travel to the GPU box):

    python tests/golden/make_golden.py

What is imported from the reference: ``verl.utils.torch_functional`` (log_probs_from_logits, masked_mean) and
``verl.trainer.core_algos`` (compute_grpo_outcome_advantage, compute_policy_loss, compute_kl). The flash-attn branch is
disabled (its Triton kernel rejects CPU tensors) and the CPU fallback's sign is corrected to the training branch's
(log p, torch_functional.py:42 vs :64) - see SURVEY.md §0. The lm_head is ``F.linear`` in fp32 and the micro-batch
arithmetic is dp_actor.py:247-278 typed out here, because ``verl.workers.actor.dp_actor`` itself needs ray/tensordict,
which are not installed.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("GRPO_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import verl.trainer.core_algos as ca  # noqa: E402
import verl.utils.torch_functional as VF  # noqa: E402

VF.FLAH_ATTN_CROSS_ENTROPY_LOSS_AVAILABLE = False
OUT = os.path.dirname(os.path.abspath(__file__))
CLIP = (0.2, 0.3, 3.0)


def ref_logp(z, labels):
    return -VF.log_probs_from_logits(z, labels)


def kat_advantage():
    scores = torch.tensor([1, 0, 0.5, 0.2, 1, 1, 0, 0, 0.7, 0.7])
    uid = np.array(list("ababababcc"), dtype=object)
    lens = [4, 3, 2, 4, 1, 4, 4, 2, 3, 4]
    t = 4
    mask = torch.zeros(10, t, dtype=torch.int64)
    rewards = torch.zeros(10, t)
    for i, n in enumerate(lens):
        mask[i, :n] = 1
        rewards[i, n - 1] = scores[i]
    adv, ret = ca.compute_grpo_outcome_advantage(rewards.clone(), mask, uid)
    assert adv is ret
    out = {"a_rewards": rewards.numpy(), "a_mask": mask.numpy(), "a_uid": uid.astype(str), "a_adv": adv.numpy()}
    # seeded, larger, permuted groups, ragged masks, n = 8 and n = 16
    g = torch.Generator().manual_seed(11)
    for tag, bsz, n, t in (("b", 64, 8, 33), ("c", 96, 16, 20)):
        lens = torch.randint(1, t + 1, (bsz,), generator=g)
        mask = (torch.arange(t)[None] < lens[:, None]).long()
        rewards = torch.zeros(bsz, t)
        rewards[torch.arange(bsz), lens - 1] = torch.rand(bsz, generator=g)
        uid = np.repeat(np.array([f"u{i}" for i in range(bsz // n)], dtype=object), n)[torch.randperm(bsz, generator=g).numpy()]
        adv, _ = ca.compute_grpo_outcome_advantage(rewards.clone(), mask, uid)
        out.update({f"{tag}_rewards": rewards.numpy(), f"{tag}_mask": mask.numpy(), f"{tag}_uid": uid.astype(str),
                    f"{tag}_adv": adv.numpy()})
    # dense rewards on every token (sum over tokens is then a real reduction)
    rewards = torch.rand(16, 12, generator=g)
    mask = torch.ones(16, 12, dtype=torch.int64)
    uid = np.repeat(np.array(["p", "q"], dtype=object), 8)
    adv, _ = ca.compute_grpo_outcome_advantage(rewards.clone(), mask, uid)
    out.update({"d_rewards": rewards.numpy(), "d_mask": mask.numpy(), "d_uid": uid.astype(str), "d_adv": adv.numpy()})
    np.savez(os.path.join(OUT, "advantage.npz"), **out)


def kat_policy_loss():
    out = {}
    x = torch.tensor([[0, 0.1, 0.3, -0.3, 0.3, -0.3, 1.2, 1.2]], requires_grad=True)
    old = torch.zeros(1, 8)
    adv = torch.tensor([[1.0, 1, 1, 1, -1, -1, -1, 1]])
    mask = torch.tensor([[1, 1, 1, 1, 1, 1, 1, 0]])
    res = ca.compute_policy_loss(old, x, adv, mask, *CLIP)
    res[0].backward()
    out.update({"a_logp": x.detach().numpy(), "a_old": old.numpy(), "a_adv": adv.numpy(), "a_mask": mask.numpy(),
                "a_out": np.array([float(r) for r in res], dtype=np.float32), "a_grad": x.grad.numpy()})
    g = torch.Generator().manual_seed(5)
    bsz, t = 6, 50
    logp = (-2.0 * torch.rand(bsz, t, generator=g)).requires_grad_(True)
    old = logp.detach() + 0.3 * torch.randn(bsz, t, generator=g)
    old[0, :5] += 1.5
    old[1, :5] -= 1.5
    adv = torch.randn(bsz, 1, generator=g).expand(bsz, t).contiguous()
    lens = torch.randint(1, t + 1, (bsz,), generator=g)
    mask = (torch.arange(t)[None] < lens[:, None]).long()
    ref = logp.detach() + 0.2 * torch.randn(bsz, t, generator=g)
    ref[2, :3] -= 6.0  # drives low_var_kl into its clamp
    res = ca.compute_policy_loss(old, logp, adv, mask, *CLIP)
    kld = ca.compute_kl(logp, ref, "low_var_kl")
    kl = VF.masked_mean(kld, mask)
    total = (res[0] + 0.01 * kl) / 4.0  # dp_actor.py:271-277 with kl_coef 1e-2, GA = 4
    total.backward()
    ent = -VF.masked_mean(logp.detach(), mask)
    out.update({"b_logp": logp.detach().numpy(), "b_old": old.numpy(), "b_adv": adv.numpy(), "b_mask": mask.numpy(),
                "b_ref": ref.numpy(), "b_out": np.array([float(r) for r in res] + [float(kl), float(ent), float(total)],
                                                        dtype=np.float32),
                "b_grad": logp.grad.numpy()})
    # every KL mode with its gradient
    lp = torch.tensor([-1.0, -2.0, -0.5, -12.0, -0.3, -3.0])
    rf = torch.tensor([-1.2, -1.0, -0.5, -1.0, -2.9, -0.1])
    for mode in ("kl", "abs", "mse", "low_var_kl", "chi2"):
        a = lp.clone().requires_grad_(True)
        v = ca.compute_kl(a, rf, mode)
        v.sum().backward()
        out[f"kl_{mode}"] = v.detach().numpy()
        out[f"kl_{mode}_grad"] = a.grad.numpy()
    out["kl_logp"], out["kl_ref"] = lp.numpy(), rf.numpy()
    out["masked_mean_zero"] = np.array([float(VF.masked_mean(torch.ones(3, 4), torch.zeros(3, 4)))], dtype=np.float32)
    out["clip_bounds"] = np.array([np.log(1 - 0.2), np.log(1 + 0.3)], dtype=np.float64)
    np.savez(os.path.join(OUT, "policy_loss.npz"), **out)


def kat_end_to_end():
    """Small lm_head -> loss -> backward cases (kept tiny so the fixture is a few hundred KB)."""
    out = {}
    for tag, (bsz, t, h, v, sigma, temp) in {"flat": (4, 24, 128, 1024, 0.02, 1.0), "peaked": (4, 24, 128, 1024, 0.5, 0.7)}.items():
        g = torch.Generator().manual_seed(21)
        hidden = torch.randn(bsz, t, h, generator=g).to(torch.bfloat16)
        weight = (sigma * torch.randn(v, h, generator=g)).to(torch.bfloat16)
        labels = torch.randint(0, v, (bsz, t), generator=g)
        lens = torch.tensor([t, t - 5, 3, t])
        mask = (torch.arange(t)[None] < lens[:, None]).long()
        hf = hidden.float().requires_grad_(True)
        wf = weight.float().requires_grad_(True)
        z = F.linear(hf, wf) / temp
        logp = ref_logp(z, labels)
        old = logp.detach() + 0.1 * torch.randn(bsz, t, generator=g)
        old[0, :4] += 1.5
        old[1, :4] -= 1.5
        ref = logp.detach() + 0.1 * torch.randn(bsz, t, generator=g)
        rewards = torch.zeros(bsz, t)
        rewards[torch.arange(bsz), lens - 1] = torch.tensor([1.0, 0.2, 0.7, 0.0])
        uid = np.array(["x", "y", "x", "y"], dtype=object)
        adv, _ = ca.compute_grpo_outcome_advantage(rewards.clone(), mask, uid)
        ent_loss = -VF.masked_mean(logp, mask)
        pg, cfh, cfl, pkl = ca.compute_policy_loss(old, logp, adv, mask, *CLIP)
        kl = VF.masked_mean(ca.compute_kl(logp, ref, "low_var_kl"), mask)
        total = pg + kl * 0.01
        (total / 2.0).backward()
        out.update({
            f"{tag}_hidden": hidden.float().numpy(), f"{tag}_weight": weight.float().numpy(), f"{tag}_labels": labels.numpy(),
            f"{tag}_mask": mask.numpy(), f"{tag}_old": old.numpy(), f"{tag}_ref": ref.numpy(), f"{tag}_rewards": rewards.numpy(),
            f"{tag}_uid": uid.astype(str), f"{tag}_adv": adv.numpy(), f"{tag}_temp": np.array([temp], dtype=np.float32),
            f"{tag}_logp": logp.detach().numpy(),
            f"{tag}_entropy": (torch.logsumexp(z, -1) - (torch.softmax(z, -1) * z).sum(-1)).detach().numpy(),
            f"{tag}_scalars": np.array([float(total), float(cfh), float(cfl), float(ent_loss), float(pkl), float(kl), float(pg)],
                                       dtype=np.float32),
            f"{tag}_dhidden": hf.grad.numpy(), f"{tag}_dweight": wf.grad.numpy(),
        })
    np.savez_compressed(os.path.join(OUT, "end_to_end.npz"), **out)


def kat_estimators():
    """RLOO / ReMax / REINFORCE++ / GAE, masked_var / masked_whiten, value loss, KL reward shaping
    - outputs of the reference's own functions; apply_kl_penalty's arithmetic (ray_trainer.py:131-142) is typed out with
    the reference's compute_kl / masked_mean because ray_trainer.py itself imports ray."""
    out = {}
    g = torch.Generator().manual_seed(33)
    for tag, bsz, n, t in (("s", 12, 4, 7), ("m", 64, 8, 45), ("l", 48, 16, 100)):
        lens = torch.randint(1, t + 1, (bsz,), generator=g)
        lens[0] = t
        mask = (torch.arange(t)[None] < lens[:, None]).long()
        sparse = torch.zeros(bsz, t)
        sparse[torch.arange(bsz), lens - 1] = torch.rand(bsz, generator=g)
        dense = torch.randn(bsz, t, generator=g) * mask  # per-token rewards (after KL shaping every token carries one)
        values = torch.randn(bsz, t, generator=g)
        baselines = torch.rand(bsz, generator=g)
        uid = np.repeat(np.array([f"u{i}" for i in range(bsz // n)], dtype=object), n)[torch.randperm(bsz, generator=g).numpy()]
        out.update({f"{tag}_mask": mask.numpy(), f"{tag}_sparse": sparse.numpy(), f"{tag}_dense": dense.numpy(),
                    f"{tag}_values": values.numpy(), f"{tag}_baselines": baselines.numpy(), f"{tag}_uid": uid.astype(str)})
        for rname, rew in (("sparse", sparse), ("dense", dense)):
            a, r = ca.compute_rloo_outcome_advantage(rew.clone(), mask, uid)
            assert a is r
            out[f"{tag}_{rname}_rloo"] = a.numpy()
            a, r = ca.compute_remax_outcome_advantage(rew.clone(), baselines, mask)
            assert a is r
            out[f"{tag}_{rname}_remax"] = a.numpy()
            for gamma in (1.0, 0.97):
                a, r = ca.compute_reinforce_plus_plus_outcome_advantage(rew.clone(), mask, gamma)
                out[f"{tag}_{rname}_rpp_adv_{gamma}"] = a.numpy()
                out[f"{tag}_{rname}_rpp_ret_{gamma}"] = r.numpy()
            for gamma, lam in ((1.0, 1.0), (0.99, 0.95)):
                a, r = ca.compute_gae_advantage_return(rew.clone(), values, mask, gamma, lam)
                out[f"{tag}_{rname}_gae_adv_{gamma}_{lam}"] = a.numpy()
                out[f"{tag}_{rname}_gae_ret_{gamma}_{lam}"] = r.numpy()
        out[f"{tag}_var"] = np.array([float(VF.masked_var(values, mask)), float(VF.masked_var(values, mask, unbiased=False))],
                                     dtype=np.float32)
        out[f"{tag}_whiten"] = VF.masked_whiten(values, mask).numpy()
        # value loss with gradient
        vp = (values + 0.4 * torch.randn(bsz, t, generator=g)).requires_grad_(True)
        ret = values + 0.5 * torch.randn(bsz, t, generator=g)
        loss, frac = ca.compute_value_loss(vp, ret, values, mask, 0.5)
        loss.backward()
        out.update({f"{tag}_vpreds": vp.detach().numpy(), f"{tag}_returns": ret.numpy(),
                    f"{tag}_vf": np.array([float(loss), float(frac)], dtype=np.float32), f"{tag}_vf_grad": vp.grad.numpy()})
        # KL reward shaping, every estimator
        old = -2.0 * torch.rand(bsz, t, generator=g)
        ref = old + 0.3 * torch.randn(bsz, t, generator=g)
        ref[1, :2] -= 6.0
        out.update({f"{tag}_old": old.numpy(), f"{tag}_ref": ref.numpy()})
        for mode in ("kl", "abs", "mse", "low_var_kl", "chi2"):
            kld = ca.compute_kl(old, ref, kl_penalty=mode) * mask
            rewards = sparse - 0.05 * kld
            cur = torch.mean(VF.masked_mean(kld, mask=mask, dim=-1), dim=0).item()
            out[f"{tag}_klrew_{mode}"] = rewards.numpy()
            out[f"{tag}_klcur_{mode}"] = np.array([cur], dtype=np.float32)
        out[f"{tag}_compute_rewards"] = ca.compute_rewards(sparse, old, ref, 0.05).numpy()
    # degenerate masks for masked_var: one valid element -> biased value returned, none -> 0
    x = torch.tensor([[1.0, 2.0, 4.0]])
    out["var_one"] = np.array([float(VF.masked_var(x, torch.tensor([[0, 1, 0]])))], dtype=np.float32)
    # KL controllers
    ctl = ca.AdaptiveKLController(init_kl_coef=0.01, target_kl=0.1, horizon=1000.0)
    trace = []
    for cur in (0.05, 0.2, 0.11, 0.0):
        ctl.update(current_kl=cur, n_steps=128)
        trace.append(ctl.kl_coef)
    out["adaptive_kl_trace"] = np.array(trace, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "estimators.npz"), **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(4)
    todo = sys.argv[1:] or ["advantage", "policy_loss", "end_to_end", "estimators"]
    for name in todo:
        {"advantage": kat_advantage, "policy_loss": kat_policy_loss, "end_to_end": kat_end_to_end,
         "estimators": kat_estimators}[name]()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
