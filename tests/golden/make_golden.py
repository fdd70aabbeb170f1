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


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(4)
    kat_advantage()
    kat_policy_loss()
    kat_end_to_end()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
