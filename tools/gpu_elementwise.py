"""HBM roofline of the elementwise / reduction kernels ("HBM only for the elementwise/advantage kernels").

    python tools/gpu_elementwise.py [--seqs 4096] [--tlen 1024] [--out gpurun_out/elementwise.json]

Each entry point's launch sequence is captured ONCE into a CUDA graph (these kernels run for 10-100 us: timed call by
call, the host's own launch overhead would be what is measured) and the replay is timed with CUDA events on the launch
stream - 3 warm-ups, 20 replays, median - on config C3's batch (4096 sequences x 1024 response slots = 4.19 M tokens).
The working sets are 17-100 MB, smaller than the 126 MB L2, so a 256 MB buffer is written between replays. ALGORITHMIC
bytes = each input read once + each output written once in its stored dtype (int64 masks are 8 B); kernels that make
several passes (mask sum + loss, the three whitening passes) are still charged one read. The denominator is `hbm_gbs`
of MEASURED_PEAKS.json (read + write bytes of a device copy).
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import spatialthinker_b200 as st  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=4096)
    ap.add_argument("--tlen", type=int, default=1024)
    ap.add_argument("--logit-rows", type=int, default=8192)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    st.load_library()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    b, t = args.seqs, args.tlen
    n = b * t
    g = torch.Generator(device=dev).manual_seed(0)
    lens = (1 + torch.floor(torch.rand(b, generator=g, device=dev) * t)).clamp(max=t).long()
    mask = (torch.arange(t, device=dev)[None] < lens[:, None]).long()
    rewards = torch.zeros(b, t, device=dev)
    rewards[torch.arange(b, device=dev), lens - 1] = torch.rand(b, generator=g, device=dev)
    dense = torch.randn(b, t, generator=g, device=dev) * mask
    values = torch.randn(b, t, generator=g, device=dev)
    baselines = torch.rand(b, generator=g, device=dev)
    uid = np.repeat(np.arange(b // 8), 8)[np.random.default_rng(0).permutation(b)].astype(str).astype(object)
    logp = -3.0 + 0.1 * torch.randn(b, t, generator=g, device=dev)
    old = logp + 0.1 * torch.randn(b, t, generator=g, device=dev)
    ref = logp + 0.1 * torch.randn(b, t, generator=g, device=dev)
    adv = torch.randn(b, 1, generator=g, device=dev).expand(b, t).contiguous()
    vocab, lrows = 151936, args.logit_rows
    logits = torch.randn(lrows, vocab, generator=g, device=dev).to(torch.bfloat16)
    labels = torch.randint(0, vocab, (lrows,), generator=g, device=dev)
    hid = torch.randn(n // 8, 3584, generator=g, device=dev).to(torch.bfloat16)  # 3.76 GB of token rows
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lp_req = logp.clone().requires_grad_(True)
    gidx, inv, cnt = st.fused.compact_index(mask.view(-1)[: n // 8])
    m_valid = int(cnt.item())

    from spatialthinker_b200 import _lib
    from spatialthinker_b200.core_algos import group_csr

    lib = _lib.load()
    order, offsets = group_csr(uid)
    order_d, offsets_d = torch.from_numpy(order).to(dev), torch.from_numpy(offsets).to(dev)
    n_groups = offsets.size - 1
    adv_out = torch.empty(b, t, device=dev)
    seq = torch.empty(2 * b, device=dev)
    dlogits = torch.empty_like(logits)
    lse = torch.empty(lrows, device=dev)
    lp_rows = torch.empty(lrows, device=dev)
    g_rows = torch.randn(lrows, generator=g, device=dev)

    def sp():
        return _lib.stream_ptr(dev)

    def grpo_adv():  # the host half of compute_grpo_outcome_advantage (uid strings -> CSR) is not device work
        _lib.check(lib.grpo_advantage(rewards.data_ptr(), mask.data_ptr(), 1, order_d.data_ptr(), offsets_d.data_ptr(), b, t,
                                      n_groups, 1e-6, adv_out.data_ptr(), seq.data_ptr(), sp()), "adv")

    def rloo_adv():
        _lib.check(lib.grpo_rloo_advantage(rewards.data_ptr(), mask.data_ptr(), 1, order_d.data_ptr(), offsets_d.data_ptr(), b,
                                           t, n_groups, adv_out.data_ptr(), seq.data_ptr(), sp()), "rloo")

    def logits_fwd():
        _lib.check(lib.grpo_logprob_from_logits(logits.data_ptr(), 1, labels.data_ptr(), lrows, vocab, vocab,
                                                lp_rows.data_ptr(), None, lse.data_ptr(), sp()), "lfl")

    def logits_bwd():
        _lib.check(lib.grpo_logprob_from_logits_bwd(logits.data_ptr(), 1, labels.data_ptr(), lse.data_ptr(), g_rows.data_ptr(),
                                                    None, None, lrows, vocab, vocab, dlogits.data_ptr(), vocab, sp()), "lflb")

    F, I = 4, 8  # fp32 / int64 bytes
    cases = [
        # name, callable, algorithmic bytes
        ("grpo_advantage", grpo_adv, n * (F + I + F)),
        ("rloo_advantage", rloo_adv, n * (F + I + F)),
        ("remax_advantage", lambda: st.compute_remax_outcome_advantage(rewards, baselines, mask), n * (F + I + F)),
        ("reinforce_pp (scan + whiten)", lambda: st.compute_reinforce_plus_plus_outcome_advantage(dense, mask, 0.99),
         n * (F + I + 2 * F)),
        ("gae (scan + whiten)", lambda: st.compute_gae_advantage_return(dense, values, mask, 0.99, 0.95), n * (2 * F + I + 2 * F)),
        ("masked_whiten", lambda: st.masked_whiten(values, mask), n * (F + I + F)),
        ("masked_mean", lambda: st.masked_mean(values, mask), n * (F + I)),
        ("compute_kl low_var_kl (fwd)", lambda: st.compute_kl(logp, ref, "low_var_kl"), n * 3 * F),
        ("kl_penalty_rewards", lambda: st.ray_trainer.kl_penalty_rewards(rewards, old, ref, mask, 0.01, "low_var_kl"),
         n * (3 * F + I + F)),
        ("compute_policy_loss (fwd + dlogp)", lambda: st.compute_policy_loss(old, lp_req, adv, mask, 0.2, 0.3, 3.0),
         n * (3 * F + I + F)),
        ("compute_value_loss (fwd)", lambda: st.compute_value_loss(values, dense, logp, mask, 0.5), n * (3 * F + I)),
        ("log_probs_from_logits bf16 (fwd)", logits_fwd, lrows * (vocab * 2 + 8 + 8)),
        ("log_probs_from_logits bf16 (bwd)", logits_bwd, lrows * vocab * 2 * 2),
        ("entropy_from_logits bf16", lambda: st.entropy_from_logits(logits), lrows * (vocab * 2 + 4)),
        ("compact_index", lambda: st.fused.compact_index(mask), n * (I + 4 + 4)),
        ("gather_rows (7 KB token rows)", lambda: st.fused.gather_rows(hid, gidx, m_valid), 2 * m_valid * 7168),
        ("scatter_rows (7 KB token rows)", lambda: st.fused.scatter_rows(hid[:m_valid], inv), (m_valid + n // 8) * 7168),
    ]
    rows = []
    side = torch.cuda.Stream(device=dev)
    for name, fn, nbytes in cases:
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        ts = []
        for _ in range(20):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / hbm, 3)})
        print(f"{name:42s} {ms:9.4f} ms  {nbytes / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {gbs / hbm:6.1%}", flush=True)
        del graph
    out = {"batch": {"sequences": b, "response_len": t, "tokens": n, "logit_rows": lrows, "vocab": vocab},
           "hbm_peak_gbs": hbm, "timing": "CUDA-graph replay of each entry point's launch sequence, CUDA events, median of 20, 256 MB L2 "
           "flush between replays", "rows": rows}
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
