"""GPU bring-up checks, one sub-process per check so that a trap in one kernel does not hide the others.

    python tools/gpu_check.py            # run everything, print a summary table
    python tools/gpu_check.py gemm 0 1 2 # one check in-process (a_mn, b_mn, cta_group)
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _gemm(a_mn, b_mn, cta, m, n, k, accumulate=0):
    import torch
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    b = torch.randn(n, k, device=dev).to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    a_in = a.t().contiguous() if a_mn else a
    b_in = b.t().contiguous() if b_mn else b
    c = torch.full((m, n), 7.0, device=dev, dtype=torch.float32) if accumulate else torch.empty(m, n, device=dev)
    if accumulate:
        ref = ref + 7.0
    rc = lib.grpo_debug_gemm(a_in.data_ptr(), b_in.data_ptr(), c.data_ptr(), m, n, k, a_mn, b_mn, cta, accumulate,
                             _lib.stream_ptr(dev))
    _lib.check(rc, "grpo_debug_gemm")
    torch.cuda.synchronize()
    err = (c - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"gemm a_mn={a_mn} b_mn={b_mn} cta={cta} {m}x{n}x{k} acc={accumulate}: max_abs_err={err:.4e} (ref max {scale:.2f})")
    if not (err <= 2e-3 * max(scale, 1.0)):
        bad = ((c - ref).abs() > 2e-3 * max(scale, 1.0)).nonzero()
        print("  first bad idx:", bad[:8].tolist(), " n_bad:", bad.shape[0])
        sys.exit(1)


def _fwd(cta, rows, h, v, temp=1.0, peaked=False):
    import torch
    from spatialthinker_b200 import _lib

    os.environ["GRPO_CTA_GROUP"] = str(cta)
    lib = _lib.load()
    torch.manual_seed(2)
    dev = torch.device("cuda:0")
    hid = torch.randn(rows, h, device=dev).to(torch.bfloat16)
    w = ((0.16 if peaked else 0.02) * torch.randn(v, h, device=dev)).to(torch.bfloat16)
    labels = torch.randint(0, v, (rows,), device=dev)
    z = (hid.float() @ w.float().t()) / temp
    lse_ref = torch.logsumexp(z, -1)
    logp_ref = z.gather(1, labels[:, None]).squeeze(1) - lse_ref
    ent_ref = lse_ref - (torch.softmax(z, -1) * z).sum(-1)
    logp = torch.empty(rows, device=dev)
    ent = torch.empty(rows, device=dev)
    lse = torch.empty(rows, device=dev)
    nbytes = lib.grpo_lmhead_fwd_workspace_bytes(rows, h, v)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = lib.grpo_lmhead_logprob_fwd(hid.data_ptr(), w.data_ptr(), labels.data_ptr(), rows, h, v, temp, logp.data_ptr(),
                                     ent.data_ptr(), lse.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr(dev))
    _lib.check(rc, "grpo_lmhead_logprob_fwd")
    torch.cuda.synchronize()
    e1 = (logp - logp_ref).abs().max().item()
    e2 = (ent - ent_ref).abs().max().item()
    e3 = (lse - lse_ref).abs().max().item()
    print(f"fwd cta={cta} rows={rows} h={h} v={v} T={temp} peaked={peaked}: logp_err={e1:.3e} ent_err={e2:.3e} lse_err={e3:.3e}")
    if not (e1 < 2e-3 and e2 < 2e-3 and e3 < 2e-3):
        sys.exit(1)


def _bwd(cta, rows, h, v, temp=1.0, peaked=False, with_ent=False):
    import torch
    from spatialthinker_b200 import _lib

    os.environ["GRPO_CTA_GROUP"] = str(cta)
    lib = _lib.load()
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    hid = torch.randn(rows, h, device=dev).to(torch.bfloat16)
    w = ((0.16 if peaked else 0.02) * torch.randn(v, h, device=dev)).to(torch.bfloat16)
    labels = torch.randint(0, v, (rows,), device=dev)
    dlogp = torch.randn(rows, device=dev) / rows
    dlogp[::7] = 0.0
    dent = (torch.randn(rows, device=dev) / rows) if with_ent else None
    hf = hid.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    z = (hf @ wf.t()) / temp
    lse_ref = torch.logsumexp(z, -1)
    logp_ref = z.gather(1, labels[:, None]).squeeze(1) - lse_ref
    loss = (logp_ref * dlogp).sum()
    if with_ent:
        ent_ref = lse_ref - (torch.softmax(z, -1) * z).sum(-1)
        loss = loss + (ent_ref * dent).sum()
    loss.backward()
    dh = torch.empty(rows, h, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(v, h, device=dev, dtype=torch.float32)
    nbytes = lib.grpo_lmhead_bwd_workspace_bytes(rows, h, v)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = lib.grpo_lmhead_bwd(hid.data_ptr(), w.data_ptr(), labels.data_ptr(), dlogp.data_ptr(),
                             dent.data_ptr() if with_ent else None, rows, h, v, temp, dh.data_ptr(), dw.data_ptr(),
                             ws.data_ptr(), nbytes, _lib.stream_ptr(dev))
    _lib.check(rc, "grpo_lmhead_bwd")
    torch.cuda.synchronize()
    r1 = ((dh.float() - hf.grad).norm() / hf.grad.norm()).item()
    r2 = ((dw - wf.grad).norm() / wf.grad.norm()).item()
    print(f"bwd cta={cta} rows={rows} h={h} v={v} T={temp} peaked={peaked} ent={with_ent}: dH rel={r1:.3e} dW rel={r2:.3e}")
    if not (r1 < 1e-2 and r2 < 1e-2):
        sys.exit(1)


CHECKS = []
for cta in (1, 2):
    for a_mn, b_mn in ((0, 0), (0, 1), (1, 1), (1, 0)):
        CHECKS.append(["gemm", a_mn, b_mn, cta, 392, 520, 224, 0])
    CHECKS.append(["gemm", 0, 0, cta, 2048, 4096, 1024, 0])
    CHECKS.append(["gemm", 1, 1, cta, 2048, 4096, 1024, 1])
    CHECKS.append(["gemm", 0, 1, cta, 4096, 2048, 1088, 0])
    CHECKS.append(["fwd", cta, 300, 256, 2048 + 128])
    CHECKS.append(["fwd", cta, 1000, 512, 151936, 0.7, 1])
    CHECKS.append(["bwd", cta, 300, 256, 2048 + 128])
    CHECKS.append(["bwd", cta, 1000, 512, 151936, 0.7, 1, 1])


def main():
    if len(sys.argv) > 1:
        kind = sys.argv[1]
        args = [float(x) if "." in x else int(x) for x in sys.argv[2:]]
        {"gemm": _gemm, "fwd": _fwd, "bwd": _bwd}[kind](*args)
        return
    results = []
    for chk in CHECKS:
        cmd = [sys.executable, os.path.abspath(__file__)] + [str(x) for x in chk]
        t0 = time.time()
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
            ok = p.returncode == 0
            out = (p.stdout + p.stderr).strip().splitlines()
            tail = out[-6:] if not ok else out[-1:]
        except subprocess.TimeoutExpired:
            ok, tail = False, ["TIMEOUT"]
        results.append((ok, chk))
        print(("PASS " if ok else "FAIL ") + " ".join(str(x) for x in chk) + f"  [{time.time() - t0:.1f}s]")
        for line in tail:
            print("    " + line)
        sys.stdout.flush()
    n_ok = sum(1 for ok, _ in results if ok)
    print(f"{n_ok}/{len(results)} checks passed")
    sys.exit(0 if n_ok == len(results) else 1)


if __name__ == "__main__":
    main()
