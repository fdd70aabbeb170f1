"""Kernel-level timing on one B200 (CUDA events, warm-up, inputs far larger than L2).

    python tools/gpu_perf.py [hidden_dim] [rows]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spatialthinker_b200 import _lib


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return min(ts), sum(ts) / len(ts)


def main():
    h = int(sys.argv[1]) if len(sys.argv) > 1 else 3584
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
    v = 151936
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = _lib.stream_ptr(dev)
    torch.manual_seed(0)
    hid = torch.randn(rows, h, device=dev).to(torch.bfloat16)
    w = (0.02 * torch.randn(v, h, device=dev)).to(torch.bfloat16)
    labels = torch.randint(0, v, (rows,), device=dev)
    logp = torch.empty(rows, device=dev)
    ent = torch.empty(rows, device=dev)
    dlogp = torch.randn(rows, device=dev) / rows
    dh = torch.empty(rows, h, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(v, h, device=dev, dtype=torch.float32)
    nbytes = lib.grpo_lmhead_bwd_workspace_bytes(rows, h, v)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    unit = 2.0 * rows * h * v  # one GEMM unit of algorithmic flops

    def fwd(want_ent):
        _lib.check(lib.grpo_lmhead_logprob_fwd(hid.data_ptr(), w.data_ptr(), labels.data_ptr(), rows, h, v, 1.0,
                                               logp.data_ptr(), ent.data_ptr() if want_ent else None, None,
                                               ws.data_ptr(), nbytes, st), "fwd")

    def bwd():
        _lib.check(lib.grpo_lmhead_bwd(hid.data_ptr(), w.data_ptr(), labels.data_ptr(), dlogp.data_ptr(), None, rows, h,
                                       v, 1.0, dh.data_ptr(), dw.data_ptr(), ws.data_ptr(), nbytes, st), "bwd")

    g = torch.randn(rows, v, device=dev).to(torch.bfloat16)
    c2 = torch.empty(rows, h, device=dev)
    c3 = torch.zeros(v, h, device=dev)
    c1 = None

    def gemm2(cta):
        _lib.check(lib.grpo_debug_gemm(g.data_ptr(), w.data_ptr(), c2.data_ptr(), rows, h, v, 0, 1, cta, 0, st), "gemm2")

    def gemm3(cta):
        _lib.check(lib.grpo_debug_gemm(g.data_ptr(), hid.data_ptr(), c3.data_ptr(), v, h, rows, 1, 1, cta, 1, st), "gemm3")

    print(f"shape rows={rows} H={h} V={v}; one GEMM unit = {unit / 1e12:.2f} TFLOP")
    mn, av = timeit(lambda: fwd(False))
    print(f"fwd (no entropy)     : {mn:8.3f} ms min {av:8.3f} avg  -> {unit / mn / 1e9:7.1f} TFLOP/s")
    mn, av = timeit(lambda: fwd(True))
    print(f"fwd (+entropy)       : {mn:8.3f} ms min {av:8.3f} avg  -> {unit / mn / 1e9:7.1f} TFLOP/s")
    mn, av = timeit(bwd, iters=3, warm=1)
    print(f"fwd-stash+bwd (3 GEMM): {mn:8.3f} ms min {av:8.3f} avg  -> {3 * unit / mn / 1e9:7.1f} TFLOP/s algorithmic")
    for cta in (1, 2):
        mn, av = timeit(lambda: gemm2(cta), iters=3, warm=1)
        print(f"dH-shaped GEMM cta={cta} : {mn:8.3f} ms min {av:8.3f} avg  -> {unit / mn / 1e9:7.1f} TFLOP/s")
        mn, av = timeit(lambda: gemm3(cta), iters=3, warm=1)
        print(f"dW-shaped GEMM cta={cta} : {mn:8.3f} ms min {av:8.3f} avg  -> {unit / mn / 1e9:7.1f} TFLOP/s")
    a = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    b = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    mn, av = timeit(lambda: torch.matmul(a, b.t()), iters=10, warm=3)
    print(f"cuBLAS 8192^3 bf16   : {mn:8.3f} ms min -> {2 * 8192 ** 3 / mn / 1e9:7.1f} TFLOP/s (library yardstick)")
    cc = torch.empty(8192, 8192, device=dev)
    for cta in (1, 2):
        mn, av = timeit(lambda: _lib.check(lib.grpo_debug_gemm(a.data_ptr(), b.data_ptr(), cc.data_ptr(), 8192, 8192,
                                                               8192, 0, 0, cta, 0, st), "g"), iters=5, warm=2)
        print(f"ours 8192^3 cta={cta}    : {mn:8.3f} ms min -> {2 * 8192 ** 3 / mn / 1e9:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
