"""Interleaved A/B timing of knob settings in ONE process (the pipeline is power-capped: wall time drifts by several
percent between boxes and minutes, so variants are alternated round-robin and compared on medians).

    python tools/gpu_ab.py "l2_hints=1" "l2_hints=0" ["fwd_panel=37,l2_hints=0" ...]  [--dent] [--rounds 6] [--iters 10]
"""
import argparse
import ctypes
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spatialthinker_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("variants", nargs="+")
ap.add_argument("--hidden", type=int, default=3584)
ap.add_argument("--rows", type=int, default=9472)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--dent", action="store_true")
args = ap.parse_args()
h, rows, v = args.hidden, args.rows, 151936
lib = _lib.load()
dev = torch.device("cuda:0")
st = _lib.stream_ptr(dev)
torch.manual_seed(0)
hid = torch.randn(rows, h, device=dev).to(torch.bfloat16)
w = (0.02 * torch.randn(v, h, device=dev)).to(torch.bfloat16)
labels = torch.randint(0, v, (rows,), device=dev)
dlogp = torch.randn(rows, device=dev) / rows
dent = torch.zeros(rows, device=dev) if args.dent else None
dh = torch.empty(rows, h, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(v, h, device=dev, dtype=torch.float32)
nbytes = 0
for spec in args.variants:  # the workspace must fit the largest chunk setting among the variants
    for kv in filter(None, spec.split(",")):
        k, val = kv.split("=")
        if k in ("chunk_rows", "ksub"):
            lib.grpo_set_option(k.encode(), int(val))
    nbytes = max(nbytes, lib.grpo_lmhead_bwd_workspace_bytes(rows, h, v))
    lib.grpo_set_option(b"chunk_rows", 0)
    lib.grpo_set_option(b"ksub", 2)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
DEFAULTS = {"cta_group": 2, "fwd_panel": 4864, "sync_fwd": 28, "sync_dh": 8, "sync_dw": 8, "l2_hints": 0, "dh_m_fast": 0, "chunk_rows": 0, "ksub": 2, "wait_hint_ns": 10000000,
            "epi_mode": 7, "dw_tma": 1, "acc_lead": 2, "clk_probe": 1, "st_hint": 3, "dw_split": 1, "epi_share": 0, "dh_split": 0, "k_serp": 0}


def apply(spec):
    opts = dict(DEFAULTS)
    for kv in filter(None, spec.split(",")):
        k, val = kv.split("=")
        opts[k] = int(val)
    for k, val in opts.items():
        _lib.check(lib.grpo_set_option(k.encode(), val), "set_option")


def bwd():
    _lib.check(lib.grpo_lmhead_bwd(hid.data_ptr(), w.data_ptr(), labels.data_ptr(), dlogp.data_ptr(),
                                   dent.data_ptr() if dent is not None else None, rows, h, v, 1.0, dh.data_ptr(),
                                   dw.data_ptr(), ws.data_ptr(), nbytes, st), "bwd")


for _ in range(3):
    bwd()
torch.cuda.synchronize()
probe_off = lib.grpo_debug_probe_offset(rows, h, v, 1)
clk = {s: [[], [], []] for s in args.variants}  # GHz the three GEMMs really ran at (clock64 / globaltimer, last chunk)
span = {s: [[], [], []] for s in args.variants}  # (kernel span, spread of group starts, spread of group ends) in ms
res = {s: [] for s in args.variants}
phases = {s: [0.0] * 6 for s in args.variants}
lib.grpo_profile_enable(1)
for rnd in range(args.rounds):
    for spec in args.variants:
        apply(spec)
        bwd()
        torch.cuda.synchronize()
        lib.grpo_profile_read(None, None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            bwd()
        e1.record()
        torch.cuda.synchronize()
        res[spec].append(e0.elapsed_time(e1) / args.iters)
        probe_off = lib.grpo_debug_probe_offset(rows, h, v, 1)  # depends on the variant's chunk size
        pr = ws[probe_off:probe_off + 3 * 8192].view(torch.int64).cpu().view(3, 1024)
        for i in range(3):
            c0, n0, c1, n1 = pr[i, :4].tolist()
            if n1 > n0:
                clk[spec][i].append((c1 - c0) / (n1 - n0))
            g = pr[i, 8:8 + 2 * 74].view(74, 2)
            span[spec][i].append(((g[:, 1].max() - g[:, 0].min()).item() / 1e6, (g[:, 0].max() - g[:, 0].min()).item() / 1e6,
                                  (g[:, 1].max() - g[:, 1].min()).item() / 1e6))
        ms = (ctypes.c_double * 6)()
        cnt = (ctypes.c_longlong * 6)()
        lib.grpo_profile_read(ms, cnt, 1)
        for i in range(6):
            phases[spec][i] += ms[i] / max(cnt[i], 1) / args.rounds
unit = 2.0 * rows * h * v
for spec in args.variants:
    t = res[spec]
    med = statistics.median(t)
    ph = "  ".join(f"{n}={phases[spec][i]:.3f}" for i, n in enumerate(_lib.PHASE_NAMES))
    ghz = " ".join(f"{statistics.median(c):.3f}" if c else "-" for c in clk[spec])
    ph += f" | GHz fwd/dh/dw {ghz} | span/start-spread/end-spread ms " + " ".join(
        "/".join(f"{statistics.median(x[j] for x in sp):.3f}" for j in range(3)) if sp else "-" for sp in span[spec])
    print(f"[{spec or 'default':28s}] median {med:7.3f} ms  min {min(t):7.3f}  max {max(t):7.3f}  -> {3 * unit / med / 1e9:6.0f} TF alg | {ph}")
