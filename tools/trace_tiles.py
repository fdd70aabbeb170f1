"""Tile-boundary timeline of CTA 0 for the three GEMM kernels (needs the -DGRPO_TRACE build):

    nvcc <flags of spatialthinker_b200/build.py> -DGRPO_TRACE -o tools/_build/libgrpo_trace.so spatialthinker_b200/csrc/grpo_b200.cu
    GRPO_B200_LIB=tools/_build/libgrpo_trace.so python tools/trace_tiles.py [knob=value,...]

Prints, per tile, the SM-clock offsets (in cycles, relative to the tile's first MMA) of the pipeline events.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spatialthinker_b200 import _lib

h, rows, v = 3584, 18944, 151936
lib = _lib.load()
dev = torch.device("cuda:0")
st = _lib.stream_ptr(dev)
for kv in filter(None, (sys.argv[1] if len(sys.argv) > 1 else "").split(",")):
    k, val = kv.split("=")
    _lib.check(lib.grpo_set_option(k.encode(), int(val)), "set_option")
lib.grpo_set_option(b"clk_probe", 1)
torch.manual_seed(0)
hid = torch.randn(rows, h, device=dev).to(torch.bfloat16)
w = (0.02 * torch.randn(v, h, device=dev)).to(torch.bfloat16)
labels = torch.randint(0, v, (rows,), device=dev)
dlogp = torch.randn(rows, device=dev) / rows
dh = torch.empty(rows, h, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(v, h, device=dev, dtype=torch.float32)
nbytes = lib.grpo_lmhead_bwd_workspace_bytes(rows, h, v)
ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
for _ in range(4):
    _lib.check(lib.grpo_lmhead_bwd(hid.data_ptr(), w.data_ptr(), labels.data_ptr(), dlogp.data_ptr(), None, rows, h, v, 1.0,
                                   dh.data_ptr(), dw.data_ptr(), ws.data_ptr(), nbytes, st), "bwd")
torch.cuda.synchronize()
off = lib.grpo_debug_probe_offset(rows, h, v, 1)
pr = ws[off:off + 3 * 8192].view(torch.int64).cpu().view(3, 1024)
names = ["mma0 first", "mma1 first", "mma0 last", "mma1 last", "epi0 full", "epi1 full", "epi0 release", "epi1 release",
         "epi0 done", "epi1 done", "tma first"]
for ki, kname in enumerate(("logits", "dhidden", "dweight")):
    tr = pr[ki, 256:256 + 16 * 11].view(11, 16)
    c0, n0, c1, n1 = pr[ki, :4].tolist()
    print(f"== {kname}: {(c1 - c0) / max(n1 - n0, 1):.3f} GHz, kernel {1e-6 * (n1 - n0):.3f} ms")
    print("tile " + " ".join(f"{n:>12s}" for n in names) + "   | tile period")
    prev = None
    for t in range(8):
        base = int(tr[0, t])
        if base == 0:
            break
        row = " ".join(f"{int(tr[e, t]) - base:12d}" for e in range(11))
        print(f"{t:4d} {row}   | {'' if prev is None else base - prev}")
        prev = base
