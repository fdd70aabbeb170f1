"""Per-phase timing of the chunk pipeline (library CUDA-event profiler) for the knobs given in the environment."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spatialthinker_b200 import _lib

h = int(sys.argv[1]) if len(sys.argv) > 1 else 3584
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
v = 151936
lib = _lib.load()
dev = torch.device("cuda:0")
st = _lib.stream_ptr(dev)
torch.manual_seed(0)
hid = torch.randn(rows, h, device=dev).to(torch.bfloat16)
w = (0.02 * torch.randn(v, h, device=dev)).to(torch.bfloat16)
labels = torch.randint(0, v, (rows,), device=dev)
dlogp = torch.randn(rows, device=dev) / rows
use_dent = os.environ.get("PHASE_DENT", "0") == "1"  # forces the stash -> dlogits transform path
dent = torch.zeros(rows, device=dev) if use_dent else None
dh = torch.empty(rows, h, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(v, h, device=dev, dtype=torch.float32)
nbytes = lib.grpo_lmhead_bwd_workspace_bytes(rows, h, v)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)


def bwd():
    _lib.check(lib.grpo_lmhead_bwd(hid.data_ptr(), w.data_ptr(), labels.data_ptr(), dlogp.data_ptr(), dent.data_ptr() if use_dent else None, rows, h, v,
                                   1.0, dh.data_ptr(), dw.data_ptr(), ws.data_ptr(), nbytes, st), "bwd")


for _ in range(2):
    bwd()
torch.cuda.synchronize()
lib.grpo_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    bwd()
e1.record()
torch.cuda.synchronize()
ms = (ctypes.c_double * 6)()
cnt = (ctypes.c_longlong * 6)()
lib.grpo_profile_read(ms, cnt, 1)
lib.grpo_profile_enable(0)
tot = e0.elapsed_time(e1) / iters
unit = 2.0 * rows * h * v
knobs = {k.replace("GRPO_", ""): os.environ[k] for k in sorted(os.environ) if k.startswith(("GRPO_", "PHASE_"))}
parts = "  ".join(f"{n}={ms[i] / max(cnt[i], 1):.3f}" for i, n in enumerate(_lib.PHASE_NAMES))
print(f"{knobs} total={tot:.3f} ms ({3 * unit / tot / 1e9:.0f} TF alg) | {parts}")
