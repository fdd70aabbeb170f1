"""Profiling target: one launch of each hot kernel at the 7B-head chunk shape (run under ncu, see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spatialthinker_b200 import _lib

h = int(sys.argv[1]) if len(sys.argv) > 1 else 3584
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
v = 151936
lib = _lib.load()
dev = torch.device("cuda:0")
st = _lib.stream_ptr(dev)
torch.manual_seed(0)
hid = torch.randn(rows, h, device=dev).to(torch.bfloat16)
w = (0.02 * torch.randn(v, h, device=dev)).to(torch.bfloat16)
labels = torch.randint(0, v, (rows,), device=dev)
logp = torch.empty(rows, device=dev)
dlogp = torch.randn(rows, device=dev) / rows
dh = torch.empty(rows, h, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(v, h, device=dev, dtype=torch.float32)
nbytes = lib.grpo_lmhead_bwd_workspace_bytes(rows, h, v)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
for _ in range(reps):
    _lib.check(lib.grpo_lmhead_logprob_fwd(hid.data_ptr(), w.data_ptr(), labels.data_ptr(), rows, h, v, 1.0,
                                           logp.data_ptr(), None, None, ws.data_ptr(), nbytes, st), "fwd")
    _lib.check(lib.grpo_lmhead_bwd(hid.data_ptr(), w.data_ptr(), labels.data_ptr(), dlogp.data_ptr(), None, rows, h, v,
                                   1.0, dh.data_ptr(), dw.data_ptr(), ws.data_ptr(), nbytes, st), "bwd")
torch.cuda.synchronize()
print("done")
