"""Where does the time outside the GEMM phases go when bench.py drives DataParallelPPOActor.update_policy?
Host wall-clock per stage of one step (synchronised) and the GPU-side totals. Usage: python tools/gpu_actor_overhead.py [sequences]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import spatialthinker_b200 as st  # noqa: E402
from spatialthinker_b200 import dp_actor  # noqa: E402

seqs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = list(bench.CONFIGS["c3"])
cfg[2] = seqs
cfg = tuple(cfg)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
x = bench.make_inputs(st, cfg, 0, 1, dev, False, False, 2)
actor = st.DataParallelPPOActor(bench.actor_config(st, x, 0, False, 0.0), x["weight"])
for _ in range(2):
    bench.run_step_device(st, actor, x, 1)
torch.cuda.synchronize()

stamps = {}


def wrap(obj, name, key):
    orig = getattr(obj, name)

    def timed(*a, **k):
        t0 = time.perf_counter()
        out = orig(*a, **k)
        stamps[key] = stamps.get(key, 0.0) + time.perf_counter() - t0
        stamps[key + "_n"] = stamps.get(key + "_n", 0) + 1
        return out

    setattr(obj, name, timed)


wrap(actor, "_micro_plan", "plan")
wrap(actor, "_valid_lengths", "valid_lengths")
wrap(actor, "_optimizer_step", "optimizer_step_host")
wrap(dp_actor, "grpo_micro_batch_step", "micro_batch_step_host")
wrap(dp_actor, "_rows_of", "rows_of_host")
wrap(st.core_algos, "compute_grpo_outcome_advantage_sharded", "advantage_host")
for it in range(2):
    stamps.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bench.run_step_device(st, actor, x, 1)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(f"step {it}: wall {wall * 1e3:.1f} ms, gpu {e0.elapsed_time(e1):.1f} ms")
    for k in sorted(stamps):
        if not k.endswith("_n"):
            print(f"   {k:28s} {stamps[k] * 1e3:9.2f} ms over {stamps[k + '_n']} calls")
# GPU cost of the row gathers
idx = torch.arange(0, 37 * 27, 27, device=dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    x["hidden"].index_select(0, idx)
e1.record()
torch.cuda.synchronize()
print(f"index_select of 37 sequences of hidden: {e0.elapsed_time(e1) / 10:.3f} ms")
