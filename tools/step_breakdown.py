"""Where a device-resident step goes: library-side CUDA-event times (group / harvest / push / finalize) and the
wall clock of reset, push and finalize for cfg2.  usage: python tools/step_breakdown.py [workload]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from alntools_b200 import synth
from alntools_b200._native import EcBuilder
wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2_diploid_30M"])
cols = synth.make_columns(wl["n_reads"], wl["n_targets"], wl["n_haps"], wl["seed"], mode=wl["mode"])
dev = {k: torch.from_numpy(cols[k]).cuda() for k in ("read_group", "target_idx", "hap_idx")}
n = len(cols["read_group"])
b = EcBuilder(wl["n_targets"], wl["n_haps"], alignments_hint=n, result_on_device=1)
for i in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); b.reset(); torch.cuda.synchronize()
    t1 = time.perf_counter(); b.push(dev["read_group"], dev["target_idx"], dev["hap_idx"]); torch.cuda.synchronize()
    t2 = time.perf_counter(); b.finalize_raw(); torch.cuda.synchronize()
    t3 = time.perf_counter()
    st = b.stats()
    print("step %d: wall reset %.3f push %.3f finalize %.3f | events: group %.3f harvest %.3f push %.3f finalize %.3f | launches %d"
          % (i, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), st["group_ms"], st["harvest_ms"], st["push_ms"], st["finalize_ms"],
             st["kernel_launches"]), flush=True)
