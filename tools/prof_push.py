import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from alntools_b200 import synth
from alntools_b200._native import EcBuilder
cols = synth.make_columns(8_000_000, 100000, 2, 1002, mode="diploid")
n = len(cols["read_group"])
host = {k: torch.from_numpy(cols[k]).pin_memory() for k in ("read_group", "target_idx", "hap_idx")}
half = n // 2
while cols["read_group"][half] == cols["read_group"][half - 1]:
    half += 1
for hint in (77_000_000 // 16, n, 4 * n):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = EcBuilder(100000, 2, alignments_hint=hint)
        t1 = time.perf_counter()
        b.push(host["read_group"][:half], host["target_idx"][:half], host["hap_idx"][:half], order_base=0)
        t2 = time.perf_counter()
        s1 = b.stats()
        b.push(host["read_group"][half:], host["target_idx"][half:], host["hap_idx"][half:], order_base=half)
        t3 = time.perf_counter()
        s2 = b.stats()
        res = b.finalize()
        t4 = time.perf_counter()
        b.close()
        t5 = time.perf_counter()
        print("hint %d: create %.1f ms, push1 %.1f ms (grows %d, overflow %d, group %.2f ms, push_ms %.1f), push2 %.1f ms (grows %d, overflow %d), finalize %.1f ms, close %.1f ms, n_ec %d"
              % (hint, 1e3 * (t1 - t0), 1e3 * (t2 - t1), s1["table_grows"], s1["overflow_reads"], s1["group_ms"], s1["push_ms"], 1e3 * (t3 - t2), s2["table_grows"], s2["overflow_reads"], 1e3 * (t4 - t3), 1e3 * (t5 - t4), res["n_ec"]))
