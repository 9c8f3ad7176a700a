import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from alntools_b200 import synth
from alntools_b200._native import EcBuilder
cols = synth.make_columns(30_000_000, 100000, 2, 2, mode="diploid")
dev = {k: torch.from_numpy(cols[k]).cuda() for k in ("read_group", "target_idx", "hap_idx")}
n = len(cols["read_group"])
b = EcBuilder(100000, 2, alignments_hint=n, result_on_device=1)
ms = []
for i in range(6):
    b.reset()
    try:
        b.push(dev["read_group"], dev["target_idx"], dev["hap_idx"])
    except Exception as exc:
        print("push error:", exc)
    ms.append(b.stats()["group_ms"])
print(os.environ.get("ECB200_LIB", "default"), "group_ms", [round(x, 4) for x in ms])
