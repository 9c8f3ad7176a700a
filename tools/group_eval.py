"""cfg2 (or another workload) on one GPU through the default library and any variant builds given:
parity of one against the other and the kernel / step times.
usage: python tools/group_eval.py [workload] [lib.so ...]   -> prints a table, appends to gpurun_out/group_eval.json"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from alntools_b200 import synth, _native
from alntools_b200._native import EcBuilder

wl_name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].endswith(".so") else "cfg2_diploid_30M"
libs = [a for a in sys.argv[1:] if a.endswith(".so")]
wl = dict(bench.WORKLOADS[wl_name])
t0 = time.time()
cols = synth.make_columns(wl["n_reads"], wl["n_targets"], wl["n_haps"], wl["seed"], mode=wl["mode"])
dev = {k: torch.from_numpy(cols[k]).cuda() for k in ("read_group", "target_idx", "hap_idx")}
n = len(cols["read_group"])
print("columns ready %.1f s, %d alignments" % (time.time() - t0, n), flush=True)
KEYS = ("a_indptr", "a_indices", "a_data", "n_data")
default_lib = _native.load_library()
runs = [("default", None, {})] + [(os.path.basename(l), l, {}) for l in libs]
base, out = None, {}
for name, lib, opts in runs:
    try:
        _native._lib = _native.load_library(lib) if lib else default_lib
        b = EcBuilder(wl["n_targets"], wl["n_haps"], alignments_hint=n, **opts)
        ms, step, same, res = [], [], "-", None
        for i in range(5):
            b.reset()
            torch.cuda.synchronize()
            t1 = time.time()
            b.push(dev["read_group"], dev["target_idx"], dev["hap_idx"])
            st = b.stats()
            ms.append(st["group_ms"])
            res = b.finalize()
            torch.cuda.synchronize()
            step.append((time.time() - t1) * 1e3)
        got = {k: np.array(res[k]) for k in KEYS}
        got["n_reads"], got["n_ec"] = res["n_reads"], res["n_ec"]
        if base is None:
            base, same = got, "reference"
        else:
            same = "SAME" if (all(np.array_equal(got[k], base[k]) for k in KEYS) and got["n_reads"] == base["n_reads"]
                              and got["n_ec"] == base["n_ec"]) else "DIFFERENT"
        out[name] = {"group_ms": ms, "step_ms": step, "parity": same, "n_ec": int(got["n_ec"]), "harvest_ms": st["harvest_ms"]}
        print("%-24s group_ms %s  harvest %.3f step_ms %s  n_ec %d n_reads %d  %s  frac %.3f" % (
            name, [round(x, 4) for x in ms], st["harvest_ms"], [round(x, 2) for x in step], got["n_ec"], got["n_reads"], same,
            12.0 * n / (min(ms[1:]) * 1e-3) / 6538.3e9), flush=True)
        b.close()
    except Exception as exc:
        out[name] = {"error": str(exc)}
        print(name, "ERROR:", exc, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/group_eval_%s.json" % wl_name, "w") as fh:
    json.dump(out, fh, indent=1)
