"""A small run of every kernel family for compute-sanitizer (memcheck / racecheck): single-sample build with
table growth + replay and the hot-EC cache, long reads (all harvest classes), the per-cell path, and the
exchange (both dispatch forms) emulated with two ranks on one GPU.  Results are checked against the oracle.
usage: compute-sanitizer --tool memcheck python tools/sanitizer_case.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from alntools_b200 import synth, multi_gpu
from alntools_b200._native import EcBuilder
from oracle import c_oracle, ec_oracle

def same(got, want):
    return all(np.array_equal(np.asarray(got[k]), w) for k, w in zip(("a_indptr", "a_indices", "a_data", "n_data"), want[:4]))

cols = synth.make_columns(40000, 3000, 2, seed=5, mode="diploid", dup_rate=0.03)
want = c_oracle.ec_from_columns(cols["read_group"], cols["target_idx"], cols["hap_idx"])
with EcBuilder(3000, 2, table_slots=1024) as b:            # growth + replay
    b.push(cols["read_group"], cols["target_idx"], cols["hap_idx"])
    assert same(b.finalize(), want)
print("single-sample (growth, replay): ok", flush=True)

heavy = synth.make_columns(1500, 2000, 8, seed=6, mode="heavy", dup_rate=0.02)
rg = heavy["read_group"].copy(); tg = heavy["target_idx"].copy(); hp = heavy["hap_idx"].copy()
giant = 3000                                               # one read of 3000 alignments: the CTA-wide harvest
rg = np.concatenate([rg, np.full(giant, rg[-1] + 1, np.int32)]); tg = np.concatenate([tg, (np.arange(giant) % 1900).astype(np.int32)])
hp = np.concatenate([hp, (np.arange(giant) % 8).astype(np.int32)])
want = c_oracle.ec_from_columns(rg, tg, hp)
with EcBuilder(2000, 8, verify_keys=1) as b:
    b.push(rg, tg, hp)
    assert same(b.finalize(), want)
print("heavy multimapping + giant read + key verification: ok", flush=True)

pushes = []
for f in range(2):
    c = synth.make_columns(6000, 500, 2, seed=20 + f, mode="light", n_cells=60, dup_rate=0.02)
    pushes.append((c["read_group"], c["target_idx"], c["hap_idx"], c["cell_idx"], True))
wantc = ec_oracle.ec_from_columns_cells(pushes, 40)
with EcBuilder(500, 2, with_cells=True, alignments_hint=sum(len(p[0]) for p in pushes)) as b:
    base = 0
    for r, t, h, cell, drop in pushes:
        b.push(r, t, h, cell, order_base=base, drop_last_group=drop)
        base += len(r)
    got = b.finalize(40)
assert all(np.array_equal(got[k], wantc[k]) for k in ("a_indptr", "a_indices", "a_data", "n_indptr", "n_indices", "n_data", "cell_order"))
print("per-cell path: ok", flush=True)

cols = synth.make_columns(30000, 2000, 2, seed=9, mode="diploid", dup_rate=0.02)
want = c_oracle.ec_from_columns(cols["read_group"], cols["target_idx"], cols["hap_idx"])
cuts = multi_gpu.shard_bounds(cols["read_group"], 2)
locals_, owners = [], []
for r in range(2):
    a, e = cuts[r], cuts[r + 1]
    lb = EcBuilder(2000, 2, alignments_hint=e - a)
    lb.push(np.ascontiguousarray(cols["read_group"][a:e]), np.ascontiguousarray(cols["target_idx"][a:e]),
            np.ascontiguousarray(cols["hap_idx"][a:e]), order_base=a)
    locals_.append(lb)
    owners.append(EcBuilder(2000, 2, alignments_hint=len(cols["read_group"])))
cap_ec = 2 * max(l.stats()["table_used"] for l in locals_) + 1024
bases = [o.arena_create(cap_ec)[1] for o in owners]
for l in locals_:
    l.export_to_arenas(bases, cap_ec)
for o in owners:
    o.import_arena()
    o.arena_reset()
for o in owners:
    o.order_dispatch(bases, cap_ec, cuts[:-1], cuts[1:])
at = 0
for r, o in enumerate(owners):
    sl = o.order_build(locals_[r], cuts[r], cuts[r + 1])
    a, e = at, at + sl["n_ec"]
    assert np.array_equal(sl["a_indptr"].cpu().numpy(), want[0][a:e + 1] - want[0][a])
    assert np.array_equal(sl["a_indices"].cpu().numpy(), want[1][want[0][a]:want[0][e]])
    assert np.array_equal(sl["n_data"].cpu().numpy(), want[3][a:e])
    at = e
assert at == len(want[3])
for x in locals_ + owners:
    x.close()
print("exchange (two ranks on one GPU, ordering dispatch): ok", flush=True)
