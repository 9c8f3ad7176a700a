"""CPU check of the rule the GPU per-cell finalisation is built on (alntools_b200/csrc/ecb_cells.cuh):
a cell's position in the output = rank of  min over its (file, EC, cell) triples of
(first position of (file, EC), first position of (file, EC, cell)).  Compared with the oracle's
literal restatement of the reference's nested insertion order (bam_utils_multisample.py:513-551)."""
import numpy as np
import pytest

from alntools_b200 import synth
from oracle import ec_oracle


@pytest.mark.parametrize("seed,n_files,n_cells", [(1, 1, 8), (2, 3, 30), (3, 5, 200)])
def test_min_key_rule_equals_nested_insertion_order(seed, n_files, n_cells):
    pushes = []
    base = 0
    triples = {}
    for f in range(n_files):
        c = synth.make_columns(1500, 60, 2, seed=seed * 10 + f, mode="light", n_cells=n_cells, dup_rate=0.05)
        rg, tg, hp, cell = c["read_group"], c["target_idx"], c["hap_idx"], c["cell_idx"]
        pushes.append((rg, tg, hp, cell, True))
        starts = np.flatnonzero(np.concatenate(([True], rg[1:] != rg[:-1])))
        ends = np.concatenate((starts[1:], [len(rg)]))
        for s, e in list(zip(starts.tolist(), ends.tolist()))[:-1]:        # last read dropped
            key = tuple(sorted(set((tg[s:e].astype(np.int64) * 64 + hp[s:e]).tolist())))
            t = (f, key, int(cell[s]))
            triples.setdefault(t, base + s)
        base += len(rg)
    fe_first = {}
    for (f, key, _cell), pos in triples.items():
        fe_first[(f, key)] = min(fe_first.get((f, key), 1 << 62), pos)
    cell_key = {}
    for (f, key, cell), pos in triples.items():
        cand = (fe_first[(f, key)], pos)
        if cell not in cell_key or cand < cell_key[cell]:
            cell_key[cell] = cand
    order = sorted(cell_key, key=lambda c: cell_key[c])
    want = ec_oracle.ec_from_columns_cells(pushes, 1)["cell_order"].tolist()
    assert order == want
