"""N>1 host path on CPU: world_size-2 (and 3) gloo process groups run the SAME orchestration as the
NCCL path (alntools_b200/multi_gpu.distributed_finalize: all-to-all of hash-partitioned ECs, owner
merge, bitmap all-reduce for global ids, scatter + all-reduce of the CSR) with a numpy stand-in for
the libecb200 kernels, and must reproduce the oracle's single-process answer on the concatenated
shards."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class FakeBuilder(object):
    """CPU stand-in with the primitive set of alntools_b200._native.EcBuilder used by multi_gpu.py."""

    def __init__(self):
        self.ecs = {}          # key bytes -> [first, count, row [(target, mask)]]
        self.min_base, self.max_end = 0, 0

    # -- local role
    def push(self, rg, tg, hp, order_base=0):
        n = len(rg)
        if n == 0:
            return
        self.min_base, self.max_end = order_base, order_base + n
        starts = np.flatnonzero(np.concatenate(([True], rg[1:] != rg[:-1])))
        ends = np.concatenate((starts[1:], [n]))
        code = tg.astype(np.int64) * 64 + hp
        for s, e in zip(starts.tolist(), ends.tolist()):
            key = np.unique(code[s:e])
            kb = key.tobytes()
            if kb not in self.ecs:
                row = {}
                for c in key.tolist():
                    row[c // 64] = row.get(c // 64, 0) | (1 << (c % 64))
                self.ecs[kb] = [order_base + s, 0, sorted(row.items())]
            self.ecs[kb][1] += 1

    @staticmethod
    def _key128(kb):
        d = hashlib.sha256(kb).digest()
        return int.from_bytes(d[:8], "little", signed=True), int.from_bytes(d[8:16], "little", signed=True)

    def export_partition(self, world):
        parts = [[] for _ in range(world)]
        for kb, (first, count, row) in self.ecs.items():
            lo, hi = self._key128(kb)
            parts[(hi >> 32) % world].append((lo, hi, first, count, row))
        meta, rows, ec_counts, row_counts = [], [], [], []
        for part in parts:
            roff = 0
            for lo, hi, first, count, row in part:
                meta.append([lo, hi, first, (count << 32) | len(row), roff])
                rows.extend(row)
                roff += len(row)
            ec_counts.append(len(part))
            row_counts.append(roff)
        meta_t = torch.tensor(meta, dtype=torch.int64).reshape(-1, 5)
        rows_t = torch.tensor(rows, dtype=torch.int32).reshape(-1, 2)
        return meta_t, rows_t, ec_counts, row_counts, self.min_base, self.max_end

    # -- owner role
    def import_entries(self, meta, rows, ec_counts, row_counts):
        self.owned = {}
        rec = 0
        row_base = 0
        for n_ec, n_rows in zip(ec_counts, row_counts):
            for i in range(rec, rec + n_ec):
                lo, hi, first, cl, roff = meta[i].tolist()
                count, length = cl >> 32, cl & 0xFFFFFFFF
                row = rows[row_base + roff:row_base + roff + length].tolist()
                e = self.owned.setdefault((lo, hi), [first, 0, row])
                e[0] = min(e[0], first)
                e[1] += count
            rec += n_ec
            row_base += n_rows

    def global_mark(self, min_base, bitmap):
        self.g_min = min_base
        for first, _, _ in self.owned.values():
            rel = first - min_base
            bitmap[rel >> 5] |= np.int32(np.uint32(1 << (rel & 31)).view(np.int32)).item()

    def global_count(self, bitmap):
        bits = np.unpackbits(bitmap.numpy().view(np.uint8), bitorder="little")
        self.rank = np.cumsum(bits) - bits
        return int(bits.sum())

    def global_lens(self, lens, counts):
        for first, count, row in self.owned.values():
            i = int(self.rank[first - self.g_min])
            lens[i] = len(row)
            counts[i] = count

    def global_indptr(self, lens):
        v = lens.clone()
        lens.copy_(torch.cumsum(v, 0).to(torch.int32) - v)
        return int(v.sum().item())

    def global_rows(self, indptr, indices, data):
        for first, _, row in self.owned.values():
            p = int(indptr[int(self.rank[first - self.g_min])])
            for j, (t, m) in enumerate(row):
                indices[p + j] = t
                data[p + j] = m


def _worker(rank, world, port, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alntools_b200 import multi_gpu, synth
    cols = synth.make_columns(6000, 300, 3, seed=seed, mode="diploid", dup_rate=0.03)
    rg, tg, hp = cols["read_group"], cols["target_idx"], cols["hap_idx"]
    cuts = multi_gpu.shard_bounds(rg, world)
    a, b = cuts[rank], cuts[rank + 1]
    local = FakeBuilder()
    local.push(rg[a:b], tg[a:b], hp[a:b], order_base=a)
    res = multi_gpu.distributed_finalize(local, FakeBuilder, torch.device("cpu"))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{k: np.asarray(v) for k, v in res.items()})
    dist.destroy_process_group()


@pytest.mark.parametrize("world,seed", [(2, 5), (3, 6)])
def test_sharded_exchange_reproduces_single_process_result(tmp_path, world, seed):
    from alntools_b200 import synth
    from oracle import ec_oracle
    port = 29500 + (os.getpid() + world) % 2000
    mp.spawn(_worker, args=(world, port, seed, str(tmp_path)), nprocs=world, join=True)
    cols = synth.make_columns(6000, 300, 3, seed=seed, mode="diploid", dup_rate=0.03)
    indptr, indices, data, counts = ec_oracle.ec_from_columns(cols["read_group"], cols["target_idx"], cols["hap_idx"])
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert np.array_equal(got["a_indptr"], indptr)
        assert np.array_equal(got["a_indices"], indices)
        assert np.array_equal(got["a_data"], data)
        assert np.array_equal(got["n_data"], counts)


def test_shard_bounds_never_split_a_read():
    from alntools_b200 import multi_gpu
    rg = np.repeat(np.arange(50), np.random.default_rng(0).integers(1, 9, 50)).astype(np.int32)
    for world in (1, 2, 3, 8, 64):
        cuts = multi_gpu.shard_bounds(rg, world)
        assert cuts[0] == 0 and cuts[-1] == len(rg) and len(cuts) == world + 1
        assert all(b >= a for a, b in zip(cuts[:-1], cuts[1:]))
        for c in cuts[1:-1]:
            assert c == len(rg) or c == 0 or rg[c] != rg[c - 1]
