"""Multi-GPU parity check, launched one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu_check.py [n_reads]

Every rank builds its contiguous shard of a seeded synthetic workload on its own B200, the ranks
exchange hash-partitioned ECs over NCCL (alntools_b200/multi_gpu.py) and every rank must end up with
exactly the matrices the C oracle computes for the whole, unsharded input."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
    from alntools_b200 import multi_gpu, synth
    from alntools_b200._native import EcBuilder
    from oracle import c_oracle

    for seed, mode, n_targets, n_haps in ((11, "diploid", 20000, 2), (12, "heavy", 3000, 8)):
        reads = n_reads if mode == "diploid" else n_reads // 20
        cols = synth.make_columns(reads, n_targets, n_haps, seed=seed, mode=mode, dup_rate=0.01)
        rg, tg, hp = cols["read_group"], cols["target_idx"], cols["hap_idx"]
        cuts = multi_gpu.shard_bounds(rg, world)
        a, b = cuts[rank], cuts[rank + 1]
        local_b = EcBuilder(n_targets, n_haps, alignments_hint=b - a, device=local)
        owner_b = EcBuilder(n_targets, n_haps, alignments_hint=b - a, device=local)
        for _ in range(passes):   # contexts are reused from pass to pass, as a long-running job would
            local_b.reset()
            owner_b.reset()
            local_b.push(np.ascontiguousarray(rg[a:b]), np.ascontiguousarray(tg[a:b]),
                         np.ascontiguousarray(hp[a:b]), order_base=a)
            res = multi_gpu.distributed_finalize(local_b, lambda: owner_b, dev)
        indptr, indices, data, counts, n_reads_o = c_oracle.ec_from_columns(rg, tg, hp)
        ok = (np.array_equal(res["a_indptr"].cpu().numpy(), indptr)
              and np.array_equal(res["a_indices"].cpu().numpy(), indices)
              and np.array_equal(res["a_data"].cpu().numpy(), data)
              and np.array_equal(res["n_data"].cpu().numpy(), counts))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("multigpu_check world=%d mode=%s reads=%d alignments=%d ECs=%d: %s"
                  % (world, mode, reads, len(rg), len(counts), "OK" if flag.item() else "MISMATCH"), flush=True)
        if not flag.item():
            dist.destroy_process_group()
            sys.exit(1)
        local_b.close()
        owner_b.close()

    # the bench's shape: every rank has its OWN seeded shard, columns resident on the device
    n_targets, n_haps = 20000, 2
    shards = [synth.make_columns(n_reads // 2, n_targets, n_haps, seed=100 + r, mode="diploid") for r in range(world)]
    sizes = [len(s["read_group"]) for s in shards]
    base = sum(sizes[:rank])
    mine = {k: torch.from_numpy(shards[rank][k]).to(dev) for k in ("read_group", "target_idx", "hap_idx")}
    local_b = EcBuilder(n_targets, n_haps, alignments_hint=sizes[rank], device=local, result_on_device=1)
    owner_b = EcBuilder(n_targets, n_haps, alignments_hint=sizes[rank], device=local)
    for _ in range(passes + 1):
        local_b.reset()
        owner_b.reset()
        local_b.push(mine["read_group"], mine["target_idx"], mine["hap_idx"], order_base=base)
        res = multi_gpu.distributed_finalize(local_b, lambda: owner_b, dev)
    rg = np.concatenate([s["read_group"] + 10 ** 8 * r for r, s in enumerate(shards)])   # reads stay distinct
    tg = np.concatenate([s["target_idx"] for s in shards])
    hp = np.concatenate([s["hap_idx"] for s in shards])
    indptr, indices, data, counts, _ = c_oracle.ec_from_columns(rg.astype(np.int32), tg, hp)
    ok = (np.array_equal(res["a_indptr"].cpu().numpy(), indptr) and np.array_equal(res["a_indices"].cpu().numpy(), indices)
          and np.array_equal(res["a_data"].cpu().numpy(), data) and np.array_equal(res["n_data"].cpu().numpy(), counts))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multigpu_check world=%d own-shards device-resident reuse: alignments=%d ECs=%d: %s"
              % (world, len(rg), len(counts), "OK" if flag.item() else "MISMATCH"), flush=True)
    # the same job with the result left partitioned by EC-id range (peer-memory exchange only)
    if os.environ.get("ECB_EXCHANGE", "p2p") != "nccl":
        for _ in range(2):
            local_b.reset()
            owner_b.reset()
            local_b.push(mine["read_group"], mine["target_idx"], mine["hap_idx"], order_base=base)
            sl = multi_gpu.distributed_finalize(local_b, lambda: owner_b, dev, result_on="slices")
        a, b = sl["id_base"], sl["id_base"] + sl["n_ec_local"]
        ok = (sl["n_ec"] == len(counts)
              and np.array_equal(sl["a_indptr"].cpu().numpy(), indptr[a:b + 1] - indptr[a])
              and np.array_equal(sl["a_indices"].cpu().numpy(), indices[indptr[a]:indptr[b]])
              and np.array_equal(sl["a_data"].cpu().numpy(), data[indptr[a]:indptr[b]])
              and np.array_equal(sl["n_data"].cpu().numpy(), counts[a:b]))
        flag2 = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag2, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("multigpu_check world=%d slices: %s" % (world, "OK" if flag2.item() else "MISMATCH"), flush=True)
        flag = torch.minimum(flag, flag2)
    local_b.close()
    owner_b.close()
    if not flag.item():
        dist.destroy_process_group()
        sys.exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
