"""Model of the two-stack queue of the dense grouping kernel (ECB_OPT_TWO_PHASE = 3, ecb_group.cuh:
closed reads parked from entry 95 downwards, misses from entry 0 upwards, both popped 32 at a time): the
index arithmetic of the kernel, replayed on the CPU with random windows, must hand every closed read to the
cache look-up exactly once, every miss to the log exactly once, keep the log blocks free of holes and never
let the two stacks touch.  (The kernel itself has its GPU parity test; this pins the bookkeeping.)"""
import numpy as np

MQ, BLOCK = 96, 256


def test_two_stacks_never_touch_and_deliver_everything_once():
    rng = np.random.default_rng(11)
    for trial in range(200):
        q = [None] * MQ                      # the warp's queue
        cn = qn = 0
        lb, lu = 0, BLOCK                    # log block base / fill
        cursor = 0
        log = {}
        probed, missed = [], []
        p_ins = float(rng.uniform(0.05, 0.97))
        p_miss = float(rng.uniform(0.0, 1.0))
        read_id = 0

        def log32(entries):                  # log_misses_priv(pair = false): 32 slots, absent lanes write an empty entry
            nonlocal lb, lu, cursor
            if lu >= BLOCK:
                lb, lu = cursor, 0
                cursor += BLOCK
            for lane in range(32):
                assert lb + lu + lane not in log
                log[lb + lu + lane] = entries[lane] if lane < len(entries) else "empty"
            lu += 32

        def dense_commit(has, idx_of_lane):
            nonlocal qn
            got = [q[idx_of_lane(l)] if has(l) else None for l in range(32)]
            misses = []
            for l in range(32):
                if has(l):
                    assert got[l] is not None and got[l][0] == "closed"
                    probed.append(got[l][1])
                    if rng.random() < p_miss:
                        misses.append(got[l][1])
            for k, r in enumerate(misses):   # lower stack
                assert qn + k < MQ - cn or True
                q[qn + k] = ("miss", r)
            # remaining closed entries live at indices >= MQ - cn: the new misses must stay below them
            assert qn + len(misses) <= MQ - cn
            qn += len(misses)
            if qn >= 32:
                qn -= 32
                batch = [q[qn + l][1] for l in range(32)]
                assert all(q[qn + l][0] == "miss" for l in range(32))
                missed.extend(batch)
                log32(batch)

        for window in range(int(rng.integers(1, 400))):
            ins = rng.random(31) < p_ins     # at most 31 closed reads per window
            k = int(ins.sum())
            for j in range(k):
                idx = (MQ - 1) - (cn + j)
                assert idx >= qn, "closed stack ran into the miss stack"
                q[idx] = ("closed", read_id)
                read_id += 1
            cn += k
            if cn >= 32:
                cn -= 32
                base = cn
                dense_commit(lambda l: True, lambda l, b=base: (MQ - 1) - (b + l))
        if cn:                               # end of the kernel
            left = cn
            dense_commit(lambda l, n=left: l < n, lambda l: (MQ - 1) - l)
            cn = 0
        if qn:
            batch = [q[l][1] for l in range(min(qn, 32))]
            missed.extend(batch)
            log32(batch)
        assert qn <= 32
        for i in range(lu, BLOCK):           # padding of the last block
            if lb + i not in log and cursor:
                log[lb + i] = "empty"
        assert sorted(probed) == list(range(read_id))
        assert len(missed) == len(set(missed))
        assert sorted(x for x in log.values() if x != "empty") == sorted(missed)
        assert sorted(log) == list(range(cursor))          # whole blocks, no holes
