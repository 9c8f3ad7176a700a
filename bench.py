#!/usr/bin/env python
"""bench.py — bam2ec EC-build throughput (alignments/s) on B200, per the driver contract.

One "step" = one pass of the hot path over one batch of synthetic, name-grouped alignment columns:
ecb_reset -> ecb_push (grouping + hash insert + row harvest) -> ecb_finalize (EC ids, CSR A, CSC N).

  value : whole-job alignments/s with the columns already resident in HBM and results left in HBM
  e2e   : the same through the public C-ABI call with HOST (pinned) columns and host results;
          H2D and D2H copies are inside the timed region
  roofline     : the grouping kernel against the measured HBM copy bandwidth
  cpu_baseline : the oracle's Python port of the reference's per-alignment loop on the host cores,
                 on a bounded sample of the same workload (rank 0, N=1 only)

`--impl reference` times that CPU port as its own arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: diploid F1 mouse, 2 haplotypes x ~100k transcripts, 30M reads, 1 B200
    "cfg2_diploid_30M": dict(n_reads=30_000_000, n_targets=100_000, n_haps=2, mode="diploid", seed=2),
    "cfg1_small_1M": dict(n_reads=1_000_000, n_targets=2_000, n_haps=2, mode="light", seed=1),
    "cfg3_do8_heavy": dict(n_reads=3_000_000, n_targets=140_000, n_haps=8, mode="heavy", seed=3),
    # BASELINE.json configs[3], scaled to one GPU: per-cell counts, 50k cells (Zipf sizes), 4 files, the
    # last read of every file dropped, cells below 1000 reads filtered out
    "cfg4_cells_50k": dict(n_reads=20_000_000, n_targets=140_000, n_haps=8, mode="diploid", seed=4,
                           n_cells=50_000, n_files=4, min_count=1000),
}


# BASELINE.json configs[4]: multimapping-degree sweep, exactly k alignments per read, 64 M alignments per GPU
for _k in (1, 2, 4, 8, 16, 32, 64):
    WORKLOADS["cfg5_sweep_k%d" % _k] = dict(n_reads=64_000_000 // _k, n_targets=140_000, n_haps=8, mode="aln%d" % _k, seed=5)


def measured_traffic(workload):
    """(DRAM bytes per launch of the grouping kernel, where the figure comes from) - from the committed ncu
    capture of the same workload, NOT from the run that prints it (ncu cannot run inside a timed bench)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            t = json.load(fh)
        if t.get("workload") != workload:
            return None, None
        return float(t["dram_bytes_per_launch"]), "profiles/traffic.json: %s" % t.get("capture", "ncu capture")
    except (OSError, ValueError, KeyError):
        return None, None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region.  NVML in-process (a query costs
    microseconds, so a few-millisecond region still gets many samples and is not disturbed); falls back
    to spawning nvidia-smi, whose queries can stall kernel launches for milliseconds."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]   # nvmlClocksEventReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index):
        self.index = index
        self.samples = []          # [sm_mhz, sm_max_mhz, flag, flag, flag, flag]
        self.source = "nvidia-smi"
        self._stop = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._max = int(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None
        self._thread = threading.Thread(target=self._run_nvml if self._nvml else self._run_smi, daemon=True)

    def _run_nvml(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
                self.samples.append([mhz, self._max] + [bool(reasons & b) for b in self.BITS])
            except Exception:
                pass
            self._stop.wait(0.002)

    def _run_smi(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    f = [x.strip() for x in out.stdout.strip().split(",")]
                    if f[0].isdigit():
                        self.samples.append([int(f[0]), int(f[1]) if f[1].isdigit() else None]
                                            + [x.lower().startswith("active") for x in f[2:6]])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def mark(self):
        """Samples taken from now on belong to the timed region."""
        self._first = len(self.samples)

    def summary(self):
        samples = self.samples[getattr(self, "_first", 0):] or self.samples
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock sample"], "source": self.source}
        mhz = sorted(s[0] for s in samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2 + i] for s in samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": samples[0][1], "reasons": reasons,
                "samples": len(samples), "source": self.source}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's port of the reference's own multiprocessing bam2ec path
# ------------------------------------------------------------------------------------------------
def _cpu_worker(bam_path):
    from alntools_b200 import bam_io
    from oracle import ec_oracle
    raw = bam_io.inflate_file(bam_path)
    header = bam_io.parse_header(raw)
    res = ec_oracle.group_chunk_single(bam_io.iter_records(raw, header.records_offset))
    return res.ec, res.valid_alignments


def _cpu_decode_only(bam_path):
    """The decode half of _cpu_worker alone: inflate + header + one pass over the records."""
    from alntools_b200 import bam_io
    raw = bam_io.inflate_file(bam_path)
    header = bam_io.parse_header(raw)
    n = 0
    for _ in bam_io.iter_records(raw, header.records_offset):
        n += 1
    return n


def make_cpu_sample(workdir, wl, sample_reads, n_procs):
    """A bounded sample of the workload as `n_procs` read-aligned chunk BAMs (the reference also
    materialises one temporary BAM per chunk, bam_utils.py:247-250)."""
    import numpy as np
    from alntools_b200 import bam_io, synth
    cols = synth.make_columns(sample_reads, wl["n_targets"], wl["n_haps"], wl["seed"] + 1000, mode=wl["mode"])
    rg = cols["read_group"]
    tids = cols["target_idx"].astype(np.int64) * wl["n_haps"] + cols["hap_idx"]
    refs = synth.reference_names(wl["n_targets"], wl["n_haps"])
    cuts = [0]
    for k in range(1, n_procs):
        c = len(rg) * k // n_procs
        while c < len(rg) and rg[c] == rg[c - 1]:
            c += 1
        cuts.append(c)
    cuts.append(len(rg))
    paths = []
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        p = os.path.join(workdir, "chunk%03d.bam" % i)
        bam_io.write_bam_columns(p, refs, rg[a:b], np.zeros(b - a, dtype=np.uint16), tids[a:b], level=1)
        paths.append(p)
    return paths, refs, len(rg)


def cpu_step(paths, refs, pool):
    """One pass of the reference algorithm: per-chunk grouping in a process pool, ordered merge,
    EC -> matrix, EC-file bytes.  Returns (alignments, seconds)."""
    from oracle import ec_oracle
    t0 = time.perf_counter()
    results = pool.map(_cpu_worker, paths) if pool is not None else [_cpu_worker(p) for p in paths]
    ec = {}
    valid = 0
    for part, v in results:
        valid += v
        for k, c in part.items():
            ec[k] = ec.get(k, 0) + c
    tables = ec_oracle.HeaderTables([r[0] for r in refs], [r[1] for r in refs])
    a_csr = ec_oracle.a_matrix_from_keys(list(ec.keys()), tables)
    n_csc = ec_oracle.n_matrix_single(list(ec.values()))
    blob = ec_oracle.ecsave2_bytes(tables.haplotypes, list(tables.main_targets.keys()), tables.lengths,
                                   ["sample"], a_csr, n_csc)
    assert len(blob) > 0
    return valid, time.perf_counter() - t0


def run_cpu_arm(wl_name, steps, warmup, sample_reads, max_seconds=240.0):
    import multiprocessing
    wl = WORKLOADS[wl_name]
    n_procs = os.cpu_count() or 1
    with tempfile.TemporaryDirectory(prefix="ecb_cpu_") as tmp:
        paths, refs, n_aln = make_cpu_sample(tmp, wl, sample_reads, n_procs)
        ctx = multiprocessing.get_context("fork")
        with ctx.Pool(n_procs) as pool:
            for _ in range(warmup):
                cpu_step(paths, refs, pool)
            total_aln, total_s = 0, 0.0
            done = 0
            for _ in range(steps):
                a, s = cpu_step(paths, refs, pool)
                total_aln += a
                total_s += s
                done += 1
                if total_s > max_seconds:
                    break
        # also ONE process, and the decode-only pass (what pysam would do in the reference) apart:
        # one chunk file of the same sample, one process, no pool
        t0 = time.perf_counter()
        _, one_aln = _cpu_worker(paths[0])
        t_full = time.perf_counter() - t0
        t0 = time.perf_counter()
        _cpu_decode_only(paths[0])
        t_dec = time.perf_counter() - t0
    value = total_aln / total_s
    group_share = max(t_full - t_dec, 1e-9) / t_full
    return {"value": value, "unit": "alignments/s", "cores": n_procs, "kind": "port",
            "sample": "%d reads / %d alignments of %s as %d chunk BAMs; decode + group + merge + matrix + EC bytes, "
                      "oracle Python port of bam_utils.convert, %d timed passes" % (sample_reads, n_aln, wl_name,
                                                                                    n_procs, done),
            "value_1core": one_aln / t_full, "decode_only_s": t_dec, "decode_and_group_s": t_full,
            "value_1core_decode_subtracted": one_aln / max(t_full - t_dec, 1e-9),
            "value_decode_subtracted": value / group_share,
            "one_core_sample": "chunk 0 of the same sample (%d alignments), one process: decode + group, and decode "
                               "alone (pure-Python BGZF inflate + record pass standing in for pysam); "
                               "value_decode_subtracted = value / (1 - decode share of that chunk)" % one_aln,
            "ms_per_step": 1e3 * total_s / max(done, 1), "steps_done": done, "sample_reads": sample_reads,
            "sample_alignments": n_aln}


def run_bam_e2e(wl_name, sample_reads, device, passes=3):
    """The reference-facing call itself on a bounded sample: alntools_b200.bam_utils.convert(BAM file ->
    EC file): native decode (libbamcols, all host threads) streamed to the GPU, EC bytes written."""
    import numpy as np  # noqa: F401
    from alntools_b200 import bam_utils, synth
    wl = WORKLOADS[wl_name]
    cols = synth.make_columns(sample_reads, wl["n_targets"], wl["n_haps"], wl["seed"] + 1000, mode=wl["mode"])
    n_aln = int(len(cols["read_group"]))
    with tempfile.TemporaryDirectory(prefix="ecb_bam_") as tmp:
        bam = os.path.join(tmp, "sample.bam")
        synth.columns_to_bam(bam, cols, wl["n_targets"], wl["n_haps"])
        times = []
        for i in range(passes + 1):
            t0 = time.perf_counter()
            bam_utils.convert(bam, os.path.join(tmp, "out.bin"), None, device=device)
            if i > 0:
                times.append(time.perf_counter() - t0)
        size = os.path.getsize(os.path.join(tmp, "out.bin"))
    best = min(times)
    return {"value": n_aln / best, "unit": "alignments/s", "ms": 1e3 * best, "threads": os.cpu_count(),
            "sample": "%d reads / %d alignments of %s as one BAM file; bam_utils.convert: BGZF inflate + record "
                      "pass (libbamcols) -> pinned columns -> GPU EC build -> EC file (%d bytes); best of %d after one warm-up "
                      "pass (the pinned column buffers of a finished convert() are reused by the next one in the process)"
                      % (sample_reads, n_aln, wl_name, size, passes)}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2_diploid_30M", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-reads", type=int, default=2_000_000)
    ap.add_argument("--bam-sample-reads", type=int, default=8_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--table-slots", type=int, default=0)
    ap.add_argument("--grid-ctas", type=int, default=0)
    ap.add_argument("--hot-cache", type=int, default=1)
    ap.add_argument("--chunk-len", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries exactly one JSON line: anything libraries print on fd 1 meanwhile (NCCL announces its
    # version there) is sent to stderr; the real stdout comes back for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    config = {"workload": args.workload, "reads_per_gpu": wl["n_reads"], "n_targets": wl["n_targets"],
              "n_haps": wl["n_haps"], "multimapping": str(wl["mode"]), "sharding": "contiguous read chunks per GPU; N > 1: + dispatch of the local ECs to their owner rank (hash of "
                                                              "the key) over peer memory, owner merge, second dispatch to the rank whose "
                                                              "shard holds the EC's first occurrence, ids ranked there: the final CSR stays "
                                                              "partitioned by EC-id range across the ranks (e2e: every rank copies its "
                                                              "range to its host)",
              "l2": "see config.l2_hygiene"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cpu = run_cpu_arm(args.workload, max(args.steps, 1), min(args.warmup, 1), args.cpu_sample_reads)
        config = dict(config, reads_per_gpu=cpu["sample_reads"], alignments_per_step=cpu["sample_alignments"],
                      sample_of=args.workload, sharding="one chunk BAM per host process (%d), ordered merge in the parent" % cpu["cores"],
                      l2="n/a (host cores)")
        line = {"impl": "reference", "metric": "bam2ec alignments/sec (EC build)", "value": cpu["value"],
                "unit": "alignments/s", "n_gpus": args.gpus, "steps": cpu["steps_done"], "warmup": min(args.warmup, 1),
                "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic", "config": config, "cpu_baseline": cpu,
                "e2e": {"value": cpu["value"], "unit": "alignments/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from alntools_b200 import synth
    from alntools_b200._native import EcBuilder

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path for the EC build")
    numa_cpus = None
    if world > 1:
        from alntools_b200 import utils as _utils
        numa_cpus = _utils.bind_to_gpu_numa_node(local_rank)   # before any pinned allocation
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic shard of this rank (weak scaling: fixed reads per GPU) --------------------------
    n_cells = int(wl.get("n_cells", 0))
    cols = synth.make_columns(wl["n_reads"], wl["n_targets"], wl["n_haps"], wl["seed"] + 7919 * rank, mode=wl["mode"],
                              n_cells=n_cells)
    n_aln = int(len(cols["read_group"]))
    names = ("read_group", "target_idx", "hap_idx") + (("cell_idx",) if n_cells else ())
    if n_cells and world > 1:
        raise SystemExit("the per-cell path is single-GPU in this round")
    # per-cell workloads arrive as several files: read-aligned pieces, one push each
    file_cuts = [0, n_aln]
    if n_cells:
        file_cuts = [0]
        for k in range(1, int(wl["n_files"])):
            c = n_aln * k // int(wl["n_files"])
            while c < n_aln and cols["read_group"][c] == cols["read_group"][c - 1]:
                c += 1
            file_cuts.append(c)
        file_cuts.append(n_aln)
    min_count = int(wl.get("min_count", 0))
    host = {k: torch.from_numpy(cols[k]).pin_memory() for k in names}
    dev = {k: host[k].cuda(non_blocking=True) for k in names}
    torch.cuda.synchronize()
    order_base = 0
    if world > 1:
        counts = torch.zeros(world, dtype=torch.int64, device="cuda")
        counts[rank] = n_aln
        dist.all_reduce(counts)
        order_base = int(counts[:rank].sum().item())
        total_aln = int(counts.sum().item())
    else:
        total_aln = n_aln

    opts = {}
    if args.table_slots:
        opts["table_slots"] = args.table_slots
    if args.grid_ctas:
        opts["grid_ctas"] = args.grid_ctas
    opts["hot_cache"] = args.hot_cache
    if args.chunk_len:
        opts["chunk_len"] = args.chunk_len
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # timing hygiene: inputs larger than L2 need no flush; smaller ones (cfg1) get a 512 MB write between steps,
    # outside the per-step event pairs
    col_bytes = 4 * len(names) * n_aln
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if col_bytes < (256 << 20) else None

    def timed_once(builder, columns, steps, finalize):
        """K steps bracketed by barrier + synchronize, CUDA events on the library's stream (one pair per step;
        the region's time is the sum of the steps, so an L2 flush between steps is not counted)."""
        barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        group_ms = []
        for i in range(steps):
            if flush_buf is not None:
                flush_buf.fill_(i & 0xFF)
            starts[i].record(stream)
            builder.reset()
            if n_cells:
                for a, b in zip(file_cuts[:-1], file_cuts[1:]):
                    builder.push(columns["read_group"][a:b], columns["target_idx"][a:b], columns["hap_idx"][a:b],
                                 columns["cell_idx"][a:b], order_base=order_base + a, drop_last_group=True)
            else:
                builder.push(columns["read_group"], columns["target_idx"], columns["hap_idx"], order_base=order_base)
            res = finalize(builder)
            group_ms.append(builder.stats()["group_ms"])
            ends[i].record(stream)
        barrier()
        per_step = [starts[i].elapsed_time(ends[i]) for i in range(steps)]
        ms = float(sum(per_step)) if flush_buf is not None else starts[0].elapsed_time(ends[-1])
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res, group_ms, per_step

    def timed(builder, columns, steps, finalize):
        """timed_once; a region disturbed by a one-off stall (one step more than 3x the median step, seen
        on shared boxes) is measured once more and the undisturbed region is kept, with a note."""
        ms, res, group_ms, per_step = timed_once(builder, columns, steps, finalize)
        note = None
        med = sorted(per_step)[len(per_step) // 2]
        flag = torch.tensor([1.0 if (steps >= 3 and max(per_step) > 3.0 * med) else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if flag.item() > 0:
            ms2, res, group_ms2, per_step2 = timed_once(builder, columns, steps, finalize)
            note = {"remeasured": True, "first_try_ms_per_step": ms / steps, "first_try_max_step_ms": max(per_step)}
            if ms2 < ms:
                ms, group_ms, per_step = ms2, group_ms2, per_step2
        return ms, res, group_ms, note

    # With several GPUs a step also runs the exchange: local ECs hash-partitioned to their owner rank
    # (NCCL all-to-all), owner-side merge, global EC ids and the final CSR on every rank
    # (alntools_b200/multi_gpu.py).  The owner context is reused from step to step like the local one.
    owner = None
    if world > 1:
        from alntools_b200 import multi_gpu
        owner = EcBuilder(wl["n_targets"], wl["n_haps"], alignments_hint=n_aln, device=local_rank, **opts)
        owner.set_stream(stream.cuda_stream)
        dev_t = torch.device("cuda", local_rank)

    pinned_out = {}

    class _Res(object):
        def __init__(self, d):
            self.n_ec = d["n_ec"]
            self.nnz_a = d["nnz_a"] if d.get("nnz_a") is not None else d["nnz_local"]

    def fin_device(b):
        if world == 1:
            return b.finalize_raw(min_count)
        owner.reset()
        return _Res(multi_gpu.distributed_finalize(b, lambda: owner, dev_t, result_on="slices"))

    def fin_host(b):
        if world == 1:
            return b.finalize_raw(min_count)
        owner.reset()
        out = multi_gpu.distributed_finalize(b, lambda: owner, dev_t, result_on="slices")
        # every rank brings its EC-id range of the matrices to its own pinned host memory (the ranks
        # write disjoint byte ranges of the EC file)
        for k in ("a_indptr", "a_indices", "a_data", "n_data"):
            t = out[k]
            if k not in pinned_out or pinned_out[k].numel() < t.numel():
                pinned_out[k] = torch.empty(int(t.numel() * 1.25) + 1, dtype=t.dtype, pin_memory=True)
            pinned_out[k][:t.numel()].copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        d2h_slices[0] = 4 * (int(out["n_ec_local"]) + 1) + 8 * int(out["nnz_local"]) + 4 * int(out["n_ec_local"])
        return _Res(out)

    d2h_slices = [0]

    # ---- device-resident arm ("value") --------------------------------------------------------------
    b_dev = EcBuilder(wl["n_targets"], wl["n_haps"], with_cells=bool(n_cells), alignments_hint=n_aln,
                      device=local_rank, result_on_device=1, **opts)
    b_dev.set_stream(stream.cuda_stream)
    # the sampler starts with the warm-up; only the samples of the timed region are reported
    with ClockSampler(local_rank) as clocks:
        timed_once(b_dev, dev, args.warmup, fin_device)
        clocks.mark()
        ms_dev, res_dev, group_ms, note_dev = timed(b_dev, dev, args.steps, fin_device)
    stats_dev = b_dev.stats()
    launches_per_step = stats_dev["kernel_launches"]  # stats are zeroed by reset(): this is the last step
    if owner is not None:
        launches_per_step += owner.stats()["kernel_launches"]
    n_ec, nnz_a = int(res_dev.n_ec), int(res_dev.nnz_a)
    if world > 1:   # the slices' non-zeros add up to the matrix's
        t = torch.tensor([nnz_a], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nnz_a = int(t.item())
    b_dev.close()

    # ---- end-to-end arm: host columns in, host matrices out ----------------------------------------
    b_e2e = EcBuilder(wl["n_targets"], wl["n_haps"], with_cells=bool(n_cells), alignments_hint=n_aln,
                      device=local_rank, result_on_device=1 if world > 1 else 0, **opts)
    b_e2e.set_stream(stream.cuda_stream)
    timed_once(b_e2e, host, args.warmup, fin_host)
    ms_e2e, res_e2e, _, note_e2e = timed(b_e2e, host, args.steps, fin_host)
    stats_e2e = b_e2e.stats()
    assert int(res_e2e.n_ec) == n_ec
    d2h_e2e = stats_e2e["d2h_bytes"]
    if world > 1:
        t = torch.tensor([d2h_slices[0]], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        d2h_e2e = d2h_e2e * world + int(t.item())            # all ranks: counters + their slice
    b_e2e.close()

    # ---- N > 1: parity of the sharded build, outside the timed region --------------------------------
    # every rank pushes a small shard of its own, the ranks merge through the same exchange as the timed
    # steps, and rank 0 compares the assembled matrices with the C oracle on the concatenated columns
    parity = None
    if world > 1 and not n_cells:
        sample_reads = max(1000, min(wl["n_reads"], 2_000_000 // world))
        sc = synth.make_columns(sample_reads, wl["n_targets"], wl["n_haps"], wl["seed"] + 104729 * (rank + 1), mode=wl["mode"])
        n_s = torch.zeros(world, dtype=torch.int64, device="cuda")
        n_s[rank] = len(sc["read_group"])
        dist.all_reduce(n_s)
        base_s = int(n_s[:rank].sum().item())
        b_par = EcBuilder(wl["n_targets"], wl["n_haps"], alignments_hint=len(sc["read_group"]), device=local_rank,
                          result_on_device=1, **opts)
        b_par.set_stream(stream.cuda_stream)
        b_par.push(sc["read_group"], sc["target_idx"], sc["hap_idx"], order_base=base_s)
        owner.reset()
        out = multi_gpu.distributed_finalize(b_par, lambda: owner, dev_t, result_on="slices")
        nl, zl = int(out["n_ec_local"]), int(out["nnz_local"])
        piece = (int(out["id_base"]), out["a_indptr"][:nl + 1].cpu().numpy(), out["a_indices"][:zl].cpu().numpy(),
                 out["a_data"][:zl].cpu().numpy(), out["n_data"][:nl].cpu().numpy(),
                 sc["read_group"], sc["target_idx"], sc["hap_idx"])
        pieces = [None] * world
        dist.all_gather_object(pieces, piece)      # indexed by rank = shard order = EC-id range order
        b_par.close()
        if rank == 0:
            from oracle import c_oracle          # the checker, never the thing measured
            assert [p[0] for p in pieces] == sorted(p[0] for p in pieces)
            got_indptr = np.concatenate([[0]] + [p[1][1:].astype(np.int64) + sum(len(q[2]) for q in pieces[:i])
                                                 for i, p in enumerate(pieces)])
            got = (got_indptr, np.concatenate([p[2] for p in pieces]), np.concatenate([p[3] for p in pieces]),
                   np.concatenate([p[4] for p in pieces]))
            # the shards' columns in rank order; read_group values only have to change between reads
            rg = np.concatenate([p[5].astype(np.int64) + 2 * sample_reads * i for i, p in enumerate(pieces)]).astype(np.int32)
            want = c_oracle.ec_from_columns(rg, np.concatenate([p[6] for p in pieces]), np.concatenate([p[7] for p in pieces]))
            equal = all(np.array_equal(g, w) for g, w in zip(got, want[:4]))
            parity = {"checked": True, "equal": bool(equal), "against": "C oracle (oracle/ec_oracle.c) on the shards' concatenated columns",
                      "sample_reads_per_gpu": sample_reads, "sample_alignments": int(len(rg)), "n_ec": int(len(got[3]))}
            if not equal:
                raise SystemExit("bench.py: the %d-GPU build differs from the oracle on the parity sample" % world)
    if owner is not None:
        owner.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_kind = measured_peak_gbs()
    gms = sorted(group_ms)[len(group_ms) // 2]
    # three (four with cells) int32 columns read once by the grouping kernel; group_ms is the LAST push's kernel
    algo_bytes = (16.0 if n_cells else 12.0) * (file_cuts[-1] - file_cuts[-2])
    achieved = algo_bytes / (gms * 1e-3) / 1e9
    traffic, traffic_source = measured_traffic(args.workload)
    line = {
        "metric": "bam2ec alignments/sec (EC build)",
        "value": total_aln * args.steps / (ms_dev * 1e-3),
        "unit": "alignments/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32",
        "data": "synthetic",
        "config": dict(config, alignments_per_gpu=n_aln, n_ec=n_ec, nnz_a=nnz_a,
                       l2_hygiene=("inputs of %d MB per GPU exceed the 126 MB L2; no flush needed" % (col_bytes >> 20)) if flush_buf is None
                       else ("inputs of %d MB per GPU fit the L2: a 512 MB buffer is written between steps, outside the timed event pairs" % (col_bytes >> 20)),
                       table_slots=stats_dev["table_slots"], table_grows=stats_dev["table_grows"]),
        "e2e": {"value": total_aln * args.steps / (ms_e2e * 1e-3), "unit": "alignments/s",
                "h2d_bytes_per_step": stats_e2e["h2d_bytes"] * world, "d2h_bytes_per_step": d2h_e2e,
                "ms_per_step": ms_e2e / args.steps,
                "host_binding": ("rank processes bound to the CPUs of their GPU's NUMA node (rank 0: %d CPUs)" % len(numa_cpus))
                if numa_cpus else "none"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "kernel": "ecb_group_insert_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_kind,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": gms,
                     "kernel_share_of_step": gms / (ms_dev / args.steps)},
        "clocks": clocks.summary(),
    }
    if parity is not None:
        line["parity"] = parity
    if note_dev:
        line["remeasured"] = note_dev
    if note_e2e:
        line["e2e"]["remeasured"] = note_e2e
    if world == 1 and not args.no_cpu_baseline:
        line["bam_e2e"] = run_bam_e2e(args.workload, args.bam_sample_reads, local_rank)
        line["cpu_baseline"] = run_cpu_arm(args.workload, 2, 0, args.cpu_sample_reads, max_seconds=60.0)
        # the like-for-like figure: BAM file -> EC file on both sides (decode, grouping, matrices, EC bytes);
        # `e2e` starts from decoded columns and is NOT comparable with the CPU arm, which decodes
        line["like_for_like"] = {"what": "bam_e2e.value / cpu_baseline.value: BAM file -> EC file on both sides",
                                 "ratio": line["bam_e2e"]["value"] / line["cpu_baseline"]["value"],
                                 "ratio_vs_1core": line["bam_e2e"]["value"] / line["cpu_baseline"]["value_1core"]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
