"""Single-sample bam2ec / bam2emase converter: same signature as alntools/bam_utils.convert
(bam_utils.py:512), with the grouping, counting, ordering and matrix build on the GPU.

Call stack: convert -> header tables (host) -> column emitter (host decode + filters + name compare)
-> EcBuilder.push (H2D + sm_100a kernels) -> EcBuilder.finalize (CSR A, CSC N) -> EC file writer.
"""
import multiprocessing
import collections
import os
import time

import numpy as np

from . import bamcols, bin_utils, emase, emitter, utils
from ._native import EcBuilder
from .header import TargetTables

LOG = utils.get_logger()


# ---- the reference's chunk-worker interface (bam_utils.py:30-62, 157-195, 198-363, 490-498) -------------
# convert() below does not use it - it decodes the file once and streams it to the GPU - but code that drives
# the reference's workers directly (a ConvertParams per process, wrapper_convert through Pool.imap, results
# merged in chunk order) finds the same names, fields and results here.
ParseRecord = collections.namedtuple("ParseRecord", ["header_size", "begin_read_offset", "begin_read_size",
                                                     "file_offset", "file_bytes", "end_read_offset", "end_read_size"])


class ConvertParams(object):
    slots = ['input_file', 'temp_dir', 'process_id', 'track_ranges', 'data']

    def __init__(self):
        self.input_file = None
        self.temp_dir = None
        self.process_id = None
        self.track_ranges = False
        self.data = []                      # tuples of (idx, ParseRecord)

    def __str__(self):
        return "Input: {}\nProcess ID: {}\nData: {}".format(self.input_file, self.process_id, self.data)


class ConvertResults(object):
    slots = ['valid_alignments', 'all_alignments', 'ec', 'unique_reads', 'init', 'tid_ranges']

    def __init__(self):
        self.valid_alignments = None
        self.all_alignments = None
        self.ec = None
        self.unique_reads = None
        self.init = False
        self.tid_ranges = None


def chunk_bam_file(bam_filename, new_filename, parse_rec):
    """A BAM file holding one chunk of `bam_filename` (bam_utils.py:157-195): the header blocks, the tail of
    the block the chunk starts in, the whole blocks in between, the head of the block it ends in, EOF."""
    from . import bam_io
    with open(bam_filename, "rb") as src, open(new_filename, "wb") as out:
        out.write(src.read(parse_rec.header_size))
        if parse_rec.begin_read_offset > 0:
            data = bam_io.read_virtual(bam_filename, parse_rec.begin_read_offset, parse_rec.begin_read_size)
            for off in range(0, len(data), 60000):
                out.write(bam_io.bgzf_block(data[off:off + 60000]))
        src.seek(parse_rec.file_offset)
        out.write(src.read(parse_rec.file_bytes))
        if parse_rec.end_read_offset > 0:
            data = bam_io.read_virtual(bam_filename, parse_rec.end_read_offset, parse_rec.end_read_size)
            for off in range(0, len(data), 60000):
                out.write(bam_io.bgzf_block(data[off:off + 60000]))
        out.write(bam_io.BGZF_EOF)


def process_convert_bam(cp, device=0):
    """One worker of the reference (bam_utils.py:198-363): the chunks of cp.data, in order, through the GPU
    EC build; returns ConvertResults with `ec` = OrderedDict 'tid,tid,...' (tids as strings, sorted as
    strings, :306) -> number of reads, in first-occurrence order, exactly as the reference's worker.
    unique_reads (log-only in the reference, every read name of the chunk) is left empty.  As in the
    reference, a chunk without a valid alignment ends the worker with what it has (:336-349)."""
    from collections import OrderedDict
    ret = ConvertResults()
    ret.valid_alignments = 0
    ret.all_alignments = 0
    ret.ec = OrderedDict()
    ret.unique_reads = {}
    ret.tid_ranges = {}
    builder = tables = None
    pushed = 0
    try:
        for idx, parse_record in cp.data:
            temp_file = os.path.join(cp.temp_dir, "_bam2ec.{}.bam".format(idx))
            utils.delete_file(temp_file)
            chunk_bam_file(cp.input_file, temp_file, parse_record)
            try:
                with bamcols.BamColumnReader(temp_file) as reader:
                    if tables is None:
                        tables = reader.build_tables(None)
                    else:
                        reader.set_tables(tables)
                    if cp.track_ranges:
                        reader.track_ranges(True)
                    cols = reader.read_all(chunk=1 << 23)
                    ret.all_alignments += reader.all_alignments
                    if cp.track_ranges:
                        lo, hi = reader.ranges()
                        for tid in np.flatnonzero(lo <= hi).tolist():        # bam_utils.py:272-286: keys are str(tid)
                            n, x = ret.tid_ranges.get(str(tid), (int(lo[tid]), int(hi[tid])))
                            ret.tid_ranges[str(tid)] = (min(n, int(lo[tid])), max(x, int(hi[tid])))
            finally:
                utils.delete_file(temp_file)
            n = len(cols["read_group"])
            if n == 0:
                LOG.error("Error: sequence item 0: expected str instance, NoneType found")   # what :348 logs
                break
            if builder is None:
                builder = EcBuilder(tables.num_targets, tables.num_haplotypes, with_cells=False, alignments_hint=0,
                                    device=device)
            builder.push(cols["read_group"], cols["target_idx"], cols["hap_idx"], order_base=pushed)
            pushed += n
            ret.valid_alignments += n
        if builder is not None and pushed:
            res = builder.finalize()
            # EC row (main target, haplotype mask) -> the tids of the reference's key
            tid_of = np.full((tables.num_targets, max(1, tables.num_haplotypes)), -1, dtype=np.int64)
            tid_of[tables.tid_target, tables.tid_hap] = np.arange(len(tables.tid_target))
            indptr, indices, data, counts = res["a_indptr"], res["a_indices"], res["a_data"], res["n_data"]
            for e in range(int(res["n_ec"])):
                tids = []
                for j in range(int(indptr[e]), int(indptr[e + 1])):
                    mask = int(data[j])
                    for h in range(tables.num_haplotypes):
                        if (mask >> h) & 1:
                            tids.append(str(int(tid_of[indices[j], h])))
                ret.ec[",".join(sorted(tids))] = int(counts[e])
    finally:
        if builder is not None:
            builder.close()
    return ret


def wrapper_convert(args):
    """As the reference's (bam_utils.py:490-498): unpack the argument tuple Pool.imap delivers."""
    return process_convert_bam(*args)


def _job_plan(num_chunks, number_processes):
    """bam_utils.py:530-544."""
    num_processes = multiprocessing.cpu_count() if number_processes <= 0 else number_processes
    if num_chunks <= 0:
        num_chunks = num_processes
    elif num_chunks > 1000:
        LOG.info("Modifying number of chunks from {} to 1000".format(num_chunks))
        num_chunks = 1000
    return num_chunks, min(num_processes, num_chunks)


def _convert_python_emitter(bam_filename, target_filename, device, start_time):
    """The same build fed by the record-level Python emitter (ALNTOOLS_B200_EMITTER=python)."""
    temp_time = time.time()
    header, records = emitter.read_bam(bam_filename)
    tables = TargetTables(header.references, header.lengths, target_filename)
    LOG.info("File parsed in {}, total time: {}".format(utils.format_time(temp_time, time.time()),
                                                        utils.format_time(start_time, time.time())))
    cols = emitter.emit_single(records, tables)
    if cols.valid_alignments == 0:
        raise RuntimeError("The shape must be a tuple of three positive integers.")
    with EcBuilder(tables.num_targets, tables.num_haplotypes, with_cells=False,
                   alignments_hint=cols.valid_alignments, device=device) as builder:
        builder.push(cols.read_group, cols.target_idx, cols.hap_idx, order_base=0)
        res = builder.finalize()
    return res, tables, cols.valid_alignments


def _requested_gpus(devices):
    """How many GPUs the caller asked for: the `devices` keyword, else the environment variable ALNTOOLS_GPUS
    (the CLI keeps the reference's options, so the environment is its only way to say it), else one."""
    if devices is not None:
        return max(1, int(devices))
    try:
        return max(1, int(os.environ.get("ALNTOOLS_GPUS", "1") or 1))
    except ValueError:
        raise ValueError("ALNTOOLS_GPUS must be a number of GPUs, found %r" % os.environ.get("ALNTOOLS_GPUS"))


def _under_launcher():
    """True inside a one-process-per-GPU job (torchrun sets RANK / WORLD_SIZE / LOCAL_RANK)."""
    return int(os.environ.get("WORLD_SIZE", "1") or 1) > 1 and "RANK" in os.environ


def _spawn_ranks(n_gpus, kwargs):
    """convert() on `n_gpus` GPUs from an ordinary call: one worker process per GPU (the contract of the
    multi-GPU path), each running convert_rank on its shard of the file; rank 0 reports the totals."""
    import json
    import socket
    import subprocess
    import sys
    import tempfile
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    with tempfile.TemporaryDirectory(prefix="alntools_b200_") as tmp:
        job = os.path.join(tmp, "job.json")
        with open(job, "w") as fh:
            json.dump(kwargs, fh)
        procs = []
        for rank in range(n_gpus):
            env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(n_gpus),
                       MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ALNTOOLS_B200_SUMMARY=os.path.join(tmp, "summary.json"),
                       ALNTOOLS_B200_VERBOSE="1" if LOG.isEnabledFor(20) else "0")
            procs.append(subprocess.Popen([sys.executable, "-m", "alntools_b200._rank_worker", job], env=env))
        # one worker that fails would leave the others waiting in a collective: when the first non-zero exit code
        # shows up the remaining workers (exactly the processes started above) are terminated
        codes = [None] * n_gpus
        while any(c is None for c in codes):
            for i, p in enumerate(procs):
                if codes[i] is None:
                    codes[i] = p.poll()
            if any(c not in (None, 0) for c in codes):
                for i, p in enumerate(procs):
                    if codes[i] is None:
                        p.terminate()
                for i, p in enumerate(procs):
                    if codes[i] is None:
                        try:
                            codes[i] = p.wait(timeout=20)
                        except subprocess.TimeoutExpired:
                            p.kill()
                            codes[i] = p.wait()
                break
            time.sleep(0.02)
        if any(codes):
            raise RuntimeError("multi-GPU bam2ec failed: worker exit codes %s" % codes)
        with open(os.path.join(tmp, "summary.json")) as fh:
            return json.load(fh)


def convert_rank(bam_filename, ec_filename, emase_filename, num_chunks=0, number_processes=-1, temp_dir=None,
                 range_filename=None, sample=None, target_filename=None):
    """One rank of the multi-GPU convert (one process per GPU; RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* in
    the environment, as torchrun sets them).  What the reference does with a process pool over chunk BAMs
    merged in chunk order (alntools/bam_utils.py:642-724) happens here as: every rank plans the same shards of
    the ONE file (virtual offsets at read boundaries, no temporary BAMs: bamcols_plan_shards), decodes and
    groups its own on its GPU, is moved to its global position in read order (ecb_rebase: the position is
    known once every rank has counted its alignments), and the ranks merge through the exchange of
    multi_gpu.distributed_finalize.  The final matrices stay partitioned by EC-id range; every rank writes its
    byte ranges of the EC file.  Returns the totals on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    from . import multi_gpu
    start_time = time.time()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    utils.bind_to_gpu_numa_node(local_rank)          # pinned buffers next to this rank's GPU
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", device_id=device)
    try:
        _, num_processes = _job_plan(num_chunks, number_processes)
        num_processes = max(1, num_processes // world)          # the host threads are shared by the ranks
        if sample is None:
            sample = os.path.basename(bam_filename)
        with bamcols.BamColumnReader(bam_filename, n_threads=num_processes) as reader:
            tables = reader.build_tables(target_filename)
            plan = reader.plan_shards(world)
            reader.set_range(plan[rank], plan[rank + 1])
            if range_filename is not None:
                reader.track_ranges(True)
            opts = dict(with_cells=False, alignments_hint=0, device=local_rank, result_on_device=1)
            with EcBuilder(tables.num_targets, tables.num_haplotypes, **opts) as builder, \
                    EcBuilder(tables.num_targets, tables.num_haplotypes, **opts) as owner:
                chunk_rows = int(min(1 << 21, max(1 << 16, os.path.getsize(bam_filename) // 2)))   # 24 MB of columns per piece
                valid = emitter.stream_single(reader, builder, chunk_rows=chunk_rows, pinned=True)
                mine = torch.tensor([valid, reader.all_alignments], dtype=torch.int64, device=device)
                every = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(every, mine)
                every = torch.stack(every).tolist()
                total_valid = sum(r[0] for r in every)
                if total_valid == 0:                              # as the single-GPU path (Sparse3DMatrix.py:45-46)
                    raise RuntimeError("The shape must be a tuple of three positive integers.")
                builder.rebase(sum(r[0] for r in every[:rank]))
                out = multi_gpu.distributed_finalize(builder, lambda: owner, device, result_on="slices")
                n_ec, id_base, n_local, nnz_local = int(out["n_ec"]), int(out["id_base"]), int(out["n_ec_local"]), int(out["nnz_local"])
                sizes = [torch.empty(2, dtype=torch.int64, device=device) for _ in range(world)]
                dist.all_gather(sizes, torch.tensor([n_local, nnz_local], dtype=torch.int64, device=device))
                sizes = torch.stack(sizes).tolist()
                nnz_base, nnz_total = sum(r[1] for r in sizes[:rank]), sum(r[1] for r in sizes)
                a_indptr = out["a_indptr"][:n_local + 1].cpu().numpy()
                a_indices = out["a_indices"][:nnz_local].cpu().numpy()
                a_data = out["a_data"][:nnz_local].cpu().numpy()
                counts = out["n_data"][:n_local].cpu().numpy()
                reads = torch.tensor([int(counts.astype(np.int64).sum())], dtype=torch.int64, device=device)
                dist.all_reduce(reads)
            section = tables.target_section()
            target_names = list(tables.main_targets.keys()) if (emase_filename or section is None) else None
            if ec_filename:
                header = bin_utils.ec_header_bytes(tables.haplotypes, target_names, tables.lengths, [sample],
                                                   target_section=section, n_targets=tables.num_targets)
                args = (ec_filename, header, n_ec, nnz_total, id_base, nnz_base, a_indptr, a_indices, a_data, counts)
                if rank == 0:
                    try:
                        os.remove(ec_filename)
                    except OSError:
                        pass
                    bin_utils.ecsave2_slice(*args, create=True)
                dist.barrier()                                    # the file is laid out
                if rank != 0:
                    bin_utils.ecsave2_slice(*args, create=False)
                dist.barrier()
            if emase_filename:                                    # one writer: the slices travel to rank 0
                parts = [None] * world
                dist.all_gather_object(parts, (a_indptr, a_indices, a_data, counts))
                if rank == 0:
                    indptr = np.concatenate([[0]] + [p[0][1:].astype(np.int64) + sum(len(q[1]) for q in parts[:i])
                                                     for i, p in enumerate(parts)]).astype(np.int32)
                    a_csr = (indptr, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]))
                    n_counts = np.concatenate([p[3] for p in parts])
                    n_csc = (np.array([0, n_ec], dtype=np.int32), np.arange(n_ec, dtype=np.int32), n_counts)
                    try:
                        os.remove(emase_filename)
                    except OSError:
                        pass
                    emase.save_emase(emase_filename, "bam2ec", (tables.num_targets, tables.num_haplotypes, n_ec),
                                     tables.haplotypes, target_names, tables.lengths, [sample], a_csr, n_csc,
                                     incidence_only=True)
            if range_filename is not None:                        # bam_utils.py:735-766, over all shards
                lo, hi = reader.ranges()
                lo_t, hi_t = torch.from_numpy(lo.copy()).to(device), torch.from_numpy(hi.copy()).to(device)
                dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
                if rank == 0:
                    utils.write_range_file(range_filename, list(tables.main_targets.keys()), tables.haplotypes,
                                           reader.references, lo_t.cpu().numpy(), hi_t.cpu().numpy())
        if rank != 0:
            return None
        LOG.info("# Valid Alignments: {:,}".format(total_valid))
        LOG.info("# Main Targets: {:,}".format(tables.num_targets))
        LOG.info("# Haplotypes: {:,}".format(tables.num_haplotypes))
        LOG.info("# Equivalence Classes: {:,}".format(n_ec))
        LOG.info("# Unique Reads: {:,}".format(int(reads.item())))
        LOG.info("{} GPUs, total time: {}".format(world, utils.format_time(start_time, time.time())))
        return {"n_ec": n_ec, "nnz_a": nnz_total, "n_samples": 1, "nnz_n": n_ec, "n_reads": int(reads.item()),
                "n_alignments": total_valid, "n_gpus": world,
                "alignments_per_gpu": [r[0] for r in every]}
    finally:
        if own_group and dist.is_initialized():
            dist.destroy_process_group()


def convert(bam_filename, ec_filename, emase_filename, num_chunks=0, number_processes=-1, temp_dir=None,
            range_filename=None, sample=None, target_filename=None, device=0, devices=None):
    """Convert a name-grouped BAM file into an EC binary file and/or an EMASE file.

    Arguments keep the reference's meaning.  num_chunks / number_processes only shape the host decode
    (the EC file does not depend on them, as in the reference); temp_dir is unused because no
    temporary BAM is written.  `device` selects the GPU.  devices=N (or the environment variable
    ALNTOOLS_GPUS=N) shards the file over N GPUs of this machine, one worker process per GPU (convert_rank);
    the same happens without either when the call is made inside a one-process-per-GPU job (torchrun).
    The EC file does not depend on the number of GPUs.
    """
    start_time = time.time()
    job = dict(bam_filename=bam_filename, ec_filename=ec_filename, emase_filename=emase_filename, num_chunks=num_chunks,
               number_processes=number_processes, temp_dir=temp_dir, range_filename=range_filename, sample=sample,
               target_filename=target_filename)
    if _under_launcher():
        if emitter.use_python_emitter():
            raise NotImplementedError("the multi-GPU path needs the native emitter (unset ALNTOOLS_B200_EMITTER)")
        return convert_rank(**job)
    if _requested_gpus(devices) > 1:
        if emitter.use_python_emitter():
            raise NotImplementedError("the multi-GPU path needs the native emitter (unset ALNTOOLS_B200_EMITTER)")
        return _spawn_ranks(_requested_gpus(devices), job)
    if range_filename is not None and emitter.use_python_emitter():
        raise NotImplementedError("--rangefile needs the native emitter (unset ALNTOOLS_B200_EMITTER)")
    num_chunks, num_processes = _job_plan(num_chunks, number_processes)
    if sample is None:                                        # bam_utils.py:552-554
        sample = os.path.basename(bam_filename)
        LOG.info("Sample not supplied, using filename: {}".format(sample))

    def save(res, tables, valid, temp_time):
        """Log the totals and write the outputs from the result arrays (bam_utils.py:726-876)."""
        LOG.info("All results combined in {}, total time: {}".format(utils.format_time(temp_time, time.time()),
                                                                     utils.format_time(start_time, time.time())))
        LOG.info("# Valid Alignments: {:,}".format(valid))
        LOG.info("# Main Targets: {:,}".format(tables.num_targets))
        LOG.info("# Haplotypes: {:,}".format(tables.num_haplotypes))
        LOG.info("# Equivalence Classes: {:,}".format(res["n_ec"]))
        LOG.info("# Unique Reads: {:,}".format(res["n_reads"]))
        a_csr = (res["a_indptr"], res["a_indices"], res["a_data"])
        n_csc = (res["n_indptr"], res["n_indices"], res["n_data"])
        section = tables.target_section() if hasattr(tables, "target_section") else None   # native tables only
        target_names = list(tables.main_targets.keys()) if (emase_filename or section is None) else None
        if emase_filename:
            LOG.info("Saving to {}...".format(emase_filename))
            try:
                os.remove(emase_filename)
            except OSError:
                pass
            emase.save_emase(emase_filename, "bam2ec", (tables.num_targets, tables.num_haplotypes, res["n_ec"]),
                             tables.haplotypes, target_names, tables.lengths, [sample], a_csr, n_csc,
                             incidence_only=True)
        if ec_filename:
            LOG.info("Saving to {}...".format(ec_filename))
            try:
                os.remove(ec_filename)
            except OSError:
                pass
            temp_time = time.time()
            bin_utils.ecsave2_arrays(ec_filename, tables.haplotypes, target_names, tables.lengths, [sample],
                                     a_csr, n_csc, target_section=section, n_targets=tables.num_targets)
            LOG.info("{} created in {}, total time: {}".format(ec_filename,
                                                               utils.format_time(temp_time, time.time()),
                                                               utils.format_time(start_time, time.time())))
        return {k: v for k, v in res.items() if not hasattr(v, "shape")}   # the scalar totals

    LOG.info("Parsing file information ...")
    temp_time = time.time()
    if emitter.use_python_emitter():
        res, tables, valid = _convert_python_emitter(bam_filename, target_filename, device, start_time)
        return save(res, tables, valid, temp_time)
    # native emitter: inflate threads + one record pass, streamed to the GPU in read-aligned chunks
    import torch
    with bamcols.BamColumnReader(bam_filename, n_threads=num_processes) as reader:
        tables = reader.build_tables(target_filename)       # native statement of header.TargetTables
        if range_filename is not None:
            reader.track_ranges(True)
        LOG.info("File parsed in {}, total time: {}".format(utils.format_time(temp_time, time.time()),
                                                            utils.format_time(start_time, time.time())))
        temp_time = time.time()
        # no alignment count is known up front: the library sizes its table from the first chunk and
        # grows it ahead of later ones; one finalize per context, so results skip the pinning cost
        with EcBuilder(tables.num_targets, tables.num_haplotypes, with_cells=False, alignments_hint=0,
                       device=device, pageable_results=1) as builder:
            chunk_rows = int(min(1 << 21, max(1 << 16, os.path.getsize(bam_filename) // 2)))   # 24 MB of columns per piece
            valid = emitter.stream_single(reader, builder, chunk_rows=chunk_rows, pinned=torch.cuda.is_available())
            if valid == 0:
                # the reference ends up with zero ECs and APM() raises (Sparse3DMatrix.py:45-46)
                raise RuntimeError("The shape must be a tuple of three positive integers.")
            # the files are written straight from the library's result buffers (no intermediate copies)
            summary = save(builder.finalize(copy=False), tables, valid, temp_time)
        if range_filename is not None:                         # bam_utils.py:735-766
            lo, hi = reader.ranges()
            utils.write_range_file(range_filename, list(tables.main_targets.keys()), tables.haplotypes,
                                   reader.references, lo, hi)
    return summary

