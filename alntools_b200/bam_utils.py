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


def convert(bam_filename, ec_filename, emase_filename, num_chunks=0, number_processes=-1, temp_dir=None,
            range_filename=None, sample=None, target_filename=None, device=0):
    """Convert a name-grouped BAM file into an EC binary file and/or an EMASE file.

    Arguments keep the reference's meaning.  num_chunks / number_processes only shape the host decode
    (the EC file does not depend on them, as in the reference); temp_dir is unused because no
    temporary BAM is written.  `device` selects the GPU.
    """
    start_time = time.time()
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
            chunk_rows = int(min(1 << 23, max(1 << 16, os.path.getsize(bam_filename) // 2)))
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
