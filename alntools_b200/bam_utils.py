"""Single-sample bam2ec / bam2emase converter: same signature as alntools/bam_utils.convert
(bam_utils.py:512), with the grouping, counting, ordering and matrix build on the GPU.

Call stack: convert -> header tables (host) -> column emitter (host decode + filters + name compare)
-> EcBuilder.push (H2D + sm_100a kernels) -> EcBuilder.finalize (CSR A, CSC N) -> EC file writer.
"""
import multiprocessing
import os
import time

import numpy as np

from . import bamcols, bin_utils, emase, emitter, utils
from ._native import EcBuilder
from .header import TargetTables

LOG = utils.get_logger()


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
