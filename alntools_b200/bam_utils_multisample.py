"""Per-cell (multisample) bam2ec / bam2emase converter: same signature as
alntools/bam_utils_multisample.convert (bam_utils_multisample.py:357).  One BAM file per task, the
cell id is field 14 of the '|||'-split read name, the last read of every file is dropped (:306-308),
cells below the minimum count and ECs left without cells are removed (:595-636).
"""
import glob
import os
import time

import numpy as np

from . import bamcols, bin_utils, emase, emitter, utils
from ._native import EcBuilder
from .header import TargetTables

LOG = utils.get_logger()


def convert_files(bam_files, ec_filename, emase_filename, minimum_count, target_filename=None, device=0,
                  range_filename=None):
    """The reference's convert() after its glob: files are merged in the given order
    (bam_utils_multisample.py:503-560)."""
    start_time = time.time()
    per_file = []
    total_valid = 0
    if emitter.use_python_emitter():
        header, _ = emitter.read_bam(bam_files[0])            # tables from the first file only (:399)
        tables = TargetTables(header.references, header.lengths, target_filename)
        cell_ids = {}
        for bam_file in bam_files:
            _, records = emitter.read_bam(bam_file)
            cols = emitter.emit_multisample(records, tables, cell_ids)
            per_file.append((cols.read_group, cols.target_idx, cols.hap_idx, cols.cell_idx))
            total_valid += cols.valid_alignments
        cell_names = [None] * len(cell_ids)
        for name, idx in cell_ids.items():
            cell_names[idx] = name
    else:
        cells = bamcols.CellDictionary()
        tables = None
        range_lo = range_hi = references = section = None
        for bam_file in bam_files:
            with bamcols.BamColumnReader(bam_file) as reader:
                if tables is None:                            # tables from the first file only (:399)
                    tables = reader.build_tables(target_filename)
                    section = tables.target_section()         # the EC file's targets block, while the reader lives
                    if range_filename is not None:
                        references = reader.references
                else:
                    reader.set_tables(tables)
                if range_filename is not None:
                    reader.track_ranges(True)
                c = reader.read_all(cells=cells, chunk=1 << 23)
                if range_filename is not None:                # merged over the files (:568-576)
                    lo, hi = reader.ranges()
                    range_lo = lo if range_lo is None else np.minimum(range_lo, lo)
                    range_hi = hi if range_hi is None else np.maximum(range_hi, hi)
            per_file.append((c["read_group"], c["target_idx"], c["hap_idx"], c["cell_idx"]))
            total_valid += len(c["read_group"])
        cell_names = cells.names()
        cells.close()
        if range_filename is not None:                        # :638-668
            utils.write_range_file(range_filename, list(tables.main_targets.keys()), tables.haplotypes,
                                   references, range_lo, range_hi)

    with EcBuilder(tables.num_targets, tables.num_haplotypes, with_cells=True,
                   alignments_hint=total_valid, device=device) as builder:
        base = 0
        for rg, tg, hp, cell in per_file:
            builder.push(rg, tg, hp, cell, order_base=base, drop_last_group=True)
            base += len(rg)
        res = builder.finalize(minimum_count)

    LOG.info("Number of alignments: {:,}".format(total_valid))
    LOG.info("Number of main targets: {:,}".format(tables.num_targets))
    LOG.info("Number of haplotypes: {:,}".format(tables.num_haplotypes))
    LOG.info("Number of ECs after filtering : {:,}".format(res["n_ec"]))
    LOG.info("Number of cells after filtering: {:,}".format(res["n_samples"]))
    sample_names = [cell_names[c] for c in res["cell_order"].tolist()]
    a_csr = (res["a_indptr"], res["a_indices"], res["a_data"])
    n_csc = (res["n_indptr"], res["n_indices"], res["n_data"])
    if emitter.use_python_emitter():
        section = None
    target_names = list(tables.main_targets.keys()) if (emase_filename or section is None) else None
    if emase_filename:
        try:
            os.remove(emase_filename)
        except OSError:
            pass
        emase.save_emase(emase_filename, "Multisample APM",
                         (tables.num_targets, tables.num_haplotypes, res["n_ec"]), tables.haplotypes,
                         target_names, tables.lengths, sample_names, a_csr, n_csc, incidence_only=False)
    if ec_filename:
        try:
            os.remove(ec_filename)
        except OSError:
            pass
        bin_utils.ecsave2_arrays(ec_filename, tables.haplotypes, target_names, tables.lengths, sample_names,
                                 a_csr, n_csc, target_section=section, n_targets=tables.num_targets)
        LOG.info("{} created, total time: {}".format(ec_filename, utils.format_time(start_time, time.time())))
    return res


def convert(bam_filename, ec_filename, emase_filename, num_chunks, minimum_count, number_processes, temp_dir,
            range_filename, target_filename, device=0):
    if os.path.isfile(bam_filename):                          # :375-377
        LOG.error('bam file must be a directory')
        return None
    bam_files = glob.glob(os.path.join(bam_filename, "*.bam"))  # :379 (filesystem order, as the reference)
    if len(bam_files) == 0:
        LOG.error('No bam files found in directory: {}'.format(bam_filename))
        return None
    if range_filename is not None and emitter.use_python_emitter():
        raise NotImplementedError("--rangefile needs the native emitter (unset ALNTOOLS_B200_EMITTER)")
    return convert_files(bam_files, ec_filename, emase_filename, minimum_count, target_filename, device,
                         range_filename)
