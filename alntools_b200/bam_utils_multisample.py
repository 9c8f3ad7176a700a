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
        with EcBuilder(tables.num_targets, tables.num_haplotypes, with_cells=True,
                       alignments_hint=total_valid, device=device) as builder:
            base = 0
            for rg, tg, hp, cell in per_file:
                builder.push(rg, tg, hp, cell, order_base=base, drop_last_group=True)
                base += len(rg)
            res = builder.finalize(minimum_count)
    else:
        # native emitter: a worker thread decodes file i + 1 while file i is pushed (one push = one file: the file
        # is part of the (file, EC, cell) key and its last read is dropped); at most two files are in host memory
        import queue
        import threading
        cells = bamcols.CellDictionary()
        state = {"tables": None, "section": None, "references": None, "lo": None, "hi": None}
        ready = queue.Queue(maxsize=1)
        stop = threading.Event()
        first_tables = threading.Event()

        def decode():
            try:
                for bam_file in bam_files:
                    if stop.is_set():
                        return
                    with bamcols.BamColumnReader(bam_file) as reader:
                        if state["tables"] is None:              # tables from the first file only (:399)
                            state["tables"] = reader.build_tables(target_filename)
                            state["section"] = state["tables"].target_section()   # while the reader lives
                            if range_filename is not None:
                                state["references"] = reader.references
                            first_tables.set()
                        else:
                            reader.set_tables(state["tables"])
                        if range_filename is not None:
                            reader.track_ranges(True)
                        c = reader.read_all(cells=cells, chunk=1 << 23)
                        if range_filename is not None:            # merged over the files (:568-576)
                            lo, hi = reader.ranges()
                            state["lo"] = lo if state["lo"] is None else np.minimum(state["lo"], lo)
                            state["hi"] = hi if state["hi"] is None else np.maximum(state["hi"], hi)
                    item = (c["read_group"], c["target_idx"], c["hap_idx"], c["cell_idx"])
                    while not stop.is_set():
                        try:
                            ready.put(item, timeout=0.05)
                            break
                        except queue.Full:
                            continue
                ready.put(None)
            except BaseException as exc:                          # re-raised by the consumer
                first_tables.set()
                ready.put(exc)

        worker = threading.Thread(target=decode, daemon=True)
        worker.start()
        try:
            first_tables.wait()
            first = ready.get()
            if isinstance(first, BaseException):
                raise first
            tables, section = state["tables"], state["section"]
            with EcBuilder(tables.num_targets, tables.num_haplotypes, with_cells=True,
                           alignments_hint=0, device=device) as builder:
                base = 0
                item = first
                while item is not None:
                    rg, tg, hp, cell = item
                    builder.push(rg, tg, hp, cell, order_base=base, drop_last_group=True)
                    base += len(rg)
                    item = ready.get()
                    if isinstance(item, BaseException):
                        raise item
                total_valid = base
                res = builder.finalize(minimum_count)
        finally:
            # whatever happened, the decode thread has left the native library before the dictionary is closed
            stop.set()
            while worker.is_alive():
                try:
                    ready.get_nowait()
                except queue.Empty:
                    pass
                worker.join(timeout=0.05)
        cell_names = cells.names()
        cells.close()
        if range_filename is not None:                        # :638-668
            utils.write_range_file(range_filename, list(tables.main_targets.keys()), tables.haplotypes,
                                   state["references"], state["lo"], state["hi"])

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
