"""Host column emitter: BAM records -> int32 columns for libecb200.

Does what alntools/bam_utils.process_convert_bam does up to (not including) the key build: decode,
the unmapped/paired-end filters (bam_utils.py:264-270), read-name trimming (:301-304) and the
name-change test (:306).  Per VALID alignment it emits
    read_group  (changes whenever the reference would start a new read)
    target_idx, hap_idx  (header lookup of the alignment's reference id)
    [cell_idx]  multisample only: dense id of field 14 of the '|||'-split group name
                (bam_utils_multisample.py:270-280)
Everything after that - grouping, de-duplication, counting, ordering, matrix build - runs on the GPU.
"""
import numpy as np

from . import bam_io


def _trim(name):
    i = name.find(" ")
    return name[:i] if i > 0 else name


def _valid(flag, tid, ntid, npos):
    if flag & 0x4:
        return False
    if flag & 0x1 and ((flag & 0x80) or not (flag & 0x2) or tid != ntid or npos < 0):
        return False
    return True


class Columns(object):
    __slots__ = ("read_group", "target_idx", "hap_idx", "cell_idx", "all_alignments", "valid_alignments",
                 "n_groups")

    def __init__(self, rg, tg, hp, cell, all_alignments, n_groups):
        self.read_group = np.asarray(rg, dtype=np.int32)
        self.target_idx = np.asarray(tg, dtype=np.int32)
        self.hap_idx = np.asarray(hp, dtype=np.int32)
        self.cell_idx = None if cell is None else np.asarray(cell, dtype=np.int32)
        self.all_alignments = all_alignments
        self.valid_alignments = len(self.read_group)
        self.n_groups = n_groups


def emit_single(records, tables):
    """records: iterable of (qname, flag, tid, pos, next_tid, next_pos); bam_utils.py:258-328."""
    rg, tids = [], []
    current = None
    group = -1
    total = 0
    for qname, flag, tid, _pos, ntid, npos in records:
        total += 1
        if not _valid(flag, tid, ntid, npos):
            continue
        name = _trim(qname)
        if name != current or group < 0:
            current = name
            group += 1
        rg.append(group)
        tids.append(tid)
    tids = np.asarray(tids, dtype=np.int64)
    return Columns(rg, tables.tid_target[tids] if len(tids) else [], tables.tid_hap[tids] if len(tids) else [],
                   None, total, group + 1)


def emit_multisample(records, tables, cell_ids):
    """One BAM file of the per-cell path (bam_utils_multisample.py:209-300).

    cell_ids: dict cell name -> dense int, shared by all files of the job and extended here.
    Quirks kept: the remembered group name is trimmed only for the first group of a file (:258-262 vs
    :292), the cell comes from the remembered name, and a name with fewer than 15 '|||' fields raises
    IndexError exactly where the reference does.  The last read is NOT dropped here; the caller pushes
    the file with drop_last_group=1 (:306-308).
    """
    rg, tids, cells = [], [], []
    current = None
    group = -1
    cell = 0
    pending = False   # the current group's cell is resolved when the reference would: at the next valid alignment
    total = 0
    for qname, flag, tid, _pos, ntid, npos in records:
        total += 1
        if not _valid(flag, tid, ntid, npos):
            continue
        if current is None:
            current = _trim(qname)
            group = 0
            cell = cell_ids.setdefault(current.split("|||")[14], len(cell_ids))
        elif pending:
            cell = cell_ids.setdefault(current.split("|||")[14], len(cell_ids))   # :270-280
            cells[-1] = cell
            pending = False
        if current != _trim(qname):                                              # :288
            current = qname                                                      # :292 (untrimmed)
            group += 1
            pending = True
            cell = 0
        rg.append(group)
        tids.append(tid)
        cells.append(cell)
    tids = np.asarray(tids, dtype=np.int64)
    return Columns(rg, tables.tid_target[tids] if len(tids) else [], tables.tid_hap[tids] if len(tids) else [],
                   cells, total, group + 1)


def use_python_emitter():
    """ALNTOOLS_B200_EMITTER=python selects the record-level Python emitter above (slow; kept as the
    readable statement of the rules and as a cross-check); the default is the native one (bamcols)."""
    import os
    return os.environ.get("ALNTOOLS_B200_EMITTER", "native").lower() == "python"


_PINNED_POOL = []      # pinned buffer sets of finished conversions, reused by the next one in this process


def _column_buffers(rows, with_cells, pinned):
    n = 4 if with_cells else 3
    if pinned:
        import torch
        # pinning host memory costs about as much as decoding it: a set that an earlier convert() of this
        # process has pinned is taken again when it is large enough
        for i, bufs in enumerate(_PINNED_POOL):
            if len(bufs) == n and int(bufs[0].shape[0]) >= rows:
                return _PINNED_POOL.pop(i)
        return [torch.empty(rows, dtype=torch.int32, pin_memory=True) for _ in range(n)]
    return [np.empty(rows, dtype=np.int32) for _ in range(n)]


def _release_buffers(sets, pinned):
    """Buffer sets of a finished stream go back to the pool (at most 8 sets are kept)."""
    if pinned:
        for bufs in sets:
            if len(_PINNED_POOL) < 8:
                _PINNED_POOL.append(bufs)


def stream_single(reader, builder, chunk_rows=1 << 21, pinned=True, depth=3):
    """Decode `reader` (bamcols.BamColumnReader, tables set) on a worker thread into a small ring of
    pinned buffers and push every filled buffer to the GPU while the next one is being decoded.
    Chunks end on read boundaries (bamcols_emit never splits a read), order_base = rows pushed so far,
    so the result equals one push of the whole file.  Returns the number of valid alignments."""
    import queue
    import threading
    free, full = queue.Queue(), queue.Queue()
    allocated = [0]
    all_sets = []
    stop = threading.Event()
    rows = [int(chunk_rows)]

    def next_buffers():
        """A free buffer set; a new one is pinned only when all existing ones are in flight."""
        while not stop.is_set():
            try:
                bufs = free.get_nowait()
            except queue.Empty:
                if allocated[0] < depth:
                    allocated[0] += 1
                    bufs = _column_buffers(rows[0], False, pinned)
                    all_sets.append(bufs)
                    return bufs
                try:
                    bufs = free.get(timeout=0.05)
                except queue.Empty:
                    continue
            if int(bufs[0].shape[0]) >= rows[0]:
                return bufs
            allocated[0] -= 1                             # a set from before the buffers grew: let it go
        return None

    def produce():
        try:
            while not stop.is_set():
                bufs = next_buffers()
                if bufs is None:
                    return
                try:
                    n, done = reader.emit(bufs[0], bufs[1], bufs[2])
                except ValueError as exc:                 # a read longer than the buffers hold: grow and retry
                    if "buffers hold" in str(exc) and rows[0] < (1 << 28):   # (as BamColumnReader.read_all does)
                        rows[0] *= 4
                        allocated[0] -= 1
                        continue
                    raise
                full.put((bufs, n, done, None))
                if done:
                    return
        except BaseException as exc:                      # re-raised by the consumer
            full.put((None, 0, True, exc))

    worker = threading.Thread(target=produce, daemon=True)
    worker.start()
    pushed = 0
    try:
        while True:
            bufs, n, done, exc = full.get()
            if exc is not None:
                raise exc
            if n:
                builder.push(bufs[0], bufs[1], bufs[2], order_base=pushed, n=n)
                pushed += n
            free.put(bufs)
            if done:
                break
    finally:
        # whatever happened, the producer has left native code before the caller may close the reader:
        # bamcols_close frees what bamcols_emit is working on
        stop.set()
        worker.join()
        _release_buffers(all_sets, pinned)
    return pushed


def read_bam(filename):
    """(BamHeader, record iterator) using the built-in reader."""
    raw = bam_io.inflate_file(filename)
    header = bam_io.parse_header(raw)
    return header, bam_io.iter_records(raw, header.records_offset)
