"""CPU ORACLE for the alntools bam2ec equivalence-class path.  *** TEST INFRASTRUCTURE ONLY ***

This module is the checker, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  `alntools_b200` must never
import anything under `oracle/`.

It restates, in plain Python/numpy, what the reference computes on the hot path.  Every function
cites the reference lines it follows (paths relative to /root/reference/alntools).

Parity pinning: the reference ships no runnable tests or golden vectors (tests/test_alntools.py
imports names that do not exist), so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
executed unmodified in the build container through import shims (oracle/run_reference.py); the
resulting EC files are committed under tests/golden/ with the script that made them
(oracle/make_golden.py) and tests/test_oracle_golden.py checks this module reproduces them
byte for byte.
"""
from collections import OrderedDict
import struct

import numpy as np


# --------------------------------------------------------------------------------------------
# A1  header -> target / haplotype tables            bam_utils.py:561-633 (multi: multisample:399-465)
# --------------------------------------------------------------------------------------------
def split_reference_name(name):
    """bam_utils.py:584-591: split at the LAST '_' unless it is at position 0 or absent."""
    i = name.rfind('_')
    if i > 0:
        return name[:i], name[i + 1:]
    return name, ''


def parse_targets(target_file):
    """utils.py:161-178."""
    targets = OrderedDict()
    with open(target_file, 'r') as fh:
        for line in fh:
            if line and line[0] == '#':
                continue
            targets[line.strip().split()[0]] = len(targets)
    return targets


class HeaderTables(object):
    """main_targets: OrderedDict name->idx; haplotypes: sorted list; tid_target/tid_hap: int32[nSQ];
    lengths: int32[T,H]."""

    def __init__(self, references, ref_lengths, target_filename=None):
        main_targets = OrderedDict()
        if target_filename:                                   # :571-579
            main_targets = parse_targets(target_filename)
        split = [split_reference_name(n) for n in references]
        haps = set()
        for target, hap in split:                             # :582-600
            haps.add(hap)
            if target not in main_targets:
                main_targets[target] = len(main_targets)
        self.haplotypes = sorted(haps)                        # :602
        hap_idx = {h: i for i, h in enumerate(self.haplotypes)}
        self.main_targets = main_targets
        self.references = list(references)
        self.lengths = np.zeros((len(main_targets), len(self.haplotypes)), dtype=np.int32)  # :605
        self.tid_target = np.zeros(len(references), dtype=np.int32)
        self.tid_hap = np.zeros(len(references), dtype=np.int32)
        for tid, (target, hap) in enumerate(split):           # :615-633
            self.lengths[main_targets[target], hap_idx[hap]] = ref_lengths[tid]
            self.tid_target[tid] = main_targets[target]
            self.tid_hap[tid] = hap_idx[hap]
        self._name_to_tid = {}
        for tid, n in enumerate(references):
            self._name_to_tid.setdefault(n, tid)

    def gettid(self, name):
        return self._name_to_tid.get(name, -1)


# --------------------------------------------------------------------------------------------
# A2  per-alignment filters and grouping            bam_utils.py:258-344
# --------------------------------------------------------------------------------------------
def alignment_is_valid(flag, tid, next_tid, next_pos):
    """bam_utils.py:264-270 (same in bam_utils_multisample.py:220-226)."""
    if flag & 0x4:                                            # is_unmapped
        return False
    if flag & 0x1:                                            # is_paired
        if (flag & 0x80) or not (flag & 0x2) or tid != next_tid or next_pos < 0:
            return False
    return True


def trim_name(name):
    """bam_utils.py:301-304: cut at the first space when its index is > 0."""
    i = name.find(' ')
    return name[:i] if i > 0 else name


class ChunkResult(object):
    def __init__(self):
        self.ec = OrderedDict()
        self.valid_alignments = 0
        self.all_alignments = 0
        self.failed = False


def _ec_key(tids):
    """bam_utils.py:307: key = ','.join(sorted(str tids)).  Only set membership matters later."""
    return ','.join(sorted(str(t) for t in tids))


def group_chunk_single(records):
    """One chunk of bam_utils.process_convert_bam (:253-344).

    records: iterable of (qname, flag, tid, pos, next_tid, next_pos).
    A chunk with no valid alignment makes the reference raise inside the final flush
    (`','.join([None])`), which its outer handler swallows (:348) -> `failed`.
    """
    res = ChunkResult()
    current = None
    tids = []
    for qname, flag, tid, _pos, ntid, npos in records:
        res.all_alignments += 1
        if not alignment_is_valid(flag, tid, ntid, npos):
            continue
        res.valid_alignments += 1
        name = trim_name(qname)
        if current is None:
            current = name
        if name != current:                                   # :306-320
            key = _ec_key(tids)
            res.ec[key] = res.ec.get(key, 0) + 1
            current = name
            tids = [tid]
        elif tid not in tids:                                 # :322-323
            tids.append(tid)
    if res.valid_alignments == 0:                             # :336-344 raises -> :348
        res.failed = True
        return res
    key = _ec_key(tids)
    res.ec[key] = res.ec.get(key, 0) + 1
    return res


def merge_single(chunk_results):
    """bam_utils.py:680-698: chunk-ordered merge; EC id = rank of first occurrence."""
    ec = OrderedDict()
    for res in chunk_results:
        for k, v in res.ec.items():
            ec[k] = ec.get(k, 0) + v
    return ec


# --------------------------------------------------------------------------------------------
# A4/A5/A7  EC dict -> A matrix (CSR of haplotype bitmasks)   bam_utils.py:788-847, bin_utils.py:208-232
# --------------------------------------------------------------------------------------------
def a_matrix_from_keys(keys, tables):
    """For each EC key (comma separated tid strings) emit the canonical CSR row.

    Follows bam_utils.py:788-825 literally (target names -> '<target>_<hap>' -> gettid ->
    membership test) and then does what scipy does at bin_utils.py:208-211: sum 2^h * data[h],
    convert to CSR with column indices ascending within a row.
    """
    indptr = [0]
    indices = []
    data = []
    inv_targets = tables.main_targets
    for key in keys:
        arr = key.split(',')
        members = set(arr)
        row = {}
        names = set()
        for t in arr:                                         # :792-794
            names.add(split_reference_name(tables.references[int(t)])[0])
        for main_target in names:                             # :797-819
            for i, hap in enumerate(tables.haplotypes):
                ref_name = main_target if len(hap) == 0 else '{}_{}'.format(main_target, hap)
                if str(tables.gettid(ref_name)) in members:
                    col = inv_targets[main_target]
                    row[col] = row.get(col, 0) + (1 << i)
        for col in sorted(row):
            indices.append(col)
            data.append(row[col])
        indptr.append(len(indices))
    return (np.asarray(indptr, dtype=np.int32), np.asarray(indices, dtype=np.int32),
            np.asarray(data, dtype=np.int32))


# --------------------------------------------------------------------------------------------
# A7  EC file bytes                                   bin_utils.py:105-277
# --------------------------------------------------------------------------------------------
def _i32(arr):
    return np.ascontiguousarray(arr, dtype='<i4').tobytes()


def ecsave2_bytes(haplotypes, target_names, lengths, sample_names, a_csr, n_csc):
    """Byte layout of bin_utils.ecsave2 (format 2)."""
    out = [struct.pack('<i', 2), struct.pack('<i', len(haplotypes))]          # :108,:131
    for hap in haplotypes:                                                     # :132-135
        out.append(struct.pack('<i', len(hap)))
        out.append(hap.encode('utf-8'))
    out.append(struct.pack('<i', len(target_names)))                           # :154
    for idx, name in enumerate(target_names):                                  # :155-159
        out.append(struct.pack('<i', len(name)))
        out.append(name.encode('utf-8'))
        out.append(_i32(lengths[idx, :len(haplotypes)]))
    out.append(struct.pack('<i', len(sample_names)))                           # :177
    for s in sample_names:                                                     # :178-180
        out.append(struct.pack('<i', len(s)))
        out.append(s.encode('utf-8'))
    for indptr, indices, data in (a_csr, n_csc):                               # :214-232, :259-275
        out.append(struct.pack('<i', len(indptr)))
        out.append(struct.pack('<i', len(indices)))
        out.append(_i32(indptr))
        out.append(_i32(indices))
        out.append(_i32(data))
    return b''.join(out)


def n_matrix_single(counts):
    """bam_utils.py:845: csc_matrix(np.matrix(counts).T) -> indptr [0,E], indices 0..E-1."""
    e = len(counts)
    return (np.asarray([0, e], dtype=np.int32), np.arange(e, dtype=np.int32),
            np.asarray(counts, dtype=np.int32))


def convert_single(references, ref_lengths, chunks_of_records, sample_name, target_filename=None):
    """Whole single-sample bam2ec path -> EC file bytes.  chunks_of_records: list of record iterables
    in chunk order.  Mirrors bam_utils.convert (:512-876) for ec_filename output."""
    tables = HeaderTables(references, ref_lengths, target_filename)
    ec = merge_single([group_chunk_single(c) for c in chunks_of_records])
    if len(ec) == 0:
        raise RuntimeError('The shape must be a tuple of three positive integers.')  # Sparse3DMatrix.py:45-46
    a_csr = a_matrix_from_keys(list(ec.keys()), tables)
    n_csc = n_matrix_single(list(ec.values()))
    return ecsave2_bytes(tables.haplotypes, list(tables.main_targets.keys()), tables.lengths,
                         [sample_name], a_csr, n_csc)


# --------------------------------------------------------------------------------------------
# A6  multisample (per-cell)                          bam_utils_multisample.py:175-321, 503-636, 702-791
# --------------------------------------------------------------------------------------------
def group_file_multisample(records):
    """bam_utils_multisample.process_convert_bam (:209-300).

    Differences from the single-sample worker, all reproduced: the cell id is field 14 of the
    '|||'-split name of the CURRENT GROUP (:270-280); after the first switch the remembered name is
    NOT trimmed (:292), so names with a space split every alignment into its own read; there is no
    flush at end of file (:306-308) so the last read is dropped.
    Returns OrderedDict key -> OrderedDict cell -> count, in first-occurrence order.
    """
    ec = OrderedDict()
    current = None
    tids = []
    for qname, flag, tid, _pos, ntid, npos in records:
        if not alignment_is_valid(flag, tid, ntid, npos):
            continue
        if current is None:
            current = trim_name(qname)
        cell = current.split('|||')[14]                       # :270-280 (IndexError propagates)
        if current != trim_name(qname):                       # :288
            key = _ec_key(tids)
            cells = ec.setdefault(key, OrderedDict())
            cells[cell] = cells.get(cell, 0) + 1
            current = qname                                   # :292 untrimmed
            tids = [tid]
        elif tid not in tids:
            tids.append(tid)
    return ec


def merge_multisample(file_results, minimum_count):
    """bam_utils_multisample.py:503-636.

    Returns (ec: OrderedDict key -> {cell: count} after filtering, cells: list of kept cell names in
    output order).  Cell order = insertion order of cr_totals (:513-551): for file, for EC in that
    file's first-occurrence order, for cell in first-occurrence order inside that (file, EC).
    """
    final = OrderedDict()
    cr_totals = OrderedDict()
    for res in file_results:
        for key, cells in res.items():
            for cell, count in cells.items():
                cr_totals[cell] = cr_totals.get(cell, 0) + count
                final.setdefault(key, {})
                final[key][cell] = final[key].get(cell, 0) + count
    if len(final) == 0:
        raise ValueError('max() arg is an empty sequence')    # :593
    if minimum_count <= 0:                                    # :596-597
        minimum_count = 1
    kept = OrderedDict()
    for cell, total in cr_totals.items():                     # :603-608
        if total >= minimum_count:
            kept[cell] = len(kept)
    out = OrderedDict()
    for key, cells in final.items():                          # :616-632
        sub = {c: n for c, n in cells.items() if c in kept}
        if sub:
            out[key] = sub
    return out, kept


def n_matrix_multisample(ec, kept):
    """bam_utils_multisample.py:737-747 (CSR rows sorted by cell idx) then .tocsc() (:791)."""
    num_cells = len(kept)
    cols = [[] for _ in range(num_cells)]
    for ec_id, cells in enumerate(ec.values()):
        for cell, count in cells.items():
            cols[kept[cell]].append((ec_id, count))
    indptr = [0]
    indices = []
    data = []
    for col in cols:
        for ec_id, count in col:                              # ec ids ascend by construction
            indices.append(ec_id)
            data.append(count)
        indptr.append(len(indices))
    return (np.asarray(indptr, dtype=np.int32), np.asarray(indices, dtype=np.int32),
            np.asarray(data, dtype=np.int32))


def convert_multisample(references, ref_lengths, files_of_records, minimum_count, target_filename=None):
    """Whole multisample bam2ec path -> EC file bytes (bam_utils_multisample.convert :357-820).
    Header tables come from the first file only (:399)."""
    tables = HeaderTables(references, ref_lengths, target_filename)
    ec, kept = merge_multisample([group_file_multisample(r) for r in files_of_records], minimum_count)
    if len(ec) == 0 or len(kept) == 0:
        raise RuntimeError('The shape must be a tuple of three positive integers.')
    a_csr = a_matrix_from_keys(list(ec.keys()), tables)
    n_csc = n_matrix_multisample(ec, kept)
    return ecsave2_bytes(tables.haplotypes, list(tables.main_targets.keys()), tables.lengths,
                         list(kept.keys()), a_csr, n_csc)


# --------------------------------------------------------------------------------------------
# Column-level restatement (the GPU kernel's contract, SURVEY 8a "GPU restatement")
# --------------------------------------------------------------------------------------------
def ec_from_columns(read_group, target_idx, hap_idx):
    """Single-sample EC build on int32 columns.

    read_group[A] non-decreasing; consecutive equal values = one read.  A read's key is its SET of
    (target, hap) pairs (duplicates collapse, bam_utils.py:322-325).  EC id = rank of the key's first
    occurrence in read order (:693-698); count = number of reads with that key.
    Returns (indptr[E+1], indices[Z], data[Z], counts[E]) int32; rows sorted by target, data = OR of
    1<<hap (bin_utils.py:208-211).
    """
    rg = np.asarray(read_group)
    t = np.asarray(target_idx, dtype=np.int64)
    h = np.asarray(hap_idx, dtype=np.int64)
    n = len(rg)
    ecs = OrderedDict()
    if n:
        starts = np.flatnonzero(np.concatenate(([True], rg[1:] != rg[:-1])))
        ends = np.concatenate((starts[1:], [n]))
        code = t * 64 + h
        for s, e in zip(starts.tolist(), ends.tolist()):
            key = np.unique(code[s:e]).tobytes()
            ecs[key] = ecs.get(key, 0) + 1
    indptr = [0]
    indices = []
    data = []
    for key in ecs:
        codes = np.frombuffer(key, dtype=np.int64)
        row = OrderedDict()
        for c in codes.tolist():                              # ascending by (target, hap)
            row[c // 64] = row.get(c // 64, 0) | (1 << (c % 64))
        indices.extend(row.keys())
        data.extend(row.values())
        indptr.append(len(indices))
    return (np.asarray(indptr, dtype=np.int32), np.asarray(indices, dtype=np.int32),
            np.asarray(data, dtype=np.int32), np.asarray(list(ecs.values()), dtype=np.int32))


def ec_from_columns_cells(pushes, minimum_count):
    """Multisample EC build on int32 columns.

    pushes: list (file order) of (read_group, target_idx, hap_idx, cell_idx, drop_last_group).
    cell_idx is per alignment; the read's cell is the value on its first alignment.
    Returns dict with A (indptr, indices, data), N csc (indptr, indices, data) and cell_order
    (original cell_idx in output column order).  Follows bam_utils_multisample.py:503-636,702-791.
    """
    per_file = []
    for rg, t, h, cell, drop_last in pushes:
        rg = np.asarray(rg)
        t = np.asarray(t, dtype=np.int64)
        h = np.asarray(h, dtype=np.int64)
        cell = np.asarray(cell)
        ec = OrderedDict()
        n = len(rg)
        if n:
            starts = np.flatnonzero(np.concatenate(([True], rg[1:] != rg[:-1])))
            ends = np.concatenate((starts[1:], [n]))
            if drop_last:
                starts, ends = starts[:-1], ends[:-1]
            code = t * 64 + h
            for s, e in zip(starts.tolist(), ends.tolist()):
                key = np.unique(code[s:e]).tobytes()
                cells = ec.setdefault(key, OrderedDict())
                c = int(cell[s])
                cells[c] = cells.get(c, 0) + 1
        per_file.append(ec)
    ec, kept = merge_multisample(per_file, minimum_count)
    indptr = [0]
    indices = []
    data = []
    for key in ec:
        codes = np.frombuffer(key, dtype=np.int64)
        row = OrderedDict()
        for c in codes.tolist():
            row[c // 64] = row.get(c // 64, 0) | (1 << (c % 64))
        indices.extend(row.keys())
        data.extend(row.values())
        indptr.append(len(indices))
    n_indptr, n_indices, n_data = n_matrix_multisample(ec, kept)
    return {
        'a_indptr': np.asarray(indptr, dtype=np.int32),
        'a_indices': np.asarray(indices, dtype=np.int32),
        'a_data': np.asarray(data, dtype=np.int32),
        'n_indptr': n_indptr, 'n_indices': n_indices, 'n_data': n_data,
        'cell_order': np.asarray(list(kept.keys()), dtype=np.int32),
    }
