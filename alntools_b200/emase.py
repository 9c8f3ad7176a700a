"""EMASE (.h5) output of the alignment-property matrix (alntools/matrix/Sparse3DMatrix.py:325-342,
AlignmentPropertyMatrix.py:507-532).  Needs PyTables, which this image does not ship: the writer is
complete but only runs where `import tables` works; otherwise it raises (never writes a partial file).
"""
import numpy as np


def split_haplotype_csc(a_indptr, a_indices, a_data, n_targets, n_haps):
    """CSR bitmask matrix (E x T) -> per-haplotype CSC incidence (indptr[T+1], indices) like
    APM.finalize() leaves apm.data[h] (Sparse3DMatrix.py:189-193)."""
    n_ec = len(a_indptr) - 1
    rows = np.repeat(np.arange(n_ec, dtype=np.int64), np.diff(a_indptr))
    out = []
    for h in range(n_haps):
        sel = (a_data >> h) & 1 == 1
        cols = a_indices[sel].astype(np.int64)
        r = rows[sel]
        order = np.lexsort((r, cols))
        indptr = np.zeros(n_targets + 1, dtype=np.int64)
        np.add.at(indptr, cols + 1, 1)
        out.append((np.cumsum(indptr), r[order]))
    return out


def save_emase(h5file, title, shape, haplotypes, target_names, lengths, sample_names, a_csr, n_csc,
               incidence_only=True, as_ecload=False):
    """as_ecload: the object the reference saves in `ec2emase` comes from `ecload` (bin_utils.py:32-95), not from
    bam2ec, and differs in what it carries: lengths as float64, haplotype names as an array, no read names, and
    for a single sample the counts as a plain float64 vector instead of a CSC group."""
    try:
        import tables
    except ImportError as exc:
        raise RuntimeError("EMASE (.h5) output needs PyTables, which is not installed: %s" % exc)
    n_targets, n_haps, n_ec = shape
    per_hap = split_haplotype_csc(a_csr[0], a_csr[1], a_csr[2], n_targets, n_haps)
    h5 = tables.open_file(h5file, "w", title=title)
    fil = tables.Filters(complevel=1, complib="zlib")
    h5.set_node_attr(h5.root, "incidence_only", incidence_only)
    h5.set_node_attr(h5.root, "mtype", "csc_matrix")
    h5.set_node_attr(h5.root, "shape", shape)
    for hid, (indptr, indices) in enumerate(per_hap):
        grp = h5.create_group(h5.root, "h%d" % hid, "Sparse matrix components for Haplotype %d" % hid)
        h5.create_carray(grp, "indptr", obj=indptr.astype("uint32"), filters=fil)
        h5.create_carray(grp, "indices", obj=indices.astype("uint32"), filters=fil)
        if not incidence_only:
            h5.create_carray(grp, "data", obj=np.ones(len(indices), dtype=float), filters=fil)
    h5.create_carray(h5.root, "lengths", obj=np.asarray(lengths, dtype=float) if as_ecload else np.asarray(lengths),
                     title="Transcript Lengths", filters=fil)
    if as_ecload and len(sample_names) == 1:
        counts = np.zeros(n_ec, dtype=float)                       # csc (E x 1) -> dense vector (bin_utils.py:91-92)
        counts[np.asarray(n_csc[1])] = np.asarray(n_csc[2])
        h5.create_carray(h5.root, "count", obj=counts, title="Equivalence Class Counts", filters=fil)
    else:
        cgrp = h5.create_group(h5.root, "count", "Sparse matrix components for N matrix")
        for name, arr in zip(("indptr", "indices", "data"), n_csc):
            h5.create_carray(cgrp, name, obj=np.asarray(arr).astype("uint32"), filters=fil)
    h5.set_node_attr(h5.root, "hname", np.array(haplotypes, dtype=str) if as_ecload else haplotypes)
    h5.create_carray(h5.root, "lname", obj=np.array(target_names, dtype=str) if as_ecload else np.array(target_names),
                     title="Locus Names", filters=fil)
    if not as_ecload:
        h5.create_carray(h5.root, "rname", obj=np.arange(n_ec).astype(str), title="Read Names", filters=fil)
    h5.create_carray(h5.root, "sname", obj=np.array(sample_names, dtype=str) if as_ecload else np.array(sample_names),
                     title="Sample Names", filters=fil)
    h5.flush()
    h5.close()
