"""The output object of the reference, rebuilt from the device-built arrays.

alntools hands its writers an AlignmentPropertyMatrix (alntools/matrix/AlignmentPropertyMatrix.py:29-51,
Sparse3DMatrix.py:41-50) whose fields `ecsave2` (bin_utils.py:105-277) and `APM.save`
(AlignmentPropertyMatrix.py:507-532) read: num_loci, num_haplotypes, num_reads, num_samples, shape, hname,
lname, rname, sname, lengths, data[h] (scipy CSC, E x T, ones), count (scipy CSC, E x S), finalized.
`ApmArrays` carries exactly those, so that code written against the reference's object - including the
reference's own, unmodified writers - can consume a result of this package.  convert() itself does not build
it: it writes the files straight from the arrays (same bytes, no scipy round trip)."""
import numpy as np


class ApmArrays(object):
    """Field-compatible stand-in for the reference's finalized AlignmentPropertyMatrix."""

    def __init__(self, haplotypes, target_names, lengths, sample_names, a_csr, n_csc):
        from scipy.sparse import csc_matrix, csr_matrix
        a_indptr, a_indices, a_data = (np.asarray(x) for x in a_csr)
        n_indptr, n_indices, n_data = (np.asarray(x) for x in n_csc)
        n_ec = len(a_indptr) - 1
        self.num_loci = len(target_names)
        self.num_haplotypes = len(haplotypes)
        self.num_reads = n_ec
        self.num_samples = len(sample_names)
        self.num_groups = 0
        self.shape = (self.num_loci, self.num_haplotypes, self.num_reads)
        self.hname = list(haplotypes)
        self.lname = np.array(target_names)
        self.rname = np.arange(n_ec).astype(str)               # bam_utils.py:829
        self.sname = np.array(sample_names)
        self.lid = dict(zip(self.lname, np.arange(self.num_loci)))
        self.hid = dict(zip(self.hname, np.arange(self.num_haplotypes)))
        self.rid = dict(zip(self.rname, np.arange(self.num_reads)))
        self.lengths = np.asarray(lengths)
        self.finalized = True
        # data[h]: E x T incidence of haplotype h, CSC like Sparse3DMatrix.finalize() leaves it (:189-193)
        self.data = []
        for h in range(self.num_haplotypes):
            sel = ((a_data >> h) & 1) == 1
            rows = np.repeat(np.arange(n_ec, dtype=np.int64), np.diff(a_indptr))[sel]
            m = csr_matrix((np.ones(int(sel.sum())), (rows, a_indices[sel].astype(np.int64))),
                           shape=(n_ec, self.num_loci))
            self.data.append(m.tocsc())
        # count: E x S (bam_utils.py:833: csc_matrix(np.matrix(counts).T) for one sample)
        self.count = csc_matrix((n_data, n_indices, n_indptr), shape=(n_ec, self.num_samples))

    @classmethod
    def from_result(cls, result, tables, sample_names):
        """result: dict of EcBuilder.finalize(); tables: header tables (haplotypes, main_targets, lengths)."""
        return cls(tables.haplotypes, list(tables.main_targets.keys()), tables.lengths, sample_names,
                   (result["a_indptr"], result["a_indices"], result["a_data"]),
                   (result["n_indptr"], result["n_indices"], result["n_data"]))
