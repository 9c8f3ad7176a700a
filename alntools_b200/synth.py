"""Seeded synthetic workloads for the bam2ec path.

Columns are what the host emitter hands to the device: per valid alignment an int32 read-group id
(non-decreasing), main-target index and haplotype index, optionally a cell index.  Transcript
popularity is Zipf(1.3) mod T; a read hits k transcripts, each on a uniformly random non-empty
subset of the H haplotypes.
"""
import numpy as np


def _reads_per_k(rng, n_reads, mode):
    """Number of transcripts per read."""
    if isinstance(mode, int):                       # fixed multimapping degree
        return np.full(n_reads, mode, dtype=np.int64)
    if mode == "light":                              # P(k>1) = 1/8, k in 2..4 then
        multi = rng.integers(0, 8, n_reads) == 0
        return np.where(multi, rng.integers(2, 5, n_reads), 1).astype(np.int64)
    if mode == "diploid":                            # cfg2: about 1.5 transcripts per read
        k = rng.geometric(0.65, n_reads)
        return np.minimum(k, 16).astype(np.int64)
    if mode == "heavy":                              # cfg3: geometric, mean about 8 transcripts, cap 256
        k = rng.geometric(0.12, n_reads)
        return np.minimum(k, 256).astype(np.int64)
    raise ValueError(mode)


def make_columns_fixed_degree(n_reads, n_targets, n_haps, seed, degree):
    """The multimapping-degree sweep (BASELINE.json configs[4]): every read has exactly `degree` alignments,
    each on a Zipf(1.3) transcript and a uniformly random haplotype (duplicates inside a read are possible
    and collapse, as in the reference)."""
    rng = np.random.default_rng(seed)
    n = int(n_reads) * int(degree)
    return {
        "read_group": np.repeat(np.arange(n_reads, dtype=np.int32), degree),
        "target_idx": (rng.zipf(1.3, n) % n_targets).astype(np.int32),
        "hap_idx": rng.integers(0, n_haps, n, dtype=np.int32),
        "n_reads": int(n_reads),
    }


def make_columns(n_reads, n_targets, n_haps, seed, mode="light", n_cells=0, dup_rate=0.0):
    """Return dict(read_group, target_idx, hap_idx[, cell_idx]) int32 arrays plus n_reads.

    dup_rate: fraction of alignments duplicated verbatim inside their read (same transcript hit at a
    second position), which the reference collapses (bam_utils.py:322-325).
    """
    if isinstance(mode, str) and mode.startswith("aln"):           # "aln<k>": exactly k alignments per read
        return make_columns_fixed_degree(n_reads, n_targets, n_haps, seed, int(mode[3:]))
    rng = np.random.default_rng(seed)
    k = _reads_per_k(rng, n_reads, mode)
    n_hits = int(k.sum())
    hit_read = np.repeat(np.arange(n_reads, dtype=np.int64), k)
    hit_target = (rng.zipf(1.3, n_hits) % n_targets).astype(np.int64)
    mask = rng.integers(1, 1 << n_haps, n_hits, dtype=np.int64)
    nbits = np.zeros(n_hits, dtype=np.int64)
    for h in range(n_haps):
        nbits += (mask >> h) & 1
    aln_hit = np.repeat(np.arange(n_hits, dtype=np.int64), nbits)
    # haplotype of each alignment: the j-th set bit of its hit's mask
    first = np.cumsum(nbits) - nbits
    j = np.arange(len(aln_hit), dtype=np.int64) - first[aln_hit]
    hap = np.zeros(len(aln_hit), dtype=np.int64)
    m = mask[aln_hit]
    seen = np.zeros(len(aln_hit), dtype=np.int64)
    for h in range(n_haps):
        bit = (m >> h) & 1
        sel = (bit == 1) & (seen == j)
        hap[sel] = h
        seen += bit
    rg = hit_read[aln_hit]
    tg = hit_target[aln_hit]
    if dup_rate > 0:
        dup = rng.random(len(rg)) < dup_rate
        reps = np.where(dup, 2, 1)
        rg, tg, hap = np.repeat(rg, reps), np.repeat(tg, reps), np.repeat(hap, reps)
    out = {
        "read_group": rg.astype(np.int32),
        "target_idx": tg.astype(np.int32),
        "hap_idx": hap.astype(np.int32),
        "n_reads": int(n_reads),
    }
    if n_cells:
        read_cell = (rng.zipf(1.2, n_reads) % n_cells).astype(np.int32)
        out["cell_idx"] = read_cell[rg]
    return out


def reference_names(n_targets, n_haps, hap_letters="ABCDEFGH"):
    """@SQ names 'ENSMUST%011d_<hap>' with tid = target * H + hap, lengths 1000 + target."""
    return [("ENSMUST%011d_%s" % (t, hap_letters[h]), 1000 + t)
            for t in range(n_targets) for h in range(n_haps)]


def columns_to_bam(filename, cols, n_targets, n_haps, level=1):
    """Materialise single-sample columns as a name-grouped BAM (read names 'read%09d')."""
    from . import bam_io
    tids = cols["target_idx"].astype(np.int64) * n_haps + cols["hap_idx"]
    bam_io.write_bam_columns(filename, reference_names(n_targets, n_haps), cols["read_group"],
                             np.zeros(len(tids), dtype=np.uint16), tids, level=level)
