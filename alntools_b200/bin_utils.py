"""EC binary file (format 2) writer/reader fed by device-built arrays.

Byte layout = alntools/bin_utils.ecsave2 (bin_utils.py:105-277): every integer a little-endian int32,
strings raw UTF-8 with an int32 length prefix, no padding:
    2 | H | H x [len, bytes] | T | T x [len, bytes, H x length] | S | S x [len, bytes]
    | A (CSR, E x T): len(indptr)=E+1, nnz, indptr, indices, data (haplotype bitmask)
    | N (CSC, E x S): len(indptr)=S+1, nnz, indptr, indices (EC ids), data (counts)
The reference packs every array with struct.pack('<{n}i', *arr) (one Python int per element);
ndarray.astype('<i4').tobytes() yields the same bytes.
"""
import struct

import numpy as np


def _i32(arr):
    """Little-endian int32 bytes of `arr` without a copy when it already is a contiguous int32 array
    (the library's result buffers are written to the file directly)."""
    a = np.ascontiguousarray(arr, dtype="<i4")
    return a.data if a.size else b""


def _name_bytes(name):
    """The bytes ecsave2 writes for a name: pack('<{len(name)}s', name.encode()) keeps len(name) BYTES, so a
    non-ASCII name is cut inside its UTF-8 encoding (bin_utils.py:131-135,153-159,177-180).  Same here, so the
    file stays byte-identical and the length prefix always matches what follows."""
    return name.encode("utf-8")[:len(name)]


def _target_section(target_names, lengths, n_haps):
    """T x [len(name), name bytes, H x length] (bin_utils.py:153-159).  ASCII names (every aligner
    index) are laid out with numpy scatters; anything else takes the per-name loop."""
    joined = "".join(target_names)
    n = len(target_names)
    if n == 0:
        return b""
    if not joined.isascii():
        parts = []
        for idx, name in enumerate(target_names):
            parts.append(struct.pack("<i", len(name)))
            parts.append(_name_bytes(name))
            parts.append(_i32(lengths[idx, :n_haps]))
        return b"".join(parts)
    nb = np.fromiter(map(len, target_names), dtype=np.int64, count=n)
    rec = 4 + nb + 4 * n_haps
    start = np.cumsum(rec) - rec
    out = np.empty(int(rec.sum()), dtype=np.uint8)
    out[start[:, None] + np.arange(4)] = nb.astype("<i4").view(np.uint8).reshape(n, 4)
    name_first = np.cumsum(nb) - nb
    out[np.repeat(start + 4 - name_first, nb) + np.arange(int(nb.sum()))] = np.frombuffer(joined.encode("ascii"),
                                                                                         dtype=np.uint8)
    if n_haps:
        lens = np.ascontiguousarray(lengths[:, :n_haps], dtype="<i4").view(np.uint8).reshape(n, 4 * n_haps)
        out[(start + 4 + nb)[:, None] + np.arange(4 * n_haps)] = lens
    return out.tobytes()


def ecsave2_arrays(ec_filename, haplotypes, target_names, lengths, sample_names, a_csr, n_csc, target_section=None,
                   n_targets=None):
    """a_csr / n_csc: (indptr, indices, data) int32 arrays.  target_section (with n_targets): the bytes of the
    targets section when the caller already has them (the native header tables build them without creating
    a Python string per target); then target_names / lengths are not looked at."""
    with open(ec_filename, "wb") as fh:
        fh.write(struct.pack("<i", 2))
        fh.write(struct.pack("<i", len(haplotypes)))
        for hap in haplotypes:
            fh.write(struct.pack("<i", len(hap)))
            fh.write(_name_bytes(hap))
        if target_section is not None:
            fh.write(struct.pack("<i", int(n_targets)))
            fh.write(target_section)
        else:
            lengths = np.asarray(lengths).astype(int)
            fh.write(struct.pack("<i", len(target_names)))
            fh.write(_target_section(target_names, lengths, len(haplotypes)))
        fh.write(struct.pack("<i", len(sample_names)))
        parts = []
        for sample in sample_names:
            parts.append(struct.pack("<i", len(sample)))
            parts.append(_name_bytes(sample))
        fh.write(b"".join(parts))
        for indptr, indices, data in (a_csr, n_csc):
            fh.write(struct.pack("<i", len(indptr)))
            fh.write(struct.pack("<i", len(indices)))
            fh.write(_i32(indptr))
            fh.write(_i32(indices))
            fh.write(_i32(data))


def ecload_arrays(ec_filename):
    """Inverse of ecsave2_arrays (what bin_utils.ecload :32-102 / bin_file.ECFile :99-402 parse)."""
    with open(ec_filename, "rb") as fh:
        buf = fh.read()
    pos = [0]

    def i32():
        v = struct.unpack_from("<i", buf, pos[0])[0]
        pos[0] += 4
        return v

    def text():
        n = i32()
        s = buf[pos[0]:pos[0] + n].decode("utf-8", errors="replace")   # (a cut non-ASCII name: see _name_bytes)
        pos[0] += n
        return s

    def arr(n):
        a = np.frombuffer(buf, dtype="<i4", count=n, offset=pos[0]).copy()
        pos[0] += 4 * n
        return a

    fmt = i32()
    if fmt != 2:
        raise ValueError("only EC format 2 is supported, found %d" % fmt)
    haplotypes = [text() for _ in range(i32())]
    n_targets = i32()
    targets = []
    lengths = np.zeros((n_targets, len(haplotypes)), dtype=np.int32)
    for t in range(n_targets):
        targets.append(text())
        lengths[t] = arr(len(haplotypes))
    samples = [text() for _ in range(i32())]
    mats = []
    for _ in range(2):
        n_ptr, nnz = i32(), i32()
        mats.append((arr(n_ptr), arr(nnz), arr(nnz)))
    if pos[0] != len(buf):
        raise ValueError("trailing bytes in EC file")
    return {"haplotypes": haplotypes, "targets": targets, "lengths": lengths, "samples": samples,
            "a": mats[0], "n": mats[1]}


def ec2emase(ec_filename, emase_filename):
    """EC binary file -> EMASE (.h5) file (alntools/bin_utils.py:979-995: ecload, then APM.save with the title
    'Converted from <ec file>' and the data arrays included).  Needs PyTables like every EMASE output."""
    import os
    from . import emase
    ec = ecload_arrays(ec_filename)
    try:
        os.remove(emase_filename)
    except OSError:
        pass
    shape = (len(ec["targets"]), len(ec["haplotypes"]), len(ec["a"][0]) - 1)
    emase.save_emase(emase_filename, "Converted from {}".format(ec_filename), shape, ec["haplotypes"], ec["targets"],
                     ec["lengths"], ec["samples"], ec["a"], ec["n"], incidence_only=False, as_ecload=True)
