"""Import shim for the subset of pysam.AlignmentFile the reference touches
(bam_utils.py:96-98,253-304,561-633,1224-1247; bam_utils_multisample.py:209-292).

TEST INFRASTRUCTURE ONLY: a pure-Python BGZF/BAM record reader so the unmodified reference can
run in a container without pysam/htslib. Restates the published BAM format (SAM spec 4.2).
`tell()` follows htslib: once the current block is exhausted it reports the next block's start.
"""
import struct
import zlib


class _BgzfStream(object):
    def __init__(self, filename):
        self.fh = open(filename, "rb")
        self.block_start = 0
        self.buf = b""
        self.pos = 0
        self.eof = False
        self._load()

    def _load(self):
        self.block_start = self.fh.tell()
        head = self.fh.read(18)
        if len(head) < 18:
            self.buf, self.pos, self.eof = b"", 0, True
            return
        size = struct.unpack("<H", head[16:18])[0] + 1
        body = self.fh.read(size - 18)
        self.buf, self.pos, self.eof = zlib.decompress(body[:-8], -15), 0, False

    def read(self, size):
        parts = []
        while size > 0:
            if self.pos >= len(self.buf):
                if self.eof:
                    break
                self._load()
                continue
            piece = self.buf[self.pos:self.pos + size]
            self.pos += len(piece)
            size -= len(piece)
            parts.append(piece)
        return b"".join(parts)

    def tell(self):
        if self.pos >= len(self.buf) and not self.eof:
            return self.fh.tell() << 16
        return (self.block_start << 16) | self.pos

    def seek(self, virtual_offset):
        self.fh.seek(virtual_offset >> 16)
        self._load()
        self.pos = virtual_offset & 0xFFFF


class AlignedSegment(object):
    __slots__ = ("query_name", "flag", "reference_id", "reference_start",
                 "next_reference_id", "next_reference_start", "reference_name")

    @property
    def is_paired(self):
        return bool(self.flag & 0x1)

    @property
    def is_proper_pair(self):
        return bool(self.flag & 0x2)

    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    @property
    def is_read2(self):
        return bool(self.flag & 0x80)


class AlignmentFile(object):
    def __init__(self, filename, mode="rb"):
        self._s = _BgzfStream(filename)
        if self._s.read(4) != b"BAM\x01":
            raise ValueError("not a BAM file: %s" % filename)
        l_text = struct.unpack("<i", self._s.read(4))[0]
        self._s.read(l_text)
        n_ref = struct.unpack("<i", self._s.read(4))[0]
        names, lengths = [], []
        for _ in range(n_ref):
            l_name = struct.unpack("<i", self._s.read(4))[0]
            names.append(self._s.read(l_name)[:-1].decode())
            lengths.append(struct.unpack("<i", self._s.read(4))[0])
        self.references = tuple(names)
        self.lengths = tuple(lengths)
        self._tid = {}
        for i, n in enumerate(names):
            self._tid.setdefault(n, i)

    def get_tid(self, name):
        return self._tid.get(name, -1)

    gettid = get_tid

    def tell(self):
        return self._s.tell()

    def seek(self, virtual_offset):
        self._s.seek(virtual_offset)

    def close(self):
        self._s.fh.close()

    def __iter__(self):
        return self

    def __next__(self):
        raw = self._s.read(4)
        if len(raw) < 4:
            raise StopIteration
        rec = self._s.read(struct.unpack("<i", raw)[0])
        (ref_id, pos, l_name, _mapq, _bin, _ncig, flag, _lseq,
         next_ref, next_pos, _tlen) = struct.unpack_from("<iiBBHHHiiii", rec, 0)
        seg = AlignedSegment()
        seg.query_name = rec[32:32 + l_name - 1].decode()
        seg.flag = flag
        seg.reference_id = ref_id
        seg.reference_start = pos
        seg.next_reference_id = next_ref
        seg.next_reference_start = next_pos
        seg.reference_name = self.references[ref_id] if ref_id >= 0 else None
        return seg

    next = __next__
