"""Import shim for the subset of Bio.bgzf the reference touches (bam_utils.py:157-195,1174-1367).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference run in a container without biopython.
Restates the published BGZF container format (SAM spec section 4.1).
"""
import struct
import zlib

_bgzf_magic = b"\x1f\x8b\x08\x04"
_bytes_BC = b"BC"
_bgzf_eof = (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00"
             b"\x1b\x00\x03\x00\x00\x00\x00\x00\x00\x00\x00\x00")
_BLOCK_PAYLOAD = 65280


def make_virtual_offset(block_start_offset, within_block_offset):
    return (block_start_offset << 16) | within_block_offset


def split_virtual_offset(virtual_offset):
    return virtual_offset >> 16, virtual_offset & 0xFFFF


def _deflate_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    return (_bgzf_magic + b"\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00"
            + struct.pack("<H", len(body) + 25) + body
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


class BgzfReader(object):
    def __init__(self, filename, mode="rb"):
        self._fh = open(filename, "rb")
        self._buf = b""
        self._pos = 0

    def _next_block(self):
        head = self._fh.read(18)
        if len(head) < 18:
            self._buf, self._pos = b"", 0
            return False
        size = struct.unpack("<H", head[16:18])[0] + 1
        body = self._fh.read(size - 18)
        self._buf, self._pos = zlib.decompress(body[:-8], -15), 0
        return True

    def seek(self, virtual_offset):
        self._fh.seek(virtual_offset >> 16)
        self._next_block()
        self._pos = virtual_offset & 0xFFFF

    def read(self, size):
        parts = []
        while size > 0:
            if self._pos >= len(self._buf):
                if not self._next_block():
                    break
                continue
            piece = self._buf[self._pos:self._pos + size]
            self._pos += len(piece)
            size -= len(piece)
            parts.append(piece)
        return b"".join(parts)

    def close(self):
        self._fh.close()


class BgzfWriter(object):
    def __init__(self, filename, mode="w"):
        self._fh = open(filename, "ab" if "a" in mode else "wb")
        self._buf = b""

    def write(self, data):
        self._buf += data
        while len(self._buf) >= _BLOCK_PAYLOAD:
            self._fh.write(_deflate_block(self._buf[:_BLOCK_PAYLOAD]))
            self._buf = self._buf[_BLOCK_PAYLOAD:]

    def close(self):
        if self._buf:
            self._fh.write(_deflate_block(self._buf))
            self._buf = b""
        self._fh.write(_bgzf_eof)
        self._fh.close()
