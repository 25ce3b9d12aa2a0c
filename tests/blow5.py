"""TEST INFRASTRUCTURE: a minimal pure-Python reader of BLOW5 files (slow5lib v1.x binary format), used as the checker
of the device-side record decode and to feed the real-data fixtures of tests/golden/ecoli to the GPU tests.

Layout (reference slow5lib/src/slow5.c:794-870, slow5_press.c): magic "BLOW5\\1", version major.minor.patch (3 bytes),
record compression (1 byte: 0 none, 1 zlib, 2 svb-zd [signal only], 3 zstd), number of read groups (uint32), from
v0.2.0 the signal compression (1 byte), padding up to byte 64, uint32 size of the text header, the text header, then
records — each a uint64 byte count followed by that many bytes (compressed as a whole with the record method) — and the
trailer "5WOLB". A decompressed record is: uint16 len + read_id, uint32 read_group, double digitisation, offset, range,
sampling_rate, uint64 len_raw_signal, the raw signal (int16 x len, or svb-zd compressed with its own length prefix from
v0.2.0 on), then the auxiliary fields.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

RECORD_NONE, RECORD_ZLIB, RECORD_SVB_ZD, RECORD_ZSTD = 0, 1, 2, 3


class Blow5:
    def __init__(self, path: str):
        self.data = open(path, "rb").read()
        d = self.data
        assert d[:6] == b"BLOW5\x01", "not a BLOW5 file"
        self.version = tuple(d[6:9])
        self.record_method = d[9]
        self.n_read_groups = struct.unpack("<I", d[10:14])[0]
        self.signal_method = d[14] if self.version >= (0, 2, 0) else 0
        hs = struct.unpack("<I", d[64:68])[0]
        self.header_text = d[68:68 + hs].decode(errors="replace")
        p = 68 + hs
        self.records = []                     # (offset of the compressed bytes, size)
        while p + 8 <= len(d) and d[p:p + 5] != b"5WOLB":
            (sz,) = struct.unpack("<Q", d[p:p + 8])
            self.records.append((p + 8, sz))
            p += 8 + sz
        assert d[p:p + 5] == b"5WOLB", "missing BLOW5 end-of-file marker"

    def record_bytes(self, i: int) -> bytes:
        o, sz = self.records[i]
        return self.data[o:o + sz]

    def decompress(self, i: int) -> bytes:
        raw = self.record_bytes(i)
        if self.record_method == RECORD_NONE:
            return raw
        if self.record_method == RECORD_ZLIB:
            return zlib.decompress(raw)
        raise NotImplementedError("record compression %d" % self.record_method)

    def read(self, i: int):
        """(read_id, digitisation, offset, range, sampling_rate, int16 signal) of record i (host decode: the checker)."""
        rec = self.decompress(i)
        (l,) = struct.unpack("<H", rec[:2])
        rid = rec[2:2 + l].rstrip(b"\0").decode()
        q = 2 + l
        rg, dig, off, rng, sr, n = struct.unpack("<IddddQ", rec[q:q + 44])
        q += 44
        if self.signal_method == 0:
            sig = np.frombuffer(rec[q:q + 2 * n], dtype="<i2").copy()
        elif self.signal_method == 1:     # svb-zd: len_raw_signal holds the BYTES of the compressed signal
            sig = svb_zd_decode(rec[q:q + n])
        else:
            raise NotImplementedError("signal compression %d" % self.signal_method)
        return rid, dig, off, rng, sr, sig

    def __len__(self):
        return len(self.records)


def read_fasta(path: str):
    out, name, buf = [], None, []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if name is not None:
                out.append((name, "".join(buf)))
            name, buf = line[1:].split()[0], []
        elif line:
            buf.append(line)
    if name is not None:
        out.append((name, "".join(buf)))
    return out


def svb_zd_decode(buf: bytes) -> np.ndarray:
    """Host restatement of slow5lib's svb-zd signal decompression (slow5_press.c:1118-1170, streamvbyte_decode.c,
    streamvbyte_zigzag.c:34-40), numpy: [uint32 count][2-bit keys, 4 per byte: bytes-1][little-endian values]; value ->
    zigzag decode -> running sum, truncated to int16."""
    (count,) = struct.unpack("<I", buf[:4])
    nkeys = (count + 3) // 4
    keys = np.frombuffer(buf[4:4 + nkeys], dtype=np.uint8)
    codes = ((keys[:, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(-1)[:count].astype(np.int64)
    lens = codes + 1
    offs = np.zeros(count, dtype=np.int64)
    np.cumsum(lens[:-1], out=offs[1:])
    data = np.frombuffer(buf[4 + nkeys:], dtype=np.uint8).astype(np.uint32)
    data = np.concatenate([data, np.zeros(4, dtype=np.uint32)])
    val = data[offs].copy()
    for j in range(1, 4):
        val |= np.where(lens > j, data[offs + j] << np.uint32(8 * j), np.uint32(0))
    assert int(offs[-1] + lens[-1]) == len(buf) - 4 - nkeys if count else True
    zz = (val >> np.uint32(1)).astype(np.int64) ^ -(val & np.uint32(1)).astype(np.int64)
    return np.cumsum(zz).astype(np.int64).astype(np.int16) if count else np.zeros(0, dtype=np.int16)
