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
        elif self.signal_method == 2:     # ex-zd, likewise
            sig = ex_zd_decode(rec[q:q + n])
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


def svb_decode_u32(buf: bytes, count: int) -> np.ndarray:
    """Plain streamvbyte (slow5lib's __slow5_streamvbyte_decode): keys, then 1..4 little-endian bytes per value."""
    nkeys = (count + 3) // 4
    keys = np.frombuffer(buf[:nkeys], dtype=np.uint8)
    codes = ((keys[:, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(-1)[:count].astype(np.int64)
    lens = codes + 1
    offs = np.zeros(count, dtype=np.int64)
    np.cumsum(lens[:-1], out=offs[1:])
    data = np.concatenate([np.frombuffer(buf[nkeys:], dtype=np.uint8).astype(np.uint32), np.zeros(4, dtype=np.uint32)])
    val = data[offs].copy()
    for j in range(1, 4):
        val |= np.where(lens > j, data[offs + j] << np.uint32(8 * j), np.uint32(0))
    assert int(offs[-1] + lens[-1]) == len(buf) - nkeys
    return val


def ex_zd_decode(buf: bytes) -> np.ndarray:
    """Host restatement of slow5lib's ex-zd signal decompression (slow5_press.c:1420-1530 ex_depress, :1646-1672
    ex_zd_depress_16, :1787-1842 the v0 framing): u8 version, u64 n, u8 q, u16 first zigzag delta, u32 number of
    exceptions, their positions (streamvbyte of the gaps - 1) and values - 256 (streamvbyte), each behind a u32 size —
    or two plain u32 when there is exactly one — then a byte per non-exception; zigzag decode, running sum in int16,
    shift left by q."""
    assert buf[0] == 0, "ex-zd version"
    n, q = struct.unpack("<QB", buf[1:10])
    (first,) = struct.unpack("<H", buf[10:12])
    (nex,) = struct.unpack("<I", buf[12:16])
    off = 16
    zd = np.zeros(n, dtype=np.uint32)
    zd[0] = first
    is_ex = np.zeros(n - 1, dtype=bool)
    if nex > 1:
        (pl,) = struct.unpack("<I", buf[off:off + 4]); off += 4
        gaps = svb_decode_u32(buf[off:off + pl], nex).astype(np.int64); off += pl
        pos = np.cumsum(gaps + 1) - 1
        (vl,) = struct.unpack("<I", buf[off:off + 4]); off += 4
        val = svb_decode_u32(buf[off:off + vl], nex); off += vl
    elif nex == 1:
        p0, v0 = struct.unpack("<II", buf[off:off + 8]); off += 8
        pos, val = np.array([p0], dtype=np.int64), np.array([v0], dtype=np.uint32)
    else:
        pos, val = np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.uint32)
    is_ex[pos] = True
    rest = np.frombuffer(buf[off:], dtype=np.uint8)
    assert rest.shape[0] == n - 1 - nex
    tail = np.zeros(n - 1, dtype=np.uint32)
    tail[pos] = val + 256
    tail[~is_ex] = rest
    zd[1:] = tail
    zz = (zd >> np.uint32(1)).astype(np.int64) ^ -(zd & np.uint32(1)).astype(np.int64)
    sig = np.cumsum(zz.astype(np.int16).astype(np.int64)).astype(np.int16)
    return (sig.astype(np.int32) << q).astype(np.int16)


def svb_encode_u32(vals) -> bytes:
    """Plain streamvbyte encoder (the inverse of svb_decode_u32), for synthetic test records."""
    vals = [int(v) for v in vals]
    keys = bytearray((len(vals) + 3) // 4)
    data = bytearray()
    for i, v in enumerate(vals):
        nb = 1 if v < (1 << 8) else 2 if v < (1 << 16) else 3 if v < (1 << 24) else 4
        keys[i // 4] |= (nb - 1) << (2 * (i % 4))
        data += v.to_bytes(4, "little")[:nb]
    return bytes(keys) + bytes(data)


def ex_zd_encode(sig: np.ndarray) -> bytes:
    """ex-zd v0 encoder restated from slow5lib (slow5_press.c:1262-1420 ex_press, :1596-1630, :1721-1776), for synthetic
    test records only: the shift q (up to 5 trailing zero bits common to all samples), zigzag deltas, exceptions > 255."""
    sig = np.asarray(sig, dtype=np.int16)
    n = len(sig)
    q = 5
    while q and np.any(sig.astype(np.int32) & ((1 << q) - 1)):
        q -= 1
    s = (sig.astype(np.int32) >> q).astype(np.int16)
    d = np.diff(np.concatenate([[0], s.astype(np.int32)])).astype(np.int16)         # int16 wrap, like the reference
    zd = ((d.astype(np.int32) * 2) ^ (d.astype(np.int32) >> 15)).astype(np.uint16)    # zigzag_one_16
    tail = zd[1:].astype(np.int64)
    pos = np.nonzero(tail > 255)[0]
    out = bytes([0]) + struct.pack("<QB", n, q) + struct.pack("<H", int(zd[0])) + struct.pack("<I", len(pos))
    if len(pos) > 1:
        gaps = np.concatenate([[pos[0]], np.diff(pos) - 1])
        a = svb_encode_u32(gaps)
        b = svb_encode_u32(tail[pos] - 256)
        out += struct.pack("<I", len(a)) + a + struct.pack("<I", len(b)) + b
    elif len(pos) == 1:
        out += struct.pack("<II", int(pos[0]), int(tail[pos[0]] - 256))
    return out + tail[tail <= 255].astype(np.uint8).tobytes()
