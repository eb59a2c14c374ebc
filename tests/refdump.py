"""TEST INFRASTRUCTURE: reader for the record files written by
oracle/ref_harness/ref_dump.c (the unmodified reference behind a recording FIFO)
and helpers shared by tests/golden/make_golden.py and the parity tests."""
import struct
import zlib

import numpy as np

TAG_HEADER, TAG_BLOCK, TAG_NAV, TAG_CODE, TAG_TABLES, TAG_END = 1, 2, 3, 4, 5, 6

CHAN_DT = np.dtype([
    ("prn", "<i4"), ("iword", "<i4"), ("ibit", "<i4"), ("icode", "<i4"),
    ("dataBit", "<i4"), ("codeCA", "<i4"),
    ("f_carr", "<f8"), ("f_code", "<f8"), ("carr_phase", "<f8"),
    ("code_phase", "<f8"), ("gain", "<f8"),
])
assert CHAN_DT.itemsize == 64

SAMPLES_PER_BLOCK = 300000
SUBCHUNKS = 30  # CRC granularity inside a block (10 000 complex samples)


def read_params(path):
    """-> dict(max_chan, sample_size, chans[nblk, C] (CHAN_DT), nav=[(block, ch, words[60])],
    codes={prn: uint8[1023]}, sin512, cos512, producer_seconds)"""
    raw = open(path, "rb").read()
    off = 0
    out = {"nav": [], "codes": {}}
    blocks = []
    while off < len(raw):
        tag, n = struct.unpack_from("<II", raw, off)
        off += 8
        p = raw[off:off + n]
        off += n
        if tag == TAG_HEADER:
            v, mc, ss, spb = struct.unpack("<IIII", p)
            out.update(max_chan=mc, sample_size=ss, samples_per_block=spb)
        elif tag == TAG_BLOCK:
            blocks.append(np.frombuffer(p, dtype=CHAN_DT, offset=8))  # u32 block + 4 pad, then 64-byte records
        elif tag == TAG_NAV:
            b, c = struct.unpack_from("<II", p)
            out["nav"].append((b, c, np.frombuffer(p, dtype="<u4", offset=8).copy()))
        elif tag == TAG_CODE:
            (prn,) = struct.unpack_from("<I", p)
            out["codes"][prn] = np.frombuffer(p, dtype=np.uint8, offset=4, count=1023).copy()
        elif tag == TAG_TABLES:
            t = np.frombuffer(p, dtype="<i4")
            out["sin512"], out["cos512"] = t[:512].copy(), t[512:].copy()
        elif tag == TAG_END:
            nb, _pad, secs = struct.unpack("<IId", p)
            out["producer_seconds"] = secs
    out["chans"] = np.stack(blocks) if blocks else np.zeros((0, out.get("max_chan", 0)), CHAN_DT)
    return out


def nav_table(par):
    """Dense NAV words per (block, channel): uint32[nblk, C, 60] from the sparse NAV records."""
    ch = par["chans"]
    nblk, C = ch.shape
    words = np.zeros((nblk, C, 60), np.uint32)
    cur = np.zeros((C, 60), np.uint32)
    by_block = {}
    for b, c, w in par["nav"]:
        by_block.setdefault(b, []).append((c, w))
    for b in range(nblk):
        for c, w in by_block.get(b, []):
            cur[c] = w
        words[b] = cur
    return words


def block_crcs(stream, elem_per_block=2 * SAMPLES_PER_BLOCK):
    """stream: 1-D int8/int16 array holding whole blocks. -> uint32[nblk, 1 + SUBCHUNKS]:
    column 0 = CRC-32 of the block, columns 1.. = CRC-32 of each 1/30 of it."""
    nblk = stream.size // elem_per_block
    s = stream[:nblk * elem_per_block].reshape(nblk, elem_per_block)
    sub = elem_per_block // SUBCHUNKS
    out = np.zeros((nblk, 1 + SUBCHUNKS), np.uint32)
    for b in range(nblk):
        row = np.ascontiguousarray(s[b])
        out[b, 0] = zlib.crc32(row.tobytes())
        for j in range(SUBCHUNKS):
            out[b, 1 + j] = zlib.crc32(row[j * sub:(j + 1) * sub].tobytes())
    return out
