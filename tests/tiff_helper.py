"""Tiny TIFF writer/reader for the CLI tests (independent of sqeazy_b200/csrc/cli/tiff_min.hpp): uncompressed grayscale
stacks, little or big endian, several strips per page."""
import struct

import numpy as np


def write_tiff(path, vol: np.ndarray, big_endian=False, rows_per_strip=None):
    vol = np.ascontiguousarray(vol)
    assert vol.ndim == 3 and vol.dtype in (np.uint8, np.uint16)
    e = ">" if big_endian else "<"
    Z, H, W = vol.shape
    bits = vol.dtype.itemsize * 8
    rps = rows_per_strip or H
    nstrips = (H + rps - 1) // rps
    out = bytearray()
    out += (b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, 0)
    prev_next_field = 4   # where the offset of the next IFD has to be patched
    for z in range(Z):
        page = vol[z].astype(vol.dtype.newbyteorder(e)).tobytes()
        row_bytes = W * bits // 8
        offs, lens = [], []
        for s in range(nstrips):
            if len(out) & 1:
                out += b"\0"
            chunk = page[s * rps * row_bytes: min(H, (s + 1) * rps) * row_bytes]
            offs.append(len(out)); lens.append(len(chunk))
            out += chunk
        if len(out) & 1:
            out += b"\0"
        tables = b""
        if nstrips > 1:
            off_tab = len(out); out += struct.pack(e + "%dI" % nstrips, *offs)
            len_tab = len(out); out += struct.pack(e + "%dI" % nstrips, *lens)
        ifd_off = len(out)
        struct.pack_into(e + "I", out, prev_next_field, ifd_off)

        def ent(tag, typ, count, value):
            if typ == 3 and count == 1:
                return struct.pack(e + "HHIHH", tag, typ, count, value, 0)
            return struct.pack(e + "HHII", tag, typ, count, value)
        ents = [ent(256, 4, 1, W), ent(257, 4, 1, H), ent(258, 3, 1, bits), ent(259, 3, 1, 1), ent(262, 3, 1, 1),
                ent(273, 4, nstrips, offs[0] if nstrips == 1 else off_tab), ent(277, 3, 1, 1), ent(278, 4, 1, rps),
                ent(279, 4, nstrips, lens[0] if nstrips == 1 else len_tab)]
        out += struct.pack(e + "H", len(ents)) + b"".join(ents)
        prev_next_field = len(out)
        out += struct.pack(e + "I", 0)
    with open(path, "wb") as f:
        f.write(bytes(out))


def read_tiff(path) -> np.ndarray:
    b = open(path, "rb").read()
    e = "<" if b[:2] == b"II" else ">"
    assert struct.unpack(e + "H", b[2:4])[0] == 42
    ifd = struct.unpack(e + "I", b[4:8])[0]
    pages = []
    while ifd:
        n = struct.unpack(e + "H", b[ifd:ifd + 2])[0]
        tags = {}
        for i in range(n):
            tag, typ, count = struct.unpack(e + "HHI", b[ifd + 2 + 12 * i: ifd + 10 + 12 * i])
            raw = b[ifd + 10 + 12 * i: ifd + 14 + 12 * i]
            val = struct.unpack(e + "H", raw[:2])[0] if (typ == 3 and count == 1) else struct.unpack(e + "I", raw)[0]
            tags[tag] = (typ, count, val)
        W, H, bits = tags[256][2], tags[257][2], tags[258][2]
        assert tags[259][2] == 1 and tags[273][1] == 1
        off, ln = tags[273][2], tags[279][2]
        dt = np.dtype(np.uint8 if bits == 8 else np.uint16).newbyteorder(e)
        pages.append(np.frombuffer(b[off:off + ln], dtype=dt).reshape(H, W).astype(dt.newbyteorder("=")))
        ifd = struct.unpack(e + "I", b[ifd + 2 + 12 * n: ifd + 6 + 12 * n])[0]
    return np.stack(pages)
