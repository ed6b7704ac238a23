"""Flat binary container for mechanism tables ("PJB200T1").

Layout (little endian):
    char   magic[8]  = "PJB200T1"
    int64  n_entries
    entry[n_entries]: char name[24]; int32 dtype (0 = float64, 1 = int32, 2 = uint16); int32 pad;
                      int64 count; int64 offset   (offset from blob start, 16-byte aligned)
    payload
Read on the C side by ``pjt_find`` (pyjac_b200/csrc/pjtable.h) and by the oracle.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

MAGIC = b'PJB200T1'
_ENTRY = struct.Struct('<24sii qq')


def pack(tables: Dict[str, np.ndarray]) -> bytes:
    names = list(tables)
    head = 8 + 8 + _ENTRY.size * len(names)
    off = (head + 15) // 16 * 16
    entries = []
    chunks = []
    for nm in names:
        a = np.ascontiguousarray(tables[nm])
        if a.dtype == np.float64:
            code = 0
        elif a.dtype == np.int32:
            code = 1
        elif a.dtype == np.uint16:
            code = 2
        else:
            raise TypeError('table %s has unsupported dtype %s' % (nm, a.dtype))
        if len(nm) > 23:
            raise ValueError('table name too long: ' + nm)
        raw = a.tobytes()
        entries.append(_ENTRY.pack(nm.encode(), code, 0, a.size, off))
        pad = (-len(raw)) % 16
        chunks.append(raw + b'\0' * pad)
        off += len(raw) + pad
    blob = MAGIC + struct.pack('<q', len(names)) + b''.join(entries)
    blob += b'\0' * ((-len(blob)) % 16)
    return blob + b''.join(chunks)


def unpack(blob: bytes) -> Dict[str, np.ndarray]:
    if blob[:8] != MAGIC:
        raise ValueError('not a PJB200T1 table blob')
    (n,) = struct.unpack_from('<q', blob, 8)
    out = {}
    for k in range(n):
        nm, code, _, count, off = _ENTRY.unpack_from(blob, 16 + k * _ENTRY.size)
        dt = (np.float64, np.int32, np.uint16)[code]
        out[nm.rstrip(b'\0').decode()] = np.frombuffer(blob, dt, count, off).copy()
    return out
