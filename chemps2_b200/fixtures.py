"""Reader for the .b2fx / .npz golden fixtures written by oracle/ref_driver.cpp (test + bench helper, no compute)."""
import struct

import numpy as np


def read_b2fx(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        (ln,) = struct.unpack_from("<i", data, pos); pos += 4
        name = data[pos:pos + ln].decode(); pos += ln
        dtype, n = struct.unpack_from("<iq", data, pos); pos += 12
        if dtype == 0:
            out[name] = np.frombuffer(data, dtype="<i4", count=n, offset=pos).copy(); pos += 4 * n
        else:
            out[name] = np.frombuffer(data, dtype="<f8", count=n, offset=pos).copy(); pos += 8 * n
    return out


def load(path):
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    return read_b2fx(path)


def split_ops(fx, prefix):
    """-> (boundary, moving_right, [(kind, i, j, data)])"""
    hdr = fx[prefix + "/hdr"]
    meta = fx[prefix + "/meta"].reshape(-1, 4)
    data = fx[prefix + "/data"]
    ops, pos = [], 0
    for kind, i, j, size in meta:
        ops.append((int(kind), int(i), int(j), data[pos:pos + size]))
        pos += size
    return int(hdr[0]), bool(hdr[1]), ops
