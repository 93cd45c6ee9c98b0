"""Reader for the binary dumps written by oracle/ref_driver.cpp (test infrastructure only)."""
import struct
import numpy as np


def read_dump(path):
    out = {}
    with open(path, "rb") as fh:
        data = fh.read()
    pos = 0
    while pos < len(data):
        (nl,) = struct.unpack_from("<i", data, pos); pos += 4
        name = data[pos:pos + nl].decode(); pos += nl
        (nd,) = struct.unpack_from("<i", data, pos); pos += 4
        dims = struct.unpack_from("<%dq" % nd, data, pos); pos += 8 * nd
        cnt = int(np.prod(dims))
        arr = np.frombuffer(data, dtype="<f8", count=cnt, offset=pos).reshape(dims).copy(); pos += 8 * cnt
        if name.startswith("order/"):
            out[name] = bytes(arr.astype(np.uint8)).decode()
        else:
            out[name] = arr
    return out
