"""B2LS container: a flat list of named arrays shared by the Python tests, the
reference harness (oracle/ref_harness.C) and the C oracle.

Layout (little endian):
    magic   8 bytes  "B2LS0001"
    int64   nEntries
    per entry: int64 nameLen; name; int64 dtype (0=i32, 1=f64, 2=u8); int64 count; data
"""
import struct

import numpy as np

_DT = {0: np.int32, 1: np.float64, 2: np.uint8}
_CODE = {np.dtype(np.int32): 0, np.dtype(np.float64): 1, np.dtype(np.uint8): 2}


def write(path, entries):
    """entries: dict name -> ndarray(int32|float64|uint8) | str | int | float."""
    with open(path, "wb") as f:
        f.write(b"B2LS0001")
        f.write(struct.pack("<q", len(entries)))
        for name, val in entries.items():
            if isinstance(val, str):
                arr = np.frombuffer(val.encode(), dtype=np.uint8)
            elif isinstance(val, (int, np.integer)):
                arr = np.array([val], dtype=np.int32)
            elif isinstance(val, float):
                arr = np.array([val], dtype=np.float64)
            else:
                arr = np.ascontiguousarray(val)
                if arr.dtype == np.bool_:
                    arr = arr.astype(np.uint8)
            code = _CODE[arr.dtype]
            nb = name.encode()
            f.write(struct.pack("<q", len(nb)))
            f.write(nb)
            f.write(struct.pack("<qq", code, arr.size))
            f.write(arr.tobytes())


def read(path):
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != b"B2LS0001":
            raise ValueError(f"{path}: bad magic")
        (n,) = struct.unpack("<q", f.read(8))
        for _ in range(n):
            (ln,) = struct.unpack("<q", f.read(8))
            name = f.read(ln).decode()
            code, count = struct.unpack("<qq", f.read(16))
            dt = np.dtype(_DT[code])
            out[name] = np.frombuffer(f.read(count * dt.itemsize), dtype=dt).copy()
    return out


def as_str(arr):
    return bytes(arr.astype(np.uint8)).decode()
