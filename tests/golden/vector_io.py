"""Exact, compact JSON container for numeric fixtures: every array is stored as
{"dtype": "<f8", "shape": [...], "b64": "..."} (little-endian raw bytes, base64), scalars and strings as plain JSON.
Used by tools/dump_reference_vectors.py (writer) and tests/test_reference_vectors.py (reader)."""
from __future__ import annotations

import base64
import json

import numpy as np


def _enc(v):
    if isinstance(v, dict):
        return {k: _enc(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)) and not (v and isinstance(v[0], (int, float, str, bool))):
        return [_enc(x) for x in v]
    if isinstance(v, np.ndarray) or hasattr(v, "__array__"):
        a = np.ascontiguousarray(np.asarray(v))
        a = a.astype(a.dtype.newbyteorder("<")) if a.dtype.byteorder == ">" else a
        return {"dtype": a.dtype.str, "shape": list(a.shape), "b64": base64.b64encode(a.tobytes()).decode("ascii")}
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    return v


def _dec(v):
    if isinstance(v, dict):
        if set(v) == {"dtype", "shape", "b64"}:
            return np.frombuffer(base64.b64decode(v["b64"]), dtype=np.dtype(v["dtype"])).reshape(v["shape"]).copy()
        return {k: _dec(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_dec(x) for x in v]
    return v


def save(path, obj):
    with open(path, "w") as f:
        json.dump(_enc(obj), f)


def load(path):
    with open(path) as f:
        return _dec(json.load(f))
