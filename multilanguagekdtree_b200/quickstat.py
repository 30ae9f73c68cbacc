"""Host-side mirror of Parallel/RustVersion/src/quickstat.rs: `quickstat_index(indices, goal, lt)` with the comparator
of its only call sites, `vals[i1] < vals[i2]` (array_kd_tree.rs:561-562, bin/bench_quickstat.rs:20), on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .array_kd_tree import KdnbError, KDTreeSim


def quickstat_index(indices: np.ndarray, goal: int, vals: np.ndarray, sim: KDTreeSim | None = None) -> float:
    """Permute `indices` (uint64 element ids into `vals`, in place) so that indices[goal] is the goal-th smallest value,
    nothing before it is larger and nothing after it smaller (quickstat.rs:9-34; post-condition :199-253).
    Returns the device milliseconds of the selection (radix select + stable three-way partition, csrc/select.cu)."""
    if indices.dtype != np.uint64 or not indices.flags.c_contiguous:
        raise TypeError("indices must be a C-contiguous uint64 array (Rust usize)")
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    own = sim is None
    s = KDTreeSim() if own else sim
    try:
        ms = C.c_double(0.0)
        rc = _lib.load().kdnb_quickstat_index(s._h, vals.ctypes.data, len(vals), indices.ctypes.data, len(indices),
                                              int(goal), C.byref(ms))
        if rc != 0:
            raise KdnbError(f"kdnb_quickstat_index failed ({rc}): {_lib.load().kdnb_last_error(s._h).decode()}")
        return ms.value
    finally:
        if own:
            s.close()
