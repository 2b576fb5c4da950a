#!/usr/bin/env python
"""The short command ncu wraps to capture the streaming path's small-slab launch: slabs of 8 transfers (48 KiB) pushed through
perseus_gpu_input_callback and flushed, on the direct route (the kernel reads the pinned slab over the link) or staged."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as G  # noqa: E402

pg = G.load_package()
staged = len(sys.argv) > 1 and sys.argv[1] == "staged"
wire = pg.synth_fill(6144 * 8)
with pg.PerseusGpu(device=0, stream_flags=pg.OUT_INT32 | pg.OUT_FLOAT, slab_bytes=6144 * 8, nslabs=4,
                   direct_bytes=pg.DIRECT_NEVER if staged else 0) as h:
    for _ in range(30):
        for k in range(8):
            h.input_callback(wire[k * 6144:].ctypes.data, 6144)
        h.flush()
    print(h.stats())
