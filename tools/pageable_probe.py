#!/usr/bin/env python
"""perseus_gpu_unpack fed from PAGEABLE host memory (malloc / numpy / mmap: what an application that never heard of CUDA
holds) against the same call fed from pinned memory, cfg2 recording, outputs staying in HBM or coming back to pageable memory."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as ge  # noqa: E402

pg = ge.load_package()
NBUF = int(sys.argv[1]) if len(sys.argv) > 1 else 174_762


def best_ms(fn, reps=4):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


def main():
    nbytes = NBUF * 6144
    ns = nbytes // 6
    fused = pg.OUT_INT32 | pg.OUT_FLOAT
    wire = pg.synth_fill(nbytes)                           # numpy: pageable
    out_f = np.zeros(ns * 2, np.float32)
    for threads, chunk_mib in ((pg.COPY_BY_RUNTIME, 32), (1, 32), (2, 32), (4, 32), (8, 32), (12, 32), (16, 32), (8, 8), (8, 16), (0, 0)):
        with pg.PerseusGpu(device=0, copy_threads=threads, chunk_bytes=chunk_mib << 20) as h:
            pin = h.host_alloc(nbytes)
            C.memmove(pin, wire.ctypes.data, nbytes)
            d_i, d_f = h.dev_alloc(ns * 8), h.dev_alloc(ns * 8)
            row = {"copy_threads": "cuda runtime" if threads == pg.COPY_BY_RUNTIME else threads or "default", "chunk_mib": chunk_mib or "default"}
            for name, src in (("pinned", pin), ("pageable", wire.ctypes.data)):
                ms = best_ms(lambda: h.unpack(src, nbytes, d_i, d_f, fused))
                row[f"{name}_in_device_out_gbs"] = round(nbytes / ms / 1e6, 2)
                ms = best_ms(lambda: h.unpack(src, nbytes, None, out_f.ctypes.data, pg.OUT_FLOAT), reps=2)
                row[f"{name}_in_pageable_float_out_msamples"] = round(ns / ms / 1e3, 1)
            d_in = h.dev_alloc(nbytes)
            h.memcpy(d_in, pin, nbytes)
            assert h.verify(d_in, nbytes, d_i, d_f, fused)[0] == 0
            ms = best_ms(lambda: h.unpack(d_in, nbytes, None, out_f.ctypes.data, pg.OUT_FLOAT), reps=2)
            row["device_in_pageable_float_out_d2h_gbs"] = round(ns * 8 / ms / 1e6, 2)
            print(json.dumps(row), flush=True)
            for p_ in (d_i, d_f, d_in):
                h.dev_free(p_)
            h.host_free(pin)
    t0 = time.perf_counter()
    pin2 = np.empty_like(wire)
    np.copyto(pin2, wire)
    t0 = time.perf_counter()
    np.copyto(pin2, wire)
    print(json.dumps({"one_thread_memcpy_gbs": round(nbytes / (time.perf_counter() - t0) / 1e9, 2)}))


if __name__ == "__main__":
    main()
