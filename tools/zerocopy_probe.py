#!/usr/bin/env python
"""Experiment: does the round trip get faster when the kernel writes its outputs straight into pinned host memory (stores
over PCIe) instead of HBM staging + copy engine?  PERSEUS_GPU_EXPERIMENT_ZEROCOPY=1 makes the library treat pinned host
pointers as device pointers.  Prints one JSON line per variant."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

BUF, NBUF = 6144, 174_762


def main():
    pg = G.load_package()
    zc = os.environ.get("PERSEUS_GPU_EXPERIMENT_ZEROCOPY") == "1"
    with pg.PerseusGpu(device=0) as h:
        n = NBUF * BUF
        ns = n // 6
        d_in = h.dev_alloc(n)
        h.generate(d_in, n)
        pin = h.host_alloc(n)
        h.memcpy(pin, d_in, n)
        po_i, po_f = h.host_alloc(ns * 8), h.host_alloc(ns * 8)
        d_i, d_f = h.dev_alloc(ns * 8), h.dev_alloc(ns * 8)
        F = pg.OUT_INT32 | pg.OUT_FLOAT
        cases = {"host_in_host_out": (pin, po_i, po_f, F), "host_in_host_out_float_only": (pin, None, po_f, pg.OUT_FLOAT),
                 "host_in_device_out": (pin, d_i, d_f, F), "device_in_host_out": (d_in, po_i, po_f, F)}
        for name, (src, oi, of, flags) in cases.items():
            h.unpack(src, n, oi, of, flags)
            t = []
            for _ in range(4):
                h.sync()
                t0 = time.perf_counter()
                h.unpack(src, n, oi, of, flags)
                t.append((time.perf_counter() - t0) * 1e3)
            ref = h.to_host(d_f + ns * 8 - 8192, 8192, "uint32") if name == "host_in_device_out" else None
            print(json.dumps({"zerocopy": zc, "case": name, "ms_best": round(min(t), 3), "ms_all": [round(x, 2) for x in t],
                              "msamples_s": round(ns / min(t) / 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
