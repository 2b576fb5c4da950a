"""TEST INFRASTRUCTURE ONLY: run with PERSEUS_GPU_LIB pointing at the host-simulation build (tests/hostsim/build.sh).
Randomised perseus_gpu_unpack calls through the product's host layer: every mix of device / pinned / pageable pointers, chunk sizes
and slot counts of the staging pipeline, copy-pool sizes (and the CUDA runtime's own staging), ASYNC, CHECKSUM, ragged sizes, odd
pointer phases, several calls per handle (slot rotation carries over).  The arithmetic is the CPU oracle's behind the CUDA
stand-in; what is under test is the pipeline's bookkeeping: every byte of every output, nothing outside, totals and statistics.

    PERSEUS_GPU_LIB=/tmp/hostsim.so python tests/hostsim/fuzz_bulk.py <seed> <handles>
"""
import ctypes as C
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402
from oracle import oracle as O  # noqa: E402

pg = G.load_package()
co = O.COracle()
assert "hostsim" in str(pg.LIB_PATH), "this script drives the host-simulation build only"
GUARD = 64


class Buf:
    """nbytes of device / pinned / pageable memory at a chosen phase, with guard bytes on both sides"""

    def __init__(self, h, kind, nbytes, phase, fill):
        self.h, self.kind, self.n = h, kind, nbytes
        total = nbytes + 2 * GUARD + 16
        if kind == "pageable":
            self.arr = np.full(total, fill, np.uint8)
            self.base = self.arr.ctypes.data
        else:
            self.base = h.dev_alloc(total) if kind == "device" else h.host_alloc(total)
            C.memset(self.base, fill, total)                   # the stand-in's "device" memory is host memory
        self.p = (self.base + GUARD + 15) // 16 * 16 + phase
        self.fill = fill

    def view(self):
        return np.ctypeslib.as_array((C.c_uint8 * (self.n + 2 * GUARD + 16)).from_address(self.base))

    def data(self):
        off = self.p - self.base
        v = self.view()
        assert (v[:off] == self.fill).all() and (v[off + self.n:] == self.fill).all(), "wrote outside the output range"
        return v[off:off + self.n].copy()

    def free(self):
        if self.kind == "device":
            self.h.dev_free(self.base)
        elif self.kind == "pinned":
            self.h.host_free(self.base)


def one_handle(rng, idx):
    chunk = rng.choice([48 * rng.randint(1, 400), 12288 * rng.randint(1, 40), 12288 * rng.randint(60, 140)])   # the last: the copy pool's helpers get slices
    cfg = dict(chunk_bytes=chunk, stage_slots=rng.randint(2, 8), copy_threads=rng.choice([0, 1, 3, pg.COPY_BY_RUNTIME]))
    h2d = d2h = 0
    with pg.PerseusGpu(device=0, **cfg) as h:
        for call in range(rng.randint(1, 5)):
            nbytes = rng.choice([0, 5, rng.randint(6, 3000), rng.randint(3000, 600000), 6144 * rng.randint(1, 80) + rng.randint(0, 5),
                                 rng.randint(600000, 5000000)])
            ns = nbytes // 6
            fmt = rng.choice([pg.OUT_INT32, pg.OUT_FLOAT, pg.OUT_FLOAT_POW2, pg.OUT_INT32 | pg.OUT_FLOAT, pg.OUT_INT32 | pg.OUT_FLOAT_POW2])
            flags = fmt | (pg.ASYNC if rng.random() < 0.4 else 0) | (pg.CHECKSUM if rng.random() < 0.4 else 0)
            kinds = [rng.choice(["device", "pinned", "pageable"]) for _ in range(3)]
            wire = co.synth_random(max(nbytes, 1), seed=idx * 100 + call)[:nbytes]
            bin_ = Buf(h, kinds[0], nbytes, rng.randint(0, 15), 0x11)
            C.memmove(bin_.p, wire.ctypes.data, nbytes)
            bi = Buf(h, kinds[1], ns * 8, 4 * rng.randint(0, 3), 0xA5) if fmt & pg.OUT_INT32 else None
            bf = Buf(h, kinds[2], ns * 8, 4 * rng.randint(0, 3), 0x5A) if fmt & (pg.OUT_FLOAT | pg.OUT_FLOAT_POW2) else None
            assert h.unpack(bin_.p, nbytes, bi.p if bi else None, bf.p if bf else None, flags) == ns, (cfg, nbytes, kinds)
            sums = h.get_checksums() if flags & pg.CHECKSUM else None
            h.sync()
            fmode = O.MODE_F32_POW2 if fmt & pg.OUT_FLOAT_POW2 else O.MODE_F32
            for b, mode in ((bi, O.MODE_I32), (bf, fmode)):
                if b is not None:
                    want = co.unpack(wire, mode).view(np.uint8).reshape(-1)
                    assert np.array_equal(b.data(), want), (cfg, nbytes, kinds, flags)
            if sums is not None:
                wi = co.checksum32(co.unpack(wire, O.MODE_I32).view(np.uint32).reshape(-1)) if bi else 0
                wf = co.checksum32(co.unpack(wire, fmode).view(np.uint32).reshape(-1)) if bf else 0
                assert sums == (wi, wf), (cfg, nbytes, kinds, flags)
                d2h += 16
            if ns:
                h2d += ns * 6 if kinds[0] != "device" else 0
                d2h += sum(ns * 8 for b, k in ((bi, kinds[1]), (bf, kinds[2])) if b is not None and k != "device")
            for b in (bin_, bi, bf):
                if b is not None:
                    b.free()
        st = h.stats()
        assert (st["h2d_bytes"], st["d2h_bytes"]) == (h2d, d2h), (cfg, st, h2d, d2h)


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    rng = random.Random(seed)
    for k in range(count):
        one_handle(rng, k)
    print(f"fuzz_bulk: {count} handles passed")


if __name__ == "__main__":
    main()
