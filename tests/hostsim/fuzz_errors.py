"""TEST INFRASTRUCTURE ONLY: run with PERSEUS_GPU_LIB pointing at the host-simulation build (tests/hostsim/build.sh).
Error paths of the product's host layer: one device / pinned allocation, chosen at random, fails in the middle of a streaming
scenario or a bulk call (the CUDA stand-in's fake_cuda_fail_alloc_in hook).  The reference ignores callback return values
(perseus-in.c:207), so the library latches the failure, counts what it drops and reports it at the next flush / sync / close.
What is checked: no crash, no hang, the error surfaces exactly once with the CUDA text, every callback is accounted for
(callbacks + dropped_callbacks), whatever WAS delivered is correct, and the handle keeps working afterwards.

    PERSEUS_GPU_LIB=/tmp/hostsim.so python tests/hostsim/fuzz_errors.py <seed> <scenarios>
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
fail_in = pg.lib().fake_cuda_fail_alloc_in
fail_in.argtypes, fail_in.restype = [C.c_long], None


def streaming(rng, idx):
    fmt = rng.choice([pg.OUT_INT32, pg.OUT_FLOAT, pg.OUT_INT32 | pg.OUT_FLOAT])
    cfg = dict(stream_flags=fmt, slab_bytes=6144 * rng.randint(1, 6), nslabs=rng.randint(2, 4), direct_bytes=rng.choice([0, pg.DIRECT_NEVER]),
               eager_gap_us=pg.EAGER_NEVER, max_latency_us=0xFFFFFFFF)
    n = rng.randint(5, 60)
    wire = co.synth_random(n * 6144, seed=50 + idx)
    want = co.unpack(wire, O.MODE_I32 if fmt & pg.OUT_INT32 else O.MODE_F32).view(np.uint32).reshape(-1)
    host, errors = [], 0
    h = pg.PerseusGpu(device=0, **cfg)
    try:
        def sink(blk, extra):
            b = blk.contents
            host.append((b.first_sample, np.ctypeslib.as_array((C.c_uint32 * (2 * b.nsamples)).from_address(b.i32 or b.f32)).copy()))
        if rng.random() < 0.7:
            h.set_host_sink(sink)
        if rng.random() < 0.3:
            fail_in(rng.randint(1, 12))
            try:
                h.prepare()
            except pg.PerseusGpuError as e:
                errors += 1
                assert e.code == pg.ERR["CUDAERR"] and "out of memory" in e.msg, e
        else:
            fail_in(rng.randint(1, 25))
        delivered = []                                   # the bytes the library did NOT drop, in order
        for k in range(n):
            before = h.stats()["dropped_bytes"]
            h.input_callback(wire[k * 6144:].ctypes.data, 6144)
            kept_bytes = 6144 - (h.stats()["dropped_bytes"] - before)     # the failing transfer may have been taken in part
            assert kept_bytes % 6 == 0
            delivered.append(wire[k * 6144:k * 6144 + kept_bytes])
            if rng.random() < 0.15:
                try:
                    h.flush()
                except pg.PerseusGpuError as e:
                    errors += 1
                    assert e.code == pg.ERR["CUDAERR"] and "out of memory" in e.msg, e
        for attempt in range(2):                         # the first flush may report the latched failure; the second must be clean
            try:
                h.flush()
                break
            except pg.PerseusGpuError as e:
                errors += 1
                assert attempt == 0 and "out of memory" in e.msg, e
        st = h.stats()
        assert st["callbacks"] + st["dropped_callbacks"] == n and st["dropped_bytes"] >= st["dropped_callbacks"] * 6144, st
        assert errors <= 1, errors                       # one injected failure: reported once
        fail_in(0)
        # the stream the library accepted is the concatenation of the callbacks it did not drop
        kept = np.concatenate(delivered) if delivered else np.empty(0, np.uint8)
        assert st["samples"] == kept.size // 6, (st, len(delivered))
        want = co.unpack(kept, O.MODE_I32 if fmt & pg.OUT_INT32 else O.MODE_F32).view(np.uint32).reshape(-1) if kept.size else np.empty(0, np.uint32)
        if host:
            got = np.concatenate([b[1] for b in host])
            assert [b[0] for b in host] == list(np.cumsum([0] + [b[1].size // 2 for b in host[:-1]])), "blocks out of order"
            assert np.array_equal(got, want[:got.size]), (cfg, "delivered content")
        # and the handle still works
        d = h.to_device(wire[:6144])
        o = h.dev_alloc(8192)
        assert h.unpack(d, 6144, o, None, pg.OUT_INT32) == 1024
        h.dev_free(d); h.dev_free(o)
    finally:
        fail_in(0)
        try:
            h.close()
        except pg.PerseusGpuError as e:
            assert "out of memory" in e.msg, e
    return errors


def bulk(rng, idx):
    nbytes = 12288 * rng.randint(1, 200) + rng.randint(0, 5)
    wire = co.synth_random(nbytes, seed=900 + idx)
    ns = nbytes // 6
    out = np.zeros(ns * 2, np.uint32)
    with pg.PerseusGpu(device=0, chunk_bytes=12288 * rng.randint(1, 30), stage_slots=rng.randint(2, 5), copy_threads=rng.choice([0, 1, 3])) as h:
        fail_in(rng.randint(1, 14))
        errors = 0
        try:
            h.unpack(wire.ctypes.data, nbytes, out.ctypes.data, None, pg.OUT_INT32)
        except pg.PerseusGpuError as e:
            errors = 1
            assert e.code == pg.ERR["CUDAERR"] and "out of memory" in e.msg, e
        fail_in(0)
        out[:] = 0
        assert h.unpack(wire.ctypes.data, nbytes, out.ctypes.data, None, pg.OUT_INT32) == ns      # the retry allocates what was missing
        assert np.array_equal(out, co.unpack(wire, O.MODE_I32).view(np.uint32).reshape(-1))
    return errors


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    rng = random.Random(seed)
    hit = 0
    for k in range(count):
        hit += streaming(rng, k) if rng.random() < 0.7 else bulk(rng, k)
    print(f"fuzz_errors: {count} scenarios passed ({hit} injected failures surfaced)")


if __name__ == "__main__":
    main()
