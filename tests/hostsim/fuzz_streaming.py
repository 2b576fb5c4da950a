"""TEST INFRASTRUCTURE ONLY: run with PERSEUS_GPU_LIB pointing at the host-simulation build (tests/hostsim/build.sh).
Randomised streaming scenarios through the product's host layer -- slab ring, both slab routes, eager submission, age bound and
watchdog, device sink, host sink (delivery thread), file sink, interleaved flush / poll / stats calls, every legal transfer size
-- with the sample arithmetic done by the CPU oracle behind the CUDA stand-in.  What is under test is the plumbing: every consumer
must see exactly the unpack of the concatenated transfers, in order, whatever the timing does.

    PERSEUS_GPU_LIB=/tmp/hostsim.so python tests/hostsim/fuzz_streaming.py <seed> <scenarios>
"""
import ctypes as C
import os
import random
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402
from oracle import oracle as O  # noqa: E402

pg = G.load_package()
co = O.COracle()
assert "hostsim" in str(pg.LIB_PATH), "this script drives the host-simulation build only"


def scenario(rng, idx, tmp):
    fmt = rng.choice([pg.OUT_INT32, pg.OUT_FLOAT, pg.OUT_FLOAT_POW2, pg.OUT_INT32 | pg.OUT_FLOAT])
    one_format = fmt in (pg.OUT_INT32, pg.OUT_FLOAT, pg.OUT_FLOAT_POW2)
    cfg = dict(stream_flags=fmt, slab_bytes=48 * rng.randint(3, 700), nslabs=rng.randint(2, 5), nstreams=rng.randint(1, 3),
               direct_bytes=rng.choice([0, pg.DIRECT_NEVER, 48 * 40]), eager_gap_us=rng.choice([0, pg.EAGER_NEVER, 30]),
               max_latency_us=rng.choice([0, 0xFFFFFFFF, 300]), options=rng.choice([0, pg.OPT_NO_WATCHDOG]))
    use_dev, use_host, use_file = rng.random() < 0.5, rng.random() < 0.7, one_format and rng.random() < 0.5
    sizes = [rng.choice([6144, 12288] + [510 * k for k in (1, 2, 7, 32)]) for _ in range(rng.randint(1, 60))]
    wire = co.synth_random(sum(sizes), seed=1000 + idx)
    dev, host = [], []
    path = os.path.join(tmp, f"s{idx}.bin")
    with pg.PerseusGpu(device=0, **cfg) as h:
        def dev_sink(blk, extra):
            b = blk.contents
            h.sync()                                    # re-enters the handle from its own sink (allowed for the device sink)
            dev.append((b.first_sample, h.to_host(b.dev_i32 or b.dev_f32, b.nsamples * 8, np.uint32)))

        def host_sink(blk, extra):
            b = blk.contents
            host.append((b.first_sample, np.ctypeslib.as_array((C.c_uint32 * (2 * b.nsamples)).from_address(b.i32 or b.f32)).copy(),
                         bool(b.i32), bool(b.f32)))

        if use_dev:
            h.set_sink(dev_sink)
        if use_host:
            h.set_host_sink(host_sink)
        if use_file:
            h.stream_to_file(path)
        if rng.random() < 0.3:
            h.prepare()
        # some scenarios switch the host sink off and on again in mid-stream: setting it flushes, so the sink must receive
        # exactly the samples pushed while it was set
        toggling = use_host and not use_file and rng.random() < 0.25
        host_on, on_from, intervals = use_host, 0, []
        dev_on, dev_from, dev_spans, bulk_samples = use_dev, 0, [], 0
        off = 0
        if not toggling and rng.random() < 0.3:
            # two threads on one handle: the receiver's thread calls back as fast as it can (lock-free fast path) while this one
            # keeps taking the handle away from it (statistics, poll, flush, sync: each goes through the ownership hand-off)
            def receiver():
                o = 0
                for n in sizes:
                    h.input_callback(wire[o:].ctypes.data, n)
                    o += n
            t = threading.Thread(target=receiver)
            t.start()
            while t.is_alive():
                rng.choice([h.stats, h.poll, h.flush, h.sync, h.prepare, lambda: h.get_geometry(fmt & 7)])()
            t.join()
            sizes_left = []
        else:
            sizes_left = sizes
        for n in sizes_left:
            h.input_callback(wire[off:].ctypes.data, n)
            off += n
            if toggling and rng.random() < 0.15:
                if host_on:
                    h.set_host_sink(None)
                    intervals.append((on_from, off // 6))
                else:
                    h.set_host_sink(host_sink)
                    on_from = off // 6
                host_on = not host_on
            r = rng.random()
            if r < 0.08:
                h.flush()
            elif r < 0.16:
                h.poll()
            elif r < 0.22:
                h.stats()
            elif r < 0.24:
                h.prepare()                             # in mid-stream: must not touch a slab that holds live data
            elif r < 0.28:
                # a bulk call on the streaming handle (it shares streams[0] with the slabs): its own result must be right and the
                # stream must not notice
                nb = rng.choice([48, 6144, 6144 * 3 + 30])
                bw = co.synth_random(nb, seed=7000 + idx)
                out = np.zeros(nb // 6 * 2, np.uint32)
                assert h.unpack(bw.ctypes.data, nb, out.ctypes.data, None, pg.OUT_INT32) == nb // 6
                assert np.array_equal(out, co.unpack(bw, O.MODE_I32).view(np.uint32).reshape(-1)), ("bulk call in mid-stream", cfg)
                bulk_samples += nb // 6
            elif r < 0.31 and use_dev and not toggling:
                # the device sink taken away and given back: the samples in between go unobserved by it, nothing else changes
                # (with it gone, small slabs of a host-only stream may be stored straight into host memory)
                if dev_on:
                    h.flush()
                    h.set_sink(None)
                    dev_spans.append((dev_from, off // 6))
                else:
                    h.flush()
                    h.set_sink(dev_sink)
                    dev_from = off // 6
                dev_on = not dev_on
            elif r < 0.34:
                time.sleep(rng.choice([0.0, 0.0001, 0.0006]))
        if rng.random() < 0.5:
            time.sleep(0.001)                           # let the watchdog / the delivery thread do the last part on their own
        closed_without_flush = rng.random() < 0.3
        if not closed_without_flush:                    # otherwise: perseus_gpu_close flushes, delivers and closes the file
            h.flush()
            if use_file:
                h.stream_to_file(None)
        st = h.stats()
    if dev_on:
        dev_spans.append((dev_from, wire.size // 6))
    if host_on:
        intervals.append((on_from, wire.size // 6))
    ns = wire.size // 6
    first_mode = O.MODE_I32 if fmt & pg.OUT_INT32 else O.MODE_F32_POW2 if fmt & pg.OUT_FLOAT_POW2 else O.MODE_F32
    want = co.unpack(wire, first_mode).view(np.uint32).reshape(-1)
    assert st["callbacks"] == len(sizes) and st["dropped_callbacks"] == 0, (cfg, st)
    assert closed_without_flush or st["samples"] == ns + bulk_samples, (cfg, st)
    for name, blocks, spans in (("device sink", dev if use_dev else None, dev_spans), ("host sink", host if use_host else None, intervals)):
        if blocks is None:
            continue
        covered = []
        for b in blocks:                                   # every block carries the samples its first_sample says it carries ...
            first, n = b[0], b[1].size // 2
            assert np.array_equal(b[1], want[2 * first:2 * (first + n)]), (name, cfg, "block content")
            if covered and covered[-1][1] == first:
                covered[-1] = (covered[-1][0], first + n)
            else:
                covered.append((first, first + n))
        # ... in stream order, and together exactly the samples pushed while the sink was set
        assert covered == [sp for sp in spans if sp[1] > sp[0]], (name, cfg, covered, spans)
    if use_host:
        assert all(b[2] == bool(fmt & pg.OUT_INT32) and b[3] == bool(fmt & (pg.OUT_FLOAT | pg.OUT_FLOAT_POW2)) for b in host)
        assert closed_without_flush or (st["host_blocks"] == len(host) and (toggling or st["host_blocks"] == st["slabs"])), (cfg, st)
    if use_file:
        assert Path(path).read_bytes() == want.tobytes(), ("file", cfg)
    return st["slabs"]


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    rng = random.Random(seed)
    slabs = 0
    with tempfile.TemporaryDirectory() as tmp:
        for k in range(count):
            slabs += scenario(rng, k, tmp)
    print(f"fuzz_streaming: {count} scenarios passed ({slabs} slabs)")


if __name__ == "__main__":
    main()
