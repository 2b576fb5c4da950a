"""TEST INFRASTRUCTURE ONLY: run with PERSEUS_GPU_LIB pointing at the host-simulation build (tests/hostsim/build.sh).
Randomised perseus_gpu_unpack_batch / perseus_gpu_plan_* calls through the product's host layer: segment tables with empty, tiny
and ragged receivers, every format, plans reused across runs and tuning changes, argument errors.  What is under test is the
tile -> (segment, tile-in-segment) map the host builds for the batched kernels (the CUDA stand-in walks it with the oracle); the
kernels themselves, and the per-segment pre-roll for odd output alignments, are tested on the B200.

    PERSEUS_GPU_LIB=/tmp/hostsim.so python tests/hostsim/fuzz_batch.py <seed> <batches>
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


def one_batch(rng, h, idx):
    fmt = rng.choice([pg.OUT_INT32, pg.OUT_FLOAT, pg.OUT_FLOAT_POW2, pg.OUT_INT32 | pg.OUT_FLOAT, pg.OUT_INT32 | pg.OUT_FLOAT_POW2])
    nseg = rng.choice([0, 1, 2, rng.randint(3, 40)])
    segs, keep = [], []
    for k in range(nseg):
        nbytes = rng.choice([0, 5, 6, rng.randint(7, 4000), 6144 * rng.randint(1, 30), 6144 * rng.randint(1, 30) + rng.randint(1, 5), 510 * rng.randint(1, 32)])
        wire = co.synth_random(max(nbytes, 1), seed=idx * 1000 + k)[:nbytes]
        ns = nbytes // 6
        d_in = h.to_device(wire) if nbytes else h.dev_alloc(16)
        d_i = h.dev_alloc(max(ns * 8, 16)) if fmt & pg.OUT_INT32 else None
        d_f = h.dev_alloc(max(ns * 8, 16)) if fmt & (pg.OUT_FLOAT | pg.OUT_FLOAT_POW2) else None
        for d in (d_i, d_f):
            if d:
                C.memset(d, 0xEE, max(ns * 8, 16))
        segs.append((d_in, nbytes, d_i, d_f))
        keep.append((wire, ns))
    total = sum(ns for _, ns in keep)
    if rng.random() < 0.5:
        assert h.unpack_batch(segs, fmt) == total
    else:
        plan = h.plan_create(segs, fmt)
        for run in range(rng.randint(1, 3)):
            if run and rng.random() < 0.5:
                h.set_tuning(stages=rng.choice([2, 4, 8]))      # a plan keeps its tile size; stages follow the handle
            assert h.plan_run(plan, pg.ASYNC if rng.random() < 0.5 else 0) == total
            h.sync()
        h.plan_destroy(plan)
        h.set_tuning()
    fmode = O.MODE_F32_POW2 if fmt & pg.OUT_FLOAT_POW2 else O.MODE_F32
    for (d_in, nbytes, d_i, d_f), (wire, ns) in zip(segs, keep):
        for d, mode in ((d_i, O.MODE_I32), (d_f, fmode)):
            if d:
                got = h.to_host(d, max(ns * 8, 16), np.uint8)
                want = co.unpack(wire, mode).view(np.uint8).reshape(-1) if ns else np.empty(0, np.uint8)
                assert np.array_equal(got[:ns * 8], want) and (got[ns * 8:] == 0xEE).all(), (idx, nbytes, fmt)
                h.dev_free(d)
        h.dev_free(d_in)
    # argument errors leave the handle usable
    if nseg and rng.random() < 0.2:
        bad = list(segs[0])
        bad[2 if fmt & pg.OUT_INT32 else 3] = None
        try:
            h.unpack_batch([tuple(bad)] if bad[1] >= 6 else [(None, 600, None, None)], fmt)
            raise AssertionError("a segment without its output was accepted")
        except pg.PerseusGpuError as e:
            assert e.code == pg.ERR["ERRPARAM"]


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    rng = random.Random(seed)
    with pg.PerseusGpu(device=0) as h:
        for k in range(count):
            one_batch(rng, h, k)
        st = h.stats()
    print(f"fuzz_batch: {count} batches passed ({st['kernel_launches']} launches)")


if __name__ == "__main__":
    main()
