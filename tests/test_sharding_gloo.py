"""world_size>1 host logic on CPU (gloo): shard ranges tile the recording, per-shard results combine to the
whole-recording result.  The per-shard unpack here is done by the ORACLE (there is no GPU in this container);
what is under test is the product's shard planner (C ABI), its host generator's random access, and the
collectives bench.py uses to combine ranks."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT

NBUF = 1001          # transfers in the recording; deliberately not divisible by 2 or 3


def _worker(rank, world, initfile, q):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    import __graft_entry__ as G
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    try:
        pg = G.load_package()
        sh = __import__("importlib").import_module("libperseus_sdr_b200.sharding")
        co = O.COracle()
        first, count = sh.rank_shard(pg, NBUF, world, rank)
        ranges = sh.gather_ranges(first, count)
        wire = pg.synth_fill(count * 6144, pg.SYNTH_RANDOM, pg.SYNTH_SEED, first * 6144)     # this rank's byte range only
        words = co.unpack(wire, O.MODE_F32).reshape(-1)
        mine = co.checksum32(words, first_index=first * 2048)
        total = sh.allreduce_sum_u64(mine)
        slowest = sh.allreduce_max(1.0 + rank)
        big = sh.allreduce_sum_u64((1 << 64) - 1 - rank)                                      # wraps modulo 2^64
        # link-weighted shards (bench.py's host-fed leg): every rank contributes its "link rate", all ranks cut the same recording
        rates = sh.allgather_float(10.0 + 7.5 * rank)
        wfirst, wcount = pg.shard_range_weighted(NBUF, rates, rank)
        wwire = pg.synth_fill(wcount * 6144, pg.SYNTH_RANDOM, pg.SYNTH_SEED, wfirst * 6144)
        wtotal = sh.allreduce_sum_u64(co.checksum32(co.unpack(wwire, O.MODE_F32).reshape(-1), first_index=wfirst * 2048))
        q.put((rank, ranges, total, slowest, big, rates, (wfirst, wcount), wtotal))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shards_combine_over_gloo(world, coracle):
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as td:
        procs = [ctx.Process(target=_worker, args=(r, world, os.path.join(td, "init"), q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    whole = coracle.unpack(coracle.synth_random(NBUF * 6144, O.SYNTH_SEED), O.MODE_F32).reshape(-1)
    want = coracle.checksum32(whole)
    wranges = {}
    for rank, ranges, total, slowest, big, rates, wrange, wtotal in res:
        assert rates == [10.0 + 7.5 * r for r in range(world)]
        assert wtotal == want                                             # another sharding of the same recording: same checksum
        wranges[rank] = wrange
        assert total == want                                              # checksum of checksums
        assert slowest == float(world)                                     # max over ranks
        assert big == (sum((1 << 64) - 1 - r for r in range(world))) % (1 << 64)
        pos = 0
        for first, count in ranges:                                        # contiguous, in rank order, no gaps
            assert first == pos
            pos += count
        assert pos == NBUF
    pos = 0
    for r in range(world):                                                 # weighted shards tile the recording too, bigger for faster links
        assert wranges[r][0] == pos
        pos += wranges[r][1]
    assert pos == NBUF and [wranges[r][1] for r in range(world)] == sorted(wranges[r][1] for r in range(world))


def test_weighted_shard_planner(pg):
    total = 8 * 174_762
    assert [pg.shard_range_weighted(total, [3.25] * 8, s) for s in range(8)] == [pg.shard_range(total, 8, s) for s in range(8)]
    rates = [23.3] * 4 + [35.6] * 4                                        # the 8-GPU box of profiles/r2_h2d_matrix_8gpu.md
    shards = [pg.shard_range_weighted(total, rates, s) for s in range(8)]
    assert sum(c for _, c in shards) == total and all(shards[k][0] + shards[k][1] == shards[k + 1][0] for k in range(7))
    assert abs(shards[7][1] / shards[0][1] - 35.6 / 23.3) < 1e-3
    assert max(c / r for (_, c), r in zip(shards, rates)) / min(c / r for (_, c), r in zip(shards, rates)) < 1.0001   # equal finish times
    assert [pg.shard_range_weighted(10, [0, 1, 0, 3], s) for s in range(4)] == [(0, 0), (0, 2), (2, 0), (2, 8)]
    for bad in ([], [0, 0], [1, -1], [float("nan"), 1]):
        with pytest.raises(pg.PerseusGpuError):
            pg.shard_range_weighted(10, bad, 0)
