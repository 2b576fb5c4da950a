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
        # weights refined from observed rates (bench.py --balance-passes): a simulated box in which the slow ranks share a path
        # that speeds up once the fast ranks have finished -- the first measurement flatters them, the passes correct it
        obs = sh.observed_rates(float(wcount), float(wcount) / (10.0 + 7.5 * rank))
        q.put((rank, ranges, total, slowest, big, rates, (wfirst, wcount), wtotal, obs))
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
    for rank, ranges, total, slowest, big, rates, wrange, wtotal, obs in res:
        assert rates == [10.0 + 7.5 * r for r in range(world)]
        assert obs == pytest.approx(rates)                                 # units / seconds of every rank, in rank order
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


def test_refined_weights_converge_to_equal_finish_times(pg):
    """The host-fed 8-GPU box of profiles/r2_h2d_matrix_8gpu.md as a model: ranks 4-7 read host memory at 35.6 each; ranks 0-3 share
    a path worth 82 while ranks 4-7 are busy and 115 once they are done.  Weights measured with equal work over-rate the slow
    group (its tail runs alone, faster), so the first weighted shards leave it finishing last; re-measuring on the weighted
    shards, as bench.py --balance-passes does, settles on the steady-state rates and the shards finish together."""
    total = 8 * 174_762

    def simulate(counts):
        left = list(map(float, counts))
        t, done = 0.0, [0.0] * 8
        while any(x > 1e-6 for x in left):
            fast_busy = any(left[k] > 1e-6 for k in range(4, 8))
            slow = [k for k in range(4) if left[k] > 1e-6]
            rate = [0.0] * 8
            for k in slow:
                rate[k] = (82.0 if fast_busy else 115.0) / len(slow)
            for k in range(4, 8):
                rate[k] = 35.6 if left[k] > 1e-6 else 0.0
            dt = min(left[k] / rate[k] for k in range(8) if rate[k] > 0)
            t += dt
            for k in range(8):
                if rate[k] > 0:
                    left[k] -= rate[k] * dt
                    if left[k] <= 1e-6:
                        left[k], done[k] = 0.0, t
        return done

    equal = [total // 8] * 8
    weights = [c / t for c, t in zip(equal, simulate(equal))]                       # the first measurement: equal work, all at once
    assert weights[0] == pytest.approx(23.4, abs=0.3)                               # flattered: the steady-state rate is 20.5
    finish = []
    for _ in range(4):
        counts = [pg.shard_range_weighted(total, weights, s)[1] for s in range(8)]
        times = simulate(counts)
        finish.append(max(times))
        weights = [c / t for c, t in zip(counts, times)]
    ideal = total / (82.0 + 4 * 35.6)
    assert finish[0] > 1.03 * ideal                                                 # cf. profiles/r2_bench_n8.json before the passes: 211 of ~224 GB/s
    assert finish[-1] == pytest.approx(ideal, rel=2e-3) and finish == sorted(finish, reverse=True)


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
