"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(include/perseus-gpu.h), against the CPU oracle / the committed golden vectors.
Bar: bit-exact for int32 AND float (the float path is one exact int->float conversion and one
correctly-rounded multiply, proven equal to the reference's division for all 2^24 codes)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FMT_CASES = None  # filled lazily (needs the package constants)


def fmt_cases(pg):
    return [("i32", pg.OUT_INT32, [O.MODE_I32, None]), ("f32", pg.OUT_FLOAT, [None, O.MODE_F32]),
            ("pow2", pg.OUT_FLOAT_POW2, [None, O.MODE_F32_POW2]), ("i32+f32", pg.OUT_INT32 | pg.OUT_FLOAT, [O.MODE_I32, O.MODE_F32]),
            ("i32+pow2", pg.OUT_INT32 | pg.OUT_FLOAT_POW2, [O.MODE_I32, O.MODE_F32_POW2])]


class DevBuf:
    """device allocation through the C ABI, freed on exit"""

    def __init__(self, h, nbytes):
        self.h, self.n = h, nbytes
        self.p = h.dev_alloc(nbytes)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.h.dev_free(self.p)


def run_unpack(pg, h, wire, flags, in_off=0, out_off=0, fill=0xA5, tail_pad=64):
    """Unpacks `wire` (numpy u8) placed at device offset in_off; returns (i32 words | None, f32 words | None).
    The output buffers are pre-filled and carry guard bytes so out-of-range writes are caught."""
    ns = wire.size // 6
    guard = 256
    with DevBuf(h, max(1, in_off + wire.size + tail_pad)) as din, DevBuf(h, out_off + ns * 8 + guard) as di, DevBuf(h, out_off + ns * 8 + guard) as df:
        if wire.size:
            h.memcpy(din.p + in_off, wire.ctypes.data, wire.size)
        h.memset(di.p, fill, di.n)
        h.memset(df.p, fill, df.n)
        want_i = bool(flags & pg.OUT_INT32)
        want_f = bool(flags & (pg.OUT_FLOAT | pg.OUT_FLOAT_POW2))
        got = h.unpack(din.p + in_off, wire.size, di.p + out_off if want_i else None, df.p + out_off if want_f else None, flags)
        assert got == ns
        res = []
        for want, d in ((want_i, di), (want_f, df)):
            raw = h.to_host(d.p, d.n, np.uint8)
            if want:
                assert (raw[:out_off] == fill).all() and (raw[out_off + ns * 8:] == fill).all(), "wrote outside the output range"
                res.append(raw[out_off:out_off + ns * 8].copy().view(np.uint32))
            else:
                assert (raw == fill).all(), "wrote to an output that was not requested"
                res.append(None)
        return res


def check_against_oracle(pg, h, coracle, wire, in_off=0, out_off=0, cases=None, tail_pad=64):
    for name, flags, modes in (cases or fmt_cases(pg)):
        got = run_unpack(pg, h, wire, flags, in_off, out_off, tail_pad=tail_pad)
        for g, m in zip(got, modes):
            if m is None:
                assert g is None
            else:
                want = coracle.unpack(wire, m).view(np.uint32).reshape(-1)
                assert np.array_equal(g, want), (name, wire.size, in_off, out_off, h.get_tuning())


@pytest.fixture(params=["stream", "direct"])
def variant(request, pg, gpu):
    gpu.set_tuning(variant=pg.VARIANT_STREAM if request.param == "stream" else pg.VARIANT_DIRECT)
    yield request.param
    gpu.set_tuning()


def test_device_is_b200_class(pg, gpu):
    L = pg.lib()
    name = C.create_string_buffer(128)
    sm, maj, mnr, mem = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
    assert L.perseus_gpu_device_info(0, name, 128, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)) == 0
    assert maj.value == 10 and sm.value >= 100, (name.value, maj.value, sm.value)


def test_golden_known_answers(pg, gpu, variant):
    kat = json.loads((GOLDEN / "kat.json").read_text())["vectors"]
    wire = np.frombuffer(b"".join(bytes.fromhex(v["wire_hex"]) + bytes.fromhex(v["q_code24"])[::-1] for v in kat), np.uint8)
    i32, f32 = run_unpack(pg, gpu, wire, pg.OUT_INT32 | pg.OUT_FLOAT)
    for k, v in enumerate(kat):
        assert f"{int(i32[2 * k]):08x}" == v["int32_hex"] and f"{int(i32[2 * k + 1]):08x}" == v["q_int32_hex"], v
        assert f"{int(f32[2 * k]):08x}" == v["float_bits_hex"] and f"{int(f32[2 * k + 1]):08x}" == v["q_float_bits_hex"], v


@pytest.mark.parametrize("name", ["xfer6144", "xfer510", "ragged1000"])
def test_golden_fixtures(pg, gpu, variant, name):
    wire = np.fromfile(GOLDEN / f"{name}.in.bin", np.uint8)
    i32, f32 = run_unpack(pg, gpu, wire, pg.OUT_INT32 | pg.OUT_FLOAT)
    assert np.array_equal(i32, np.fromfile(GOLDEN / f"{name}.i32.bin", np.uint32))
    assert np.array_equal(f32, np.fromfile(GOLDEN / f"{name}.f32.bin", np.uint32))


def test_exhaustive_2p24_all_formats(pg, gpu, coracle, variant):
    """Every 24-bit code in I, a permutation of them in Q: device ramp generator == oracle ramp, every
    format bit-exact vs the oracle, int32/float hashes == tests/golden/exhaustive.json (from oracle/_ref)."""
    g = json.loads((GOLDEN / "exhaustive.json").read_text())
    n = (1 << 24) * 6
    ramp = coracle.synth_ramp(1 << 24)
    with DevBuf(gpu, n) as din:
        gpu.generate(din.p, n, pg.SYNTH_RAMP, 0, 0)
        assert np.array_equal(gpu.to_host(din.p, n, np.uint8), ramp)
    assert f"{coracle.fnv1a64(ramp):016x}" == g["input_fnv1a64"]
    for name, flags, modes in fmt_cases(pg):
        got = run_unpack(pg, gpu, ramp, flags)
        for out, m in zip(got, modes):
            if m is None:
                continue
            assert np.array_equal(out, coracle.unpack(ramp, m, nthreads=O.host_threads()).view(np.uint32).reshape(-1)), name
            if m == O.MODE_I32:
                assert f"{coracle.fnv1a64(out):016x}" == g["int32_fnv1a64"]
            if m == O.MODE_F32:
                assert f"{coracle.fnv1a64(out):016x}" == g["float_fnv1a64"]
    # The ramp puts even codes of I in the first sample of every 12-byte unit and odd codes in the second; shifting the
    # ramp by one sample swaps them, so every code has now been through every one of the four output positions.
    shifted = coracle.synth_ramp(1 << 24, first_sample=1)
    with DevBuf(gpu, n) as din:
        gpu.generate(din.p, n, pg.SYNTH_RAMP, 0, 6)
        assert np.array_equal(gpu.to_host(din.p, n, np.uint8), shifted)
    i32, f32 = run_unpack(pg, gpu, shifted, pg.OUT_INT32 | pg.OUT_FLOAT)
    assert np.array_equal(i32, coracle.unpack(shifted, O.MODE_I32, nthreads=O.host_threads()).view(np.uint32).reshape(-1))
    assert np.array_equal(f32, coracle.unpack(shifted, O.MODE_F32, nthreads=O.host_threads()).view(np.uint32).reshape(-1))


def test_multiply_equals_ieee_division_on_device(pg, gpu):
    """SURVEY.md F2 re-proved on the B200: the product's I2F + FMUL(0x30000001) vs the verify kernel's
    I2F + correctly rounded division by 2147483392.0f, for all 2^24 codes in both fields."""
    n = (1 << 24) * 6
    with DevBuf(gpu, n) as din, DevBuf(gpu, (1 << 24) * 8) as di, DevBuf(gpu, (1 << 24) * 8) as df:
        gpu.generate(din.p, n, pg.SYNTH_RAMP, 0, 0)
        gpu.unpack(din.p, n, di.p, df.p, pg.OUT_INT32 | pg.OUT_FLOAT)
        assert gpu.verify(din.p, n, di.p, df.p, pg.OUT_INT32 | pg.OUT_FLOAT) == (0, 2 ** 64 - 1)
        gpu.unpack(din.p, n, None, df.p, pg.OUT_FLOAT_POW2)
        assert gpu.verify(din.p, n, None, df.p, pg.OUT_FLOAT_POW2)[0] == 0
        bad, first = gpu.verify(din.p, n, None, df.p, pg.OUT_FLOAT)   # POW2 output is NOT the reference scale
        assert bad == 2 * (1 << 24) - 2 and first == 2               # all but the two zero fields (I and Q of sample 0)


def test_verify_detects_single_corruption(pg, gpu, coracle):
    wire = coracle.synth_random(6144 * 7, seed=21)
    ns = wire.size // 6
    with DevBuf(gpu, wire.size) as din, DevBuf(gpu, ns * 8) as di:
        gpu.memcpy(din.p, wire.ctypes.data, wire.size)
        gpu.unpack(din.p, wire.size, di.p, None, pg.OUT_INT32)
        assert gpu.verify(din.p, wire.size, di.p, None, pg.OUT_INT32)[0] == 0
        word = np.array([0xDEADBE00], np.uint32)
        gpu.memcpy(di.p + 4 * 4321, word.ctypes.data, 4)
        assert gpu.verify(din.p, wire.size, di.p, None, pg.OUT_INT32) == (1, 4321)


SIZES = sorted(set(list(range(0, 100)) + [510 * k for k in range(1, 33)] + [510, 1020, 16320, 6143, 6144, 6145, 6150, 12287, 12288, 12289, 12294, 12300, 12288 + 48,
                                          24575, 24576, 24582, 3 * 12288 - 6, 3 * 12288 + 18, 510 * 7, 49152, 100_003]))


def test_every_small_and_ragged_size(pg, gpu, coracle, variant):
    """Empty, sub-sample, odd sample counts, sizes straddling every tile size; n%6 trailing bytes ignored."""
    big = coracle.synth_random(max(SIZES) + 16, seed=5)
    cases = [c for c in fmt_cases(pg) if c[0] in ("i32", "i32+f32")]
    for n in SIZES:
        check_against_oracle(pg, gpu, coracle, big[:n], cases=cases)


def test_unaligned_pointers_stay_exact(pg, gpu, coracle, variant):
    """Legacy 510-byte transfers and ring seams give pointers with any alignment: wire pointers at any byte,
    outputs at any multiple of 4, through the pipeline (128-bit, pre-rolled 128-bit or 32-bit stores) and the register-only kernel."""
    wire = coracle.synth_random(510 * 9 + 4, seed=6)
    for in_off in (0, 1, 2, 3, 4, 6, 8, 12, 15):
        for out_off in (0, 4, 8, 12):
            check_against_oracle(pg, gpu, coracle, wire, in_off=in_off, out_off=out_off,
                                 cases=[c for c in fmt_cases(pg) if c[0] in ("i32+f32", "pow2")])


@pytest.mark.parametrize("tile", [6144, 12288, 24576])
def test_stream_kernel_any_wire_alignment(pg, gpu, coracle, tile):
    """The TMA pipeline copies from the 16-byte boundary below the wire pointer and shifts inside shared memory, so
    every misalignment 0..15 takes the fast kernel.  The wire data ends exactly at the end of its allocation
    (tail_pad=0): with compute-sanitizer memcheck this proves the bulk copies never read past the caller's buffer."""
    gpu.set_tuning(variant=pg.VARIANT_STREAM, tile_bytes=tile, stages=3, ctas_per_sm=2)
    big = coracle.synth_random(tile * 5 + 4000, seed=tile)
    cases = [c for c in fmt_cases(pg) if c[0] == "i32+f32"]
    for delta in range(16):
        for n in (tile * 3, tile * 3 + 6, tile * 2 + 4000, tile - 6, tile + 48, tile * 4 + 16 - delta, 30, 6 * 7):
            check_against_oracle(pg, gpu, coracle, big[:n], in_off=delta, cases=cases, tail_pad=0)
    gpu.set_tuning()


@pytest.mark.parametrize("out_off", [4, 8, 12])
@pytest.mark.parametrize("tile", [6144, 12288])
def test_stream_kernel_outputs_off_16_byte_alignment(pg, gpu, coracle, tile, out_off):
    """An {I,Q} array is naturally 8-byte aligned (packed per-receiver outputs, 85-sample legacy transfers), and any multiple of
    4 is legal.  Such outputs still get 128-bit stores: every output word is made of its own 3 wire bytes, so the buffer is
    treated as if it began 1..3 words earlier, which re-aligns every store; those pre-roll words are neither read from before
    the caller's wire buffer (small in_off puts them before the allocation: memcheck) nor written before the caller's output
    (guard bytes in run_unpack).  Sizes include every tail length (words mod 4) and buffers smaller than the pre-roll unit."""
    gpu.set_tuning(variant=pg.VARIANT_STREAM, tile_bytes=tile, stages=3, ctas_per_sm=2)
    big = coracle.synth_random(tile * 5 + 4000, seed=tile + out_off)
    for delta in range(16):
        for n in (tile * 3, tile * 3 + 6, tile * 2 + 4000, tile - 6, tile, tile + 6, tile + 12, tile + 18, 6, 12, 18, 24, 30, 510):
            check_against_oracle(pg, gpu, coracle, big[:n], in_off=delta, out_off=out_off, tail_pad=0,
                                 cases=[c for c in fmt_cases(pg) if c[0] in ("i32+f32", "f32")])
    gpu.set_tuning()
    # the two outputs at different phases (0 and 8): no single pre-roll serves both -> 32-bit stores, still exact
    wire = big[:tile * 2 + 30]
    ns = wire.size // 6
    with DevBuf(gpu, wire.size) as din, DevBuf(gpu, ns * 8 + 32) as di, DevBuf(gpu, ns * 8 + 32) as df:
        gpu.memcpy(din.p, wire.ctypes.data, wire.size)
        gpu.unpack(din.p, wire.size, di.p, df.p + 8, pg.OUT_INT32 | pg.OUT_FLOAT)
        assert np.array_equal(gpu.to_host(di.p, ns * 8, np.uint32), coracle.unpack(wire, O.MODE_I32).view(np.uint32).reshape(-1))
        assert np.array_equal(gpu.to_host(df.p + 8, ns * 8, np.uint32), coracle.unpack(wire, O.MODE_F32).view(np.uint32).reshape(-1))


@pytest.mark.parametrize("tile", [6144, 9216, 12288, 18432, 24576])
@pytest.mark.parametrize("stages", [2, 3, 8])
def test_every_pipeline_geometry(pg, gpu, coracle, tile, stages):
    wire = coracle.synth_random(24576 * 3 * 148 + 6144 * 5 + 30, seed=tile + stages)   # > one wave, ragged end
    for ctas in (1, 2, 4):
        for store in (1, 2):
            try:
                gpu.set_tuning(variant=pg.VARIANT_STREAM, tile_bytes=tile, stages=stages, ctas_per_sm=ctas, store_mode=store)
            except pg.PerseusGpuError as e:
                assert e.code == pg.ERR["ERRPARAM"] and stages * tile * ctas > 150 * 1024
                continue
            check_against_oracle(pg, gpu, coracle, wire, cases=[c for c in fmt_cases(pg) if c[0] == "i32+f32"])
    gpu.set_tuning()


def test_randomised_sizes_alignments_formats_geometries(pg, gpu, coracle):
    """Seeded fuzz: 300 random (size, input misalignment, output misalignment, format, kernel variant, geometry) cases."""
    rng = np.random.default_rng(20261017)
    big = coracle.synth_random(400_000, seed=77)
    cases = fmt_cases(pg)
    tiles = [0, 6144, 9216, 12288, 18432, 24576]
    for it in range(300):
        n = int(rng.integers(0, 400_000 - 64)) if it % 3 else int(rng.integers(0, 700))
        in_off = int(rng.integers(0, 32)) if it % 2 else 0
        out_off = 4 * int(rng.integers(0, 8)) if it % 4 == 1 else 0
        tile, stages, ctas = int(rng.choice(tiles)), int(rng.choice([0, 2, 3, 5, 8])), int(rng.choice([0, 1, 2, 4]))
        try:
            gpu.set_tuning(variant=int(rng.integers(0, 3)), tile_bytes=tile, stages=stages, ctas_per_sm=ctas, store_mode=int(rng.integers(0, 3)))
        except pg.PerseusGpuError:
            gpu.set_tuning()
        start = int(rng.integers(0, 64))
        check_against_oracle(pg, gpu, coracle, big[start:start + n], in_off=in_off, out_off=out_off, cases=[cases[it % len(cases)]])
    gpu.set_tuning()


def test_offsets_beyond_4_gib(pg, gpu, coracle):
    """Byte offsets exceed 2^32 in cfg4-sized recordings: unpack 4.5 GiB in one call and compare windows around the
    4 GiB input boundary, the 4 GiB output boundary and the end against the oracle; everything else by the verify kernel."""
    nbuf = 750_000                                    # 4 608 000 000 bytes in, 6 144 000 000 bytes out
    n = nbuf * 6144
    with DevBuf(gpu, n) as din, DevBuf(gpu, n // 6 * 8) as df:
        gpu.generate(din.p, n, pg.SYNTH_RANDOM, O.SYNTH_SEED, 0)
        assert gpu.unpack(din.p, n, None, df.p, pg.OUT_FLOAT) == n // 6
        assert gpu.verify(din.p, n, None, df.p, pg.OUT_FLOAT)[0] == 0
        win = 6144 * 4
        for centre in ((1 << 32), (1 << 32) * 6 // 8, n - win // 2):      # input boundary, output boundary (in input bytes), end
            off = (centre - win // 2) // 6144 * 6144
            wire = gpu.to_host(din.p + off, win, np.uint8)
            assert np.array_equal(wire, coracle.synth_random(win, O.SYNTH_SEED, off))
            got = gpu.to_host(df.p + off // 6 * 8, win // 6 * 8, np.uint32)
            assert np.array_equal(got, coracle.unpack(wire, O.MODE_F32).view(np.uint32).reshape(-1))


def test_device_generator_matches_oracle_definition(pg, gpu, coracle):
    for off, n in ((0, 6144 * 3), (8, 100_000), (3, 1000), (12345, 7777), ((1 << 36) + 11, 4099)):
        for misalign in (0, 1, 4):
            with DevBuf(gpu, n + 16) as d:
                gpu.memset(d.p, 0xEE, n + 16)
                gpu.generate(d.p + misalign, n, pg.SYNTH_RANDOM, O.SYNTH_SEED, off)
                raw = gpu.to_host(d.p, n + 16, np.uint8)
                assert np.array_equal(raw[misalign:misalign + n], coracle.synth_random(n, O.SYNTH_SEED, off))
                assert (raw[:misalign] == 0xEE).all() and (raw[misalign + n:] == 0xEE).all()


def test_checksum_kernel_matches_oracle_and_is_shard_additive(pg, gpu, coracle):
    w = coracle.unpack(coracle.synth_random(6 * 300_001, seed=8), O.MODE_I32).reshape(-1)
    d = gpu.to_device(w)
    total = gpu.checksum(d, w.size)
    assert total == coracle.checksum32(w)
    cut = 123_457
    assert (gpu.checksum(d, cut) + gpu.checksum(d + 4 * cut, w.size - cut, first_index=cut)) % (1 << 64) == total
    gpu.dev_free(d)


def test_host_pointers_pageable_and_pinned(pg, coracle):
    """perseus_gpu_unpack with HOST buffers: staged H2D -> kernel -> D2H in chunks (chunk smaller than the input)."""
    wire = coracle.synth_random(6144 * 37 + 13, seed=9)
    ns = wire.size // 6
    want_i = coracle.unpack(wire, O.MODE_I32).view(np.uint32).reshape(-1)
    want_f = coracle.unpack(wire, O.MODE_F32).view(np.uint32).reshape(-1)
    with pg.PerseusGpu(device=0, chunk_bytes=6144 * 5, nstreams=3) as h:
        oi, of = np.zeros(ns * 2, np.uint32), np.zeros(ns * 2, np.uint32)
        assert h.unpack(wire.ctypes.data, wire.size, oi.ctypes.data, of.ctypes.data, 0) == ns      # pageable in and out
        assert np.array_equal(oi, want_i) and np.array_equal(of, want_f)
        pin = h.host_alloc(wire.size)
        pout = h.host_alloc(ns * 8)
        C.memmove(pin, wire.ctypes.data, wire.size)
        assert h.unpack(pin, wire.size, pout, None, pg.OUT_INT32) == ns                             # pinned in and out
        assert np.array_equal(np.ctypeslib.as_array((C.c_uint32 * (ns * 2)).from_address(pout)), want_i)
        with DevBuf(h, ns * 8) as df:                                                              # pinned in, device out (cfg5 shape)
            assert h.unpack(pin, wire.size, None, df.p, pg.OUT_FLOAT | pg.ASYNC) == ns
            h.sync()
            assert np.array_equal(h.to_host(df.p, ns * 8, np.uint32), want_f)
        st = h.stats()
        assert st["h2d_bytes"] == 3 * ns * 6 and st["d2h_bytes"] == 3 * ns * 8 and st["kernel_launches"] >= 3 * 8
        h.host_free(pin)
        h.host_free(pout)


@pytest.mark.parametrize("threads", [0, 1, 5, 0xFFFFFFFF])
def test_pageable_buffers_through_the_bounce_pipeline(pg, coracle, threads):
    """Ordinary application memory (numpy = malloc) in and out: chunks large enough for the helper threads to split, a ragged
    tail, more chunks than slots, mixed pointer kinds, misaligned pageable pointers, ASYNC; copy_threads = automatic, the
    caller alone, five participants, and the CUDA runtime's own staging."""
    chunk = 12288 * 90                                                # ~1.1 MB: 4 slices of >= 256 KiB
    raw = coracle.synth_random(chunk * 7 + 6144 * 3 + 17 + 1, seed=29)
    wire = raw[1:]                                                    # a pageable pointer that is not even 2-byte aligned
    ns = wire.size // 6
    want_i = coracle.unpack(wire, O.MODE_I32).view(np.uint32).reshape(-1)
    want_f = coracle.unpack(wire, O.MODE_F32).view(np.uint32).reshape(-1)
    with pg.PerseusGpu(device=0, chunk_bytes=chunk, stage_slots=3, copy_threads=threads) as h:
        oi, of = np.zeros(ns * 2 + 1, np.uint32)[1:], np.zeros(ns * 2, np.uint32)          # out_i32 only 4-byte aligned
        assert h.unpack(wire.ctypes.data, wire.size, oi.ctypes.data, of.ctypes.data, pg.ASYNC) == ns
        assert np.array_equal(oi, want_i) and np.array_equal(of, want_f)                    # complete at return although ASYNC
        with DevBuf(h, ns * 8) as df:                                                       # pageable in, device out
            assert h.unpack(wire.ctypes.data, wire.size, None, df.p, pg.OUT_FLOAT | pg.CHECKSUM) == ns
            assert h.get_checksums() == (0, coracle.checksum32(want_f))
            assert np.array_equal(h.to_host(df.p, ns * 8, np.uint32), want_f)
            of[:] = 0
            assert h.unpack(df.p - 0 + 0, 0, None, of.ctypes.data, pg.OUT_FLOAT) == 0       # empty call: nothing touched
            assert not of.any()
        with DevBuf(h, wire.size) as din:                                                   # device in, pageable out
            h.memcpy(din.p, wire.ctypes.data, wire.size)
            oi[:] = 0
            assert h.unpack(din.p, wire.size, oi.ctypes.data, None, pg.OUT_INT32) == ns
            assert np.array_equal(oi, want_i)
        st = h.stats()
        assert st["h2d_bytes"] == 2 * ns * 6 and st["d2h_bytes"] == 3 * ns * 8 + 16        # + the two checksums read back


def test_overlapped_checksum_flag(pg, coracle):
    """PERSEUS_GPU_CHECKSUM: per-piece checksums queued behind each piece's kernel add up to the oracle's checksum of the
    whole output, for the one-launch device path and the chunked host path, and restart with every call."""
    wire = coracle.synth_random(6144 * 41 + 30, seed=13)
    ns = wire.size // 6
    want_i = coracle.checksum32(coracle.unpack(wire, O.MODE_I32))
    want_f = coracle.checksum32(coracle.unpack(wire, O.MODE_F32))
    with pg.PerseusGpu(device=0, chunk_bytes=6144 * 6, nstreams=3) as h, DevBuf(h, ns * 8) as di, DevBuf(h, ns * 8) as df:
        d_in = h.to_device(wire)
        h.unpack(d_in, wire.size, di.p, df.p, pg.OUT_INT32 | pg.OUT_FLOAT | pg.CHECKSUM)
        assert h.get_checksums() == (want_i, want_f)
        h.unpack(wire.ctypes.data, wire.size, None, df.p, pg.OUT_FLOAT | pg.CHECKSUM | pg.ASYNC)      # host input, 7 chunks
        assert h.get_checksums() == (0, want_f)
        host_out = np.zeros(ns * 2, np.uint32)
        h.unpack(wire.ctypes.data, wire.size, host_out.ctypes.data, None, pg.OUT_INT32 | pg.CHECKSUM)  # host in and out
        assert h.get_checksums() == (want_i, 0) and coracle.checksum32(host_out) == want_i
        h.dev_free(d_in)


def test_checksum_totals_restart_even_for_an_empty_call(pg, coracle):
    """perseus_gpu_get_checksums reports the most recent CHECKSUM call -- also when that call held no whole sample."""
    wire = coracle.synth_random(6144 * 3, seed=5)
    with pg.PerseusGpu(device=0) as h, DevBuf(h, wire.size) as din, DevBuf(h, wire.size // 6 * 8) as di:
        h.memcpy(din.p, wire.ctypes.data, wire.size)
        h.unpack(din.p, wire.size, di.p, None, pg.OUT_INT32 | pg.CHECKSUM)
        assert h.get_checksums() == (coracle.checksum32(coracle.unpack(wire, O.MODE_I32)), 0)
        assert h.unpack(din.p, 5, di.p, None, pg.OUT_INT32 | pg.CHECKSUM) == 0
        assert h.get_checksums() == (0, 0)


def test_event_slots_belong_to_the_caller(pg, coracle):
    """All 32 timing slots are the caller's: autotune and the probes time with private events."""
    with pg.PerseusGpu(device=0) as h, DevBuf(h, 6144 * 64) as din, DevBuf(h, 8192 * 64) as di:
        h.event_record(30)
        h.unpack(din.p, din.n, di.p, None, pg.OUT_INT32)
        h.event_record(31)
        before = h.event_elapsed_ms(30, 31)
        h.probe_hbm(0, 64 << 20, 2)
        h.autotune()
        h.probe_pcie(pg.PCIE_DUPLEX, 8 << 20, 0, 1)
        assert h.event_elapsed_ms(30, 31) == before > 0


def test_pcie_probe_reports_both_directions(pg):
    with pg.PerseusGpu(device=0) as h:
        up, none = h.probe_pcie(pg.PCIE_H2D, 64 << 20)
        none2, down = h.probe_pcie(pg.PCIE_D2H, 64 << 20)
        dup_up, dup_down = h.probe_pcie(pg.PCIE_DUPLEX, 48 << 20, 128 << 20)
        assert none == none2 == 0.0 and 5 < up < 70 and 5 < down < 70, (up, down)       # PCIe Gen4/5 x16 territory
        assert 0.3 * up < dup_up <= 1.05 * up and 0.3 * down < dup_down <= 1.05 * down, (up, down, dup_up, dup_down)
        with pytest.raises(pg.PerseusGpuError):
            h.probe_pcie(7, 64 << 20)


@pytest.mark.parametrize("slots,chunk", [(2, 48 * 500), (3, 48 * 1000), (5, 48 * 333), (8, 6144 * 7)])
def test_host_pointer_pipeline_rotates_its_slots(pg, coracle, slots, chunk):
    """perseus_gpu_unpack with host pointers is a three-stream pipeline over `stage_slots` staging slots; with small chunks
    every slot is reused many times, across back-to-back ASYNC calls too.  Every pointer-kind combination stays exact."""
    wire = coracle.synth_random(chunk * 23 + 510 + 4, seed=slots)
    ns = wire.size // 6
    want_i = coracle.unpack(wire, O.MODE_I32).view(np.uint32).reshape(-1)
    want_f = coracle.unpack(wire, O.MODE_F32).view(np.uint32).reshape(-1)
    with pg.PerseusGpu(device=0, stage_slots=slots, chunk_bytes=chunk) as h:
        pin = h.host_alloc(wire.size)
        C.memmove(pin, wire.ctypes.data, wire.size)
        po_i, po_f = h.host_alloc(ns * 8), h.host_alloc(ns * 8)
        d_in = h.to_device(wire)
        d_i, d_f = h.dev_alloc(ns * 8), h.dev_alloc(ns * 8)
        page_i, page_f = np.zeros(ns * 2, np.uint32), np.zeros(ns * 2, np.uint32)
        host = lambda p: np.ctypeslib.as_array((C.c_uint32 * (ns * 2)).from_address(p))
        F = pg.OUT_INT32 | pg.OUT_FLOAT
        # pinned in -> pinned out, twice back to back without a sync in between (slot events carry over)
        h.unpack(pin, wire.size, po_i, po_f, F | pg.ASYNC)
        h.unpack(pin, wire.size, po_i, po_f, F | pg.ASYNC)
        h.sync()
        assert np.array_equal(host(po_i), want_i) and np.array_equal(host(po_f), want_f)
        # pageable in -> device out;  device in -> pageable out;  pinned in -> (device int32, pinned float)
        h.unpack(wire.ctypes.data, wire.size, d_i, d_f, F)
        assert np.array_equal(h.to_host(d_i, ns * 8, np.uint32), want_i) and np.array_equal(h.to_host(d_f, ns * 8, np.uint32), want_f)
        h.unpack(d_in, wire.size, page_i.ctypes.data, page_f.ctypes.data, F)
        assert np.array_equal(page_i, want_i) and np.array_equal(page_f, want_f)
        h.memset(d_i, 0, ns * 8); C.memset(po_f, 0, ns * 8)
        h.unpack(pin, wire.size, d_i, po_f, F | pg.CHECKSUM)
        assert np.array_equal(h.to_host(d_i, ns * 8, np.uint32), want_i) and np.array_equal(host(po_f), want_f)
        assert h.get_checksums() == (coracle.checksum32(want_i), coracle.checksum32(want_f))
        st = h.stats()
        assert st["h2d_bytes"] == 4 * (wire.size // 6 * 6) and st["d2h_bytes"] == (2 * 2 + 2 + 1) * ns * 8 + 16   # + the checksum read-back
        for p_ in (pin, po_i, po_f):
            h.host_free(p_)
        for p_ in (d_in, d_i, d_f):
            h.dev_free(p_)


def test_autotune_measures_and_keeps_a_geometry(pg, coracle):
    with pg.PerseusGpu(device=0) as h:
        assert h.get_geometry(pg.OUT_FLOAT) == {"tile_bytes": 6144, "stages": 8, "ctas_per_sm": 1}               # B200 defaults
        assert h.get_geometry(pg.OUT_INT32 | pg.OUT_FLOAT) == {"tile_bytes": 12288, "stages": 3, "ctas_per_sm": 1}
        single, fused = h.autotune()
        assert single > 4000 and fused > 4000, (single, fused)            # GB/s of the winners on a B200-class device
        tuned = (h.get_geometry(pg.OUT_INT32), h.get_geometry(pg.OUT_INT32 | pg.OUT_FLOAT_POW2))
        assert all(g["tile_bytes"] in (6144, 12288, 18432, 24576) and 2 <= g["stages"] <= 8 for g in tuned)
        check_against_oracle(pg, h, coracle, coracle.synth_random(12288 * 300 + 30, seed=3))                      # still bit-exact
        h.set_tuning(stages=2)                                            # explicit settings win over measured ones
        assert h.get_geometry(pg.OUT_FLOAT)["stages"] == 2
        h.set_tuning()
        assert (h.get_geometry(pg.OUT_INT32), h.get_geometry(pg.OUT_INT32 | pg.OUT_FLOAT_POW2)) == tuned


def test_argument_errors(pg, gpu):
    with DevBuf(gpu, 6144) as d, DevBuf(gpu, 8192) as o:
        for args, code in (((d.p, 6144, o.p, o.p, pg.OUT_FLOAT | pg.OUT_FLOAT_POW2), "ERRPARAM"),
                           ((d.p, 6144, None, None, 0), "ERRPARAM"),
                           ((d.p, 6144, None, o.p, pg.OUT_INT32), "ERRPARAM"),
                           ((d.p, 6144, o.p, None, 0x8000), "ERRPARAM"),
                           ((None, 6144, o.p, None, 0), "ERRPARAM"),
                           ((d.p, 6144, o.p + 2, None, 0), "ERRPARAM")):
            with pytest.raises(pg.PerseusGpuError) as e:
                gpu.unpack(*args)
            assert e.value.code == pg.ERR[code], args
        assert gpu.unpack(d.p, 5, o.p, None, 0) == 0 and gpu.unpack(d.p, 0, o.p, None, 0) == 0
    for bad in (dict(tile_bytes=1000), dict(stages=1), dict(stages=9), dict(ctas_per_sm=9), dict(tile_bytes=24576, stages=8, ctas_per_sm=2),
                dict(variant=7), dict(store_mode=5)):
        with pytest.raises(pg.PerseusGpuError) as e:
            gpu.set_tuning(**bad)
        assert e.value.code == pg.ERR["ERRPARAM"], bad
    gpu.set_tuning()


# ------------------------------------------------------------------ BASELINE configs

def test_cfg2_one_gib_2ms_layout_bit_exact_over_entire_output(pg, gpu, coracle):
    """BASELINE config 2: 174 762 transfers x 6144 B generated on the device, unpacked to int32 and float,
    compared word for word with the CPU oracle over the ENTIRE output, plus size-independent properties."""
    nbuf = 174_762
    n = nbuf * 6144
    ns = n // 6
    threads = O.host_threads()
    with DevBuf(gpu, n) as din, DevBuf(gpu, ns * 8) as di, DevBuf(gpu, ns * 8) as df:
        gpu.generate(din.p, n, pg.SYNTH_RANDOM, O.SYNTH_SEED, 0)
        assert gpu.unpack(din.p, n, di.p, df.p, pg.OUT_INT32 | pg.OUT_FLOAT) == ns
        wire = gpu.to_host(din.p, n, np.uint8)
        # the device generator wrote the stream the oracle defines (spot ranges, incl. the end)
        for off in (0, 6144 * 1000 + 5, n - 4099):
            assert np.array_equal(wire[off:off + 4099], coracle.synth_random(4099, O.SYNTH_SEED, off))
        want = np.empty((ns, 2), np.int32)
        coracle.unpack(wire, O.MODE_I32, nthreads=threads, out=want)
        got = gpu.to_host(di.p, ns * 8, np.uint32)
        assert np.array_equal(got, want.view(np.uint32).reshape(-1))
        cs_i = gpu.checksum(di.p, ns * 2)
        assert cs_i == coracle.checksum32(got)
        wantf = want.view(np.float32)
        coracle.unpack(wire, O.MODE_F32, nthreads=threads, out=wantf)
        got = gpu.to_host(df.p, ns * 8, np.uint32)
        assert np.array_equal(got, wantf.view(np.uint32).reshape(-1))
        del want, wantf, got
        # properties that do not need the oracle: independent per-sample kernel agrees everywhere,
        assert gpu.verify(din.p, n, di.p, df.p, pg.OUT_INT32 | pg.OUT_FLOAT)[0] == 0
        # the fused pass equals the single-format passes, and every shard split reproduces the whole
        gpu.memset(di.p, 0, ns * 8)
        for shards in (2, 8):
            for s in range(shards):
                first, count = pg.shard_range(nbuf, shards, s)
                gpu.unpack(din.p + first * 6144, count * 6144, di.p + first * 8192, None, pg.OUT_INT32 | pg.ASYNC)
            gpu.sync()
            assert gpu.checksum(di.p, ns * 2) == cs_i
            parts = [gpu.checksum(di.p + 8192 * pg.shard_range(nbuf, shards, s)[0], 2048 * pg.shard_range(nbuf, shards, s)[1],
                                  first_index=2048 * pg.shard_range(nbuf, shards, s)[0]) for s in range(shards)]
            assert sum(parts) % (1 << 64) == cs_i                         # checksum of checksums


def mixed_rate_segments(nrx, window_s):
    """cfg3 layout: receiver r runs at {48k,96k,192k,500k,1M,2M}[r%6]; window_s seconds of 6144-byte transfers."""
    rates = [48000, 96000, 192000, 500000, 1000000, 2000000]
    return [max(1, round(rates[r % 6] * window_s / 1024)) for r in range(nrx)]


def test_cfg3_mixed_rate_batch_small_against_oracle(pg, gpu, coracle):
    """Same structure as BASELINE config 3 at a size the oracle checks completely; ragged and unaligned segments included."""
    nbufs = mixed_rate_segments(96, 0.064)          # 3..125 transfers per receiver
    sizes = [b * 6144 for b in nbufs]
    sizes[5] += 510                                  # ragged: not a multiple of the tile, not of 16
    sizes[11] = 6 * 333                              # smaller than one tile
    sizes[17] = 0                                    # an idle receiver
    sizes[23] += 4                                   # trailing bytes that are not a sample
    for flags, modes in ((pg.OUT_INT32 | pg.OUT_FLOAT, (O.MODE_I32, O.MODE_F32)), (pg.OUT_FLOAT_POW2, (None, O.MODE_F32_POW2))):
        for misalign in (0, 2):                      # misalign != 0 forces the byte-load batched kernel
            wires = [coracle.synth_random(s, seed=O.SYNTH_SEED + r) for r, s in enumerate(sizes)]
            bufs = []
            segs = []
            for w in wires:
                din = DevBuf(gpu, w.size + 64); di = DevBuf(gpu, w.size // 6 * 8 + 64); df = DevBuf(gpu, w.size // 6 * 8 + 64)
                if w.size:
                    gpu.memcpy(din.p + misalign, w.ctypes.data, w.size)
                gpu.memset(di.p, 0x5A, di.n); gpu.memset(df.p, 0x5A, df.n)
                bufs.append((din, di, df))
                segs.append((din.p + misalign, w.size, di.p if modes[0] is not None else None, df.p))
            launches0 = gpu.stats()["kernel_launches"]
            total = gpu.unpack_batch(segs, flags)
            assert total == sum(s // 6 for s in sizes)
            assert gpu.stats()["kernel_launches"] == launches0 + 1           # ONE launch for all receivers
            for w, (din, di, df) in zip(wires, bufs):
                ns = w.size // 6
                for d, m in ((di, modes[0]), (df, modes[1])):
                    raw = gpu.to_host(d.p, d.n, np.uint8)
                    if m is None:
                        assert (raw == 0x5A).all()
                        continue
                    assert np.array_equal(raw[:ns * 8].view(np.uint32), coracle.unpack(w, m).view(np.uint32).reshape(-1))
                    assert (raw[ns * 8:] == 0x5A).all()
            for t in bufs:
                for b in t:
                    gpu.dev_free(b.p)


def test_batch_plan_edge_cases(pg, gpu, coracle):
    """Empty batches, and a plan outliving a change of the handle's tuning (it keeps the tile size it was built with)."""
    assert gpu.unpack_batch([], pg.OUT_INT32) == 0
    wires = [coracle.synth_random(6144 * (5 + 3 * r) + 6 * r, seed=900 + r) for r in range(7)]
    devs = [(gpu.to_device(w), DevBuf(gpu, w.size // 6 * 8 + 16)) for w in wires]
    segs = [(d, w.size, o.p, None) for (d, o), w in zip(devs, wires)]
    plan = gpu.plan_create(segs, pg.OUT_INT32)
    for tune in (dict(), dict(tile_bytes=6144, stages=5), dict(tile_bytes=24576, stages=2, ctas_per_sm=2), dict(variant=pg.VARIANT_DIRECT)):
        gpu.set_tuning(**tune)
        for _, o in devs:
            gpu.memset(o.p, 0, o.n)
        assert gpu.plan_run(plan) == sum(w.size // 6 for w in wires)
        for (d, o), w in zip(devs, wires):
            assert np.array_equal(gpu.to_host(o.p, w.size // 6 * 8, np.uint32), coracle.unpack(w, O.MODE_I32).view(np.uint32).reshape(-1)), tune
    gpu.set_tuning()
    gpu.plan_destroy(plan)
    for d, o in devs:
        gpu.dev_free(d)
        gpu.dev_free(o.p)
    with pytest.raises(pg.PerseusGpuError) as e:
        gpu.unpack_batch([(devs[0][0], 6144, None, None)], pg.OUT_INT32)          # output missing
    assert e.value.code == pg.ERR["ERRPARAM"]


def test_batch_with_misaligned_receivers_stays_one_launch(pg, gpu, coracle):
    """Outputs that are only 8- or 4-byte aligned cannot take 16-byte stores as they are.  Only THOSE receivers' segments are
    treated differently (pre-rolled by 1..3 output words), inside the same launch of the same pipeline kernel; everyone else is
    unaffected."""
    sizes = [6144 * (20 + 7 * r) + (510 if r == 3 else 0) for r in range(12)]
    odd = {5: 4, 7: 8, 9: 12, 2: 8}                    # receiver -> byte offset of its outputs (pre-rolled by 1, 2, 3, 2 words)
    wires = [coracle.synth_random(n, seed=7000 + r) for r, n in enumerate(sizes)]
    flags = pg.OUT_INT32 | pg.OUT_FLOAT
    bufs, segs = [], []
    for r, w in enumerate(wires):
        din = DevBuf(gpu, w.size + 64); di = DevBuf(gpu, w.size // 6 * 8 + 64); df = DevBuf(gpu, w.size // 6 * 8 + 64)
        gpu.memcpy(din.p + (r % 3), w.ctypes.data, w.size)          # wire pointers at any alignment, as always
        gpu.memset(di.p, 0x5A, di.n); gpu.memset(df.p, 0x5A, df.n)
        bufs.append((din, di, df))
        segs.append((din.p + (r % 3), w.size, di.p + odd.get(r, 0), df.p + odd.get(r, 0)))
    l0 = gpu.stats()["kernel_launches"]
    assert gpu.unpack_batch(segs, flags) == sum(n // 6 for n in sizes)
    assert gpu.stats()["kernel_launches"] == l0 + 1                  # still ONE launch for all receivers
    for r, (w, (din, di, df)) in enumerate(zip(wires, bufs)):
        ns, o = w.size // 6, odd.get(r, 0)
        for d, m in ((di, O.MODE_I32), (df, O.MODE_F32)):
            raw = gpu.to_host(d.p, d.n, np.uint8)
            assert np.array_equal(raw[o:o + ns * 8].copy().view(np.uint32), coracle.unpack(w, m).view(np.uint32).reshape(-1)), r
            assert (raw[:o] == 0x5A).all() and (raw[o + ns * 8:] == 0x5A).all(), r
    # tuned to the register-only kernel (the A/B variant), the whole batch goes there: also one launch, same bytes
    gpu.set_tuning(variant=pg.VARIANT_DIRECT)
    for _, di, df in bufs:
        gpu.memset(di.p, 0x5A, di.n); gpu.memset(df.p, 0x5A, df.n)
    l0 = gpu.stats()["kernel_launches"]
    gpu.unpack_batch(segs, flags)
    gpu.set_tuning()
    assert gpu.stats()["kernel_launches"] == l0 + 1
    for r, (w, (din, di, df)) in enumerate(zip(wires, bufs)):
        ns, o = w.size // 6, odd.get(r, 0)
        raw = gpu.to_host(df.p, df.n, np.uint8)
        assert np.array_equal(raw[o:o + ns * 8].copy().view(np.uint32), coracle.unpack(w, O.MODE_F32).view(np.uint32).reshape(-1)), r
        assert (raw[:o] == 0x5A).all() and (raw[o + ns * 8:] == 0x5A).all(), r
    for t in bufs:
        for b in t:
            gpu.dev_free(b.p)


def test_cfg3_full_size_1024_receivers_one_launch(pg, gpu, coracle):
    """BASELINE config 3 at full size: 1024 receivers, 652 956 transfers = 4 011 761 664 B in one launch.
    Checked on the device by the independent per-sample kernel over everything, and against the CPU oracle
    on a sample of receivers (first, last, one per rate)."""
    nbufs = mixed_rate_segments(1024, 1.024)
    assert sorted(set(nbufs)) == [48, 96, 192, 500, 1000, 2000] and sum(nbufs) == 652_956
    total_in = sum(nbufs) * 6144
    with DevBuf(gpu, total_in) as din, DevBuf(gpu, total_in // 6 * 8) as di, DevBuf(gpu, total_in // 6 * 8) as df:
        segs, off = [], 0
        for r, b in enumerate(nbufs):
            gpu.generate(din.p + off, b * 6144, pg.SYNTH_RANDOM, O.SYNTH_SEED + r, 0)     # own seed per receiver
            segs.append((din.p + off, b * 6144, di.p + off // 6 * 8, df.p + off // 6 * 8))
            off += b * 6144
        plan = gpu.plan_create(segs, pg.OUT_INT32 | pg.OUT_FLOAT)
        l0 = gpu.stats()["kernel_launches"]
        assert gpu.plan_run(plan) == total_in // 6
        assert gpu.stats()["kernel_launches"] == l0 + 1
        gpu.plan_destroy(plan)
        assert gpu.verify(din.p, total_in, di.p, df.p, pg.OUT_INT32 | pg.OUT_FLOAT)[0] == 0
        offs = np.concatenate([[0], np.cumsum(nbufs)]) * 6144
        for r in (0, 1, 2, 3, 4, 5, 511, 1023):
            n = nbufs[r] * 6144
            wire = coracle.synth_random(n, seed=O.SYNTH_SEED + r)
            o = int(offs[r]) // 6 * 8
            assert np.array_equal(gpu.to_host(di.p + o, n // 6 * 8, np.uint32), coracle.unpack(wire, O.MODE_I32).view(np.uint32).reshape(-1))
            assert np.array_equal(gpu.to_host(df.p + o, n // 6 * 8, np.uint32), coracle.unpack(wire, O.MODE_F32).view(np.uint32).reshape(-1))


# ------------------------------------------------------------------ the drop-in boundary: callback trampoline

def collect_stream(pg, h, run):
    """Installs a sink that records block descriptors, runs `run()`, returns the concatenated device outputs."""
    blocks = []
    h.set_sink(lambda blk, extra: blocks.append((blk.contents.first_sample, blk.contents.nsamples, blk.contents.dev_i32,
                                                 blk.contents.dev_f32, blk.contents.stream)))
    run()
    return blocks


SLAB_ROUTES = {"direct": 0, "staged": 0xFFFFFFFF}      # perseus_gpu_config.direct_bytes: small slabs read over the link by the kernel / copied first


@pytest.mark.parametrize("route", list(SLAB_ROUTES))
@pytest.mark.parametrize("buffersize,ep", [(6144, 512), (12288, 512), (510, 510), (510 * 32, 510)])
def test_callback_trampoline_driven_by_virtual_receiver(pg, coracle, buffersize, ep, route):
    """perseus_vrx_* plays perseus_start_async_input + the libusb queue (8-slot pageable ring, in-order
    callbacks, buffer reused right after return); the registered callback IS perseus_gpu_input_callback,
    passed as a C function pointer exactly as a libperseus-sdr user would."""
    ntransfers = 203
    slab = 48 * 1000                                   # not a multiple of the transfer: slabs split transfers
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_INT32 | pg.OUT_FLOAT, slab_bytes=slab, nslabs=3, nstreams=2,
                       max_latency_us=0xFFFFFFFF, direct_bytes=SLAB_ROUTES[route]) as h:    # slabs are counted below: size-bounded only
        outs_i, outs_f = [], []

        def sink(blk, extra):
            b = blk.contents
            h.sync()                                   # test-only: a real sink would enqueue work on b.stream instead
            outs_i.append((b.first_sample, h.to_host(b.dev_i32, b.nsamples * 8, np.uint32)))
            outs_f.append((b.first_sample, h.to_host(b.dev_f32, b.nsamples * 8, np.uint32)))

        h.set_sink(sink)
        v = pg.VirtualReceiver(sample_rate=2_000_000, ep_max_packet=ep, seed=77)
        cb, extra = h.callback
        st = v.run(buffersize, cb, extra, ntransfers)
        h.flush()
        v.close()
        assert st["delivered"] == ntransfers
        wire = coracle.synth_random(ntransfers * buffersize, seed=77)
        ns = wire.size // 6
        pos = 0
        for first, arr in outs_i:
            assert first == pos
            pos += arr.size // 2
        assert pos == ns
        assert np.array_equal(np.concatenate([a for _, a in outs_i]), coracle.unpack(wire, O.MODE_I32).view(np.uint32).reshape(-1))
        assert np.array_equal(np.concatenate([a for _, a in outs_f]), coracle.unpack(wire, O.MODE_F32).view(np.uint32).reshape(-1))
        s = h.stats()
        assert s["callbacks"] == ntransfers and s["samples"] == ns and s["h2d_bytes"] == wire.size
        assert s["slabs"] == -(-wire.size // slab)


def test_callback_with_dropped_transfers_leaves_a_gap_like_the_reference(pg, coracle):
    """Dropped transfers never reach the callback (perseus-in.c:209-216): the consumer sees the stream minus them."""
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_INT32, slab_bytes=6144 * 4, nslabs=2) as h:
        got = []

        def sink(blk, extra):
            h.sync()
            got.append(h.to_host(blk.contents.dev_i32, blk.contents.nsamples * 8, np.uint32))

        h.set_sink(sink)
        v = pg.VirtualReceiver(drop_every=4, seed=3)
        st = v.run(6144, *h.callback, 16)
        h.flush()
        v.close()
        assert st["delivered"] == 12 and st["dropped_short"] == 4
        wire = coracle.synth_random(16 * 6144, seed=3).reshape(16, 6144)
        kept = wire[[k for k in range(16) if (k + 1) % 4]].reshape(-1)
        assert np.array_equal(np.concatenate(got), coracle.unpack(kept, O.MODE_I32).view(np.uint32).reshape(-1))


@pytest.mark.parametrize("route", list(SLAB_ROUTES))
@pytest.mark.parametrize("fmt,mode,refmode", [("OUT_INT32", O.MODE_I32, O.MODE_I32), ("OUT_FLOAT", O.MODE_F32, O.MODE_F32)])
def test_stream_file_is_byte_identical_to_perseustest_output(pg, coracle, tmp_path, fmt, mode, refmode, route):
    """`perseustest -o file [-p]` writes a raw headerless stream of 8-byte samples (perseustest.c:337-343,457,499).
    The GPU path's file sink must produce the same bytes; compared with what the reference's own callbacks
    fwrite (oracle/_ref) when that library travelled here, else with the restated oracle."""
    path = tmp_path / "perseusdata"
    ntransfers = 97
    with pg.PerseusGpu(device=0, stream_flags=getattr(pg, fmt), slab_bytes=6144 * 10, nslabs=3, direct_bytes=SLAB_ROUTES[route],
                       eager_gap_us=pg.EAGER_NEVER) as h:                 # slabs are counted below
        h.stream_to_file(str(path))
        v = pg.VirtualReceiver(sample_rate=250000, seed=11)
        v.run(6144, *h.callback, ntransfers)
        h.flush()
        h.stream_to_file(None)
        v.close()
        st = h.stats()
        assert st["d2h_bytes"] == ntransfers * 8192 and st["host_blocks"] == st["slabs"] == 10
    wire = coracle.synth_random(ntransfers * 6144, seed=11)
    want = O.Ref().unpack(wire, refmode, chunk=6144) if O.Ref.available() else coracle.unpack(wire, mode)
    assert path.read_bytes() == want.tobytes()


def collect_host_blocks(got):
    """A host sink that copies each block out of the pinned memory it is only lent for the call."""
    def sink(blk, extra):
        b = blk.contents
        cp = lambda p: np.ctypeslib.as_array((C.c_uint32 * (2 * b.nsamples)).from_address(p)).copy() if p else None
        got.append((b.first_sample, b.nsamples, cp(b.i32), cp(b.f32)))
    return sink


@pytest.mark.parametrize("route", list(SLAB_ROUTES))
@pytest.mark.parametrize("fmtname", ["OUT_INT32", "OUT_FLOAT", "both"])
def test_host_sink_gets_every_block_in_stream_order(pg, coracle, route, fmtname):
    """The consumer that stays on the CPU: unpacked samples handed over in pinned host memory, once per slab, in stream
    order although the slabs rotate over two CUDA streams -- what the reference's callbacks hold when they fwrite
    (perseustest.c:457,499).  Both slab routes: read over the link by the kernel and stored straight into host memory,
    or staged through HBM by the copy engines."""
    fmt = pg.OUT_INT32 | pg.OUT_FLOAT if fmtname == "both" else getattr(pg, fmtname)
    ntransfers, slab = 157, 48 * 700
    got = []
    with pg.PerseusGpu(device=0, stream_flags=fmt, slab_bytes=slab, nslabs=3, nstreams=2, max_latency_us=0xFFFFFFFF,
                       direct_bytes=SLAB_ROUTES[route]) as h:
        h.set_host_sink(collect_host_blocks(got))
        v = pg.VirtualReceiver(sample_rate=1_000_000, seed=91)
        v.run(6144, *h.callback, ntransfers)
        h.flush()
        v.close()
        st = h.stats()
        h.set_host_sink(None)
    wire = coracle.synth_random(ntransfers * 6144, seed=91)
    ns = wire.size // 6
    nfmt = 2 if fmtname == "both" else 1
    assert st["host_blocks"] == st["slabs"] == len(got) == -(-wire.size // slab) and st["d2h_bytes"] == ns * 8 * nfmt
    assert [g[0] for g in got] == list(np.cumsum([0] + [g[1] for g in got[:-1]])) and sum(g[1] for g in got) == ns
    for idx, bit, mode in ((2, pg.OUT_INT32, O.MODE_I32), (3, pg.OUT_FLOAT, O.MODE_F32)):
        if fmt & bit:
            assert np.array_equal(np.concatenate([g[idx] for g in got]), coracle.unpack(wire, mode).view(np.uint32).reshape(-1))
        else:
            assert all(g[idx] is None for g in got)


def test_host_sink_is_served_without_a_flush(pg, coracle):
    """Delivery is driven by the device, not by the next call into the library: one full slab, nobody calls anything."""
    import time
    wire = coracle.synth_random(6144 * 2, seed=5).reshape(2, 6144)
    got = []
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_FLOAT, slab_bytes=6144, nslabs=4) as h:
        h.set_host_sink(collect_host_blocks(got))
        h.input_callback(wire[0].ctypes.data, 6144)       # allocates the slabs (tens of ms): not timed
        h.flush()
        t0 = time.perf_counter()
        h.input_callback(wire[1].ctypes.data, 6144)
        while len(got) < 2 and time.perf_counter() - t0 < 1.0:
            time.sleep(0.0002)
        dt = time.perf_counter() - t0
        assert len(got) == 2 and dt < 0.05, dt
        assert np.array_equal(got[1][3], coracle.unpack(wire[1], O.MODE_F32).view(np.uint32).reshape(-1))


def test_device_sink_host_sink_and_file_together(pg, coracle, tmp_path):
    """All three consumers of one stream see the same samples (with a device sink the slabs always land in HBM)."""
    path = tmp_path / "perseusdata"
    got, dev = [], []
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_FLOAT, slab_bytes=6144 * 3, nslabs=3) as h:
        def sink(blk, extra):
            h.sync()
            dev.append(h.to_host(blk.contents.dev_f32, blk.contents.nsamples * 8, np.uint32))
        h.set_sink(sink)
        h.set_host_sink(collect_host_blocks(got))
        h.stream_to_file(str(path))
        v = pg.VirtualReceiver(sample_rate=500_000, seed=8)
        v.run(6144, *h.callback, 20)
        h.flush()
        h.stream_to_file(None)
        v.close()
    want = coracle.unpack(coracle.synth_random(20 * 6144, seed=8), O.MODE_F32)
    assert np.array_equal(np.concatenate(dev), want.view(np.uint32).reshape(-1))
    assert np.array_equal(np.concatenate([g[3] for g in got]), want.view(np.uint32).reshape(-1))
    assert path.read_bytes() == want.tobytes()


@pytest.mark.skipif(not O.RefQueue.available(), reason="oracle/_ref/libperseus_refqueue.so not built")
@pytest.mark.parametrize("faults", [{}, {"drop_every": 9}, {"swap_every": 13}])
def test_reference_own_queue_code_drives_the_gpu_callback(pg, coracle, faults):
    """/root/reference/perseus-in.c, compiled UNMODIFIED over a fake libusb device, registers
    perseus_gpu_input_callback (a C function pointer, extra = perseus_gpu*) exactly as perseus_start_async_input does
    (perseus-sdr.c:683).  The GPU output must equal the oracle's unpack of exactly the transfers the reference's queue
    chose to deliver (same code applied with a recording Python callback)."""
    n, seed = 120, 4242
    delivered = []
    rq = O.RefQueue(seed=seed, **faults)
    rq.start(6144, lambda b, s, e: delivered.append(bytes((C.c_ubyte * s).from_address(b))) or 0)
    rq.pump(n)
    rq.close()
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_INT32 | pg.OUT_FLOAT, slab_bytes=6144 * 7, nslabs=3) as h:
        oi, of = [], []

        def sink(blk, extra):
            h.sync()
            oi.append(h.to_host(blk.contents.dev_i32, blk.contents.nsamples * 8, np.uint32))
            of.append(h.to_host(blk.contents.dev_f32, blk.contents.nsamples * 8, np.uint32))

        h.set_sink(sink)
        rq = O.RefQueue(seed=seed, **faults)
        rq.start(6144, *h.callback)
        assert rq.pump(n) == n
        assert rq.stop() == rq_bytes(n, faults)
        rq.close()
        h.flush()
        assert h.stats()["callbacks"] == len(delivered)
    wire = np.frombuffer(b"".join(delivered), np.uint8)
    assert len(delivered) < n or not faults
    assert np.array_equal(np.concatenate(oi), coracle.unpack(wire, O.MODE_I32).view(np.uint32).reshape(-1))
    assert np.array_equal(np.concatenate(of), coracle.unpack(wire, O.MODE_F32).view(np.uint32).reshape(-1))


def rq_bytes(n, faults):
    """bytes_received as perseus-in.c:202 counts it: every completed transfer, short ones with their short length."""
    d = faults.get("drop_every", 0)
    return n * 6144 - (6 * (n // d) if d else 0)


def test_streaming_latency_bound_submits_partial_slabs(pg, coracle):
    """A real receiver delivers a 6144-byte transfer every 10.8 ms at 95 kS/s: slabs must go out on time, not when full."""
    import time
    wire = coracle.synth_random(6144 * 6, seed=17).reshape(6, 6144)
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_FLOAT, slab_bytes=8 << 20, max_latency_us=2000, options=pg.OPT_NO_WATCHDOG,
                       eager_gap_us=pg.EAGER_NEVER) as h:                 # this test is about the age check alone
        blocks = []
        h.set_sink(lambda blk, extra: blocks.append((blk.contents.first_sample, blk.contents.nsamples, blk.contents.dev_f32)))
        for k in range(6):
            h.input_callback(wire[k].ctypes.data, 6144)
            time.sleep(0.004 if k % 2 else 0.0)          # every second transfer arrives after the bound has passed
        assert len(blocks) >= 2 and sum(b[1] for b in blocks) < 6 * 1024 + 1     # submitted before any flush
        h.flush()
        assert sum(b[1] for b in blocks) == 6 * 1024 and [b[0] for b in blocks] == list(np.cumsum([0] + [b[1] for b in blocks[:-1]]))
        got = np.concatenate([h.to_host(d, n * 8, np.uint32) for _, n, d in blocks])
        assert np.array_equal(got, coracle.unpack(wire.reshape(-1), O.MODE_F32).view(np.uint32).reshape(-1))
    with pg.PerseusGpu(device=0, slab_bytes=8 << 20, max_latency_us=0xFFFFFFFF) as h:   # bound disabled: only full slabs or flush
        blocks = []
        h.set_sink(lambda blk, extra: blocks.append(blk.contents.nsamples))
        for k in range(3):
            h.input_callback(wire[k].ctypes.data, 6144)
            time.sleep(0.003)
        assert blocks == []
        h.flush()
        assert blocks == [3 * 1024]


def test_prepare_takes_the_allocations_out_of_the_first_callback(pg, coracle):
    """perseus_gpu_prepare: the first callback of a real receiver must not stall libperseus-sdr's poll thread for the tens of
    milliseconds pinned allocations and the first kernel launch take (8 transfers of ring cover 4 ms at 2 MS/s)."""
    import time
    wire = coracle.synth_random(6144 * 3, seed=52).reshape(3, 6144)
    got = []
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_INT32 | pg.OUT_FLOAT) as h:         # default slabs: 4 x 8 MiB pinned + device buffers
        h.set_host_sink(collect_host_blocks(got))
        h.prepare()
        assert h.stats()["samples"] == 0 and h.stats()["slabs"] == 0                     # the warm-up launch delivers and counts nothing
        times = []
        for k in range(3):
            t0 = time.perf_counter()
            h.input_callback(wire[k].ctypes.data, 6144)
            times.append(time.perf_counter() - t0)
            time.sleep(0.001)
        h.flush()
        assert times[0] < 0.008 and sorted(times)[1] < 0.004, times                       # typically 2-30 us each; unprepared: 17-51 ms, then 4-9 ms
        assert np.array_equal(np.concatenate([g[2] for g in got]), coracle.unpack(wire.reshape(-1), O.MODE_I32).view(np.uint32).reshape(-1))
        assert np.array_equal(np.concatenate([g[3] for g in got]), coracle.unpack(wire.reshape(-1), O.MODE_F32).view(np.uint32).reshape(-1))
        h.prepare()                                                                        # idempotent


def test_slow_stream_goes_out_transfer_by_transfer_fast_stream_fills_slabs(pg, coracle):
    """perseus_gpu_config.eager_gap_us: a transfer that arrives a while after the previous one (a real receiver: one every
    0.5-10.8 ms) is submitted at once, with the handle's DEFAULT 8 MiB slabs and 50 ms bound; the same handle fed back to back
    (a replayed recording) batches into slabs again."""
    import time
    wire = coracle.synth_random(6144 * 12, seed=41).reshape(12, 6144)
    got = []
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_INT32) as h:
        h.set_host_sink(collect_host_blocks(got))
        h.input_callback(wire[0].ctypes.data, 6144)       # first callback ever: allocations, and no idle time to go by
        h.flush()
        for k in range(1, 12):
            time.sleep(0.002)
            h.input_callback(wire[k].ctypes.data, 6144)
            t0 = time.perf_counter()
            while len(got) < k + 1 and time.perf_counter() - t0 < 0.5:
                time.sleep(0.0001)
            assert len(got) == k + 1, (k, len(got))       # in host memory without a flush, a poll, or the 50 ms bound
        st = h.stats()
        assert st["slabs"] == 12 and st["watchdog_submits"] == 0 and [g[1] for g in got] == [1024] * 12
        assert np.array_equal(np.concatenate([g[2] for g in got]), coracle.unpack(wire.reshape(-1), O.MODE_I32).view(np.uint32).reshape(-1))
        # back to back: ~3000 transfers (18 MB) from the virtual receiver's loop -> three slabs, give or take a scheduling hiccup
        h.set_host_sink(None)
        v = pg.VirtualReceiver(sample_rate=2_000_000, seed=6)
        v.run(6144, *h.callback, 3000)
        h.flush()
        v.close()
        assert 3 <= h.stats()["slabs"] - 12 <= 40, h.stats()          # 3 when nothing interrupts the loop; each hiccup of the box adds one


def test_latency_bound_holds_without_a_following_callback(pg, coracle):
    """The stream stalls after three transfers (USB error, or the tail before perseus_stop_async_input): the partial slab
    must still reach the device within max_latency_us -- by the handle's watchdog thread, or by perseus_gpu_poll when the
    application disabled the watchdog."""
    import time
    wire = coracle.synth_random(6144 * 3, seed=23).reshape(3, 6144)
    want = coracle.unpack(wire.reshape(-1), O.MODE_F32).view(np.uint32).reshape(-1)
    bound = 0.020
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_FLOAT, slab_bytes=8 << 20, max_latency_us=int(bound * 1e6),
                       eager_gap_us=pg.EAGER_NEVER) as h:                 # a burst that stops: the watchdog's case
        h.input_callback(wire[0].ctypes.data, 6144)       # first callback of a handle allocates its slabs (tens of ms): not timed
        h.flush()
        blocks = []
        h.set_sink(lambda blk, extra: blocks.append((time.perf_counter(), blk.contents.first_sample, blk.contents.nsamples, blk.contents.dev_f32)))
        t0 = time.perf_counter()
        for k in range(3):
            h.input_callback(wire[k].ctypes.data, 6144)
        assert blocks == []                               # not full, not over age: nothing submitted yet
        while not blocks and time.perf_counter() - t0 < 1.0:
            time.sleep(0.001)                             # NO further callback, no flush
        assert len(blocks) == 1, "the watchdog did not submit the partial slab"
        t, first, n, dev = blocks[0]
        assert (first, n) == (1024, 3 * 1024) and bound * 0.9 <= t - t0 < bound + 0.05, t - t0   # slack for a busy box, not for the library
        h.sync()
        assert np.array_equal(h.to_host(dev, n * 8, np.uint32), want)
        st = h.stats()
        assert st["watchdog_submits"] == 1 and st["slabs"] == 2
        # the stream resumes: sample numbering continues, the watchdog goes back to sleep
        h.input_callback(wire[0].ctypes.data, 6144)
        h.flush()
        assert [(b[1], b[2]) for b in blocks] == [(1024, 3072), (4096, 1024)]
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_FLOAT, slab_bytes=8 << 20, max_latency_us=int(bound * 1e6),
                       options=pg.OPT_NO_WATCHDOG, eager_gap_us=pg.EAGER_NEVER) as h:
        blocks = []
        h.set_sink(lambda blk, extra: blocks.append((blk.contents.first_sample, blk.contents.nsamples, blk.contents.dev_f32)))
        for k in range(3):
            h.input_callback(wire[k].ctypes.data, 6144)
        assert h.poll() == 0                              # not over age yet
        time.sleep(bound * 3)
        assert blocks == []                               # nobody is watching: that is what the option asks for
        assert h.poll() == 1 and len(blocks) == 1 and h.poll() == 0
        h.sync()
        assert np.array_equal(h.to_host(blocks[0][2], 3072 * 8, np.uint32), want)
        assert h.stats()["watchdog_submits"] == 1


@pytest.mark.skipif(not O.Ref.available(), reason="oracle/_ref not built")
def test_cfg1_95k_paced_float_file_equals_reference(pg, coracle, tmp_path):
    """BASELINE config 1 on the GPU as specified: perseus95k24v31 (95 000 S/s), NBUF=6 x 1024-byte buffers = 6144-byte
    transfers (perseustest.c:100-102,349), float output (-p), paced in real time, the handle's DEFAULT configuration
    (8 MiB slabs, 50 ms latency bound, watchdog).  A transfer arrives every 10.8 ms, so slabs are cut by time, not by
    size.  The file must equal what the reference's own float callback writes for the delivered transfers."""
    import time
    path = tmp_path / "perseusdata"
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_FLOAT) as h:
        h.stream_to_file(str(path))
        v = pg.VirtualReceiver(sample_rate=95000, realtime=True, seed=95)
        assert v.sample_rate == 95000 and pg.bitstream_name(v.sample_rate) == "perseus95k24v31"
        v.start_async_input(6 * 1024, *h.callback)
        time.sleep(0.6)
        st = v.stop_async_input()
        mid = h.stats()
        h.flush()
        h.stream_to_file(None)
        v.close()
        n = st["delivered"]
        assert 40 <= n <= 95 and 80 < st["ksamples_per_s"] < 110, st       # 0.6 s (more if the box is busy) at 92.8 transfers/s
        # a transfer every 10.8 ms goes out on arrival (eager_gap_us), long before any flush or age bound; transfers the receiver's
        # pacing thread delivers back to back to catch up with its schedule (after the first callback, which allocates the slabs;
        # whenever the box oversleeps) share a slab, as they should -- how many do is the box's business, so only the age bound's
        # floor is asserted here (test_slow_stream_goes_out_transfer_by_transfer... pins the eager path with its own pacing)
        assert 8 <= mid["slabs"] <= n and mid["callbacks"] == n and mid["samples"] >= (n - 8) * 1024, mid
    wire = coracle.synth_random(n * 6144, seed=95)
    assert path.read_bytes() == O.Ref().unpack(wire, O.MODE_F32, chunk=6144).tobytes()


def test_streaming_while_the_application_hammers_the_same_handle(pg, coracle, tmp_path):
    """One handle, three parties at once: the receiver's thread calling perseus_gpu_input_callback as fast as it can (lock-free
    fast path), the application thread polling statistics / perseus_gpu_poll / perseus_gpu_sync (each of which takes the handle
    away from the callback through the membarrier handshake), and the latency watchdog.  The file must still be the unpack of
    every transfer, in order, and no transfer may be lost or counted twice."""
    import threading
    n = 30_000                                           # 184 MB of wire, a few hundred hand-offs
    path = tmp_path / "perseusdata"
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_INT32, slab_bytes=1 << 20, nslabs=4, max_latency_us=1000) as h:
        h.stream_to_file(str(path))
        v = pg.VirtualReceiver(sample_rate=2_000_000, seed=31337)
        cbp, cbx = h.callback
        done, res = threading.Event(), {}

        def feed():
            res["st"] = v.run(6144, cbp, cbx, n)
            done.set()

        t = threading.Thread(target=feed)
        t.start()
        polls = 0
        while not done.is_set():
            s = h.stats()
            assert s["callbacks"] <= n and s["dropped_callbacks"] == 0
            h.poll()
            if polls % 5 == 0:
                h.sync()
            polls += 1
        t.join()
        h.flush()
        h.stream_to_file(None)
        v.close()
        s = h.stats()
        assert res["st"]["delivered"] == n and s["callbacks"] == n and s["samples"] == n * 1024 and polls > 20, (s, polls)
    got = np.fromfile(path, np.uint32)
    want = coracle.unpack(coracle.synth_random(n * 6144, seed=31337), O.MODE_I32, nthreads=4).view(np.uint32).reshape(-1)
    assert got.size == want.size and np.array_equal(got, want)


def test_sink_on_the_watchdog_thread_may_use_its_handle(pg, coracle):
    """The sink runs on whichever thread submits the slab -- here the watchdog -- and may call the handle's plumbing from there."""
    import time
    wire = coracle.synth_random(6144 * 2, seed=77).reshape(2, 6144)
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_FLOAT, max_latency_us=5000, eager_gap_us=pg.EAGER_NEVER) as h:
        got, seen = [], []

        def sink(blk, extra):
            h.sync()
            got.append(h.to_host(blk.contents.dev_f32, blk.contents.nsamples * 8, np.uint32))
            seen.append(h.stats()["callbacks"])

        h.set_sink(sink)
        for k in range(2):
            h.input_callback(wire[k].ctypes.data, 6144)
        t0 = time.perf_counter()
        while not got and time.perf_counter() - t0 < 2.0:
            time.sleep(0.001)
        assert len(got) == 1 and seen == [2] and h.stats()["watchdog_submits"] == 1
        assert np.array_equal(got[0], coracle.unpack(wire.reshape(-1), O.MODE_F32).view(np.uint32).reshape(-1))


def test_two_handles_on_two_threads(pg, coracle):
    """Handles are independent: distinct handles may be driven from distinct threads at the same time."""
    import threading
    wire = [coracle.synth_random(6144 * 300 + 6 * k, seed=30 + k) for k in range(2)]
    out, err = [None, None], []

    def work(k):
        try:
            with pg.PerseusGpu(device=0) as h:
                d = h.to_device(wire[k])
                ns = wire[k].size // 6
                o = h.dev_alloc(ns * 8)
                for _ in range(20):
                    h.unpack(d, wire[k].size, o, None, pg.OUT_INT32)
                out[k] = h.to_host(o, ns * 8, np.uint32)
                h.dev_free(d); h.dev_free(o)
        except Exception as e:  # pragma: no cover
            err.append(e)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not err, err
    for k in range(2):
        assert np.array_equal(out[k], coracle.unpack(wire[k], O.MODE_I32).view(np.uint32).reshape(-1))


def test_several_receivers_served_by_one_thread(pg, coracle):
    """libperseus-sdr serves up to 8 receivers from ONE poll thread (perseus-sdr.c:43,736-770): their callbacks interleave
    on that thread, each with its own `extra`.  One perseus_gpu handle per receiver; streams and slabs are independent."""
    nrx, ntransfers = 4, 37
    wires = [coracle.synth_random(ntransfers * 6144, seed=500 + r).reshape(ntransfers, 6144) for r in range(nrx)]
    handles = [pg.PerseusGpu(device=0, stream_flags=pg.OUT_FLOAT if r % 2 else pg.OUT_INT32, slab_bytes=6144 * (3 + r), nslabs=2,
                             max_latency_us=0xFFFFFFFF) for r in range(nrx)]
    outs = [[] for _ in range(nrx)]

    def make_sink(r):
        def sink(blk, extra):
            b = blk.contents
            handles[r].sync()
            outs[r].append(handles[r].to_host(b.dev_f32 if r % 2 else b.dev_i32, b.nsamples * 8, np.uint32))
        return sink

    for r, h in enumerate(handles):
        h.set_sink(make_sink(r))
    for k in range(ntransfers):                      # the poll thread's view: transfers of all receivers, interleaved
        for r, h in enumerate(handles):
            h.input_callback(wires[r][k].ctypes.data, 6144)
    for r, h in enumerate(handles):
        h.flush()
        want = coracle.unpack(wires[r].reshape(-1), O.MODE_F32 if r % 2 else O.MODE_I32).view(np.uint32).reshape(-1)
        assert np.array_equal(np.concatenate(outs[r]), want), r
        assert h.stats()["callbacks"] == ntransfers
        h.close()


def test_callback_errors_are_latched_and_surface_at_flush(pg, coracle):
    """The reference ignores callback return values (perseus-in.c:207), so the trampoline returns 0 even when it fails;
    the failure is reported by the next perseus_gpu_flush / close, once, and the handle stays usable."""
    h = pg.PerseusGpu(device=0, slab_bytes=1 << 44, nslabs=2)            # 16 TiB pinned slabs cannot be allocated
    buf = coracle.synth_random(6144, seed=1)
    assert h.input_callback(buf.ctypes.data, 6144) == 0
    assert h.input_callback(buf.ctypes.data, 6144) == 0                  # further transfers are dropped, and counted
    assert h.input_callback(buf.ctypes.data, 6000 + 5) == 0
    st = h.stats()
    # the transfer that hit the failure never made it into a slab: it is counted with the ones dropped after it
    assert (st["callbacks"], st["dropped_callbacks"], st["dropped_bytes"]) == (0, 3, 6144 + 6144 + 6000)
    with pytest.raises(pg.PerseusGpuError) as e:
        h.flush()
    assert e.value.code == pg.ERR["CUDAERR"] and "cudaHostAlloc" in e.value.msg
    h.flush()                                                            # reported once
    with DevBuf(h, 6144) as d, DevBuf(h, 8192) as o:                     # bulk path still works on the same handle
        h.memcpy(d.p, buf.ctypes.data, 6144)
        assert h.unpack(d.p, 6144, o.p, None, pg.OUT_INT32) == 1024
        assert np.array_equal(h.to_host(o.p, 8192, np.uint32), coracle.unpack(buf, O.MODE_I32).view(np.uint32).reshape(-1))
    h.close()


def test_stream_to_file_rejects_two_formats(pg):
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_INT32 | pg.OUT_FLOAT) as h:
        with pytest.raises(pg.PerseusGpuError) as e:
            h.stream_to_file("/tmp/never")
        assert e.value.code == pg.ERR["ERRPARAM"]


def test_sharded_recording_reassembles(pg, coracle):
    """cfg4 structure on one GPU: each shard generates ITS byte range of the recording on the device
    (random access generator) and unpacks it; shard checksums add up to the oracle's whole-recording checksum."""
    nbuf = 4099
    wire = coracle.synth_random(nbuf * 6144, seed=O.SYNTH_SEED)
    want = coracle.checksum32(coracle.unpack(wire, O.MODE_F32))
    for shards in (2, 4, 8):
        acc = 0
        for s in range(shards):
            first, count = pg.shard_range(nbuf, shards, s)
            with pg.PerseusGpu(device=0) as h, DevBuf(h, count * 6144) as din, DevBuf(h, count * 8192) as df:
                h.generate(din.p, count * 6144, pg.SYNTH_RANDOM, O.SYNTH_SEED, first * 6144)
                h.unpack(din.p, count * 6144, None, df.p, pg.OUT_FLOAT)
                acc += h.checksum(df.p, count * 2048, first_index=first * 2048)
        assert acc % (1 << 64) == want
