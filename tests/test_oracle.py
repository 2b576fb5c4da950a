"""CPU tests that PIN the oracle: restated C == numpy == the reference's own callbacks
(oracle/_ref, verbatim /root/reference/examples/perseustest.c:432-502) == tests/golden/."""
import json

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as O

# Known answers produced by the reference's own code during the survey.
SURVEY_KAT = [  # (24-bit code, int32 bits, float bits)
    (0x000000, 0x00000000, 0x00000000),
    (0x000001, 0x00000100, 0x34000001),
    (0x7FFFFF, 0x7FFFFF00, 0x3F800000),
    (0x800000, 0x80000000, 0xBF800001),
    (0xFFFFFF, 0xFFFFFF00, 0xB4000001),
    (0x123456, 0x12345600, 0x3E11A2B1),
]


def wire(code_i, code_q=0):
    return bytes((code_i & 255, code_i >> 8 & 255, code_i >> 16 & 255, code_q & 255, code_q >> 8 & 255, code_q >> 16 & 255))


def impls(coracle, ref_or_none):
    out = {"c": lambda b, m: coracle.unpack(b, m), "numpy": O.np_unpack}
    if ref_or_none is not None:
        out["ref"] = lambda b, m: ref_or_none.unpack(b, m, chunk=max(6, len(bytes(b)))) if m < 2 else None
    return out


@pytest.fixture(scope="module")
def maybe_ref():
    return O.Ref() if O.Ref.available() else None


def test_survey_known_answers(coracle, maybe_ref):
    buf = np.frombuffer(b"".join(wire(c, c) for c, _, _ in SURVEY_KAT), np.uint8)
    for name, fn in impls(coracle, maybe_ref).items():
        i32 = fn(buf, O.MODE_I32).view(np.uint32)
        f32 = fn(buf, O.MODE_F32).view(np.uint32)
        for k, (_, want_i, want_f) in enumerate(SURVEY_KAT):
            assert int(i32[k, 0]) == want_i and int(i32[k, 1]) == want_i, (name, k)
            assert int(f32[k, 0]) == want_f and int(f32[k, 1]) == want_f, (name, k)


def test_golden_kat(coracle):
    kat = json.loads((GOLDEN / "kat.json").read_text())["vectors"]
    buf = np.frombuffer(b"".join(bytes.fromhex(v["wire_hex"]) + wire(int(v["q_code24"], 16))[:3] for v in kat), np.uint8)
    for fn in (lambda b, m: coracle.unpack(b, m), O.np_unpack):
        i32 = fn(buf, O.MODE_I32).view(np.uint32)
        f32 = fn(buf, O.MODE_F32).view(np.uint32)
        for k, v in enumerate(kat):
            assert f"{int(i32[k, 0]):08x}" == v["int32_hex"] and f"{int(i32[k, 1]):08x}" == v["q_int32_hex"]
            assert f"{int(f32[k, 0]):08x}" == v["float_bits_hex"] and f"{int(f32[k, 1]):08x}" == v["q_float_bits_hex"]


@pytest.mark.parametrize("name", ["xfer6144", "xfer510", "ragged1000"])
def test_golden_fixtures(coracle, name):
    meta = json.loads((GOLDEN / "fixtures.json").read_text())[name]
    b = np.fromfile(GOLDEN / f"{name}.in.bin", np.uint8)
    assert b.size == meta["in_bytes"]
    for mode, ext in ((O.MODE_I32, "i32"), (O.MODE_F32, "f32")):
        want = np.fromfile(GOLDEN / f"{name}.{ext}.bin", np.uint32)
        assert want.size == 2 * meta["samples"] == 2 * (b.size // 6)
        assert np.array_equal(coracle.unpack(b, mode).view(np.uint32).reshape(-1), want)
        assert np.array_equal(O.np_unpack(b, mode).view(np.uint32).reshape(-1), want)


def test_golden_stream_hashes(coracle):
    meta = json.loads((GOLDEN / "fixtures.json").read_text())["stream96"]
    b = coracle.synth_random(meta["in_bytes"], seed=int(meta["seed"], 16))
    assert f"{coracle.fnv1a64(b):016x}" == meta["input_fnv1a64"]
    assert f"{coracle.fnv1a64(coracle.unpack(b, O.MODE_I32)):016x}" == meta["int32_fnv1a64"]
    assert f"{coracle.fnv1a64(coracle.unpack(b, O.MODE_F32)):016x}" == meta["float_fnv1a64"]


def test_exhaustive_2p24_against_golden_hashes(coracle):
    """All 2^24 codes of the I field (and a permutation of them in Q): hashes from oracle/_ref."""
    g = json.loads((GOLDEN / "exhaustive.json").read_text())
    ramp = coracle.synth_ramp(1 << 24)
    assert f"{coracle.fnv1a64(ramp):016x}" == g["input_fnv1a64"]
    for name, mode in (("int32", O.MODE_I32), ("float", O.MODE_F32)):
        out = coracle.unpack(ramp, mode, nthreads=4)
        assert out.nbytes == g[f"{name}_nbytes"]
        assert f"{coracle.fnv1a64(out):016x}" == g[f"{name}_fnv1a64"]
        assert f"{coracle.checksum32(out):016x}" == g[f"{name}_checksum32"]


def test_exhaustive_2p24_restated_equals_reference_verbatim(coracle, ref):
    """The pin itself: the reference's callbacks, 6144 bytes per call, vs the restatement, every code."""
    ramp = coracle.synth_ramp(1 << 24)
    for mode in (O.MODE_I32, O.MODE_F32):
        a = ref.unpack(ramp, mode, chunk=6144).view(np.uint32)
        b = coracle.unpack(ramp, mode).view(np.uint32)
        assert np.array_equal(a, b)
        assert np.array_equal(O.np_unpack(ramp, mode).view(np.uint32), b)


def test_float_scale_facts():
    """SURVEY.md F1/F2/F3 over every 24-bit code (numpy, IEEE RN)."""
    v = np.arange(1 << 24, dtype=np.uint32)
    x = (v << np.uint32(8)).view(np.int32)                     # F3: MSB aligned == v << 8
    b = np.zeros((1 << 24, 6), np.uint8)
    for k in range(3):
        b[:, k] = (v >> np.uint32(8 * k)).astype(np.uint8)
    assert np.array_equal(O.np_unpack_i32(b.reshape(-1))[:, 0], x)
    xf = x.astype(np.float32)
    assert np.array_equal(xf.astype(np.int64), x.astype(np.int64))   # (float)int32 exact: 24 significant bits
    div = xf / np.float32(2147483392.0)
    recip = np.uint32(0x30000001).view(np.float32)
    assert np.array_equal(div.view(np.uint32), (xf * recip).view(np.uint32))          # F2
    pow2 = xf * np.float32(2.0 ** -31)
    assert int((div.view(np.uint32) != pow2.view(np.uint32)).sum()) == (1 << 24) - 1   # F1: differs except at 0
    assert float(div.max()) == 1.0 and div.min() < -1.0


@pytest.mark.parametrize("nbytes", [0, 1, 5, 6, 7, 11, 12, 47, 48, 509, 510, 1020, 6143, 6144, 6145, 12288, 16320])
def test_ragged_sizes_ignore_trailing_bytes(coracle, maybe_ref, nbytes):
    """perseustest.c:443: nSamples = buf_size/6 — the tail is dropped, never read as a sample."""
    b = coracle.synth_random(nbytes, seed=7)
    for mode in (O.MODE_I32, O.MODE_F32, O.MODE_F32_POW2):
        c = coracle.unpack(b, mode)
        assert c.shape == (nbytes // 6, 2)
        assert np.array_equal(c.view(np.uint32), O.np_unpack(b, mode).view(np.uint32))
        if maybe_ref is not None and mode != O.MODE_F32_POW2 and nbytes:
            assert np.array_equal(maybe_ref.unpack(b, mode, chunk=nbytes).view(np.uint32), c.view(np.uint32))


def test_every_legal_transfer_size(coracle, ref):
    """perseus-sdr.c:662-680: 6144 and 12288 with the shipped firmware, 510*k (k = 1..32, <= 16320) with the legacy one."""
    wire = coracle.synth_random(16320 * 3, seed=12)
    for size in [6144, 12288] + [510 * k for k in range(1, 33)]:
        assert size % 6 == 0 and size <= 16320
        for mode in (O.MODE_I32, O.MODE_F32):
            a = ref.unpack(wire[: size * 3], mode, chunk=size).view(np.uint32)       # three transfers of that size
            assert np.array_equal(a, coracle.unpack(wire[: size * 3], mode).view(np.uint32))


def test_per_transfer_chunking_is_stateless(coracle, ref):
    """Every legal transfer size is a whole number of samples (perseus-sdr.c:671-676), so
    calling the callback per transfer == unpacking the concatenation."""
    b = coracle.synth_random(510 * 6 * 4, seed=3)
    whole = coracle.unpack(b, O.MODE_I32)
    for chunk in (510, 1020, 3060, 6120):
        assert np.array_equal(ref.unpack(b, O.MODE_I32, chunk=chunk), whole)


def test_mt_equals_st(coracle):
    b = coracle.synth_random(6 * 100_003, seed=11)
    for mode in (0, 1, 2):
        assert np.array_equal(coracle.unpack(b, mode, nthreads=1).view(np.uint32),
                              coracle.unpack(b, mode, nthreads=7).view(np.uint32))


def test_synth_generators_c_equals_numpy(coracle):
    for off, n in ((0, 100), (1, 99), (7, 64), (8, 1000), (12345, 6144), ((1 << 33) + 5, 77)):
        assert np.array_equal(coracle.synth_random(n, O.SYNTH_SEED, off), O.np_synth_random(n, O.SYNTH_SEED, off))
    whole = coracle.synth_random(5000, 99, 0)
    assert np.array_equal(whole[1234:2345], coracle.synth_random(1111, 99, 1234))      # random access
    assert np.array_equal(coracle.synth_ramp(5000, 16777000), O.np_synth_ramp(5000, 16777000))
    # every 24-bit field of the random stream is roughly uniform: both signs, all byte values
    f = O.np_unpack_i32(coracle.synth_random(6 * 200_000, 5))
    assert 0.45 < float((f < 0).mean()) < 0.55


def test_checksum_is_shard_additive(coracle):
    w = coracle.unpack(coracle.synth_random(6 * 10_000, 1), O.MODE_I32).reshape(-1)
    total = coracle.checksum32(w)
    assert total == O.np_checksum32(w)
    parts = [coracle.checksum32(w[a:b], first_index=a) for a, b in ((0, 3333), (3333, 12000), (12000, 20000))]
    assert sum(parts) % (1 << 64) == total
    w2 = w.copy(); w2[[5, 6]] = w2[[6, 5]]
    assert coracle.checksum32(w2) != total or w[5] == w[6]
    assert coracle.fnv1a64(b"perseus") == O.fnv1a64_py(b"perseus")
