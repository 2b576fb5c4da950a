"""CPU-side tests of the C-ABI library: it loads, exports every symbol include/perseus-gpu.h declares,
its host-only entry points (generator, shard planner, virtual receiver) behave like the reference's,
and it fails LOUDLY (no fallback) when there is no sm_100 device."""
import ctypes as C
import re
import subprocess
import tempfile
import time

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O


def test_library_exports_every_declared_symbol(pg):
    L = pg.lib()
    names = pg.declared_symbols()
    assert len(names) >= 40 and "perseus_gpu_open" in names and "perseus_gpu_unpack" in names and "perseus_gpu_close" in names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.perseus_gpu_version()


def test_header_is_plain_c_and_cites_reference():
    """The boundary must compile as C (no CUDA / C++ types) and must not pull in libusb."""
    hdr = (ROOT / "include" / "perseus-gpu.h").read_text()
    assert "#include <libusb" not in hdr and "#include \"perseus-sdr.h\"" not in hdr
    assert "perseus-sdr.h:81" in hdr and "perseustest.c:432-460" in hdr
    with tempfile.TemporaryDirectory() as td:
        src = f"{td}/t.c"
        open(src, "w").write('#include "perseus-gpu.h"\n'
                             "static int cb(void *b, int n, void *e) { (void)b; (void)n; (void)e; return 0; }\n"
                             "int main(void) { perseus_gpu_input_fn f = cb; perseus_gpu_input_fn g = perseus_gpu_input_callback;"
                             " return f == g; }\n")
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(ROOT / "include"), "-c", src, "-o", f"{td}/t.o"],
                       check=True)


@pytest.mark.parametrize("order", [("perseus-sdr.h", "perseus-gpu.h"), ("perseus-gpu.h", "perseus-sdr.h")])
def test_header_coexists_with_the_reference_header(order):
    """INTEGRATION.md step 1 compiles with the two headers in EITHER order: the reference's real perseus-sdr.h (libusb
    stubbed) and perseus-gpu.h, perseus_gpu_input_callback handed to perseus_start_async_input's own prototype
    (perseus-sdr.h:247-248), and a perseus_input_callback handed to the virtual receiver's prototype.  -pedantic
    -Werror: a re-declared typedef would be an error in C99."""
    import os
    if not os.path.exists("/root/reference/perseus-sdr.h"):
        pytest.skip("/root/reference not mounted")
    with tempfile.TemporaryDirectory() as td:
        open(f"{td}/t.c", "w").write(f'#include "{order[0]}"\n#include "{order[1]}"\n'
                                     "int start(perseus_descr *d, perseus_gpu *g)\n"
                                     "{ return perseus_start_async_input(d, 6144, perseus_gpu_input_callback, g); }\n"
                                     "int vstart(perseus_vrx *v, perseus_input_callback cb)\n"
                                     "{ perseus_gpu_input_fn same = cb; return perseus_vrx_start_async_input(v, 6144, same, 0); }\n")
        subprocess.run(["gcc", "-std=gnu99", "-pedantic", "-Wall", "-Werror", "-Wno-variadic-macros", "-I", str(ROOT / "oracle" / "stub"),
                        "-I", "/root/reference", "-I", str(ROOT / "include"), "-c", f"{td}/t.c", "-o", f"{td}/t.o"], check=True)


def test_struct_layouts_match_header(pg):
    """ctypes mirrors vs the C compiler's view of the structs."""
    with tempfile.TemporaryDirectory() as td:
        open(f"{td}/s.c", "w").write('#include <stdio.h>\n#include "perseus-gpu.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                                     "sizeof(perseus_gpu_config),sizeof(perseus_gpu_tuning),sizeof(perseus_gpu_seg),sizeof(perseus_gpu_block),"
                                     "sizeof(perseus_gpu_stats),sizeof(perseus_vrx_config),sizeof(perseus_vrx_stats),"
                                     "offsetof(perseus_gpu_config,tuning),sizeof(perseus_gpu_host_block),offsetof(perseus_gpu_config,direct_bytes),"
                                     "offsetof(perseus_gpu_config,copy_threads),offsetof(perseus_gpu_stats,host_blocks));return 0;}\n")
        subprocess.run(["gcc", "-I", str(ROOT / "include"), f"{td}/s.c", "-o", f"{td}/s"], check=True)
        sizes = [int(x) for x in subprocess.run([f"{td}/s"], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(pg.Config), C.sizeof(pg.Tuning), C.sizeof(pg.Seg), C.sizeof(pg.Block), C.sizeof(pg.Stats),
                     C.sizeof(pg.VrxConfig), C.sizeof(pg.VrxStats), pg.Config.tuning.offset, C.sizeof(pg.HostBlock),
                     pg.Config.direct_bytes.offset, pg.Config.copy_threads.offset, pg.Stats.host_blocks.offset]
    assert C.sizeof(pg.Config) == 88 + 16 and C.sizeof(pg.Stats) == 96     # ABI 2 sizes + the appended eager_gap_us / reserved2


def test_no_fallback_without_device(pg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    with pytest.raises(pg.PerseusGpuError) as e:
        pg.PerseusGpu(device=0)
    assert e.value.code in (pg.ERR["NODEVICE"], pg.ERR["CUDAERR"])
    assert e.value.msg


def test_config_struct_size_is_checked_before_anything_else(pg):
    """The config struct grows by struct_size: an ABI-2 caller's 88 bytes are accepted (the call then fails for lack of a device
    here, not for its size), sizes the library cannot know are refused."""
    L, h = pg.lib(), C.c_void_p()
    for size, refused in ((88, False), (C.sizeof(pg.Config), False), (4, True), (C.sizeof(pg.Config) + 8, True)):
        cfg = pg.Config()
        cfg.struct_size = size
        rc = L.perseus_gpu_open(C.byref(h), C.byref(cfg))
        if rc == 0:                                                    # a GPU box: the handle opened
            assert not refused
            L.perseus_gpu_close(h)
        else:
            assert (rc == pg.ERR["ERRPARAM"] and b"struct_size" in L.perseus_gpu_errorstr()) == refused, (size, rc)


def test_null_handle_errors(pg):
    L = pg.lib()
    assert L.perseus_gpu_close(None) == pg.ERR["NULLHANDLE"]
    assert L.perseus_gpu_unpack(None, None, 0, None, None, 0) == pg.ERR["NULLHANDLE"]
    assert b"null" in L.perseus_gpu_errorstr()
    assert L.perseus_gpu_input_callback(None, 6144, None) == 0          # reference callbacks always return 0
    assert L.perseus_vrx_stop_async_input(None) == pg.ERR["NULLHANDLE"]


def test_host_generator_matches_oracle_definition(pg, coracle):
    for off, n in ((0, 6144), (3, 1000), (8, 4096), (12345, 777), ((1 << 35) + 6, 600)):
        assert np.array_equal(pg.synth_fill(n, pg.SYNTH_RANDOM, O.SYNTH_SEED, off), coracle.synth_random(n, O.SYNTH_SEED, off))
    assert np.array_equal(pg.synth_fill(6 * 1000, pg.SYNTH_RAMP, 0, 6 * 16777000), coracle.synth_ramp(1000, 16777000))
    with pytest.raises(pg.PerseusGpuError):
        pg.synth_fill(60, pg.SYNTH_RAMP, 0, 5)


def test_shard_range_tiles_the_recording(pg):
    total = 11_184_810                       # cfg4: 64 GiB of 6144-byte transfers
    for n in (1, 2, 3, 4, 7, 8):
        pos = 0
        for s in range(n):
            first, count = pg.shard_range(total, n, s)
            assert first == pos and first == s * total // n
            pos += count
        assert pos == total
    assert pg.shard_range(5, 8, 7) == (4, 1) and pg.shard_range(0, 4, 2) == (0, 0)
    big = (1 << 63) + 12345
    assert sum(pg.shard_range(big, 8, s)[1] for s in range(8)) == big
    with pytest.raises(pg.PerseusGpuError):
        pg.shard_range(10, 0, 0)


# ------------------------------------------------------------------ virtual receiver (reference delivery semantics)

def reference_nearest_rate(xsr, table):
    """Restatement of getFpgaFile, /root/reference/perseus-sdr.c:776-811 (returns the rate, not the index)."""
    prev, index, vs = 0, -1, len(table)
    for i in range(vs):
        if xsr > table[i]:
            if i < vs - 1:
                prev = table[i]
                continue
            index = i
            break
        m = (table[i] + prev) // 2
        if xsr <= m:
            return table[0] if i == 0 else table[i - 1]
        return table[i]
    return table[index]


def test_rate_table_and_nearest_rate(pg):
    rates = pg.sampling_rates()
    assert rates == [48000, 95000, 96000, 125000, 192000, 250000, 500000, 1000000, 1600000, 2000000]   # perseus-sdr.h:282-285
    probes = set(rates) | {0, 1, 47999, 48001, 71500, 71501, 95499, 95500, 95501, 110500, 110501, 1299999, 1300000, 1300001,
                           1800000, 1800001, 2000001, 5_000_000}
    probes |= {(a + b) // 2 + d for a, b in zip(rates, rates[1:]) for d in (-1, 0, 1)}
    for x in sorted(probes):
        assert pg.nearest_rate(x) == reference_nearest_rate(x, rates), x
    # bitstream names = the reference's *.rbs files; the rate is encoded in the name (generate_fpga_code.sh:71-97)
    def rate_from_name(name):
        m = re.fullmatch(r"perseus(\d+)(d(\d+))?([km])24v\d+", name)
        mant = float(m.group(1) + ("." + m.group(3) if m.group(3) else ""))
        return int(round(mant * (1000 if m.group(4) == "k" else 1_000_000)))
    assert [rate_from_name(pg.bitstream_name(r)) for r in rates] == rates
    assert pg.bitstream_name(2_000_000) == "perseus2m24v21" and pg.bitstream_name(95_000) == "perseus95k24v31"   # BASELINE configs 2 and 1
    assert pg.bitstream_name(44_100) is None
    import os
    if os.path.isdir("/root/reference"):
        assert sorted(pg.bitstream_name(r) + ".rbs" for r in rates) == sorted(f for f in os.listdir("/root/reference") if f.endswith(".rbs"))
    buf = (C.c_int * 4)()
    assert pg.lib().perseus_vrx_get_sampling_rates(buf, 4) == pg.ERR["BUFFERSIZE"]   # perseus-sdr.c:825
    assert pg.lib().perseus_vrx_get_sampling_rates(buf, 0) == pg.ERR["ERRPARAM"]     # perseus-sdr.c:830


def test_vrx_buffersize_validation_matches_reference(pg):
    """perseus-sdr.c:662-680: <= 16320; %6144 with 512-byte endpoints, %510 with 510-byte ones."""
    v = pg.VirtualReceiver(sample_rate=95000)
    noop = lambda b, n, e: 0
    for size, code in ((16321, "ERRPARAM"), (20000, "ERRPARAM"), (1024, "BUFFERSIZE"), (510, "BUFFERSIZE"), (6143, "BUFFERSIZE")):
        with pytest.raises(pg.PerseusGpuError) as e:
            v.run(size, noop, None, 1)
        assert e.value.code == pg.ERR[code], size
    assert v.run(6144, noop, None, 3)["delivered"] == 3 and v.run(12288, noop, None, 2)["delivered"] == 2
    v.close()
    v = pg.VirtualReceiver(ep_max_packet=510)
    for k in (1, 2, 12, 32):
        assert v.run(510 * k, noop, None, 2)["delivered"] == 2
    with pytest.raises(pg.PerseusGpuError) as e:
        v.run(6144, noop, None, 1)
    assert e.value.code == pg.ERR["BUFFERSIZE"] and "510" in e.value.msg
    with pytest.raises(pg.PerseusGpuError):
        v.run(510 * 33, noop, None, 1)                                   # 16830 > 16320
    v.close()


def test_vrx_ring_order_lifetime_and_content(pg, coracle):
    """8-slot contiguous ring (perseus-in.c:68, perseus-sdr.c:683), cyclic in-order delivery, buffer
    reused after the callback returns (perseus-in.c:263), transfer n carries stream bytes [n*size,(n+1)*size)."""
    seen, addrs = [], []

    def cb(buf, n, extra):
        addrs.append(buf)
        seen.append(bytes((C.c_ubyte * n).from_address(buf)))
        return 0

    v = pg.VirtualReceiver(sample_rate=95000, seed=0x1234)
    st = v.run(6144, cb, None, 20)
    assert st["delivered"] == 20 and st["bytes_received"] == 20 * 6144 and st["dropped_short"] == st["dropped_sequence"] == 0
    base = addrs[0]
    assert [a - base for a in addrs] == [(k % 8) * 6144 for k in range(20)]
    want = coracle.synth_random(20 * 6144, seed=0x1234)
    assert b"".join(seen) == want.tobytes()
    v.close()


def test_vrx_cfg1_replay_through_reference_style_callback(pg, coracle):
    """BASELINE config 1: 95 kS/s, NBUF=6 x 1024-byte buffers, float unpack, on the CPU oracle."""
    v = pg.VirtualReceiver(sample_rate=95000)
    assert v.sample_rate == 95000
    chunks = []

    def cb(b, n, e):
        chunks.append(coracle.unpack(np.ctypeslib.as_array((C.c_ubyte * n).from_address(b)).copy(), O.MODE_F32))
        return 0

    v.run(6 * 1024, cb, None, 16)
    got = np.concatenate(chunks)
    want = coracle.unpack(coracle.synth_random(16 * 6144), O.MODE_F32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    v.close()


def model_reference_queue(n, drop_every=0, swap_every=0):
    """The reference's completion handler as a model (perseus-in.c:199-216,260-263): FIFO of resubmitted slots,
    deliver only the expected slot at full length, idx_expected follows the last completed slot."""
    fifo, expected, done, pos = list(range(8)), 0, 0, 0
    delivered = short = seq = 0
    order = []

    def complete(idx, is_short):
        nonlocal expected, delivered, short, seq
        if idx == expected:
            if is_short:
                short += 1
            else:
                delivered += 1
                order.append(idx)
        else:
            seq += 1
        expected = (idx + 1) % 8
        fifo.append(idx)

    while done < n:
        a = fifo.pop(0)
        if swap_every and (pos + 1) % swap_every == 0 and done + 1 < n:
            b = fifo.pop(0)
            pos += 2
            complete(b, False)
            complete(a, False)
            done += 2
            continue
        is_short = bool(drop_every) and (pos + 1) % drop_every == 0
        pos += 1
        complete(a, is_short)
        done += 1
    return delivered, short, seq, order


def test_vrx_fault_injection_follows_reference_drop_rules(pg):
    """perseus-in.c:204-216: short or out-of-sequence transfers are counted in bytes_received but never delivered.
    After one out-of-order completion the resubmission order stays permuted (perseus-in.c:263 resubmits in completion
    order), so sequence errors keep recurring: the model above and tests/test_refqueue_cpu.py (the reference's real
    code) pin that behaviour."""
    slots = []

    def cb(b, n, e):
        slots.append(b)
        return 0

    v = pg.VirtualReceiver(drop_every=5)
    st = v.run(6144, cb, None, 20)
    assert (st["delivered"], st["dropped_short"], st["dropped_sequence"]) == model_reference_queue(20, drop_every=5)[:3] == (16, 4, 0)
    assert st["bytes_received"] == 20 * 6144 - 4 * 6
    v.close()
    for swap_every, n in ((10, 30), (7, 61), (3, 40)):
        slots.clear()
        v = pg.VirtualReceiver(swap_every=swap_every)
        st = v.run(6144, cb, None, n)
        d, sh, sq, order = model_reference_queue(n, swap_every=swap_every)
        assert (st["delivered"], st["dropped_short"], st["dropped_sequence"]) == (d, sh, sq) and sq >= 3
        assert st["bytes_received"] == n * 6144
        base = min(slots)
        assert [(a - base) // 6144 for a in slots] == order or len(set(slots)) < 8
        v.close()


def test_vrx_replay_mode_repeats_the_first_ring(pg, coracle):
    got = []

    def cb(b, n, e):
        got.append(bytes((C.c_ubyte * n).from_address(b)))
        return 0

    v = pg.VirtualReceiver(seed=5, replay=True)
    v.run(6144, cb, None, 20)
    v.close()
    first = coracle.synth_random(8 * 6144, 5).reshape(8, 6144)
    assert got == [first[k % 8].tobytes() for k in range(20)]


def test_vrx_async_thread_start_stop_and_stats(pg):
    n = [0]

    def cb(b, size, e):
        n[0] += 1
        return 0

    v = pg.VirtualReceiver(sample_rate=2_000_000, realtime=True)
    v.start_async_input(6144, cb)
    with pytest.raises(pg.PerseusGpuError) as e:
        v.start_async_input(6144, cb)
    assert e.value.code == pg.ERR["ASYNCSTARTED"]                          # perseus-sdr.c:659-660
    time.sleep(0.25)
    st = v.stop_async_input()
    assert st["delivered"] == n[0] > 0
    # paced at 2 MS/s: kS/s as perseus_stop_async_input prints it (perseus-sdr.c:719-722)
    assert 1500 < st["ksamples_per_s"] < 2300, st
    with pytest.raises(pg.PerseusGpuError) as e:
        v.stop_async_input()
    assert e.value.code == pg.ERR["ASYNCSTARTED"]                          # "async input not started", perseus-sdr.c:704-706
    v.close()
