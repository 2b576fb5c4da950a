"""SURVEY.md §8(f) row n1 as written, on the CPU: the reference's WHOLE library -- perseus-sdr.c, perseusfx2.c,
perseus-in.c, perseuserr.c compiled UNMODIFIED into oracle/_ref/libperseus_sdr_ref.so -- runs over a synthetic receiver
behind a fake libusb (oracle/fakeusb.c), driven through its public API exactly as examples/perseustest.c drives it.

Two kinds of test:
  * against tests/golden/reflib.json, which tests/golden/make_golden_reflib.py wrote by EXECUTING that library: the
    product's nearest-rate rule vs the reference's static getFpgaFile (perseus-sdr.c:776-811), its rate table vs
    perseus_get_sampling_rates, and perseus_vrx_start_async_input's codes and messages vs perseus_start_async_input's
    (perseus-sdr.c:662-680).  These run wherever the repository is (the GPU box has no /root/reference);
  * live, when oracle/_ref travelled here: the same comparisons against the running reference, the whole bring-up
    sequence, its poll thread, every transfer status of perseus-in.c:199-257 against the product's virtual receiver.
"""
import ctypes as C
import json
import threading
import time

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import oracle as O

GOLD = json.loads((GOLDEN / "reflib.json").read_text())
live = pytest.mark.skipif(not O.RefLib.available(), reason="oracle/_ref/libperseus_sdr_ref.so not built")
RATES = [48000, 95000, 96000, 125000, 192000, 250000, 500000, 1000000, 1600000, 2000000]


# ------------------------------------------------------------------ frozen outputs of the reference (run anywhere)

def test_nearest_rate_equals_the_reference_getFpgaFile_golden(pg):
    """perseus_vrx_nearest_rate vs what the reference's getFpgaFile made perseus_set_sampling_rate send to the FPGA."""
    assert len(GOLD["rate_choice"]) >= 60
    for x, want in GOLD["rate_choice"].items():
        assert pg.nearest_rate(int(x)) == want, x
    for a, b in zip(RATES, RATES[1:]):                       # every midpoint is covered, +-1
        assert {str((a + b) // 2 + d) for d in (-1, 0, 1)} <= set(GOLD["rate_choice"])


def test_rate_table_and_bitstream_names_equal_the_reference_table(pg):
    assert GOLD["sampling_rates"]["16"]["values"][:10] == pg.sampling_rates() == RATES
    buf = (C.c_int * 16)()
    for size, g in GOLD["sampling_rates"].items():            # same codes, values and messages for short buffers
        rc = pg.lib().perseus_vrx_get_sampling_rates(buf, int(size))
        assert rc == g["rc"], size
        assert list(buf)[:int(size)] == g["values"]
        if rc < 0:
            assert pg.errorstr() == g["message"]
    names = {b["rate"]: b["name"] for b in GOLD["bitstreams"]}
    assert sorted(names) == RATES
    for r in RATES:                                           # FpgaImages.name carries the .rbs extension
        assert pg.bitstream_name(r) + ".rbs" == names[r]


def vrx_start(pg, ep, size):
    """(rc, message, second_start_rc) of the product's virtual receiver for one buffer size."""
    noop = pg.INPUT_CALLBACK(lambda b, n, e: 0)
    try:
        v = pg.VirtualReceiver(sample_rate=2_000_000, ep_max_packet=ep)
    except pg.PerseusGpuError as e:                           # an endpoint size the reference rejects at start is rejected at open
        return e.code, e.msg, None
    L = pg.lib()
    rc = L.perseus_vrx_start_async_input(v.v, size, C.cast(noop, C.c_void_p), None)
    msg = pg.errorstr() if rc < 0 else ""
    second = None
    if rc == 0:
        second = L.perseus_vrx_start_async_input(v.v, size, C.cast(noop, C.c_void_p), None)
        assert L.perseus_vrx_stop_async_input(v.v) == 0
    stop_rc = L.perseus_vrx_stop_async_input(v.v)
    stop_msg = pg.errorstr()
    v.close()
    return rc, msg, second, stop_rc, stop_msg


def test_start_async_input_codes_and_messages_equal_the_reference_golden(pg):
    seen = 0
    for key, g in GOLD["start_codes"].items():
        ep, size = key.split(":")
        if size == "stop_when_not_started":
            continue
        got = vrx_start(pg, int(ep), int(size))
        if int(ep) == 64 and int(size) <= 16320:              # rejected by perseus_vrx_open with the reference's words
            assert got[:2] == (g["rc"], g["message"]), key
            continue
        if int(ep) == 64:                                     # reference: size check first; here the endpoint fails first
            assert got[0] == g["rc"] == pg.ERR["ERRPARAM"]
            continue
        assert got[0] == g["rc"] and got[1] == g["message"], (key, got, g)
        if g["rc"] == 0:
            assert got[2] == g["second_start_rc"] == pg.ERR["ASYNCSTARTED"]
        stop = GOLD["start_codes"][f"{ep}:stop_when_not_started"]
        assert (got[3], got[4]) == (stop["rc"], stop["message"])
        seen += 1
    assert seen >= 30


# ------------------------------------------------------------------ the running reference

@pytest.fixture(scope="module")
def reflib():
    if not O.RefLib.available():
        pytest.skip("oracle/_ref/libperseus_sdr_ref.so not built")
    return O.RefLib()


def recorder(store):
    return lambda buf, size, extra: store.append((buf, size, bytes((C.c_ubyte * size).from_address(buf)))) or 0


@live
def test_bring_up_sequence_of_perseustest_runs_on_the_reference(reflib, coracle):
    """perseus_init -> open -> firmware_download -> set_sampling_rate -> set_attenuator / adc / ddc -> start -> stop -> exit
    (perseustest.c:188-390) against the synthetic receiver; the callback is made by the reference's own poll thread."""
    L, main = reflib.L, threading.get_ident()
    calls, threads = [], set()

    def cb(buf, size, extra):
        threads.add(threading.get_ident())
        calls.append(bytes((C.c_ubyte * size).from_address(buf)))
        return 0

    with reflib.session(bring_up=False, limit=40, seed=1234, serial=4711) as d:
        assert L.reflib_descr_firmware_downloaded(d) == 1
        assert L.perseus_start_async_input(d, 6144, None, None) == O.PERSEUS_ERR["FPGANOTCFGD"]     # perseus-sdr.c:656-657
        assert L.perseus_firmware_download(d, None) == 0
        pid = O.EepromProdId()
        assert L.perseus_get_product_id(d, C.byref(pid)) == 0 and (pid.sn, pid.prodcode) == (4711, 0x8014)
        assert L.perseus_is_preserie(d, None) == 0
        assert L.perseus_set_sampling_rate(d, 95000) == 0 and L.reflib_descr_fpga_configured(d) == 1
        st = reflib.state()
        g = next(b for b in GOLD["bitstreams"] if b["rate"] == 95000)
        assert (st["fpga_rate"], st["fpga_bytes"], f"{st['fpga_hash']:016x}") == (95000, g["size"], g["fnv1a64"])   # perseus95k24v31, whole
        assert L.perseus_set_attenuator_n(d, 1) == 0 and reflib.state()["porte"] >> 4 == 1         # perseus-sdr.c:514
        assert L.perseus_set_adc(d, 1, 0) == 0 and reflib.state()["sio_ctl"] & 0x06 == 0x02         # dither on, preamp off
        assert L.perseus_set_ddc_center_freq(d, 7_050_000.0, 1) == 0
        assert reflib.state()["sio_freg"] == int(7_050_000.0 / 80_000_000.0 * 4294967296.0)        # perseus-sdr.c:584
        assert L.perseus_start_async_input(d, 6 * 1024, reflib.callback_pointer(cb), None) == 0    # perseustest.c:349
        assert reflib.state()["fifo_enabled"] == 1                                                 # perseus-sdr.c:687-688
        assert L.perseus_start_async_input(d, 6144, None, None) == O.PERSEUS_ERR["ASYNCSTARTED"]
        reflib.wait_stream_pos(40)
        assert L.perseus_stop_async_input(d) == 0                                                  # cancel loop, perseus-sdr.c:714-716
        st = reflib.state()
        assert (st["cancelled"], st["fifo_enabled"], st["completed_ok"]) == (8, 0, 40)
        assert L.reflib_descr_bytes_received(d) == 40 * 6144
        assert st["events_calls"] > 0 and len(threads) == 1 and main not in threads               # its poll thread, not ours
        assert L.perseus_stop_async_input(d) == O.PERSEUS_ERR["ASYNCSTARTED"]
    assert b"".join(calls) == coracle.synth_random(40 * 6144, 1234).tobytes()
    st = reflib.state()
    assert st["shutdowns"] == 1 and st["closes"] == st["opens"] == 1 and st["exits"] == 1          # perseus_close: perseus-sdr.c:306-307


@live
def test_blank_fx2_gets_the_reference_firmware(reflib):
    """A receiver with a blank EEPROM enumerates as a bare Cypress FX2 (04B4:8613): perseus_firmware_download writes the 306
    Intel-HEX records of perseus24v41_512 into its RAM, releases the 8051 and re-opens the re-enumerated device
    (perseus-sdr.c:381-468, perseusfx2.c:164-202; includes the reference's own 4 s re-enumeration sleep)."""
    L = reflib.L
    with reflib.session(bring_up=False, blank_eeprom=1, preserie=1) as d:
        assert L.reflib_descr_firmware_downloaded(d) == 0
        assert L.perseus_set_sampling_rate(d, 95000) == O.PERSEUS_ERR["FWNOTLOADED"]
        t0 = time.monotonic()
        assert L.perseus_firmware_download(d, None) == 0, reflib.errorstr()
        assert time.monotonic() - t0 >= 4.0
        st = reflib.state()
        fw = GOLD["firmware"]
        assert (st["fw_records"], st["fw_bytes"], f"{st['fw_hash']:016x}") == (fw["records"], fw["bytes"], fw["fnv1a64"])
        assert st["cpu_resets"] == 1 and st["firmware_loaded"] == 1 and st["opens"] == 2
        assert L.reflib_descr_firmware_downloaded(d) == 1 and L.reflib_descr_is_preserie(d) == 1
        assert L.perseus_is_preserie(d, None) == O.PERSEUS_ERR["SNNOTAVAILABLE"]
        assert L.perseus_set_sampling_rate(d, 2_000_000) == 0 and reflib.state()["fpga_rate"] == 2_000_000


@live
def test_nearest_rate_equals_the_running_reference(pg, reflib):
    """Live version of the golden test: perseus_set_sampling_rate(x) on the reference, the bitstream the fake FPGA received,
    against perseus_vrx_nearest_rate(x)."""
    with reflib.session(rate=0) as d:
        for x in [1, 48000, 71500, 71501, 95500, 95501, 110500, 110501, 1300000, 1300001, 1800000, 1800001, 2_000_000, 9_999_999]:
            assert reflib.L.perseus_set_sampling_rate(d, x) == 0
            assert reflib.state()["fpga_rate"] == pg.nearest_rate(x), x
        for n, r in enumerate(RATES):                          # perseus_set_sampling_rate_n, perseus-sdr.c:869-892
            assert reflib.L.perseus_set_sampling_rate_n(d, n) == 0 and reflib.state()["fpga_rate"] == r
        assert reflib.L.perseus_set_sampling_rate_n(d, 10) == O.PERSEUS_ERR["ERRPARAM"]


FAULTS = [{}, {"drop_every": 5}, {"swap_every": 7}, {"timeout_every": 6}, {"fail_at": 11, "fail_status": 1},
          {"fail_at": 3, "fail_status": 4}, {"fail_at": 20, "fail_status": 5, "drop_every": 4}, {"fail_at": 9, "fail_status": 6, "swap_every": 5},
          {"timeout_every": 4, "swap_every": 9, "drop_every": 7}]


@live
@pytest.mark.parametrize("faults", FAULTS, ids=lambda f: "-".join(f"{k}{v}" for k, v in f.items()) or "clean")
@pytest.mark.parametrize("buffersize,ep", [(6144, 512), (510 * 3, 510)])
def test_virtual_receiver_equals_the_reference_library_under_every_status(pg, reflib, faults, buffersize, ep):
    """Same synthetic stream, same fault schedule: the reference (perseus_start_async_input + its poll thread + perseus-in.c's
    completion handler, statuses COMPLETED / short / out of sequence / TIMED_OUT / ERROR / STALL / NO_DEVICE / OVERFLOW,
    perseus-in.c:199-257) and perseus_vrx_* must make the same callbacks: ring-slot offsets, order, bytes, byte count."""
    n = 61
    ref_calls, calls = [], []
    with reflib.session(rate=2_000_000, ep_max_packet=ep, limit=n, seed=0xABCD, **faults) as d:
        assert reflib.L.perseus_start_async_input(d, buffersize, reflib.callback_pointer(recorder(ref_calls)), None) == 0
        base = reflib.L.reflib_descr_ring(d)
        st = reflib.wait_stream_pos(n)
        assert reflib.L.perseus_stop_async_input(d) == 0
        ref_bytes = reflib.L.reflib_descr_bytes_received(d)
        st = reflib.state()
    v = pg.VirtualReceiver(sample_rate=2_000_000, ep_max_packet=ep, seed=0xABCD, **faults)
    vst = v.run(buffersize, recorder(calls), None, n)
    v.close()
    assert calls, "nothing delivered"
    # transfer 1 is never faulted in these schedules: it lands in slot 0 of either ring
    vbase = calls[0][0]
    assert [(a - vbase, s, b) for a, s, b in calls] == [(a - base, s, b) for a, s, b in ref_calls]
    assert vst["delivered"] == len(ref_calls) and vst["bytes_received"] == ref_bytes
    assert vst["timed_out"] == st["timed_out"] and vst["retired"] == st["failed"]
    assert vst["delivered"] + vst["dropped_short"] + vst["dropped_sequence"] + vst["timed_out"] + vst["retired"] == n
    if "fail_at" in faults:                                   # the retired slot's ring position is skipped from then on: sequence errors recur
        assert vst["retired"] == 1 and vst["dropped_sequence"] >= (n - faults["fail_at"]) // 8


@live
def test_a_stalled_stream_times_out_and_is_rearmed_by_the_reference(reflib):
    """After `limit` transfers the device sends nothing: each pending transfer ends LIBUSB_TRANSFER_TIMED_OUT after the
    reference's 8 x 80 ms (perseus-in.c:35,59); the handler logs it and resubmits (perseus-in.c:218-221,263), no callback."""
    got = []
    with reflib.session(rate=2_000_000, limit=5) as d:
        assert reflib.L.perseus_start_async_input(d, 6144, reflib.callback_pointer(recorder(got)), None) == 0
        reflib.wait_stream_pos(5)
        time.sleep(0.75)
        st = reflib.state()
        assert st["timed_out"] >= 8 and st["submits"] >= 8 + 5 + 8 and len(got) == 5
        assert reflib.L.perseus_stop_async_input(d) == 0
        assert reflib.L.reflib_descr_bytes_received(d) == 5 * 6144


@live
def test_reference_poll_thread_paces_like_the_bitstream(reflib):
    """realtime: the fake FPGA fills a 6144-byte transfer every 1024 samples at the configured rate; the reference's stop
    statistics (perseus-sdr.c:719-722: bytes_received / elapsed / 6000) then read the sample rate."""
    with reflib.session(rate=1_000_000, realtime=1) as d:
        n = []
        assert reflib.L.perseus_start_async_input(d, 6144, reflib.callback_pointer(lambda b, s, e: n.append(s) or 0), None) == 0
        t0 = time.monotonic()
        time.sleep(0.3)
        assert reflib.L.perseus_stop_async_input(d) == 0
        dt = time.monotonic() - t0
        ksps = reflib.L.reflib_descr_bytes_received(d) / dt / 6000.0
        assert 800 < ksps < 1100, ksps
        assert len(n) == reflib.L.reflib_descr_bytes_received(d) // 6144


APP = ROOT / "oracle" / "_ref" / "perseustest_ref"


def run_reference_app(out_file, float_output=False, limit=200, seed=77, rate=250000, extra_env=None):
    """Runs the reference's own application, unmodified, over the synthetic receiver; returns its exit status."""
    import os
    import subprocess
    env = dict(os.environ, FAKEUSB_AUTOPLUG="1", FAKEUSB_LIMIT=str(limit), FAKEUSB_SEED=str(seed), **(extra_env or {}))
    args = [str(APP), "-a", "-d", "0", "-s", str(rate), "-n", "6", "-b", "1024", "-t", "1", "-o", str(out_file)] + (["-p"] if float_output else [])
    return subprocess.run(args, env=env, capture_output=True, text=True, timeout=60)


@pytest.mark.skipif(not APP.exists(), reason="oracle/_ref/perseustest_ref not built")
@pytest.mark.parametrize("float_output,mode,name", [(False, O.MODE_I32, "int32"), (True, O.MODE_F32, "float")])
def test_reference_application_output_file(coracle, tmp_path, float_output, mode, name):
    """examples/perseustest.c's main(), unmodified, from argument parsing to fclose: its output file is what the callbacks
    alone write for the same transfers (oracle/_ref/libperseus_ref.so), and matches the hash frozen in reflib.json."""
    out = tmp_path / "perseusdata"
    r = run_reference_app(out, float_output)
    assert "Perseus receivers found" in r.stderr and "Bye" in r.stderr, r.stderr[-2000:]
    data = out.read_bytes()
    wire = coracle.synth_random(200 * 6144, 77)
    assert data == O.Ref().unpack(wire, mode, chunk=6144).tobytes()
    g = GOLD["perseustest_app"]
    assert (len(data), f"{coracle.fnv1a64(data):016x}") == (g[f"{name}_nbytes"], g[f"{name}_fnv1a64"])


def test_simple_c_application_wrote_the_same_int32_file(coracle):
    """examples/simple.c carries its own copy of the int32 callback (simple.c:33-61) and is the reference's `make check` program.
    It runs 10 s, so it is executed only by tests/golden/make_golden_reflib.py; its frozen output must be what perseustest wrote
    for the same stream, and what the restated oracle computes."""
    g, p = GOLD["simple_app"], GOLD["perseustest_app"]
    assert (g["limit"], g["seed"]) == (p["limit"], p["seed"])
    assert (g["int32_nbytes"], g["int32_fnv1a64"]) == (p["int32_nbytes"], p["int32_fnv1a64"])
    want = coracle.unpack(coracle.synth_random(g["limit"] * 6144, g["seed"]), O.MODE_I32)
    assert f"{coracle.fnv1a64(want):016x}" == g["int32_fnv1a64"]
