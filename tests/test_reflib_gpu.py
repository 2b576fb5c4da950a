"""SURVEY.md §8(f) row n1 on the GPU: the reference's WHOLE library (perseus-sdr.c, perseusfx2.c, perseus-in.c, perseuserr.c,
compiled unmodified into oracle/_ref/libperseus_sdr_ref.so over a fake libusb) performs its real bring-up and then
    perseus_start_async_input(descr, 6144, perseus_gpu_input_callback, gpu)
so the product's trampoline is called, as a plain C function pointer, on the reference's OWN poll thread (SCHED_FIFO when
the box allows it, perseus-sdr.c:736-774), fed by the reference's own transfer queue, and stopped through the reference's
own cancel handshake (perseus-sdr.c:708-716).  What the GPU path writes must be byte-identical to what the reference's own
callbacks (oracle/_ref/libperseus_ref.so, examples/perseustest.c compiled verbatim) write for the same transfers."""
import ctypes as C
import os
import subprocess
import time

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (O.RefLib.available() and O.Ref.available()), reason="oracle/_ref not built")]


@pytest.fixture(scope="module")
def reflib():
    return O.RefLib()


def reference_file(wire, mode, chunk):
    return O.Ref().unpack(wire, mode, chunk=chunk).tobytes()


@pytest.mark.parametrize("fmt,mode", [("OUT_INT32", O.MODE_I32), ("OUT_FLOAT", O.MODE_F32)])
@pytest.mark.parametrize("buffersize,ep,rate", [(6144, 512, 2_000_000), (12288, 512, 500_000), (510 * 4, 510, 250_000)])
def test_reference_library_drives_the_gpu_callback_on_its_poll_thread(pg, coracle, reflib, tmp_path, fmt, mode, buffersize, ep, rate):
    n, seed = 300, 20261017
    path = tmp_path / "perseusdata"
    with pg.PerseusGpu(device=0, stream_flags=getattr(pg, fmt), slab_bytes=6144 * 11, nslabs=3) as h:
        h.stream_to_file(str(path))                            # perseustest -o path [-p]
        cb, extra = h.callback
        with reflib.session(rate=rate, ep_max_packet=ep, limit=n, seed=seed) as d:
            assert reflib.L.perseus_start_async_input(d, buffersize, cb, extra) == 0, reflib.errorstr()
            reflib.wait_stream_pos(n)
            assert reflib.L.perseus_stop_async_input(d) == 0   # returns once all 8 transfers reported cancelled
            st = reflib.state()
            assert (st["cancelled"], st["completed_ok"], st["fpga_rate"]) == (8, n, rate)
            assert reflib.L.reflib_descr_bytes_received(d) == n * buffersize
        h.flush()
        h.stream_to_file(None)
        s = h.stats()
        assert s["callbacks"] == n and s["samples"] == n * buffersize // 6 and s["dropped_callbacks"] == 0
    wire = coracle.synth_random(n * buffersize, seed=seed)
    assert path.read_bytes() == reference_file(wire, mode, buffersize)


@pytest.mark.parametrize("faults", [{"drop_every": 9}, {"swap_every": 13}, {"timeout_every": 7}, {"fail_at": 50, "fail_status": 1},
                                    {"fail_at": 17, "fail_status": 6, "drop_every": 5}],
                         ids=lambda f: "-".join(f"{k}{v}" for k, v in f.items()))
def test_gpu_callback_sees_exactly_what_the_reference_queue_delivers_under_faults(pg, coracle, reflib, faults):
    """Short, out-of-sequence, timed-out and failed transfers never reach the callback (perseus-in.c:209-257): the GPU output is
    the reference callbacks' output for exactly the transfers the reference's queue delivered to a recording callback."""
    n, seed = 160, 99
    delivered = []
    with reflib.session(rate=2_000_000, limit=n, seed=seed, **faults) as d:
        rec = reflib.callback_pointer(lambda b, s, e: delivered.append(bytes((C.c_ubyte * s).from_address(b))) or 0)
        assert reflib.L.perseus_start_async_input(d, 6144, rec, None) == 0
        reflib.wait_stream_pos(n)
        reflib.L.perseus_stop_async_input(d)
    assert 0 < len(delivered) < n
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_INT32 | pg.OUT_FLOAT, slab_bytes=6144 * 7, nslabs=3) as h:
        oi, of = [], []

        def sink(blk, extra):
            h.sync()
            oi.append(h.to_host(blk.contents.dev_i32, blk.contents.nsamples * 8, np.uint32))
            of.append(h.to_host(blk.contents.dev_f32, blk.contents.nsamples * 8, np.uint32))

        h.set_sink(sink)
        with reflib.session(rate=2_000_000, limit=n, seed=seed, **faults) as d:
            assert reflib.L.perseus_start_async_input(d, 6144, *h.callback) == 0
            reflib.wait_stream_pos(n)
            assert reflib.L.perseus_stop_async_input(d) == 0
        h.flush()
        assert h.stats()["callbacks"] == len(delivered)
    wire = np.frombuffer(b"".join(delivered), np.uint8)
    assert np.concatenate(oi).tobytes() == reference_file(wire, O.MODE_I32, 6144)
    assert np.concatenate(of).tobytes() == reference_file(wire, O.MODE_F32, 6144)


@pytest.mark.parametrize("eager", [True, False])
def test_cfg1_through_the_reference_library_paced_at_95k(pg, coracle, reflib, tmp_path, eager):
    """BASELINE config 1 end to end: perseus95k24v31 selected by the reference's perseus_set_sampling_rate(95000), 6 x 1024
    byte transfers, float (-p), paced in real time by the fake FPGA, the product handle at its DEFAULT configuration.
    By default every transfer goes out on arrival (eager_gap_us: the stream is slower than the GPU path); with that switched off
    slabs are cut by the 50 ms bound, and when the stream stops arriving (the queue is cancelled) the tail still reaches the
    file: latency watchdog + flush."""
    path = tmp_path / "perseusdata"
    with pg.PerseusGpu(device=0, stream_flags=pg.OUT_FLOAT, eager_gap_us=0 if eager else pg.EAGER_NEVER) as h:
        h.stream_to_file(str(path))
        with reflib.session(rate=95000, realtime=1, seed=5) as d:
            assert reflib.state()["fpga_rate"] == 95000
            assert reflib.L.perseus_start_async_input(d, 6 * 1024, *h.callback) == 0
            time.sleep(0.5)
            assert reflib.L.perseus_stop_async_input(d) == 0
            n = reflib.state()["completed_ok"]
            assert 35 <= n <= 90, n                              # 0.5 s (more if the box is busy) at 92.8 transfers/s
        time.sleep(0.08)                                         # nothing arrives any more: the watchdog (50 ms) submits the tail
        st = h.stats()
        # perseus_input_queue_cancel clears callback_fn on the application thread (perseus-in.c:131) while the poll thread may be
        # completing one more transfer: the fake device has counted it, the reference's handler no longer delivers it
        # (perseus-in.c:204-207).  The reference itself decides which of the two happens.
        ncb = st["callbacks"]
        assert n - 1 <= ncb <= n and st["samples"] == ncb * 1024, (n, st)
        # eager: (nearly) one slab per transfer, on arrival -- transfers the fake FPGA delivers back to back to catch up share one;
        # not eager: slabs cut by the 50 ms bound, the tail by the watchdog
        assert (8 <= st["slabs"] <= ncb) if eager else (st["watchdog_submits"] >= 1 and st["slabs"] <= 14), st
        h.flush()
        h.stream_to_file(None)
    assert path.read_bytes() == reference_file(coracle.synth_random(ncb * 6144, seed=5), O.MODE_F32, 6144)


DEMO = ROOT / "oracle" / "_ref" / "perseus_gpu_libperseus_demo"


@pytest.mark.skipif(not DEMO.exists(), reason="oracle/_ref/perseus_gpu_libperseus_demo not built")
@pytest.mark.parametrize("flag,mode", [((), O.MODE_I32), (("-p",), O.MODE_F32)])
def test_c_program_linking_libperseus_sdr_and_libperseus_gpu(coracle, tmp_path, flag, mode):
    """tests/integration/perseus_gpu_libperseus.c is INTEGRATION.md §1 as a program: plain C against the reference's perseus-sdr.h and
    perseus-gpu.h, linked with the reference library (over the fake libusb) and libperseus_gpu.so."""
    out = tmp_path / "perseusdata"
    env = dict(os.environ, LD_LIBRARY_PATH=f"{ROOT / 'oracle' / '_ref'}:{ROOT / 'libperseus-sdr_b200' / 'lib'}:" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(DEMO), "-s", "250000", "-n", "6", "-b", "1024", "-t", "200", "-o", str(out), *flag], env=env, capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "200 transfers" in r.stdout
    wire = coracle.synth_random(200 * 6144, seed=O.SYNTH_SEED)
    assert out.read_bytes() == reference_file(wire, mode, 6144)


APP = ROOT / "oracle" / "_ref" / "perseustest_ref"


@pytest.mark.skipif(not (DEMO.exists() and APP.exists()), reason="oracle/_ref programs not built")
@pytest.mark.parametrize("flag", [(), ("-p",)])
def test_gpu_program_writes_the_same_file_as_the_reference_application(tmp_path, flag):
    """Two programs, the same synthetic receiver, the same command line letters: the reference's own perseustest (unmodified,
    CPU callbacks) and tests/integration/perseus_gpu_libperseus.c (the same flow with perseus_gpu_input_callback).  Their output
    files must be byte-identical."""
    ref_out, gpu_out = tmp_path / "ref.bin", tmp_path / "gpu.bin"
    env = dict(os.environ, FAKEUSB_AUTOPLUG="1", FAKEUSB_LIMIT="200", FAKEUSB_SEED=str(O.SYNTH_SEED),
               LD_LIBRARY_PATH=f"{ROOT / 'oracle' / '_ref'}:{ROOT / 'libperseus-sdr_b200' / 'lib'}:" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(APP), "-a", "-d", "0", "-s", "250000", "-n", "6", "-b", "1024", "-t", "1", "-o", str(ref_out), *flag], env=env,
                       capture_output=True, text=True, timeout=60)
    assert "Bye" in r.stderr, r.stderr[-2000:]
    env.pop("FAKEUSB_AUTOPLUG")
    g = subprocess.run([str(DEMO), "-s", "250000", "-n", "6", "-b", "1024", "-t", "200", "-o", str(gpu_out), *flag], env=env, capture_output=True,
                       text=True, timeout=120)
    assert g.returncode == 0, g.stderr + g.stdout
    assert ref_out.stat().st_size == 200 * 8192 and ref_out.read_bytes() == gpu_out.read_bytes()
