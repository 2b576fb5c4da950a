"""SURVEY.md §8(f) row n1 on the CPU: the reference's OWN transfer queue (perseus-in.c, compiled unmodified into
oracle/_ref/libperseus_refqueue.so over a fake libusb device) against the product's virtual receiver
(perseus_vrx_*, libperseus_gpu.so).  Same synthetic stream, same fault schedule -> the two must make the same
callbacks: same ring-slot addresses in the same order, same bytes, same drop accounting."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.RefQueue.available(), reason="oracle/_ref/libperseus_refqueue.so not built")


def record_reference(buffersize, n, **faults):
    calls = []
    rq = O.RefQueue(seed=0xABCD, **faults)
    rq.start(buffersize, lambda buf, size, extra: calls.append((buf, size, bytes((C.c_ubyte * size).from_address(buf)))) or 0)
    base = rq.ring
    assert rq.pump(n) == n
    received = rq.bytes_received
    stopped = rq.stop()
    rq.close()
    return [(a - base, s, b) for a, s, b in calls], received, stopped


def record_vrx(pg, buffersize, n, ep=0, **faults):
    calls = []
    v = pg.VirtualReceiver(sample_rate=2_000_000, ep_max_packet=ep, seed=0xABCD, **faults)
    st = v.run(buffersize, lambda buf, size, extra: calls.append((buf, size, bytes((C.c_ubyte * size).from_address(buf)))) or 0, None, n)
    v.close()
    base = calls[0][0] - 0 if calls else 0
    return calls, st


@pytest.mark.parametrize("buffersize,ep", [(6144, 512), (12288, 512), (510 * 3, 510)])
@pytest.mark.parametrize("faults", [{}, {"drop_every": 5}, {"swap_every": 7}, {"drop_every": 4, "swap_every": 9},
                                    {"timeout_every": 5}, {"fail_at": 10, "fail_status": 1}, {"fail_at": 2, "fail_status": 6, "swap_every": 6}],
                         ids=lambda f: "-".join(f"{k}{v}" for k, v in f.items()) or "clean")
def test_virtual_receiver_equals_reference_queue(pg, buffersize, ep, faults):
    """Statuses other than COMPLETED included (perseus-in.c:218-257): TIMED_OUT re-arms the slot, ERROR / OVERFLOW retire it."""
    n = 61
    ref_calls, ref_received, ref_stopped = record_reference(buffersize, n, **faults)
    calls, st = record_vrx(pg, buffersize, n, ep=ep, **faults)
    # slot 0 of each ring is where the first delivered transfer (stream position 0, never faulted here) landed
    vbase = calls[0][0]
    assert [(a - vbase, s, b) for a, s, b in calls] == ref_calls
    assert st["delivered"] == len(ref_calls)
    assert st["bytes_received"] == ref_received == ref_stopped       # cancelled transfers add nothing (perseus-in.c:191-196)
    assert st["dropped_short"] + st["dropped_sequence"] + st["timed_out"] + st["retired"] == n - len(ref_calls)
    assert st["timed_out"] == (n // faults["timeout_every"] if "timeout_every" in faults else 0)
    assert st["retired"] == (1 if "fail_at" in faults else 0)


def test_retired_slot_leaves_seven_transfers_in_flight_and_stop_still_completes():
    """ERROR / STALL / NO_DEVICE / OVERFLOW: the handler marks the slot cancelled and does not resubmit (perseus-in.c:222-257),
    so 7 transfers stay in flight; perseus_stop_async_input's completion check (perseus-in.c:143-158) still terminates."""
    rq = O.RefQueue(fail_at=4, fail_status=5)            # LIBUSB_TRANSFER_NO_DEVICE
    got = []
    rq.start(6144, lambda b, s, e: got.append(s) or 0)
    assert rq.pending == 8
    rq.pump(20)
    assert rq.pending == 7 and 0 < len(got) < 19
    rq.stop()
    rq.close()


def test_reference_queue_cancel_handshake():
    """perseus_stop_async_input's loop (perseus-sdr.c:714-716) terminates: all 8 transfers report cancelled."""
    rq = O.RefQueue()
    got = []
    rq.start(6144, lambda b, s, e: got.append(s) or 0)
    rq.pump(11)
    assert rq.stop() == 11 * 6144 and len(got) == 11
    rq.close()


def test_reference_queue_feeds_oracle_unpack(coracle):
    """End-to-end on the CPU: reference queue -> reference-style callback (the restated oracle) == unpack of the stream."""
    chunks = []

    def cb(buf, size, extra):
        chunks.append(coracle.unpack(np.ctypeslib.as_array((C.c_ubyte * size).from_address(buf)).copy(), O.MODE_I32))
        return 0

    rq = O.RefQueue(seed=99)
    rq.start(6144, cb)
    rq.pump(24)
    rq.close()
    want = coracle.unpack(coracle.synth_random(24 * 6144, 99), O.MODE_I32)
    assert np.array_equal(np.concatenate(chunks), want)


# ------------------------------------------------------------------ property test: arbitrary fault schedules

from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(1, 150), drop=st.integers(0, 11), swap=st.integers(0, 11), timeout=st.integers(0, 11),
       fail_at=st.integers(0, 60), status=st.sampled_from([1, 4, 5, 6]), size=st.sampled_from([(6144, 512), (12288, 512), (510, 510), (510 * 7, 510)]))
def test_virtual_receiver_equals_reference_queue_for_any_fault_schedule(pg, n, drop, swap, timeout, fail_at, status, size):
    """Hypothesis: any combination of short / swapped / timed-out transfers and one failing transfer, any length of run:
    perseus_vrx_* makes exactly the callbacks the reference's completion handler (perseus-in.c:187-264) makes."""
    buffersize, ep = size
    faults = dict(drop_every=drop, swap_every=swap, timeout_every=timeout, fail_at=fail_at, fail_status=status if fail_at else 0)
    ref_calls, calls = [], []
    rq = O.RefQueue(seed=7, **faults)
    rq.start(buffersize, lambda buf, sz, extra: ref_calls.append((buf, sz, bytes((C.c_ubyte * sz).from_address(buf)))) or 0)
    pumped = rq.pump(n)
    received = rq.bytes_received
    rq.stop()
    rq.close()
    v = pg.VirtualReceiver(sample_rate=2_000_000, ep_max_packet=ep, seed=7, **faults)
    st_ = v.run(buffersize, lambda buf, sz, extra: calls.append((buf, sz, bytes((C.c_ubyte * sz).from_address(buf)))) or 0, None, n)
    v.close()
    assert pumped == n
    assert len(calls) == len(ref_calls) == st_["delivered"]
    if calls:
        assert [(a - calls[0][0], s, b) for a, s, b in calls] == [(a - ref_calls[0][0], s, b) for a, s, b in ref_calls]
    assert st_["bytes_received"] == received
    assert st_["delivered"] + st_["dropped_short"] + st_["dropped_sequence"] + st_["timed_out"] + st_["retired"] == n
