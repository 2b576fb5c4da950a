"""TEST INFRASTRUCTURE ONLY — Python face of the CPU oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``libperseus-sdr_b200/``) never does.

Three independent statements of the reference's I/Q unpack
(``/root/reference/examples/perseustest.c:432-460`` int32, ``:466-502`` float;
duplicate int32 callback ``examples/simple.c:33-61``):

* ``Ref``     – the reference's own callbacks compiled verbatim (``oracle/_ref``,
                built by ``oracle/Makefile`` from ``/root/reference`` where it lies);
* ``COracle`` – our C restatement (``oracle/perseus_oracle.c``), also the timed CPU port;
* ``np_*``    – a numpy restatement, used to cross-check the other two.

Parity pinning: the reference has no golden vectors for this path; the restatements
are pinned against ``Ref`` over all 2**24 codes (tests/test_oracle.py) and against
``tests/golden/`` fixtures generated from ``Ref`` (tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_ORACLE = HERE / "liboracle.so"
LIB_REF = HERE / "_ref" / "libperseus_ref.so"
LIB_REFLIB = HERE / "_ref" / "libperseus_sdr_ref.so"
LIB_REFQUEUE = LIB_REFLIB   # the queue code is part of the whole reference library

MODE_I32, MODE_F32, MODE_F32_POW2 = 0, 1, 2
#: (float)(INT_MAX - 256), perseustest.c:496 — exactly representable in binary32.
REF_DIVISOR = np.float32(2147483392.0)
SYNTH_SEED = 0x5045525345555300  # "PERSEUS\0", SURVEY.md §8(d)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/ when /root/reference is mounted); make decides staleness."""
    # stdout -> stderr: bench.py must print exactly one JSON line on stdout
    subprocess.run(["make", "-s", "-C", str(HERE)] + (["-B"] if force else []), check=True, stdout=sys.stderr)


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(-1)


class COracle:
    """ctypes binding of oracle/perseus_oracle.c."""

    def __init__(self) -> None:
        build()
        L = C.CDLL(str(LIB_ORACLE))
        vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
        L.perseus_oracle_unpack.restype = sz
        L.perseus_oracle_unpack.argtypes = [C.c_int, vp, sz, vp]
        L.perseus_oracle_unpack_mt.restype = sz
        L.perseus_oracle_unpack_mt.argtypes = [C.c_int, vp, sz, vp, C.c_int]
        L.perseus_oracle_fnv1a64.restype = u64
        L.perseus_oracle_fnv1a64.argtypes = [vp, sz, u64]
        L.perseus_oracle_checksum32.restype = u64
        L.perseus_oracle_checksum32.argtypes = [vp, sz, u64]
        L.perseus_oracle_synth_random.restype = None
        L.perseus_oracle_synth_random.argtypes = [vp, sz, u64, u64]
        L.perseus_oracle_synth_ramp.restype = None
        L.perseus_oracle_synth_ramp.argtypes = [vp, sz, u64]
        self.L = L

    def unpack(self, buf, mode: int = MODE_I32, nthreads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
        b = _u8(buf)
        ns = b.size // 6
        dt = np.int32 if mode == MODE_I32 else np.float32
        if out is None:
            out = np.empty((ns, 2), dtype=dt)
        assert out.nbytes >= ns * 8 and out.flags.c_contiguous
        n = self.L.perseus_oracle_unpack_mt(mode, b.ctypes.data, b.size, out.ctypes.data, nthreads)
        assert n == ns, (n, ns)
        return out

    def unpack_raw(self, mode: int, in_ptr: int, nbytes: int, out_ptr: int, nthreads: int) -> int:
        return self.L.perseus_oracle_unpack_mt(mode, in_ptr, nbytes, out_ptr, nthreads)

    def fnv1a64(self, data, h: int = 0) -> int:
        b = _u8(data)
        return int(self.L.perseus_oracle_fnv1a64(b.ctypes.data, b.size, h))

    def checksum32(self, words, first_index: int = 0) -> int:
        w = np.ascontiguousarray(words).view(np.uint32).reshape(-1)
        return int(self.L.perseus_oracle_checksum32(w.ctypes.data, w.size, first_index))

    def synth_random(self, nbytes: int, seed: int = SYNTH_SEED, byte_offset: int = 0) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        self.L.perseus_oracle_synth_random(out.ctypes.data, nbytes, seed, byte_offset)
        return out

    def synth_ramp(self, nsamples: int, first_sample: int = 0) -> np.ndarray:
        out = np.empty(nsamples * 6, dtype=np.uint8)
        self.L.perseus_oracle_synth_ramp(out.ctypes.data, nsamples, first_sample)
        return out


class Ref:
    """The reference's own callbacks (verbatim), when oracle/_ref was built."""

    @staticmethod
    def available() -> bool:
        if not LIB_REF.exists() and Path("/root/reference/examples/perseustest.c").exists():
            try:
                build()
            except Exception:
                return False
        return LIB_REF.exists()

    def __init__(self) -> None:
        if not self.available():
            raise FileNotFoundError(f"{LIB_REF} not built (needs /root/reference at build time)")
        L = C.CDLL(str(LIB_REF))
        vp, sz = C.c_void_p, C.c_size_t
        for f in (L.perseus_ref_unpack_i32, L.perseus_ref_unpack_f32):
            f.restype = sz
            f.argtypes = [vp, sz, sz, vp, sz]
        L.perseus_ref_unpack_mt.restype = sz
        L.perseus_ref_unpack_mt.argtypes = [C.c_int, vp, sz, sz, vp, sz, C.c_int]
        L.perseus_ref_build_info.restype = C.c_char_p
        self.L = L

    def unpack(self, buf, mode: int = MODE_I32, chunk: int = 6144) -> np.ndarray:
        """Feed `buf` to the reference callback `chunk` bytes per call (one USB transfer
        per call, perseus-in.c:206-207) and return what it wrote to its FILE*."""
        assert mode in (MODE_I32, MODE_F32)
        b = _u8(buf)
        # samples are counted per call: a short last call drops its own remainder
        full, rem = divmod(b.size, chunk)
        ns = full * (chunk // 6) + rem // 6
        out = np.empty((ns, 2), dtype=np.int32 if mode == MODE_I32 else np.float32)
        fn = self.L.perseus_ref_unpack_i32 if mode == MODE_I32 else self.L.perseus_ref_unpack_f32
        n = fn(b.ctypes.data, b.size, chunk, out.ctypes.data, out.nbytes)
        assert n == out.nbytes, (n, out.nbytes)
        return out

    def unpack_mt_raw(self, want_float: bool, in_ptr: int, nbytes: int, chunk: int, out_ptr: int, out_cap: int,
                      nthreads: int) -> int:
        n = self.L.perseus_ref_unpack_mt(int(want_float), in_ptr, nbytes, chunk, out_ptr, out_cap, nthreads)
        if n == C.c_size_t(-1).value:
            raise RuntimeError("reference callback sink overflow")
        return n

    def build_info(self) -> str:
        return self.L.perseus_ref_build_info().decode()


class RefQueue:
    """The reference's own transfer queue (/root/reference/perseus-in.c, unmodified) over the fake libusb device
    of oracle/fakeusb.c.  start() == what perseus_start_async_input does at perseus-sdr.c:683; pump(n) completes
    n transfers (each runs the reference's input_queue_callback); stop() == perseus-sdr.c:708-726."""

    @staticmethod
    def available() -> bool:
        return LIB_REFQUEUE.exists()

    def __init__(self, seed: int = SYNTH_SEED, drop_every: int = 0, swap_every: int = 0, timeout_every: int = 0,
                 fail_at: int = 0, fail_status: int = 0) -> None:
        if not self.available():
            raise FileNotFoundError(f"{LIB_REFQUEUE} not built (needs /root/reference at build time)")
        L = C.CDLL(str(LIB_REFQUEUE))
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.fakeusb_open.restype, L.fakeusb_open.argtypes = vp, [u64, u32, u32]
        L.fakeusb_set_faults.restype, L.fakeusb_set_faults.argtypes = None, [vp, u32, u32, u32]
        L.fakeusb_pending.restype, L.fakeusb_pending.argtypes = u64, [vp]
        L.refq_completed.restype, L.refq_completed.argtypes = C.c_int, [vp]
        L.fakeusb_close.restype, L.fakeusb_close.argtypes = None, [vp]
        L.fakeusb_pump.restype, L.fakeusb_pump.argtypes = u64, [vp, u64]
        L.refq_start.restype, L.refq_start.argtypes = vp, [vp, C.c_int, vp, vp]
        L.refq_stop.restype, L.refq_stop.argtypes = u64, [vp, vp]
        L.refq_bytes_received.restype, L.refq_bytes_received.argtypes = u64, [vp]
        L.refq_ring.restype, L.refq_ring.argtypes = vp, [vp]
        L.refq_idx_expected.restype, L.refq_idx_expected.argtypes = C.c_int, [vp]
        self.L = L
        self.dev = L.fakeusb_open(seed, drop_every, swap_every)
        L.fakeusb_set_faults(self.dev, timeout_every, fail_at, fail_status)
        self.q = None
        self._keep = None

    @property
    def pending(self) -> int:
        """Transfers currently submitted to the device (8 minus the retired slots)."""
        return int(self.L.fakeusb_pending(self.dev))

    def start(self, buffersize: int, callback, extra=None) -> None:
        """callback: a C function pointer (int/c_void_p) or a Python callable (buf, size, extra) -> int."""
        if callable(callback) and not isinstance(callback, (int, C.c_void_p)):
            self._keep = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)(callback)
            callback = C.cast(self._keep, C.c_void_p)
        self.q = self.L.refq_start(self.dev, buffersize, callback, extra)
        if not self.q:
            raise RuntimeError("perseus_input_queue_create failed")

    def pump(self, n: int) -> int:
        return int(self.L.fakeusb_pump(self.dev, n))

    @property
    def bytes_received(self) -> int:
        return int(self.L.refq_bytes_received(self.q))

    @property
    def ring(self) -> int:
        return int(self.L.refq_ring(self.q))

    def stop(self) -> int:
        n = int(self.L.refq_stop(self.dev, self.q))
        self.q = None
        return n

    def close(self) -> None:
        if self.q:
            self.stop()
        if self.dev:
            self.L.fakeusb_close(self.dev)
            self.dev = None


class FakeUsbConfig(C.Structure):
    """oracle/fakeusb.h fakeusb_config"""
    _fields_ = [("struct_size", C.c_uint32), ("blank_eeprom", C.c_uint32), ("ep_max_packet", C.c_uint32), ("realtime", C.c_uint32),
                ("seed", C.c_uint64), ("limit", C.c_uint64), ("drop_every", C.c_uint32), ("swap_every", C.c_uint32),
                ("timeout_every", C.c_uint32), ("fail_at", C.c_uint32), ("fail_status", C.c_uint32), ("preserie", C.c_uint32),
                ("serial", C.c_uint32), ("reserved", C.c_uint32)]


class FakeUsbState(C.Structure):
    """oracle/fakeusb.h fakeusb_state"""
    _fields_ = [("inits", C.c_uint64), ("exits", C.c_uint64), ("opens", C.c_uint64), ("closes", C.c_uint64), ("device_refs", C.c_int64),
                ("claimed", C.c_int32), ("firmware_loaded", C.c_int32),
                ("cpu_resets", C.c_uint64), ("fw_records", C.c_uint64), ("fw_bytes", C.c_uint64), ("fw_hash", C.c_uint64),
                ("fpga_resets", C.c_uint64), ("fpga_bytes", C.c_uint64), ("fpga_hash", C.c_uint64),
                ("fpga_rate", C.c_int32), ("fifo_enabled", C.c_int32), ("sio_freg", C.c_uint32), ("sio_ctl", C.c_uint8), ("porte", C.c_uint8),
                ("pad", C.c_uint8 * 2), ("sio_writes", C.c_uint64), ("porte_writes", C.c_uint64), ("eeprom_reads", C.c_uint64),
                ("shutdowns", C.c_uint64), ("commands", C.c_uint64),
                ("submits", C.c_uint64), ("stream_pos", C.c_uint64), ("completed_ok", C.c_uint64), ("cancelled", C.c_uint64),
                ("timed_out", C.c_uint64), ("failed", C.c_uint64), ("events_calls", C.c_uint64),
                ("events_thread_policy", C.c_int32), ("events_thread_priority", C.c_int32), ("events_thread_is_fifo", C.c_int32),
                ("pad2", C.c_int32)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("pad")}


class EepromProdId(C.Structure):
    """perseus-sdr.h:66-74 eeprom_prodid (packed)"""
    _pack_ = 1
    _fields_ = [("sn", C.c_uint16), ("prodcode", C.c_uint16), ("hwrel", C.c_uint8), ("hwver", C.c_uint8), ("signature", C.c_uint8 * 6)]


# perseus-sdr.h:317-343
PERSEUS_ERR = {"NOERROR": 0, "INVALIDDEV": -1, "NULLDESCR": -2, "ALREADYOPEN": -3, "LIBUSBERR": -4, "DEVNOTOPEN": -5, "FNNOTAVAIL": -9,
               "DEVNOTFOUND": -10, "EEPROMREAD": -11, "IOERROR": -13, "FWNOTLOADED": -16, "FPGACFGERROR": -17, "FPGANOTCFGD": -18,
               "ASYNCSTARTED": -19, "NOMEM": -20, "CANTCREAT": -21, "ERRPARAM": -22, "BUFFERSIZE": -24, "ATTERROR": -25, "SNNOTAVAILABLE": -26}


class RefLib:
    """The reference's WHOLE library (/root/reference/perseus-sdr.c, perseusfx2.c, perseus-in.c, perseuserr.c, unmodified)
    over the synthetic receiver of oracle/fakeusb.c.  Methods are the reference's public API (perseus-sdr.h), called
    through ctypes exactly as a C application would; `plug()` / `state()` talk to the fake device.

    The library keeps global state (descriptor list, one poll thread: perseus-sdr.c:45-61), so there is one instance
    per process and sessions must not overlap: use `with reflib.session(...) as descr:`."""

    _inst = None

    @staticmethod
    def available() -> bool:
        return LIB_REFLIB.exists()

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
            cls._inst._load()
        return cls._inst

    def _load(self) -> None:
        if not self.available():
            raise FileNotFoundError(f"{LIB_REFLIB} not built (needs /root/reference at build time)")
        L = C.CDLL(str(LIB_REFLIB))
        vp, ci, u32 = C.c_void_p, C.c_int, C.c_uint32
        sig = {
            "perseus_init": (ci, []), "perseus_exit": (ci, []), "perseus_open": (vp, [ci]), "perseus_close": (ci, [vp]),
            "perseus_set_debug": (None, [ci]), "perseus_errorstr": (C.c_char_p, []),
            "perseus_firmware_download": (ci, [vp, C.c_char_p]), "perseus_get_product_id": (ci, [vp, C.POINTER(EepromProdId)]),
            "perseus_is_preserie": (ci, [vp, C.POINTER(ci)]),
            "perseus_set_sampling_rate": (ci, [vp, ci]), "perseus_set_sampling_rate_n": (ci, [vp, C.c_uint]),
            "perseus_get_sampling_rates": (ci, [vp, C.POINTER(ci), C.c_uint]),
            "perseus_set_attenuator_n": (ci, [vp, ci]), "perseus_set_attenuator_in_db": (ci, [vp, ci]),
            "perseus_set_adc": (ci, [vp, ci, ci]), "perseus_set_ddc_center_freq": (ci, [vp, C.c_double, ci]),
            "perseus_start_async_input": (ci, [vp, u32, vp, vp]), "perseus_stop_async_input": (ci, [vp]),
            "fakeusb_plug": (ci, [C.POINTER(FakeUsbConfig)]), "fakeusb_unplug": (None, []), "fakeusb_get_state": (None, [C.POINTER(FakeUsbState)]),
            "reflib_descr_firmware_downloaded": (ci, [vp]), "reflib_descr_fpga_configured": (ci, [vp]), "reflib_descr_is_preserie": (ci, [vp]),
            "reflib_descr_ring": (vp, [vp]), "reflib_descr_bytes_received": (C.c_uint64, [vp]), "reflib_descr_queue_active": (ci, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        self.L = L
        self._keep = None

    # -- fake device
    def plug(self, **kw) -> None:
        cfg = FakeUsbConfig()
        cfg.struct_size = C.sizeof(FakeUsbConfig)
        cfg.seed = SYNTH_SEED
        for k, v in kw.items():
            setattr(cfg, k, v)
        if self.L.fakeusb_plug(C.byref(cfg)) != 0:
            raise RuntimeError("fake receiver still has pending transfers")

    def unplug(self) -> None:
        self.L.fakeusb_unplug()

    def state(self) -> dict:
        st = FakeUsbState()
        self.L.fakeusb_get_state(C.byref(st))
        return st.asdict()

    def errorstr(self) -> str:
        return self.L.perseus_errorstr().decode(errors="replace")

    def wait_stream_pos(self, n: int, timeout_s: float = 20.0) -> dict:
        """Polls the fake device until it has completed `n` stream transfers."""
        import time
        t_end = time.monotonic() + timeout_s
        while time.monotonic() < t_end:
            st = self.state()
            if st["stream_pos"] >= n:
                return st
            time.sleep(0.001)
        raise TimeoutError(f"fake receiver stuck at {self.state()['stream_pos']} of {n} transfers")

    def callback_pointer(self, callback):
        """C pointer for `callback`: an int / c_void_p is passed through, a Python callable is wrapped (and kept alive)."""
        if callable(callback) and not isinstance(callback, (int, C.c_void_p)):
            self._keep = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)(callback)
            return C.cast(self._keep, C.c_void_p)
        return callback

    class _Session:
        def __init__(self, lib, bring_up, rate, plug):
            self.lib, self.bring_up, self.rate, self.plug_kw = lib, bring_up, rate, plug
            self.descr = None

        def __enter__(self):
            lib, L = self.lib, self.lib.L
            lib.plug(**self.plug_kw)
            n = L.perseus_init()                                        # perseustest.c:188
            assert n == 1, (n, lib.errorstr())
            self.descr = L.perseus_open(0)                              # :205
            assert self.descr, lib.errorstr()
            if self.bring_up:
                assert L.perseus_firmware_download(self.descr, None) == 0, lib.errorstr()      # :213
                if self.rate:
                    assert L.perseus_set_sampling_rate(self.descr, self.rate) == 0, lib.errorstr()   # :266
            return self.descr

        def __exit__(self, *exc):
            L = self.lib.L
            if self.descr and L.reflib_descr_queue_active(self.descr):
                L.perseus_stop_async_input(self.descr)
            L.perseus_exit()                                            # joins the poll thread, closes the device
            self.lib.unplug()
            return False

    def session(self, bring_up: bool = True, rate: int = 95000, **plug):
        """plug -> perseus_init -> perseus_open(0) [-> firmware_download -> set_sampling_rate]; perseus_exit on the way out."""
        return RefLib._Session(self, bring_up, rate, plug)


# --------------------------------------------------------------------------- numpy

def np_unpack_i32(buf) -> np.ndarray:
    """perseustest.c:447-457 in numpy: bytes (b0,b1,b2) -> b0<<8 | b1<<16 | b2<<24."""
    b = _u8(buf)
    ns = b.size // 6
    f = b[: ns * 6].reshape(ns, 2, 3).astype(np.uint32)
    u = (f[..., 0] << np.uint32(8)) | (f[..., 1] << np.uint32(16)) | (f[..., 2] << np.uint32(24))
    return u.view(np.int32)


def np_unpack_f32(buf) -> np.ndarray:
    """perseustest.c:496-497: (float)x / (INT_MAX-256), IEEE round-to-nearest."""
    return np_unpack_i32(buf).astype(np.float32) / REF_DIVISOR


def np_unpack_f32_pow2(buf) -> np.ndarray:
    return np_unpack_i32(buf).astype(np.float32) * np.float32(2.0 ** -31)


def np_unpack(buf, mode: int) -> np.ndarray:
    return (np_unpack_i32, np_unpack_f32, np_unpack_f32_pow2)[mode](buf)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def np_synth_random(nbytes: int, seed: int = SYNTH_SEED, byte_offset: int = 0) -> np.ndarray:
    """64-bit LE word w of the stream = splitmix64(seed + w)."""
    w0 = byte_offset >> 3
    w1 = (byte_offset + nbytes + 7) >> 3
    with np.errstate(over="ignore"):
        idx = np.arange(w0, w1, dtype=np.uint64) + np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
    words = _splitmix64(idx).astype("<u8")
    s = byte_offset - (w0 << 3)
    return words.view(np.uint8)[s: s + nbytes].copy()


def np_synth_ramp(nsamples: int, first_sample: int = 0) -> np.ndarray:
    """SURVEY.md §0 exhaustive pattern: I = v & 0xFFFFFF, Q = (uint32(v*2654435761)) >> 8."""
    v = (np.arange(first_sample, first_sample + nsamples, dtype=np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    with np.errstate(over="ignore"):
        q = (v * np.uint32(2654435761)) >> np.uint32(8)
    i = v & np.uint32(0xFFFFFF)
    out = np.empty((nsamples, 6), dtype=np.uint8)
    for k in range(3):
        out[:, k] = (i >> np.uint32(8 * k)).astype(np.uint8)
        out[:, 3 + k] = (q >> np.uint32(8 * k)).astype(np.uint8)
    return out.reshape(-1)


def np_checksum32(words, first_index: int = 0) -> int:
    w = np.ascontiguousarray(words).view(np.uint32).reshape(-1).astype(np.uint64)
    with np.errstate(over="ignore"):
        idx = np.arange(w.size, dtype=np.uint64) + np.uint64(first_index)
        m = _splitmix64(idx) | np.uint64(1)
        return int((m * (w + np.uint64(1))).sum(dtype=np.uint64))


def fnv1a64_py(data: bytes, h: int = 0) -> int:
    """Pure-Python FNV-1a-64 (small inputs only)."""
    if h == 0:
        h = 0xCBF29CE484222325
    for byte in data:
        h = ((h ^ byte) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1
