"""ctypes face of lib/libperseus_gpu.so — the B200-native Perseus I/Q unpack (include/perseus-gpu.h).

This module is plumbing only: it loads the C-ABI library, declares the prototypes and wraps
handles in small classes so tests and bench.py read like C callers of the reference
(``perseus_start_async_input(descr, 6144, callback, extra)``, perseus-sdr.h:247-248).
It contains no sample arithmetic and NO fallback: if the CUDA library is missing or no
sm_100 device is present, calls raise ``PerseusGpuError``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO = PKG_DIR.parent
LIB_PATH = Path(os.environ.get("PERSEUS_GPU_LIB", PKG_DIR / "lib" / "libperseus_gpu.so"))   # override: A/B builds in tools/
HEADER = REPO / "include" / "perseus-gpu.h"

# include/perseus-gpu.h
OUT_INT32, OUT_FLOAT, OUT_FLOAT_POW2, ASYNC, CHECKSUM = 0x1, 0x2, 0x4, 0x100, 0x200
VARIANT_AUTO, VARIANT_STREAM, VARIANT_DIRECT = 0, 1, 2
SYNTH_RANDOM, SYNTH_RAMP = 0, 1
SYNTH_SEED = 0x5045525345555300
ERR = {"NOERROR": 0, "NULLHANDLE": -2, "IOERROR": -13, "ASYNCSTARTED": -19, "NOMEM": -20, "ERRPARAM": -22,
       "BUFFERSIZE": -24, "CUDAERR": -40, "NODEVICE": -41, "BADARCH": -42, "MISMATCH": -43}
VRX_QUEUE_SIZE, VRX_MAX_BUFFER = 8, 16320
OPT_NO_WATCHDOG = 0x1
PCIE_H2D, PCIE_D2H, PCIE_DUPLEX = 0, 1, 2
# enum libusb_transfer_status values, PERSEUS_VRX_STATUS_*
STATUS = {"COMPLETED": 0, "ERROR": 1, "TIMED_OUT": 2, "CANCELLED": 3, "STALL": 4, "NO_DEVICE": 5, "OVERFLOW": 6}

INPUT_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)   # perseus-sdr.h:81


class PerseusGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"perseus-gpu error {code}: {msg}")
        self.code = code
        self.msg = msg


class Tuning(C.Structure):
    _fields_ = [("variant", C.c_int), ("tile_bytes", C.c_int), ("stages", C.c_int), ("ctas_per_sm", C.c_int),
                ("store_mode", C.c_int), ("reserved", C.c_int * 3)]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("stream_flags", C.c_uint32), ("nslabs", C.c_uint32),
                ("slab_bytes", C.c_uint64), ("nstreams", C.c_uint32), ("max_latency_us", C.c_uint32), ("chunk_bytes", C.c_uint64),
                ("tuning", Tuning), ("options", C.c_uint32), ("stage_slots", C.c_uint32), ("direct_bytes", C.c_uint32),
                ("copy_threads", C.c_uint32), ("eager_gap_us", C.c_uint32), ("reserved2", C.c_uint32 * 3)]


class Seg(C.Structure):
    _fields_ = [("in_", C.c_void_p), ("nbytes", C.c_size_t), ("out_i32", C.c_void_p), ("out_f32", C.c_void_p)]


class Block(C.Structure):
    _fields_ = [("first_sample", C.c_uint64), ("nsamples", C.c_uint64), ("dev_i32", C.c_void_p), ("dev_f32", C.c_void_p),
                ("stream", C.c_void_p)]


SINK = C.CFUNCTYPE(None, C.POINTER(Block), C.c_void_p)


class HostBlock(C.Structure):
    _fields_ = [("first_sample", C.c_uint64), ("nsamples", C.c_uint64), ("i32", C.c_void_p), ("f32", C.c_void_p)]


HOST_SINK = C.CFUNCTYPE(None, C.POINTER(HostBlock), C.c_void_p)
DIRECT_NEVER = 0xFFFFFFFF
EAGER_NEVER = 0xFFFFFFFF              # Config.eager_gap_us: slabs go out full or over age only
COPY_BY_RUNTIME = 0xFFFFFFFF          # Config.copy_threads: leave pageable buffers to the CUDA runtime


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kernel_launches", "samples", "bytes_in", "h2d_bytes", "d2h_bytes", "callbacks",
                                          "slabs", "stalls", "dropped_callbacks", "dropped_bytes", "watchdog_submits",
                                          "host_blocks")]

    def asdict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class VrxConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("sample_rate", C.c_int32), ("ep_max_packet", C.c_int32), ("pattern", C.c_int32),
                ("seed", C.c_uint64), ("realtime", C.c_int32), ("drop_every", C.c_uint32), ("swap_every", C.c_uint32),
                ("replay", C.c_uint32), ("timeout_every", C.c_uint32), ("fail_at", C.c_uint32), ("fail_status", C.c_uint32),
                ("reserved", C.c_uint32 * 1)]


class VrxStats(C.Structure):
    _fields_ = [("bytes_received", C.c_uint64), ("delivered", C.c_uint64), ("dropped_short", C.c_uint64),
                ("dropped_sequence", C.c_uint64), ("elapsed_s", C.c_double), ("ksamples_per_s", C.c_double),
                ("timed_out", C.c_uint64), ("retired", C.c_uint64)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def declared_symbols() -> list[str]:
    """Every function include/perseus-gpu.h declares (used by the export test)."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(perseus_(?:gpu|vrx|synth)_\w+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    """Loads the C-ABI library (no CUDA call is made by loading it)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PerseusGpuError(ERR["NODEVICE"], f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`"
                              " (there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    vp, sz, u64, i64, u32, ci = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int64, C.c_uint32, C.c_int
    P = C.POINTER
    proto = {
        "perseus_gpu_open": (ci, [P(vp), P(Config)]),
        "perseus_gpu_close": (ci, [vp]),
        "perseus_gpu_unpack": (i64, [vp, vp, sz, vp, vp, C.c_uint]),
        "perseus_gpu_sync": (ci, [vp]),
        "perseus_gpu_get_checksums": (ci, [vp, P(u64), P(u64)]),
        "perseus_gpu_unpack_batch": (i64, [vp, P(Seg), ci, C.c_uint]),
        "perseus_gpu_plan_create": (ci, [vp, P(Seg), ci, C.c_uint, P(vp)]),
        "perseus_gpu_plan_run": (i64, [vp, vp, C.c_uint]),
        "perseus_gpu_plan_destroy": (ci, [vp, vp]),
        "perseus_gpu_input_callback": (ci, [vp, ci, vp]),
        "perseus_gpu_set_sink": (ci, [vp, SINK, vp]),
        "perseus_gpu_set_host_sink": (ci, [vp, HOST_SINK, vp]),
        "perseus_gpu_stream_to_file": (ci, [vp, C.c_char_p]),
        "perseus_gpu_flush": (ci, [vp]),
        "perseus_gpu_poll": (ci, [vp]),
        "perseus_gpu_prepare": (ci, [vp]),
        "perseus_gpu_get_stats": (ci, [vp, P(Stats)]),
        "perseus_gpu_autotune": (ci, [vp, P(C.c_double), P(C.c_double)]),
        "perseus_gpu_get_geometry": (ci, [vp, C.c_uint, P(ci), P(ci), P(ci)]),
        "perseus_gpu_set_tuning": (ci, [vp, P(Tuning)]),
        "perseus_gpu_get_tuning": (ci, [vp, P(Tuning)]),
        "perseus_gpu_errorstr": (C.c_char_p, []),
        "perseus_gpu_version": (C.c_char_p, []),
        "perseus_gpu_device_count": (ci, []),
        "perseus_gpu_device_info": (ci, [ci, C.c_char_p, sz, P(ci), P(ci), P(ci), P(u64)]),
        "perseus_gpu_dev_alloc": (vp, [vp, sz]),
        "perseus_gpu_dev_free": (ci, [vp, vp]),
        "perseus_gpu_host_alloc": (vp, [vp, sz]),
        "perseus_gpu_host_free": (ci, [vp, vp]),
        "perseus_gpu_memcpy": (ci, [vp, vp, vp, sz]),
        "perseus_gpu_memset": (ci, [vp, vp, ci, sz]),
        "perseus_gpu_get_stream": (vp, [vp, ci]),
        "perseus_gpu_event_record": (ci, [vp, ci]),
        "perseus_gpu_event_elapsed_ms": (ci, [vp, ci, ci, P(C.c_float)]),
        "perseus_gpu_generate": (ci, [vp, vp, sz, ci, u64, u64]),
        "perseus_synth_fill": (ci, [vp, sz, ci, u64, u64]),
        "perseus_gpu_checksum": (ci, [vp, vp, sz, u64, P(u64)]),
        "perseus_gpu_verify": (ci, [vp, vp, sz, vp, vp, C.c_uint, P(u64), P(u64)]),
        "perseus_gpu_shard_range": (ci, [u64, ci, ci, P(u64), P(u64)]),
        "perseus_gpu_shard_range_weighted": (ci, [u64, ci, P(C.c_double), ci, P(u64), P(u64)]),
        "perseus_gpu_probe_hbm": (ci, [vp, ci, sz, ci, P(C.c_double)]),
        "perseus_gpu_probe_pcie": (ci, [vp, ci, sz, sz, ci, P(C.c_double), P(C.c_double)]),
        "perseus_vrx_open": (ci, [P(vp), P(VrxConfig)]),
        "perseus_vrx_close": (ci, [vp]),
        "perseus_vrx_get_sampling_rates": (ci, [P(ci), C.c_uint]),
        "perseus_vrx_nearest_rate": (ci, [ci]),
        "perseus_vrx_bitstream_name": (C.c_char_p, [ci]),
        "perseus_vrx_get_sampling_rate": (ci, [vp]),
        "perseus_vrx_start_async_input": (ci, [vp, u32, vp, vp]),
        "perseus_vrx_stop_async_input": (ci, [vp]),
        "perseus_vrx_run": (ci, [vp, u32, vp, vp, u64]),
        "perseus_vrx_get_stats": (ci, [vp, P(VrxStats)]),
    }
    for name, (res, args) in proto.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def errorstr() -> str:
    return lib().perseus_gpu_errorstr().decode(errors="replace")


def check(rc: int) -> int:
    if rc < 0:
        raise PerseusGpuError(int(rc), errorstr())
    return int(rc)


def shard_range(total_buffers: int, nshards: int, shard: int) -> tuple[int, int]:
    a, n = C.c_uint64(), C.c_uint64()
    check(lib().perseus_gpu_shard_range(total_buffers, nshards, shard, C.byref(a), C.byref(n)))
    return a.value, n.value


def shard_range_weighted(total_buffers: int, weights, shard: int) -> tuple[int, int]:
    """Shards proportional to `weights` (e.g. each GPU's measured host-link rate)."""
    a, n = C.c_uint64(), C.c_uint64()
    w = (C.c_double * len(weights))(*weights)
    check(lib().perseus_gpu_shard_range_weighted(total_buffers, len(weights), w, shard, C.byref(a), C.byref(n)))
    return a.value, n.value


def synth_fill(nbytes: int, pattern: int = SYNTH_RANDOM, seed: int = SYNTH_SEED, byte_offset: int = 0):
    """Host-side synthetic wire bytes (the generator the virtual receiver uses)."""
    import numpy as np
    out = np.empty(nbytes, dtype=np.uint8)
    check(lib().perseus_synth_fill(out.ctypes.data, nbytes, pattern, seed, byte_offset))
    return out


def sampling_rates() -> list[int]:
    buf = (C.c_int * 16)()
    check(lib().perseus_vrx_get_sampling_rates(buf, 16))
    return [v for v in buf if v]


def bitstream_name(rate: int) -> str | None:
    n = lib().perseus_vrx_bitstream_name(rate)
    return n.decode() if n else None


def nearest_rate(requested: int) -> int:
    return lib().perseus_vrx_nearest_rate(requested)


class PerseusGpu:
    """perseus_gpu handle.  Pointers are plain integers (device or host addresses)."""

    def __init__(self, device: int = 0, stream_flags: int = 0, nslabs: int = 0, slab_bytes: int = 0, nstreams: int = 0,
                 chunk_bytes: int = 0, max_latency_us: int = 0, options: int = 0, stage_slots: int = 0, direct_bytes: int = 0,
                 copy_threads: int = 0, eager_gap_us: int = 0, **tuning):
        L = lib()
        cfg = Config()
        cfg.struct_size = C.sizeof(Config)
        cfg.device, cfg.stream_flags, cfg.nslabs, cfg.slab_bytes = device, stream_flags, nslabs, slab_bytes
        cfg.nstreams, cfg.chunk_bytes, cfg.max_latency_us = nstreams, chunk_bytes, max_latency_us
        cfg.options, cfg.stage_slots, cfg.direct_bytes, cfg.copy_threads = options, stage_slots, direct_bytes, copy_threads
        cfg.eager_gap_us = eager_gap_us
        for k, v in tuning.items():
            setattr(cfg.tuning, k, v)
        self.h = C.c_void_p()
        self.L = L
        check(L.perseus_gpu_open(C.byref(self.h), C.byref(cfg)))
        self.device = device

    # -- life cycle
    def close(self) -> None:
        if self.h:
            try:
                check(self.L.perseus_gpu_close(self.h))     # flushes: sinks may still be called in here, and may use this object
            finally:
                self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the path
    def unpack(self, buf: int, nbytes: int, out_i32: int | None, out_f32: int | None, flags: int = 0) -> int:
        return check(self.L.perseus_gpu_unpack(self.h, buf, nbytes, out_i32, out_f32, flags))

    def sync(self) -> None:
        check(self.L.perseus_gpu_sync(self.h))

    def get_checksums(self) -> tuple[int, int]:
        """(int32 checksum, float checksum) accumulated by the last unpack(..., CHECKSUM); waits for it."""
        a, b = C.c_uint64(), C.c_uint64()
        check(self.L.perseus_gpu_get_checksums(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def unpack_batch(self, segs: list[tuple[int, int, int | None, int | None]], flags: int = 0) -> int:
        arr = (Seg * max(1, len(segs)))(*[Seg(a, n, oi, of) for a, n, oi, of in segs])
        return check(self.L.perseus_gpu_unpack_batch(self.h, arr, len(segs), flags))

    def plan_create(self, segs, flags: int = 0) -> C.c_void_p:
        arr = (Seg * max(1, len(segs)))(*[Seg(a, n, oi, of) for a, n, oi, of in segs])
        p = C.c_void_p()
        check(self.L.perseus_gpu_plan_create(self.h, arr, len(segs), flags, C.byref(p)))
        return p

    def plan_run(self, plan, flags: int = 0) -> int:
        return check(self.L.perseus_gpu_plan_run(self.h, plan, flags))

    def plan_destroy(self, plan) -> None:
        check(self.L.perseus_gpu_plan_destroy(self.h, plan))

    # -- streaming hand-off
    @property
    def callback(self):
        """(function pointer, extra) to hand to perseus_start_async_input / perseus_vrx_start_async_input."""
        return C.cast(self.L.perseus_gpu_input_callback, C.c_void_p), self.h

    def input_callback(self, buf: int, nbytes: int) -> int:
        return self.L.perseus_gpu_input_callback(buf, nbytes, self.h)

    def set_sink(self, fn) -> None:
        new = SINK(fn) if fn else C.cast(None, SINK)
        check(self.L.perseus_gpu_set_sink(self.h, new, None))
        self._sink = new

    def set_host_sink(self, fn) -> None:
        """fn(HostBlock pointer, extra) is called on a CUDA runtime thread, once per slab, in stream order."""
        new = HOST_SINK(fn) if fn else C.cast(None, HOST_SINK)
        check(self.L.perseus_gpu_set_host_sink(self.h, new, None))   # flushes: the previous sink may still be called in here ...
        self._host_sink = new                                         # ... so its trampoline is let go of only now

    def stream_to_file(self, path: str | None) -> None:
        check(self.L.perseus_gpu_stream_to_file(self.h, path.encode() if path else None))

    def flush(self) -> None:
        check(self.L.perseus_gpu_flush(self.h))

    def prepare(self) -> None:
        """Allocates and warms up the streaming path now instead of inside the first callback."""
        check(self.L.perseus_gpu_prepare(self.h))

    def poll(self) -> int:
        """Submits the partial slab if it is over age; returns how many slabs that submitted (0 or 1)."""
        return check(self.L.perseus_gpu_poll(self.h))

    def stats(self) -> dict:
        s = Stats()
        check(self.L.perseus_gpu_get_stats(self.h, C.byref(s)))
        return s.asdict()

    def set_tuning(self, **kw) -> None:
        t = Tuning()
        for k, v in kw.items():
            setattr(t, k, v)
        check(self.L.perseus_gpu_set_tuning(self.h, C.byref(t)))

    def autotune(self) -> tuple[float, float]:
        """Measures candidate geometries on this device; returns (GB/s one format, GB/s fused) of the winners."""
        a, b = C.c_double(), C.c_double()
        check(self.L.perseus_gpu_autotune(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_geometry(self, flags: int) -> dict:
        t, s, c = C.c_int(), C.c_int(), C.c_int()
        check(self.L.perseus_gpu_get_geometry(self.h, flags, C.byref(t), C.byref(s), C.byref(c)))
        return {"tile_bytes": t.value, "stages": s.value, "ctas_per_sm": c.value}

    def get_tuning(self) -> dict:
        t = Tuning()
        check(self.L.perseus_gpu_get_tuning(self.h, C.byref(t)))
        return {k: getattr(t, k) for k in ("variant", "tile_bytes", "stages", "ctas_per_sm", "store_mode")}

    # -- plumbing
    def dev_alloc(self, nbytes: int) -> int:
        p = self.L.perseus_gpu_dev_alloc(self.h, nbytes)
        if not p:
            raise PerseusGpuError(ERR["NOMEM"], errorstr())
        return p

    def dev_free(self, p: int) -> None:
        check(self.L.perseus_gpu_dev_free(self.h, p))

    def host_alloc(self, nbytes: int) -> int:
        p = self.L.perseus_gpu_host_alloc(self.h, nbytes)
        if not p:
            raise PerseusGpuError(ERR["NOMEM"], errorstr())
        return p

    def host_free(self, p: int) -> None:
        check(self.L.perseus_gpu_host_free(self.h, p))

    def memcpy(self, dst: int, src: int, nbytes: int) -> None:
        check(self.L.perseus_gpu_memcpy(self.h, dst, src, nbytes))

    def memset(self, dev: int, byte: int, nbytes: int) -> None:
        check(self.L.perseus_gpu_memset(self.h, dev, byte, nbytes))

    def stream(self, idx: int = 0) -> int:
        return self.L.perseus_gpu_get_stream(self.h, idx)

    def event_record(self, slot: int) -> None:
        check(self.L.perseus_gpu_event_record(self.h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        check(self.L.perseus_gpu_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def generate(self, dev: int, nbytes: int, pattern: int = SYNTH_RANDOM, seed: int = SYNTH_SEED, byte_offset: int = 0) -> None:
        check(self.L.perseus_gpu_generate(self.h, dev, nbytes, pattern, seed, byte_offset))

    def checksum(self, dev_words: int, nwords: int, first_index: int = 0) -> int:
        s = C.c_uint64()
        check(self.L.perseus_gpu_checksum(self.h, dev_words, nwords, first_index, C.byref(s)))
        return s.value

    def probe_hbm(self, kind: int, nbytes: int = 1 << 30, reps: int = 10) -> float:
        """GB/s this device sustains on a pure read (0), pure write (1) or copy (2) stream."""
        g = C.c_double()
        check(self.L.perseus_gpu_probe_hbm(self.h, kind, nbytes, reps, C.byref(g)))
        return g.value

    def probe_pcie(self, kind: int, nbytes: int = 256 << 20, d2h_nbytes: int = 0, reps: int = 3) -> tuple[float, float]:
        """(H2D GB/s, D2H GB/s) of plain pinned copies: one direction alone, or both at once (PCIE_DUPLEX)."""
        a, b = C.c_double(), C.c_double()
        check(self.L.perseus_gpu_probe_pcie(self.h, kind, nbytes, d2h_nbytes, reps, C.byref(a), C.byref(b)))
        return a.value, b.value

    def verify(self, dev_in: int, nbytes: int, dev_i32: int | None, dev_f32: int | None, flags: int = 0) -> tuple[int, int]:
        """Returns (mismatching words, first bad word); raises only on errors other than MISMATCH."""
        n, first = C.c_uint64(), C.c_uint64()
        rc = self.L.perseus_gpu_verify(self.h, dev_in, nbytes, dev_i32, dev_f32, flags, C.byref(n), C.byref(first))
        if rc < 0 and rc != ERR["MISMATCH"]:
            check(rc)
        return n.value, first.value

    # -- numpy conveniences for tests (host arrays <-> device memory through the C ABI)
    def to_device(self, arr) -> int:
        import numpy as np
        a = np.ascontiguousarray(arr)
        d = self.dev_alloc(max(a.nbytes, 1))
        self.memcpy(d, a.ctypes.data, a.nbytes)
        return d

    def to_host(self, dev: int, nbytes: int, dtype):
        import numpy as np
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        self.memcpy(out.ctypes.data, dev, nbytes)
        return out


class VirtualReceiver:
    """perseus_vrx handle: the reference's transfer delivery (8-slot ring, in-order callbacks) over synthetic data."""

    def __init__(self, sample_rate: int = 95000, ep_max_packet: int = 0, pattern: int = SYNTH_RANDOM, seed: int = 0,
                 realtime: bool = False, drop_every: int = 0, swap_every: int = 0, replay: bool = False,
                 timeout_every: int = 0, fail_at: int = 0, fail_status: int = 0):
        cfg = VrxConfig()
        cfg.struct_size = C.sizeof(VrxConfig)
        cfg.sample_rate, cfg.ep_max_packet, cfg.pattern, cfg.seed = sample_rate, ep_max_packet, pattern, seed
        cfg.realtime, cfg.drop_every, cfg.swap_every, cfg.replay = int(realtime), drop_every, swap_every, int(replay)
        cfg.timeout_every, cfg.fail_at, cfg.fail_status = timeout_every, fail_at, fail_status
        self.v = C.c_void_p()
        self.L = lib()
        check(self.L.perseus_vrx_open(C.byref(self.v), C.byref(cfg)))
        self._keep = None

    def close(self) -> None:
        if self.v:
            v, self.v = self.v, C.c_void_p()
            check(self.L.perseus_vrx_close(v))

    @property
    def sample_rate(self) -> int:
        return check(self.L.perseus_vrx_get_sampling_rate(self.v))

    def _cb(self, callback, extra):
        if callable(callback) and not isinstance(callback, (int, C.c_void_p)):
            self._keep = INPUT_CALLBACK(callback)
            return C.cast(self._keep, C.c_void_p), extra
        return callback, extra

    def start_async_input(self, buffersize: int, callback, extra=None) -> None:
        cb, ex = self._cb(callback, extra)
        check(self.L.perseus_vrx_start_async_input(self.v, buffersize, cb, ex))

    def stop_async_input(self) -> dict:
        check(self.L.perseus_vrx_stop_async_input(self.v))
        return self.stats()

    def run(self, buffersize: int, callback, extra, ntransfers: int) -> dict:
        cb, ex = self._cb(callback, extra)
        check(self.L.perseus_vrx_run(self.v, buffersize, cb, ex, ntransfers))
        return self.stats()

    def stats(self) -> dict:
        s = VrxStats()
        check(self.L.perseus_vrx_get_stats(self.v, C.byref(s)))
        return s.asdict()
