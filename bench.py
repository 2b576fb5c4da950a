#!/usr/bin/env python
"""bench.py — the headline benchmark of the B200-native Perseus I/Q unpack.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA path through the C ABI)
    python bench.py --impl reference [...]                          the reference's own CPU callbacks (oracle/_ref)
    torchrun --nproc-per-node N bench.py --gpus N ...               N>1: one rank per GPU, no data-path collective

Workload (BASELINE.json configs[1]): the 2 MS/s bitstream layout — per GPU 174 762
transfers x 6144 B = 1 073 737 728 wire bytes (178 956 288 complex samples) of synthetic data, unpacked to
int32 AND float.  One "step" = one pass of the hot path over that batch.  At N GPUs every rank owns the
contiguous transfer range perseus_gpu_shard_range() gives it of an N x 174 762-transfer recording (weak scaling).

Prints ONE JSON line (rank 0).  `value` = complex Msamples/s, whole job, inputs resident in HBM;
`e2e` = same metric through perseus_gpu_unpack() with PINNED HOST input (H2D inside the timed region,
outputs left on the device as north_star's end-to-end mode specifies, plus a D2H read of the step's
result: the on-device checksums of both outputs); `e2e_roundtrip` additionally copies both outputs back.
`workloads` carries BASELINE.json's other GPU configurations in the same run, same schema per record:
cfg3 (1024 mixed-rate receivers, one launch, per GPU) and cfg4 (the 64 GiB recording sharded over the N
GPUs: strong scaling, whole-recording checksum identical for every N).
oracle/ is used here only by the cpu_baseline leg and by --impl reference (timed CPU baselines) — never on the measured
path; the untimed correctness gate before timing uses the library's own independent verify kernel.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BUF = 6144                      # bytes per transfer (1024 complex samples), perseustest.c:100-102 / perseus-sdr.c:671
CFG2_BUFFERS = 174_762          # 1 GiB of wire data, SURVEY.md §8d
BYTES_PER_SAMPLE_FUSED = 6 + 8 + 8     # algorithmic HBM traffic per complex sample, int32+float in one pass
BYTES_PER_SAMPLE_SINGLE = 6 + 8
METRIC, UNIT = "complex_msamples_per_s_unpacked", "Msamples/s"
PCIE_GEN5_X16_GBS = 63.0        # 32 GT/s * 16 lanes * 128/130 / 8
CFG4_BUFFERS = 11_184_810       # 64 GiB of 6144-byte transfers, SURVEY.md §8d cfg4


def workload_config(world: int, nbuf: int = CFG2_BUFFERS) -> dict:
    """The `config` object of the JSON line.  Both arms (ours and --impl reference) print exactly this, so the driver can
    see they ran the same workload; everything specific to one arm (kernel geometry, NUMA binding) lives under `setup`."""
    nbytes = nbuf * BUF
    return {"workload": f"cfg2: perseus2m24v21 (2 MS/s) layout, {nbuf} transfers x {BUF} B = {nbytes} wire bytes per GPU "
                        f"({nbytes // 6} complex samples), each unpacked to int32 AND float; {world} GPU(s): one recording of {world}x that "
                        f"size, sharded by contiguous transfer range (HBM-resident legs: rank r owns [r*{nbuf},(r+1)*{nbuf}); host-fed leg "
                        f"at N>1: ranges proportional to each GPU's host-link rate)",
            "transfers_per_gpu": nbuf, "transfer_bytes": BUF, "outputs": "int32+float", "n_gpus": world,
            "generator": "splitmix64(seed + word index), seed 0x5045525345555300"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line.  Libraries (NCCL prints its version banner on stdout, make, ...) write to
# file descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON line goes to
# a private duplicate of the original stdout.
_JSON_FD = None


def claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ----------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        self.samples, self.reasons, self.ok = [], set(), False
        self.period, self._stop = period_s, threading.Event()
        self.max_mhz = None
        try:
            import pynvml as N
            N.nvmlInit()
            self.N = N
            self.dev = N.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(self.dev, N.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            log(f"[bench] NVML unavailable ({e}); clocks will be null")
        self.t = threading.Thread(target=self._run, daemon=True)

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost", 0x100: "display_clock_setting"}

    def _run(self):
        N = self.N
        while not self._stop.is_set():
            try:
                self.samples.append(N.nvmlDeviceGetClockInfo(self.dev, N.NVML_CLOCK_SM))
                try:
                    r = N.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = N.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in self.NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(index: int):
    """Best effort: run (and therefore first-touch pinned memory) on the CPUs next to this GPU.  Tries NVML's ideal
    CPU set first, then sysfs.  Returns a short description for the JSON line, or None when the host hides its topology."""
    try:
        import pynvml as N
        N.nvmlInit()
        dev = N.nvmlDeviceGetHandleByIndex(index)
    except Exception:
        return None
    allowed = os.sched_getaffinity(0)
    try:
        words = (max(allowed) // 64) + 1
        mask = N.nvmlDeviceGetCpuAffinity(dev, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"nvml cpu affinity: {len(cpus)} of {len(allowed)} cpus"
    except Exception:
        pass
    try:
        bus = N.nvmlDeviceGetPciInfo(dev).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"sysfs numa node {node}: {len(cpus)} cpus"
    except Exception:
        return None
    return None


# ----------------------------------------------------------------------------------------- CPU legs (oracle = checker / baseline)

def cpu_baseline_leg(budget_s: float = 12.0):
    """Times the CPU on a bounded sample of the same workload: the restated port (1 thread and all threads)
    and, when oracle/_ref travelled here, the reference's own callbacks verbatim (1 thread)."""
    import numpy as np
    from oracle import oracle as O
    co = O.COracle()
    threads = O.host_threads()
    nbuf = 16_384                                     # 96 MiB of wire, 16.8 M samples, >> CPU caches
    wire = co.synth_random(nbuf * BUF, seed=O.SYNTH_SEED)
    ns = wire.size // 6
    out = np.empty((ns, 2), np.int32)

    def rate(fn, min_reps=2, share=0.3):
        fn()                                          # warm (page faults)
        t_end = time.perf_counter() + budget_s * share
        reps, t0 = 0, time.perf_counter()
        while reps < min_reps or time.perf_counter() < t_end:
            fn(); reps += 1
            if reps >= 200:
                break
        dt = time.perf_counter() - t0
        return ns * reps / dt / 1e6

    def both(nthreads):
        co.unpack_raw(O.MODE_I32, wire.ctypes.data, wire.size, out.ctypes.data, nthreads)
        co.unpack_raw(O.MODE_F32, wire.ctypes.data, wire.size, out.ctypes.data, nthreads)

    port_all = rate(lambda: both(threads))
    port_one = rate(lambda: both(1), share=0.2)
    res = {"value": round(port_all, 1), "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{nbuf} transfers x {BUF} B ({wire.size / 2**20:.0f} MiB wire, {ns} samples), int32 pass + float pass per rep, "
                     f"oracle/perseus_oracle.c memory->memory, {threads} threads (static split)",
           "port_1_thread": round(port_one, 1)}
    if O.Ref.available():
        ref = O.Ref()
        small = 2048 * BUF                            # the verbatim callbacks fwrite per sample: ~60 MS/s
        sink = np.empty(small // 6 * 2, np.int32)

        def verbatim():
            ref.unpack_mt_raw(False, wire.ctypes.data, small, BUF, sink.ctypes.data, sink.nbytes, 1)
            ref.unpack_mt_raw(True, wire.ctypes.data, small, BUF, sink.ctypes.data, sink.nbytes, 1)
        t_end, reps, t0 = time.perf_counter() + budget_s * 0.2, 0, time.perf_counter()
        while reps < 2 or time.perf_counter() < t_end:
            verbatim(); reps += 1
        res["reference_verbatim_1_thread"] = round(small // 6 * reps / (time.perf_counter() - t0) / 1e6, 1)
    return res


def reference_arm(args):
    """--impl reference: the reference's OWN CPU implementation of the path (user_data_callback_c_u + _c_f,
    compiled verbatim into oracle/_ref) on all host threads, each thread an independent receiver stream fed
    6144-byte transfers; falls back to the restated port when oracle/_ref is not present.
    Same workload as our arm: one step = the whole N x cfg2 recording through the int32 callback and the float
    callback.  The N shards are unpacked one after the other from the same 1 GiB of synthetic wire bytes (the
    content does not change the work; it keeps host memory at 2.5 GB for every N)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import oracle as O
    co = O.COracle()
    threads = O.host_threads()
    use_ref = O.Ref.available()
    world = max(1, args.gpus)
    nbuf = args.buffers
    wire = co.synth_random(nbuf * BUF, seed=O.SYNTH_SEED)
    ns = nbuf * 1024
    out = np.empty(ns * 2, np.int32)
    if use_ref:
        ref = O.Ref()

        def shard():
            ref.unpack_mt_raw(False, wire.ctypes.data, wire.size, BUF, out.ctypes.data, out.nbytes, threads)
            ref.unpack_mt_raw(True, wire.ctypes.data, wire.size, BUF, out.ctypes.data, out.nbytes, threads)
        kind, what = "reference", "oracle/_ref: examples/perseustest.c callbacks verbatim (fwrite per sample into a memory FILE*)"
    else:
        def shard():
            co.unpack_raw(O.MODE_I32, wire.ctypes.data, wire.size, out.ctypes.data, threads)
            co.unpack_raw(O.MODE_F32, wire.ctypes.data, wire.size, out.ctypes.data, threads)
        kind, what = "port", "oracle/perseus_oracle.c restatement (oracle/_ref not present)"

    def step():
        for _ in range(world):
            shard()
    # check the arm against the restated oracle on the first transfers (untimed)
    chk = co.unpack(wire[: 8 * BUF], O.MODE_F32).view(np.uint32).reshape(-1)
    shard()
    assert np.array_equal(out[: chk.size].view(np.uint32), chk), "reference arm output differs from the oracle"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = ns * world * args.steps / dt / 1e6
    sample = (f"the full workload: {world} x {nbuf} transfers x {BUF} B per step ({world * wire.size / 2**20:.0f} MiB wire), int32 callback + "
              f"float callback on {threads} host threads, {what}")
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
            "config": workload_config(world, nbuf),
            "cpu_baseline": {"value": round(val, 2), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": round(val, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------- our arm

def spread_device(local: int, local_world: int, visible: int):
    """Which GPU a rank uses.  The shards are independent and host-fed, so when fewer ranks than GPUs run, they are spread
    over the box (stride visible // ranks: 4 ranks on an 8-GPU box use GPUs 0, 2, 4, 6) instead of packed onto GPUs 0..N-1:
    on these hosts GPUs 0-3 share one path to host memory (115 GB/s of pinned reads for the four of them, measured by
    tools/h2d_matrix.py, profiles/r2_h2d_matrix_8gpu.md), GPUs 4-7 another.  Kernel-only numbers do not depend on it."""
    if os.environ.get("PERSEUS_BENCH_PACKED") or local_world >= visible or visible % local_world:
        return local, "packed: rank r -> GPU r"
    stride = visible // local_world
    return local * stride, f"spread: rank r -> GPU r*{stride} of {visible}"


class Ctx:
    """Per-rank plumbing shared by the legs: handle, distributed helpers, timing."""

    def __init__(self, args):
        import importlib
        import torch
        import torch.distributed as dist
        import __graft_entry__ as G
        self.args, self.torch, self.dist = args, torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            log(f"[bench] WORLD_SIZE={self.world} but --gpus {args.gpus}; using {self.world}")
        self.device, self.device_map = spread_device(self.local, int(os.environ.get("LOCAL_WORLD_SIZE", self.world)), torch.cuda.device_count())
        self.numa = bind_to_gpu_numa_node(self.device)
        torch.cuda.set_device(self.device)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.device))
        if self.local == 0:
            G.ensure_built()
        self.barrier()
        self.pg = G.load_package()
        self.sharding = importlib.import_module("libperseus_sdr_b200.sharding")
        self.allmax = self.sharding.allreduce_max
        self.peak, self.peak_src = measured_peak()
        self.h = self.pg.PerseusGpu(device=self.device, chunk_bytes=args.chunk_mib << 20, stage_slots=args.slots)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def timed(self, fn, steps, warmup, sampler=None, h=None):
        """W untimed steps, then exactly K steps between barrier+sync, CUDA events on the launching stream, max over ranks."""
        h = h or self.h
        for _ in range(warmup):
            fn()
        h.sync(); self.torch.cuda.synchronize(); self.barrier()
        with (sampler if sampler is not None else _Null()):
            h.event_record(0)
            for _ in range(steps):
                fn()
            h.event_record(1)
            h.sync(); self.torch.cuda.synchronize()
            ms = h.event_elapsed_ms(0, 1)
        self.barrier()
        return self.allmax(ms) / steps

    def free_bytes(self):
        return self.torch.cuda.mem_get_info(self.device)[0]

    def close(self):
        self.h.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def mixed_rate_buffers(nrx=1024, window_s=1.024):
    rates = [48000, 96000, 192000, 500000, 1000000, 2000000]
    return [max(1, round(rates[r % 6] * window_s / 1024)) for r in range(nrx)]


def workload_cfg3(cx, steps, warmup):
    """BASELINE config 3: 1024 virtual receivers at six rates, every receiver its own input/output segment and seed,
    unpacked to int32 AND float by ONE launch per step (per GPU; at N GPUs every rank runs its own 1024)."""
    pg, h = cx.pg, cx.h
    nbufs = mixed_rate_buffers()
    total_in = sum(nbufs) * BUF
    ns = total_in // 6
    pad = 64                                                            # room to push one receiver's outputs off 16-byte alignment
    d_in, d_i, d_f = h.dev_alloc(total_in), h.dev_alloc(ns * 8 + pad), h.dev_alloc(ns * 8 + pad)
    segs, off = [], 0
    for r, b in enumerate(nbufs):
        h.generate(d_in + off, b * BUF, pg.SYNTH_RANDOM, pg.SYNTH_SEED + cx.rank * 1024 + r, 0)
        segs.append((d_in + off, b * BUF, d_i + off // 6 * 8, d_f + off // 6 * 8))
        off += b * BUF
    flags = pg.OUT_INT32 | pg.OUT_FLOAT
    plan = h.plan_create(segs, flags)
    h.plan_run(plan)
    bad, _ = h.verify(d_in, total_in, d_i, d_f, flags)
    assert bad == 0, bad
    sampler = ClockSampler(cx.device)
    l0 = h.stats()["kernel_launches"]
    ms = cx.timed(lambda: h.plan_run(plan, pg.ASYNC), steps, warmup, sampler)
    launches = h.stats()["kernel_launches"] - l0 - warmup
    h.plan_destroy(plan)
    # the same batch with ONE receiver's outputs off 16-byte alignment (the last 2 MS/s receiver, its segment shortened by
    # one transfer and its outputs moved up inside their own range): by 8 bytes -- the natural alignment of an {I,Q}
    # array -- and by 4 bytes.  Same kernel, same single launch, 128-bit stores after a pre-roll of 2 / 1 output words
    odd = max(r for r in range(len(nbufs)) if nbufs[r] == max(nbufs))
    misaligned = {}
    for shift in (8, 4):
        segs2 = list(segs)
        a, n, oi, of = segs[odd]
        segs2[odd] = (a, n - BUF, oi + shift, of + shift)
        plan2 = h.plan_create(segs2, flags)
        l0 = h.stats()["kernel_launches"]
        ms_odd = cx.timed(lambda: h.plan_run(plan2, pg.ASYNC), steps, warmup)
        per_step = (h.stats()["kernel_launches"] - l0) // (steps + warmup)
        h.plan_destroy(plan2)
        ns2 = ns - BUF // 6
        misaligned[f"outputs_plus_{shift}_bytes"] = {"ms_per_step": round(ms_odd, 4), "launches_per_step": int(per_step),
                                                     "time_per_sample_vs_all_aligned": round((ms_odd / ns2) / (ms / ns), 4)}
    for p_ in (d_in, d_i, d_f):
        h.dev_free(p_)
    achieved = BYTES_PER_SAMPLE_FUSED * ns / (ms * 1e-3) / 1e9
    return {"metric": METRIC, "value": round(ns * cx.world / (ms * 1e-3) / 1e6, 1), "unit": UNIT, "n_gpus": cx.world, "steps": steps,
            "warmup": warmup, "ms_per_step": round(ms, 4), "scaling": "weak", "dtype": "int32+f32",
            "config": {"workload": f"cfg3: 1024 virtual receivers at 48k/96k/192k/500k/1M/2M S/s, 1.024 s window = {sum(nbufs)} transfers x {BUF} B "
                                   f"({total_in} wire bytes) per GPU, own seed per receiver, unpacked to int32 AND float in ONE launch"},
            "roofline": {"bound": "hbm", "kernel": "unpack24_stream_kernel<I32|F32, batched>", "achieved": round(achieved, 1), "peak": cx.peak,
                         "unit": "GB/s", "frac": round(achieved / cx.peak, 4), "bytes_per_sample": BYTES_PER_SAMPLE_FUSED},
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "one_misaligned_receiver": dict(misaligned, what=f"receiver {odd} ({nbufs[odd]} transfers) writes to outputs that are not 16-byte "
                                            "aligned; the other 1023 receivers are unaffected either way")}


def workload_cfg4(cx, steps, warmup):
    """BASELINE config 4: the 64 GiB recording sharded by contiguous transfer range over the N GPUs (strong scaling),
    generated on the device, unpacked to float, kernel-only.  Returns None when a shard does not fit this GPU."""
    pg, h = cx.pg, cx.h
    first, count = pg.shard_range(CFG4_BUFFERS, cx.world, cx.rank)
    nbytes = count * BUF
    ns = nbytes // 6
    need = nbytes + ns * 8 + (1 << 30)
    fits = cx.sharding.allreduce_max(0.0 if cx.free_bytes() >= need else 1.0) == 0.0
    if not fits:
        return {"skipped": f"a shard needs {need / 2**30:.1f} GiB of device memory, {cx.free_bytes() / 2**30:.1f} GiB free"}
    d_in, d_f = h.dev_alloc(nbytes), h.dev_alloc(ns * 8)
    h.generate(d_in, nbytes, pg.SYNTH_RANDOM, pg.SYNTH_SEED, first * BUF)
    flags = pg.OUT_FLOAT
    h.unpack(d_in, nbytes, None, d_f, flags)
    bad, _ = h.verify(d_in, nbytes, None, d_f, flags)
    assert bad == 0, bad
    # position-weighted checksum of the whole recording's float output: shard sums add up (mod 2^64), so it must be
    # the same number whatever N is
    checksum = cx.sharding.allreduce_sum_u64(h.checksum(d_f, ns * 2, first_index=first * 2048))
    sampler = ClockSampler(cx.device)
    l0 = h.stats()["kernel_launches"]
    ms = cx.timed(lambda: h.unpack(d_in, nbytes, None, d_f, flags | pg.ASYNC), steps, warmup, sampler)
    launches = h.stats()["kernel_launches"] - l0 - warmup
    h.dev_free(d_in); h.dev_free(d_f)
    achieved = BYTES_PER_SAMPLE_SINGLE * ns / (ms * 1e-3) / 1e9          # per GPU; the slowest rank's time
    return {"metric": METRIC, "value": round(CFG4_BUFFERS * 1024 / (ms * 1e-3) / 1e6, 1), "unit": UNIT, "n_gpus": cx.world, "steps": steps,
            "warmup": warmup, "ms_per_step": round(ms, 4), "scaling": "strong", "dtype": "f32",
            "config": {"workload": f"cfg4: 64 GiB recording ({CFG4_BUFFERS} transfers x {BUF} B) sharded by contiguous transfer range over "
                                   f"{cx.world} GPU(s), rank 0 holds {count} transfers; generated on the device; float output; kernel-only"},
            "roofline": {"bound": "hbm", "kernel": "unpack24_stream_kernel<F32>", "achieved": round(achieved, 1), "peak": cx.peak, "unit": "GB/s",
                         "frac": round(achieved / cx.peak, 4), "bytes_per_sample": BYTES_PER_SAMPLE_SINGLE, "what": "per GPU, max-over-ranks time"},
            "recording_float_checksum": f"{checksum:016x}", "gpu_launches": int(launches), "clocks": sampler.summary()}


def ours(args):
    import numpy as np
    cx = Ctx(args)
    pg, h, rank, world, local = cx.pg, cx.h, cx.rank, cx.world, cx.device
    barrier, allmax, timed = cx.barrier, cx.allmax, cx.timed
    sharding = cx.sharding
    if args.tile or args.stages or args.ctas or args.variant or args.store:
        h.set_tuning(variant=args.variant, tile_bytes=args.tile, stages=args.stages, ctas_per_sm=args.ctas, store_mode=args.store)
    if args.autotune:
        h.autotune()
    tuning = h.get_tuning()
    tuning["geometry_fused"] = h.get_geometry(pg.OUT_INT32 | pg.OUT_FLOAT)
    tuning["geometry_single"] = h.get_geometry(pg.OUT_FLOAT)
    tuning["autotuned"] = bool(args.autotune)
    tuning["auto_rule"] = "0 = per format: int32+float fused -> tile 12288 B x 3 stages; one format -> tile 6144 B x 8 stages; 1 CTA/SM"

    nbuf = args.buffers
    first, count = sharding.rank_shard(pg, nbuf * world, world, rank)   # this rank's transfers of the N x cfg2 recording
    assert count == nbuf
    nbytes = nbuf * BUF
    ns = nbytes // 6
    d_in = h.dev_alloc(nbytes)
    d_i32 = h.dev_alloc(ns * 8)
    d_f32 = h.dev_alloc(ns * 8)
    h.generate(d_in, nbytes, pg.SYNTH_RANDOM, pg.SYNTH_SEED, first * BUF)
    FUSED = pg.OUT_INT32 | pg.OUT_FLOAT

    # -- untimed correctness gate: the whole output against the library's independent per-sample kernel (byte loads, IEEE division)
    h.unpack(d_in, nbytes, d_i32, d_f32, FUSED)
    bad, where = h.verify(d_in, nbytes, d_i32, d_f32, FUSED)
    if bad:
        raise SystemExit(f"[bench] rank {rank}: {bad} output words differ from the per-sample recomputation (first {where})")
    recording_checksum = sharding.allreduce_sum_u64(h.checksum(d_f32, ns * 2, first_index=first * 2048))   # shard checksums add up

    # -- headline: HBM-resident, fused int32+float pass
    sampler = ClockSampler(local)
    l0 = h.stats()["kernel_launches"]
    ms_step = timed(lambda: h.unpack(d_in, nbytes, d_i32, d_f32, FUSED | pg.ASYNC), args.steps, args.warmup, sampler)
    launches = h.stats()["kernel_launches"] - l0 - args.warmup
    clocks = sampler.summary()
    total_samples = ns * world
    value = total_samples / (ms_step * 1e-3) / 1e6
    peak, peak_src = cx.peak, cx.peak_src
    achieved = BYTES_PER_SAMPLE_FUSED * ns / (ms_step * 1e-3) / 1e9      # per GPU: each rank runs the same launch
    traffic = None
    try:
        traffic = json.loads((ROOT / "profiles" / "roofline_traffic.json").read_text()).get("fused_traffic_bytes_per_launch")
    except Exception:
        pass

    # -- context for the roofline: what this very device sustains on one-directional streams, same run
    probe = None
    if not args.no_probe:
        rd, wr, cp = (h.probe_hbm(k, 1 << 30, 10) for k in (0, 1, 2))
        t_bound = 6 * ns / (rd * 1e9) + 16 * ns / (wr * 1e9)             # reads at the pure-read rate, then writes at the pure-write rate
        probe = {"read_gbs": round(rd, 1), "write_gbs": round(wr, 1), "copy_gbs": round(cp, 1),
                 "fused_mix_bound_gbs": round(BYTES_PER_SAMPLE_FUSED * ns / t_bound / 1e9, 1),
                 "what": "plain 16-byte streaming kernels on 1 GiB (perseus_gpu_probe_hbm), best of 4 grid sizes; mix bound = "
                         "6 B/sample at the read rate + 16 B/sample at the write rate, no overlap credit"}

    # -- supporting numbers: single-format kernels (14 B/sample)
    extra = {}
    short = max(3, min(args.steps, 50))
    for name, flags, oi, of in (() if args.no_single else (("int32_only", pg.OUT_INT32, d_i32, None), ("float_only", pg.OUT_FLOAT, None, d_f32))):
        ms = timed(lambda: h.unpack(d_in, nbytes, oi, of, flags | pg.ASYNC), short, 3)
        gbs = BYTES_PER_SAMPLE_SINGLE * ns / (ms * 1e-3) / 1e9
        extra[name] = {"msamples_per_s": round(total_samples / (ms * 1e-3) / 1e6, 1), "ms_per_step": round(ms, 4),
                       "hbm_gbs_per_gpu": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4)}

    e2e = e2e_bal = e2e_rt = e2e_pageable = None
    if not args.no_e2e:
        # -- end to end: pinned host wire -> H2D -> unpack (outputs stay on device) -> D2H of the step's result
        e2e_steps = max(3, min(args.steps, args.e2e_steps))
        wc_rt = None
        if args.pinned_wc:   # experiment: write-combined pinned memory (DMA reads need not snoop the CPU caches)
            wc_rt = C.CDLL("libcudart.so.12")
            pp = C.c_void_p()
            assert wc_rt.cudaHostAlloc(C.byref(pp), C.c_size_t(nbytes), C.c_uint(0x04 | 0x01)) == 0   # WriteCombined | Portable
            pin = pp.value
        else:
            pin = h.host_alloc(nbytes)
        h.memcpy(pin, d_in, nbytes)                                           # the synthetic recording, now in pinned host memory

        def pcie(kind, up=256 << 20, down=0):
            """In-run PCIe roofline of THIS GPU alone: plain pinned copies by the library's probe (ranks take turns)."""
            res = None
            for r in range(world):
                if r == rank:
                    res = h.probe_pcie(kind, up, down, 3)
                barrier()
            return res

        own_rate = {}

        def all_ranks_at_once(dst, src, n, key=None):
            """The same plain copy with every rank copying at the same time, between buffers that already exist: GB/s of the slowest
            rank (this rank's own rate is kept in own_rate[key])."""
            if world == 1:
                return None
            times, mine = [], []
            for _ in range(2):
                barrier()
                h.event_record(2); h.memcpy(dst, src, n); h.event_record(3)
                mine.append(h.event_elapsed_ms(2, 3))
                times.append(allmax(mine[-1]))
            if key:
                own_rate[key] = n / (min(mine) * 1e-3) / 1e9
            return n / (min(times) * 1e-3) / 1e9

        h2d_gbs, _ = pcie(pg.PCIE_H2D)
        h2d_conc = all_ranks_at_once(d_in, pin, nbytes, "h2d")
        sums = []

        def e2e_step():
            # chunked H2D + unpack kernels + per-chunk checksum kernels, overlapped on the handle's streams ...
            h.unpack(pin, nbytes, d_i32, d_f32, FUSED | pg.ASYNC | pg.CHECKSUM)
            sums.append(h.get_checksums())                                   # ... then wait and read the 16-byte result

        s0 = h.stats()
        ms_e2e = timed(e2e_step, e2e_steps, 2)
        s1 = h.stats()
        assert len(set(sums)) == 1, "end-to-end results changed between steps"
        assert sums[0] == (h.checksum(d_i32, ns * 2), h.checksum(d_f32, ns * 2)), "overlapped checksum differs from the whole-output checksum"
        h2d_per_step = (s1["h2d_bytes"] - s0["h2d_bytes"]) // (e2e_steps + 2)
        e2e_val = total_samples / (ms_e2e * 1e-3) / 1e6
        e2e_gbs = 6 * ns / (ms_e2e * 1e-3) / 1e9
        e2e = {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": 16,
               "ms_per_step": round(ms_e2e, 3), "steps": e2e_steps,
               "what": "perseus_gpu_unpack(pinned host wire -> device int32+float, PERSEUS_GPU_CHECKSUM): H2D in chunks on the copy-in stream, "
                       "unpack and checksum kernels behind them on the launch stream, then the checksums of both outputs read back (outputs stay "
                       "in HBM, as north_star's end-to-end mode specifies)",
               "h2d_gbs_per_gpu": round(e2e_gbs, 2), "pcie_h2d_gbs_measured": round(h2d_gbs, 2),
               "frac_of_measured_pcie": round(e2e_gbs / h2d_gbs, 4),
               "frac_of_gen5_x16_theory": round(e2e_gbs / PCIE_GEN5_X16_GBS, 4),
               "pinned_memory": "write-combined" if args.pinned_wc else "default (cudaHostAlloc portable)"}
        if h2d_conc:
            e2e["pcie_h2d_gbs_all_ranks_at_once"] = round(h2d_conc, 2)
            e2e["frac_of_concurrent_pcie"] = round(e2e_gbs / h2d_conc, 4)

        # -- the same N x cfg2 recording, host-fed, with shards proportional to each GPU's measured host-link rate: on a box whose
        #    GPUs do not all reach host memory equally fast (profiles/r2_h2d_matrix_8gpu.md) equal shards wait for the slowest link
        e2e_bal = None
        if world > 1 and not args.no_balanced:
            rates = sharding.allgather_float(own_rate["h2d"])
            cap_b = (min(nbuf * world, nbuf * 2) + 16) * BUF                # no rank is ever given more than twice an equal share
            d_gen = h.dev_alloc(cap_b)
            pin_b = h.host_alloc(cap_b)
            d_ib, d_fb = h.dev_alloc(cap_b // 6 * 8), h.dev_alloc(cap_b // 6 * 8)
            history = []
            for it in range(args.balance_passes + 1):
                rates = [max(r, 1e-3) for r in rates]
                lim = 2.0 * sum(rates) / world
                rates = [min(r, lim) for r in rates]
                first_b, count_b = pg.shard_range_weighted(nbuf * world, rates, rank)
                nb_b = count_b * BUF
                assert nb_b <= cap_b, (count_b, nbuf)
                ns_b = nb_b // 6
                h.generate(d_gen, nb_b, pg.SYNTH_RANDOM, pg.SYNTH_SEED, first_b * BUF)
                h.memcpy(pin_b, d_gen, nb_b)
                sums_b = []

                def bal_step():
                    h.unpack(pin_b, nb_b, d_ib, d_fb, FUSED | pg.ASYNC | pg.CHECKSUM)
                    sums_b.append(h.get_checksums())

                history.append([round(r, 1) for r in rates])
                if it == args.balance_passes:
                    break
                # a pass of all ranks at once with these shards; each rank's own bytes / own time are the next weights
                bal_step()
                barrier()
                t0 = time.perf_counter()
                bal_step()
                rates = sharding.observed_rates(nb_b / 1e9, time.perf_counter() - t0)
            ms_bal = timed(bal_step, e2e_steps, 2)
            assert len(set(sums_b)) == 1 and sums_b[0] == (h.checksum(d_ib, ns_b * 2), h.checksum(d_fb, ns_b * 2))
            # a different sharding of the SAME recording: the shard checksums must add up to the same whole-recording checksum
            assert sharding.allreduce_sum_u64(h.checksum(d_fb, ns_b * 2, first_index=first_b * 2048)) == recording_checksum
            shares = sharding.allgather_float(float(count_b))
            e2e_bal = {"value": round(total_samples / (ms_bal * 1e-3) / 1e6, 1), "unit": UNIT, "ms_per_step": round(ms_bal, 3),
                       "h2d_bytes_per_step_all_gpus": int(nbytes * world), "d2h_bytes_per_step": 16,
                       "aggregate_h2d_gbs": round(nbytes * world / (ms_bal * 1e-3) / 1e9, 1),
                       "link_gbs_per_rank": history[-1], "weights_history": history, "transfers_per_rank": [int(c) for c in shares],
                       "what": "the same call on the same N x cfg2 recording, sharded with perseus_gpu_shard_range_weighted: weights start from each "
                               "rank's rate in the concurrent plain pinned copy measured just before and are refined over "
                               f"{args.balance_passes} untimed pass(es) from each rank's own bytes / own time (shards that finish together see "
                               "steady-state rates); shard checksums add up to the equal-shard recording checksum"}
            h.host_free(pin_b); h.dev_free(d_gen); h.dev_free(d_ib); h.dev_free(d_fb)

        e2e_rt = None
        if not args.no_roundtrip:
            po_i, po_f = h.host_alloc(ns * 8), h.host_alloc(ns * 8)
            rt_steps = max(2, min(e2e_steps, 5))
            _, d2h_gbs = pcie(pg.PCIE_D2H)
            d2h_conc = all_ranks_at_once(po_f, d_f32, ns * 8)
            # plain copies of the round trip's own traffic, both directions at once: the whole recording up (6 B/sample) while the
            # whole output comes down (16 B/sample fused, 8 B/sample one format), each as ONE cudaMemcpyAsync on its own stream
            dup16_up, dup16_down = pcie(pg.PCIE_DUPLEX, nbytes, ns * 16)
            dup8_up, dup8_down = pcie(pg.PCIE_DUPLEX, nbytes, ns * 8)

            def copy_bound_ms(up_gbs, down_gbs, down_bytes_per_sample):
                return max(6 * ns / (up_gbs * 1e9), down_bytes_per_sample * ns / (down_gbs * 1e9)) * 1e3

            s0 = h.stats()
            ms_rt = timed(lambda: h.unpack(pin, nbytes, po_i, po_f, FUSED), rt_steps, 1)
            s1 = h.stats()
            tail = np.ctypeslib.as_array((C.c_uint32 * 2048).from_address(po_f + ns * 8 - 8192))
            assert np.array_equal(tail, h.to_host(d_f32 + ns * 8 - 8192, 8192, np.uint32)), "round-trip output differs"
            ms_rt1 = timed(lambda: h.unpack(pin, nbytes, None, po_f, pg.OUT_FLOAT), rt_steps, 1)
            d2h_rate = 16 * ns / (ms_rt * 1e-3) / 1e9
            e2e_rt = {"value": round(total_samples / (ms_rt * 1e-3) / 1e6, 1), "unit": UNIT, "ms_per_step": round(ms_rt, 3),
                      "h2d_bytes_per_step": int((s1["h2d_bytes"] - s0["h2d_bytes"]) // (rt_steps + 1)),
                      "d2h_bytes_per_step": int((s1["d2h_bytes"] - s0["d2h_bytes"]) // (rt_steps + 1)),
                      "what": "same call with pinned HOST outputs: both formats copied back (16 B/sample D2H, full duplex with the 6 B/sample H2D); "
                              "copy-in, kernels and copy-out each on their own stream, three staging slots",
                      "d2h_gbs_per_gpu": round(d2h_rate, 2), "pcie_d2h_gbs_measured": round(d2h_gbs, 2),
                      "frac_of_d2h_peak": round(d2h_rate / d2h_gbs, 4),
                      "duplex_plain_copies_gbs": {"h2d": round(dup16_up, 2), "d2h": round(dup16_down, 2),
                                                  "what": "the same bytes as two plain pinned copies queued at once: the recording up, both outputs' worth down (perseus_gpu_probe_pcie)"},
                      "frac_of_duplex_copy_bound": round(copy_bound_ms(dup16_up, dup16_down, 16) / ms_rt, 4),
                      "single_format": {"value": round(total_samples / (ms_rt1 * 1e-3) / 1e6, 1), "ms_per_step": round(ms_rt1, 3),
                                        "d2h_gbs_per_gpu": round(8 * ns / (ms_rt1 * 1e-3) / 1e9, 2),
                                        "frac_of_d2h_peak": round(8 * ns / (ms_rt1 * 1e-3) / 1e9 / d2h_gbs, 4),
                                        "frac_of_duplex_copy_bound": round(copy_bound_ms(dup8_up, dup8_down, 8) / ms_rt1, 4),
                                        "what": "float only: 8 B/sample D2H against 6 B/sample H2D"}}
            if d2h_conc:
                e2e_rt["pcie_d2h_gbs_all_ranks_at_once"] = round(d2h_conc, 2)
                e2e_rt["frac_of_concurrent_d2h"] = round(d2h_rate / d2h_conc, 4)
            h.host_free(po_i); h.host_free(po_f)
        # -- the same call fed from PAGEABLE memory (malloc / numpy / mmap: what an application that never heard of CUDA holds):
        #    the handle's copy pool moves each chunk into pinned bounce buffers while the copy engine works on the previous one
        if world == 1 and not args.no_pageable:
            page = np.empty(nbytes, np.uint8)
            C.memmove(page.ctypes.data, pin, nbytes)
            sums_p = []

            def page_step():
                h.unpack(page.ctypes.data, nbytes, d_i32, d_f32, FUSED | pg.ASYNC | pg.CHECKSUM)
                sums_p.append(h.get_checksums())

            pg_steps = max(3, min(e2e_steps, 5))
            ms_pg = timed(page_step, pg_steps, 2)
            assert set(sums_p) == {sums[0]}, "pageable-fed result differs from the pinned-fed one"
            with pg.PerseusGpu(device=local, copy_threads=pg.COPY_BY_RUNTIME) as h_rt:
                ms_rt_stage = timed(lambda: h_rt.unpack(page.ctypes.data, nbytes, d_i32, d_f32, FUSED), 2, 1)
            e2e_pageable = {"value": round(total_samples / (ms_pg * 1e-3) / 1e6, 1), "unit": UNIT, "ms_per_step": round(ms_pg, 3),
                            "h2d_gbs_per_gpu": round(nbytes / (ms_pg * 1e-3) / 1e9, 2), "frac_of_pinned_e2e": round(ms_e2e / ms_pg, 4),
                            "staged_by_cuda_runtime_msamples_per_s": round(total_samples / (ms_rt_stage * 1e-3) / 1e6, 1),
                            "what": "perseus_gpu_unpack(PAGEABLE host wire -> device int32+float, CHECKSUM): default copy pool "
                                    "(min(8, cores/2) threads incl. the caller) into pinned bounce buffers, overlapped with the copy engine; "
                                    "beside it the same call with copy_threads = 0xFFFFFFFF (the CUDA runtime stages pageable memory itself)"}
            del page
        if wc_rt is not None:
            wc_rt.cudaFreeHost(C.c_void_p(pin))
        else:
            h.host_free(pin)

    # -- the literal drop-in: 6144-byte transfers through perseus_gpu_input_callback (one host thread, like the
    #    reference's poll thread), delivered by the virtual receiver from its 8-slot pageable ring
    e2e_cb = None
    if not args.no_callback:
        hs = pg.PerseusGpu(device=local, stream_flags=FUSED, slab_bytes=16 << 20, nslabs=4, nstreams=2)
        v = pg.VirtualReceiver(sample_rate=2_000_000, replay=True)
        cbp, cbx = hs.callback
        v.run(BUF, cbp, cbx, 4096); hs.flush()                            # warm: slabs allocated, pages touched
        barrier()
        t0 = time.perf_counter()
        st = v.run(BUF, cbp, cbx, nbuf)
        hs.flush()
        dt = allmax(time.perf_counter() - t0)
        hst = hs.stats()
        e2e_cb = {"value": round(ns * world / dt / 1e6, 1), "unit": UNIT, "wire_gbs_per_gpu": round(nbytes / dt / 1e9, 2),
                  "transfers": nbuf, "callbacks": int(st["delivered"]), "slab_stalls": hst["stalls"], "slabs": hst["slabs"],
                  "what": "perseus_vrx_run -> perseus_gpu_input_callback per 6144-byte transfer (memcpy into a pinned 16 MiB slab, "
                          "H2D + fused unpack per slab on 2 streams), one host thread; bounded by that thread's memcpy"}
        v.close(); hs.close()
        # hand-off latency of ONE transfer: callback + flush (slab of 1024 samples: H2D 6 KiB, one kernel, sync)
        hl = pg.PerseusGpu(device=local, stream_flags=FUSED, slab_bytes=BUF * 8, nslabs=2, nstreams=1)
        one = np.frombuffer(pg.synth_fill(BUF).tobytes(), np.uint8).copy()
        for _ in range(20):
            hl.input_callback(one.ctypes.data, BUF); hl.flush()
        lat = []
        for _ in range(200):
            t0 = time.perf_counter()
            hl.input_callback(one.ctypes.data, BUF); hl.flush()
            lat.append((time.perf_counter() - t0) * 1e6)
        hl.close()
        e2e_cb["one_transfer_latency_us"] = {"median": round(statistics.median(lat), 1), "p95": round(sorted(lat)[189], 1),
                                             "what": "wall time of perseus_gpu_input_callback(6144 B) + perseus_gpu_flush: samples resident on the device"}

    for p in (d_in, d_i32, d_f32):
        h.dev_free(p)

    # -- BASELINE.json's other GPU configurations, in the same driver-run line
    workloads = None
    if not args.no_workloads:
        wsteps = max(3, min(args.steps, 20))
        workloads = {"cfg3": workload_cfg3(cx, wsteps, 3), "cfg4": workload_cfg4(cx, max(3, min(args.steps, 10)), 3)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_leg(args.cpu_budget_s)

    cx.close()
    if rank != 0:
        return 0

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32+f32", "data": "synthetic",
        "config": workload_config(world, nbuf),
        "setup": {"l2": "no flush needed: inputs (1.07 GB) and outputs (2.86 GB) per step are far larger than the 126 MB L2",
                  "generated": "on the device (perseus_gpu_generate), rank r at byte offset r * shard size of the recording",
                  "pass": "int32 AND float written by one fused kernel launch per step",
                  "parallelism": f"{world} independent shard(s) (perseus_gpu_shard_range), no data-path collective", "tuning": tuning,
                  "device_map": cx.device_map, "numa_node": cx.numa},
        "roofline": {"bound": "hbm", "kernel": "unpack24_stream_kernel<I32|F32>", "achieved": round(achieved, 1), "peak": peak,
                     "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE_FUSED * ns, "bytes_per_sample": BYTES_PER_SAMPLE_FUSED,
                     "launch_ms": round(ms_step, 4), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4), "in_run_probe": probe},
        "single_format": extra,
        "recording_float_checksum": f"{recording_checksum:016x}",
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if e2e_bal:
        # Host-fed, multi-GPU: the link-weighted sharding is the product's way to feed a box whose GPUs do not reach host memory
        # equally fast, so it is the headline end-to-end number; the equal-shard run of the same recording stays beside it.
        line["e2e_equal_shards"] = e2e
        line["e2e"] = dict(e2e_bal, h2d_bytes_per_step=int(nbytes), pcie_h2d_gbs_measured=e2e["pcie_h2d_gbs_measured"],
                           pcie_h2d_gbs_all_ranks_at_once=e2e.get("pcie_h2d_gbs_all_ranks_at_once"),
                           h2d_bytes_per_step_note="average per GPU; shards differ in size, see transfers_per_rank")
    if e2e_rt:
        line["e2e_roundtrip"] = e2e_rt
    if e2e_pageable:
        line["e2e_pageable"] = e2e_pageable
    if e2e_cb:
        line["e2e_callback"] = e2e_cb
    if workloads:
        line["workloads"] = workloads
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    return 0


def other_workload(args):
    """--workload cfg3 / cfg4 alone: the sub-record of the default run as the whole line (e2e: null)."""
    cx = Ctx(args)
    rec = (workload_cfg3 if args.workload == "cfg3" else workload_cfg4)(cx, args.steps, args.warmup)
    cx.close()
    if cx.rank == 0:
        rec.update({"higher_is_better": True, "vs_baseline": None, "data": "synthetic", "e2e": None})
        rec.setdefault("roofline", {}).update({"traffic": None, "peak_source": cx.peak_src})
        rec.setdefault("config", {})["l2"] = "no flush needed: working set per step is far larger than the 126 MB L2"
        emit(rec)
    return 0


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--buffers", type=int, default=CFG2_BUFFERS, help="transfers per GPU (default: cfg2, 1 GiB)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--chunk-mib", type=int, default=0, help="staging chunk for host pointers (0 = library default, 32 MiB)")
    ap.add_argument("--slots", type=int, default=0, help="staging slots of the host-pointer pipeline (0 = library default, 3)")
    ap.add_argument("--no-workloads", action="store_true", help="skip the cfg3 / cfg4 sub-records")
    ap.add_argument("--no-roundtrip", action="store_true")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-memory end-to-end leg (N = 1)")
    ap.add_argument("--no-e2e", action="store_true", help="skip every host-fed leg (the line then has e2e: null)")
    ap.add_argument("--no-single", action="store_true", help="skip the one-format kernels")
    ap.add_argument("--no-balanced", action="store_true", help="skip the link-weighted sharding of the end-to-end leg (N > 1)")
    ap.add_argument("--balance-passes", type=int, default=3, help="untimed passes that refine the shard weights from observed rates (N > 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-callback", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="only the HBM-resident fused leg (the short command ncu wraps for the launch list)")
    ap.add_argument("--autotune", action="store_true", help="let the library measure its pipeline geometry on this device first")
    ap.add_argument("--pinned-wc", action="store_true", help="experiment: end-to-end input in write-combined pinned memory")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4"],
                    help="cfg2 (default, the headline): 1 GiB per GPU, fused; cfg3: 1024 mixed-rate receivers in one launch; "
                         "cfg4: 64 GiB recording sharded over the GPUs, float only")
    ap.add_argument("--cpu-budget-s", type=float, default=12.0)
    for k in ("variant", "tile", "stages", "ctas", "store"):
        ap.add_argument(f"--{k}", type=int, default=0)
    args = ap.parse_args()
    if args.headline_only:
        args.no_probe = args.no_roundtrip = args.no_pageable = args.no_cpu = args.no_callback = args.no_workloads = args.no_balanced = True
        args.no_e2e = args.no_single = True
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                   # timing rule: at least 3 warm-up steps
    if args.gpus > 1 and "RANK" not in os.environ and args.impl == "ours":        # convenience: relaunch ourselves under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 2000), __file__] + sys.argv[1:]
        return subprocess.call(cmd)
    claim_stdout()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args) if args.workload == "cfg2" else other_workload(args)


if __name__ == "__main__":
    sys.exit(main())
