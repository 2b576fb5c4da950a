#!/usr/bin/env python
"""Sustained-throughput check: the fused unpack launched back to back for ~30 s, reported in 2-second windows with the
SM clock, power and throttle reasons NVML shows at the end of each window.

    sustained.py [seconds]              kernel only, recording resident in HBM
    sustained.py [seconds] e2e          BASELINE config 5: the pinned-host recording replayed through perseus_gpu_unpack (chunked
                                        H2D overlapped with the kernels, outputs stay in HBM), GB/s of wire per window
    sustained.py [seconds] callback     config 5 variant B: 6144-byte transfers through perseus_gpu_input_callback on one host thread"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402


def nvml_fields(N, dev):
    return {"sm_mhz": N.nvmlDeviceGetClockInfo(dev, N.NVML_CLOCK_SM), "mem_mhz": N.nvmlDeviceGetClockInfo(dev, N.NVML_CLOCK_MEM),
            "power_w": round(N.nvmlDeviceGetPowerUsage(dev) / 1000.0, 1), "temp_c": N.nvmlDeviceGetTemperature(dev, N.NVML_TEMPERATURE_GPU),
            "reasons": hex(N.nvmlDeviceGetCurrentClocksEventReasons(dev))}


def main_e2e(seconds, mode):
    import pynvml as N
    pg = G.load_package()
    N.nvmlInit()
    dev = N.nvmlDeviceGetHandleByIndex(0)
    n = 174_762 * 6144
    ns = n // 6
    flags = pg.OUT_INT32 | pg.OUT_FLOAT
    with pg.PerseusGpu(device=0, stream_flags=flags, slab_bytes=16 << 20, nslabs=4, nstreams=2) as h:
        d_in, d_i, d_f = h.dev_alloc(n), h.dev_alloc(ns * 8), h.dev_alloc(ns * 8)
        h.generate(d_in, n)
        pin = h.host_alloc(n)
        h.memcpy(pin, d_in, n)
        h.unpack(d_in, n, d_i, d_f, flags | pg.CHECKSUM)
        want = h.get_checksums()
        v = pg.VirtualReceiver(sample_rate=2_000_000, replay=True) if mode == "callback" else None
        t_end = time.time() + seconds
        while time.time() < t_end:
            t0, reps = time.perf_counter(), 0
            while time.perf_counter() - t0 < 1.8:
                if v is None:
                    h.unpack(pin, n, d_i, d_f, flags | pg.ASYNC | pg.CHECKSUM)
                    assert h.get_checksums() == want
                else:
                    v.run(6144, *h.callback, 174_762)
                    h.flush()
                reps += 1
            dt = time.perf_counter() - t0
            print(json.dumps({"mode": mode, "passes": reps, "wire_gbs": round(reps * n / dt / 1e9, 2), "msamples_per_s": round(reps * ns / dt / 1e6, 1),
                              **nvml_fields(N, dev)}), flush=True)
        if v is not None:
            st = h.stats()
            print(json.dumps({"callbacks": st["callbacks"], "slabs": st["slabs"], "slab_stalls": st["stalls"]}))
            v.close()


def main(seconds=30.0):
    import pynvml as N
    pg = G.load_package()
    N.nvmlInit()
    dev = N.nvmlDeviceGetHandleByIndex(0)
    with pg.PerseusGpu(device=0) as h:
        n = 174_762 * 6144
        ns = n // 6
        d_in, d_i, d_f = h.dev_alloc(n), h.dev_alloc(ns * 8), h.dev_alloc(ns * 8)
        h.generate(d_in, n)
        flags = pg.OUT_INT32 | pg.OUT_FLOAT
        t_end = time.time() + seconds
        while time.time() < t_end:
            h.event_record(0)
            for _ in range(3000):                      # ~1.8 s of back-to-back launches
                h.unpack(d_in, n, d_i, d_f, flags | pg.ASYNC)
            h.event_record(1)
            h.sync()
            ms = h.event_elapsed_ms(0, 1) / 3000
            print(json.dumps({"ms_per_launch": round(ms, 4), "gbs": round(22 * ns / ms / 1e6, 1),
                              "sm_mhz": N.nvmlDeviceGetClockInfo(dev, N.NVML_CLOCK_SM), "mem_mhz": N.nvmlDeviceGetClockInfo(dev, N.NVML_CLOCK_MEM),
                              "power_w": round(N.nvmlDeviceGetPowerUsage(dev) / 1000.0, 1),
                              "temp_c": N.nvmlDeviceGetTemperature(dev, N.NVML_TEMPERATURE_GPU),
                              "reasons": hex(N.nvmlDeviceGetCurrentClocksEventReasons(dev))}), flush=True)
        assert h.verify(d_in, n, d_i, d_f, flags)[0] == 0


if __name__ == "__main__":
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    if len(sys.argv) > 2:
        main_e2e(secs, sys.argv[2])
    else:
        main(secs)
