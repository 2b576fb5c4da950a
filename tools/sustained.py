#!/usr/bin/env python
"""Sustained-throughput check: the fused unpack launched back to back for ~30 s, reported in 2-second windows with the
SM clock, power and throttle reasons NVML shows at the end of each window."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402


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
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 30.0)
