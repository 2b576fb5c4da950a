#!/usr/bin/env python
"""Kernel-only probe used for tuning sweeps and as the short command ncu wraps.

    python tools/kernel_probe.py --fmt both --reps 20                       one configuration
    python tools/kernel_probe.py --sweep > gpurun_out/sweep.jsonl            grid over tile/stages/CTAs/store/variant

Times perseus_gpu_unpack on the cfg2 recording (1 GiB wire, device resident) with CUDA events on the
launching stream; verifies the output with the independent per-sample kernel before timing."""
import argparse
import itertools
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

BUF, NBUF = 6144, 174_762


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fmt", default="both", choices=["i32", "f32", "both", "pow2"])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--buffers", type=int, default=NBUF)
    ap.add_argument("--in-offset", type=int, default=0, help="misalign the wire pointer by this many bytes (exercises the direct kernel)")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--out-skew", type=int, default=0, help="move the float output this many bytes (multiple of 16) further from the int32 output: "
                    "does the distance between the two write streams matter to the DRAM?")
    ap.add_argument("--fine", action="store_true", help="few CTAs/SM, all tile sizes, every ring depth, 3 repeats each")
    for k in ("variant", "tile", "stages", "ctas", "store"):
        ap.add_argument(f"--{k}", type=int, default=0)
    a = ap.parse_args()
    pg = G.load_package()
    peak = 6461.2
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    with pg.PerseusGpu(device=0) as h:
        n = a.buffers * BUF
        ns = n // 6
        d_in0, d_i, d_f0 = h.dev_alloc(n + 64), h.dev_alloc(ns * 8), h.dev_alloc(ns * 8 + a.out_skew)
        d_in, d_f = d_in0 + a.in_offset, d_f0 + a.out_skew
        h.generate(d_in, n)
        fmts = {"i32": (pg.OUT_INT32, d_i, None, 14), "f32": (pg.OUT_FLOAT, None, d_f, 14), "pow2": (pg.OUT_FLOAT_POW2, None, d_f, 14),
                "both": (pg.OUT_INT32 | pg.OUT_FLOAT, d_i, d_f, 22)}

        def run(fmt, **tune):
            flags, oi, of, bps = fmts[fmt]
            try:
                h.set_tuning(**tune)
            except pg.PerseusGpuError:
                return None
            h.unpack(d_in, n, oi, of, flags)
            bad, _ = h.verify(d_in, n, oi, of, flags)
            assert bad == 0, (fmt, tune, bad)
            for _ in range(3):
                h.unpack(d_in, n, oi, of, flags | pg.ASYNC)
            h.sync()
            h.event_record(0)
            for _ in range(a.reps):
                h.unpack(d_in, n, oi, of, flags | pg.ASYNC)
            h.event_record(1)
            h.sync()
            ms = h.event_elapsed_ms(0, 1) / a.reps
            gbs = bps * ns / (ms * 1e-3) / 1e9
            return {"fmt": fmt, **h.get_tuning(), "out_skew": a.out_skew, "f32_minus_i32": d_f - d_i, "ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4),
                    "gsamples_s": round(ns / ms / 1e6, 1)}

        if a.fine:
            for fmt in ("both", "f32", "i32"):
                for ctas, tile, stages in itertools.product((1, 2), (6144, 9216, 12288, 18432, 24576), range(2, 9)):
                    rs = [run(fmt, variant=pg.VARIANT_STREAM, tile_bytes=tile, stages=stages, ctas_per_sm=ctas, store_mode=1) for _ in range(3)]
                    if rs[0]:
                        best = max(rs, key=lambda r: r["gbs"])
                        best["gbs_all"] = sorted(r["gbs"] for r in rs)
                        print(json.dumps(best), flush=True)
        elif not a.sweep:
            print(json.dumps(run(a.fmt, variant=a.variant, tile_bytes=a.tile, stages=a.stages, ctas_per_sm=a.ctas, store_mode=a.store)))
        else:
            for fmt in ("both", "f32", "i32"):
                print(json.dumps(run(fmt, variant=pg.VARIANT_DIRECT, store_mode=1)), flush=True)
                print(json.dumps(run(fmt, variant=pg.VARIANT_DIRECT, store_mode=2)), flush=True)
                for tile, stages, ctas, store in itertools.product((6144, 12288, 24576), (2, 3, 4, 6, 8), (1, 2, 3, 4, 6), (1, 2)):
                    if fmt == "i32" and store == 2:
                        continue
                    r = run(fmt, variant=pg.VARIANT_STREAM, tile_bytes=tile, stages=stages, ctas_per_sm=ctas, store_mode=store)
                    if r:
                        print(json.dumps(r), flush=True)
        for p in (d_in0, d_i, d_f0):
            h.dev_free(p)


if __name__ == "__main__":
    main()
