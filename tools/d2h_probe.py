#!/usr/bin/env python
"""Where do the last percent of the round trip go?  Plain device->pinned-host copies of the cfg2 outputs (2 x 1.43 GB) in the
shapes the library's pipeline produces them: one big copy, back-to-back 42.7 MiB chunks into one array, chunks alternating
between the int32 and the float array, and the same with an event wait between chunks.  torch streams/events only."""
import json

import torch

import sys

NS = 178_956_288
CH = int(sys.argv[1]) if len(sys.argv) > 1 else (32 << 20) // 6 * 8          # output bytes of one 32 MiB wire chunk


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    nbytes = NS * 8
    hi = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    hf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    di = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    df = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    stage = [torch.ones(CH, dtype=torch.uint8, device="cuda") for _ in range(6)]
    hi.fill_(0); hf.fill_(0)
    offs = list(range(0, nbytes, CH))

    def one_big():
        hf.copy_(df, non_blocking=True)

    def two_big():
        hi.copy_(di, non_blocking=True); hf.copy_(df, non_blocking=True)

    def chunks_one_array():
        for o in offs:
            n = min(CH, nbytes - o)
            hf[o:o + n].copy_(df[o:o + n], non_blocking=True)

    def chunks_alternating():
        for o in offs:
            n = min(CH, nbytes - o)
            hi[o:o + n].copy_(di[o:o + n], non_blocking=True)
            hf[o:o + n].copy_(df[o:o + n], non_blocking=True)

    def chunks_alternating_from_staging():
        for k, o in enumerate(offs):
            n = min(CH, nbytes - o)
            hi[o:o + n].copy_(stage[(2 * k) % 6][:n], non_blocking=True)
            hf[o:o + n].copy_(stage[(2 * k + 1) % 6][:n], non_blocking=True)

    for name, fn, total in (("one_big_1.43GB", one_big, nbytes), ("two_big_2.86GB", two_big, 2 * nbytes), ("chunks_one_array", chunks_one_array, nbytes),
                            ("chunks_alternating_two_arrays", chunks_alternating, 2 * nbytes),
                            ("chunks_alternating_from_3_staging_slots", chunks_alternating_from_staging, 2 * nbytes)):
        ms = timed(fn)
        print(json.dumps({"chunk": CH, "case": name, "ms": round(ms, 3), "gbs": round(total / ms / 1e6, 2)}), flush=True)


if __name__ == "__main__":
    main()
