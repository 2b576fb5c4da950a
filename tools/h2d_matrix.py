#!/usr/bin/env python
"""Maps the HOST-side ceiling of the end-to-end mode on a multi-GPU box (VERDICT r1 "weak #3").

The end-to-end path is `pinned host memory -> cudaMemcpyAsync -> unpack`; at 4 and 8 GPUs round 1 measured only 0.5 / 0.4
of the single-GPU rate per GPU and could only say "the library reaches 0.99 of what plain copies reach".  This tool times
PLAIN pinned copies, no library involved, for

    GPU subsets   {0} {0,1} {0,4} {0..3} {4..7} {0,2,4,6} all        (which GPUs share an uplink / root complex?)
    host memory   cudaHostAlloc default | cudaHostAlloc write-combined | 2 MiB-aligned mmap + MADV_HUGEPAGE + cudaHostRegister
    chunk size    8 / 32 / 128 MiB per cudaMemcpyAsync
    direction     H2D, D2H, and both at once

one process, one stream per GPU, all copies of a measurement queued before any is waited for, CUDA events per GPU; a
measurement's aggregate is total bytes / (latest end - earliest start).  Output: one JSON line per measurement on stdout and
a markdown table (--md FILE).  PyTorch is used for streams/events/pinned tensors only.

    python tools/h2d_matrix.py --md gpurun_out/h2d_matrix.md > gpurun_out/h2d_matrix.jsonl
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import mmap
import subprocess
import sys
import time

import torch

GIB = 1 << 30


def cudart():
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            pass
    import glob
    for path in glob.glob(sys.prefix + "/lib/python*/site-packages/nvidia/cuda_runtime/lib/libcudart.so*"):
        return C.CDLL(path)
    raise OSError("libcudart not found")


class HostBuf:
    """`nbytes` of pinned host memory of one kind, viewed as a torch uint8 tensor."""

    def __init__(self, kind: str, nbytes: int, rt):
        self.kind, self.nbytes, self.rt, self.ptr, self.map = kind, nbytes, rt, None, None
        if kind == "default":
            self.t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        elif kind == "write_combined":
            p = C.c_void_p()
            rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(0x04 | 0x01))       # WriteCombined | Portable
            if rc:
                raise RuntimeError(f"cudaHostAlloc(write-combined) -> {rc}")
            self.ptr = p.value
            self.t = torch.frombuffer((C.c_uint8 * nbytes).from_address(self.ptr), dtype=torch.uint8)
        elif kind == "hugepage_registered":
            self.map = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            base = C.addressof(C.c_char.from_buffer(self.map))
            self.ptr = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
            libc = C.CDLL(None, use_errno=True)
            libc.madvise(C.c_void_p(self.ptr), C.c_size_t(nbytes), C.c_int(14))                 # MADV_HUGEPAGE
            C.memset(self.ptr, 1, nbytes)                                                        # fault the pages in (as huge pages if THP allows)
            rc = rt.cudaHostRegister(C.c_void_p(self.ptr), C.c_size_t(nbytes), C.c_uint(0x01))   # Portable
            if rc:
                raise RuntimeError(f"cudaHostRegister -> {rc}")
            self.t = torch.frombuffer((C.c_uint8 * nbytes).from_address(self.ptr), dtype=torch.uint8)
        else:
            raise ValueError(kind)
        if kind != "hugepage_registered":
            self.t.fill_(1)

    def close(self):
        t, self.t = self.t, None
        del t
        if self.kind == "write_combined":
            self.rt.cudaFreeHost(C.c_void_p(self.ptr))
        elif self.kind == "hugepage_registered":
            self.rt.cudaHostUnregister(C.c_void_p(self.ptr))
            # the mmap object is released with the tensor's exporter; nothing else to do


def thp_state():
    try:
        return open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
    except OSError:
        return "unknown"


def anon_huge_kb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("AnonHugePages"):
                return int(line.split()[1])
    except OSError:
        pass
    return -1


def measure(devs, host, dev_in, dev_out, streams, chunk, per_gpu_bytes, direction):
    """Queues ceil(per_gpu_bytes / chunk) copies per GPU (round-robin over the GPUs, so they start together), both
    directions when direction == 'both'; returns per-GPU GB/s per direction and the aggregate."""
    nch = max(1, per_gpu_bytes // chunk)
    dirs = ["h2d", "d2h"] if direction == "both" else [direction]
    ev = {(d, k): (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for d in devs for k in dirs}
    for d in devs:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for d in devs:
        for k in dirs:
            ev[(d, k)][0].record(streams[(d, k)])
    for c in range(nch):
        for d in devs:
            off = (c * chunk) % (host[d].nbytes - chunk + 1)
            off -= off % 4096
            for k in dirs:
                with torch.cuda.stream(streams[(d, k)]):
                    if k == "h2d":
                        dev_in[d][:chunk].copy_(host[d].t[off:off + chunk], non_blocking=True)
                    else:
                        host[d].t[off:off + chunk].copy_(dev_out[d][:chunk], non_blocking=True)
    for d in devs:
        for k in dirs:
            ev[(d, k)][1].record(streams[(d, k)])
    for d in devs:
        torch.cuda.synchronize(d)
    wall = time.perf_counter() - t0
    res = {}
    for k in dirs:
        per = {d: nch * chunk / (ev[(d, k)][0].elapsed_time(ev[(d, k)][1]) * 1e-3) / 1e9 for d in devs}
        res[k] = {"per_gpu_gbs": {str(d): round(v, 2) for d, v in per.items()}, "min_gbs": round(min(per.values()), 2),
                  "sum_gbs": round(sum(per.values()), 1)}
    res["wall_aggregate_gbs"] = round(len(dirs) * len(devs) * nch * chunk / wall / 1e9, 1)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu-gib", type=float, default=2.0, help="bytes copied per GPU per measurement")
    ap.add_argument("--host-gib", type=float, default=1.0, help="size of each GPU's host buffer")
    ap.add_argument("--md", default=None)
    ap.add_argument("--kinds", default="default,write_combined,hugepage_registered")
    ap.add_argument("--chunks-mib", default="8,32,128")
    a = ap.parse_args()
    n = torch.cuda.device_count()
    subsets = [[0]]
    if n >= 2:
        subsets.append([0, 1])
    if n >= 8:
        subsets += [[0, 4], [0, 1, 2, 3], [4, 5, 6, 7], [0, 2, 4, 6]]
    elif n >= 4:
        subsets += [[0, 2], [0, 1, 2, 3]]
    if n > 2 and list(range(n)) not in subsets:
        subsets.append(list(range(n)))
    rt = cudart()
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
    chunks = [int(c) << 20 for c in a.chunks_mib.split(",")]
    host_bytes = int(a.host_gib * GIB)
    per_gpu = int(a.per_gpu_gib * GIB)
    dev_in, dev_out, streams = {}, {}, {}
    for d in range(n):
        with torch.cuda.device(d):
            dev_in[d] = torch.empty(max(chunks), dtype=torch.uint8, device=f"cuda:{d}")
            dev_out[d] = torch.ones(max(chunks), dtype=torch.uint8, device=f"cuda:{d}")
            for k in ("h2d", "d2h"):
                streams[(d, k)] = torch.cuda.Stream(d)
    try:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # pragma: no cover
        topo = f"nvidia-smi topo -m unavailable: {e}"
    rows = []
    meta = {"gpus": n, "thp": thp_state(), "per_gpu_gib": a.per_gpu_gib, "host_gib": a.host_gib}
    for kind in a.kinds.split(","):
        huge0 = anon_huge_kb()
        try:
            host = {d: HostBuf(kind, host_bytes, rt) for d in range(n)}
        except Exception as e:
            print(json.dumps({"kind": kind, "skipped": str(e)}), flush=True)
            continue
        note = ""
        if kind == "hugepage_registered":
            note = f"AnonHugePages grew by {(anon_huge_kb() - huge0) // 1024} MiB for {n * host_bytes >> 20} MiB requested (THP: {thp_state()})"
        for sub in subsets:
            for chunk in chunks:
                for direction in ("h2d", "d2h", "both"):
                    if direction == "both" and chunk != chunks[len(chunks) // 2]:
                        continue
                    measure(sub, host, dev_in, dev_out, streams, chunk, min(per_gpu, 256 << 20), direction)     # warm
                    r = measure(sub, host, dev_in, dev_out, streams, chunk, per_gpu, direction)
                    row = {"kind": kind, "gpus": sub, "chunk_mib": chunk >> 20, "direction": direction, **r, "note": note}
                    rows.append(row)
                    print(json.dumps(row), flush=True)
        for hb in host.values():
            hb.close()
    if a.md:
        with open(a.md, "w") as f:
            f.write(f"# Plain pinned-copy matrix (tools/h2d_matrix.py): {json.dumps(meta)}\n\n```\n{topo}```\n\n")
            f.write("per-GPU GB/s = that GPU's bytes / its own CUDA-event time while every GPU of the subset copies; min = slowest GPU\n\n")
            f.write("| memory | GPUs | chunk MiB | dir | H2D min / sum GB/s | D2H min / sum GB/s | per-GPU H2D | per-GPU D2H |\n|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                h, d = r.get("h2d"), r.get("d2h")
                f.write(f"| {r['kind']} | {','.join(map(str, r['gpus']))} | {r['chunk_mib']} | {r['direction']} | "
                        f"{(str(h['min_gbs']) + ' / ' + str(h['sum_gbs'])) if h else '-'} | {(str(d['min_gbs']) + ' / ' + str(d['sum_gbs'])) if d else '-'} | "
                        f"{' '.join(str(v) for v in h['per_gpu_gbs'].values()) if h else '-'} | {' '.join(str(v) for v in d['per_gpu_gbs'].values()) if d else '-'} |\n")
            for kind in sorted({r["kind"] for r in rows}):
                notes = {r["note"] for r in rows if r["kind"] == kind and r["note"]}
                for nt in notes:
                    f.write(f"\n{kind}: {nt}\n")


if __name__ == "__main__":
    main()
