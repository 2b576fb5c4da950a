#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the small text summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r1_fused.ncu-rep > profiles/r1_ncu_fused.md
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_issue_stalled_membar.ratio",
    # launches that read / write pinned host memory themselves (the streaming path's direct route): 32-byte sectors over the link
    "syslts__t_sectors_srcunit_tex_aperture_sysmem_op_read_lookup_miss.sum", "syslts__t_sectors_srcunit_tex_aperture_sysmem_op_write_lookup_miss.sum",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{path}`\n")
    for n, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")]
        print(f"## launch {n}: `{name}`  grid {r[hdr.index('Grid Size')]}  block {r[hdr.index('Block Size')]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd *= scale[units[hdr.index("dram__bytes_read.sum")]]
        wr *= scale[units[hdr.index("dram__bytes_write.sum")]]
        print(f"| **traffic = dram read + write** | {rd + wr:.0f} | byte |\n")
    stalls = [(h, i) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    if stalls and len(rows) > 2:
        r = rows[2]
        top = sorted(((float(r[i] or 0), h) for h, i in stalls), reverse=True)[:8]
        print("## warp stall reasons, launch 0 (warps stalled per issue-active cycle)\n")
        for v, h in top:
            print(f"- {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')}: {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
