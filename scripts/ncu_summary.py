#!/usr/bin/env python
"""Digest of an `ncu --set full` capture (.ncu-rep) into the handful of numbers the roofline discussion uses; writes a .summary.txt.
   python scripts/ncu_summary.py capture.ncu-rep "units processed by the launch" "unit name" > profiles/ncu/NAME.summary.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (warps active)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("launch__registers_per_thread", "registers per thread"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps per issue)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global load requests"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    unit_name = sys.argv[3] if len(sys.argv) > 3 else "unit"
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit_row = rows[0], rows[1]
    for r in rows[2:]:
        get = lambda k: (r[hdr.index(k)], unit_row[hdr.index(k)]) if k in hdr else (None, None)
        print(f"capture: {rep.split('/')[-1]}  (ncu --set full --clock-control none --import-source on, one launch)")
        print(f"kernel : {get('Kernel Name')[0]}")
        vals = {}
        for key, label in WANT:
            v, u = get(key)
            if v is None:
                continue
            vals[key] = (float(v.replace(',', '')), u)
            print(f"  {label:45s} {v} {u}")
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            rd = vals["dram__bytes_read.sum"][0] * scale[vals["dram__bytes_read.sum"][1]]
            wr = vals["dram__bytes_write.sum"][0] * scale[vals["dram__bytes_write.sum"][1]]
            tscale = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
            t = vals["gpu__time_duration.sum"][0] * tscale[vals["gpu__time_duration.sum"][1]]
            print(f"  {'DRAM traffic / duration':45s} {(rd + wr) / t / 1e9:.0f} GB/s")
            if units:
                print(f"  {'DRAM bytes per ' + unit_name:45s} {(rd + wr) / units:.1f}  ({units:.4g} {unit_name}s in this launch)")
                print(f"  {'warp instructions per ' + unit_name:45s} {vals['smsp__inst_executed.sum'][0] / units:.2f}")
        except Exception as ex:
            print("  (derived figures unavailable:", ex, ")")
        print()


if __name__ == "__main__":
    main()
