#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: headline metrics per captured launch, warp-stall sample distribution and the
dynamic instruction mix (from the source page).   python scripts/ncu_summary.py <report.ncu-rep> [units per launch]"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_sample_count"]


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    hdr, unit = rows[0], rows[1]
    print("# %s" % rep)
    for r in rows[2:]:
        print("\n## %s" % r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                print("%-72s %s %s" % (k, r[hdr.index(k)], unit[hdr.index(k)]))
        try:
            tr = float(r[hdr.index("dram__bytes_read.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit[hdr.index("dram__bytes_read.sum")]]
            tw = float(r[hdr.index("dram__bytes_write.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit[hdr.index("dram__bytes_write.sum")]]
            print("%-72s %.0f byte" % ("dram traffic per launch (read + write)", tr + tw))
            if units:
                print("%-72s %.0f byte" % ("dram traffic per unit (%g units per launch)" % units, (tr + tw) / units))
                print("%-72s %.0f" % ("warp instructions per unit", float(r[hdr.index("smsp__inst_executed.sum")]) / units))
        except (ValueError, KeyError):
            pass
        st = [(k.replace("smsp__pcsamp_warps_issue_stalled_", ""), int(float(r[i]))) for i, k in enumerate(hdr)
              if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and r[i] not in ("", "n/a")]
        tot = sum(v for _, v in st) or 1
        print("warp stall samples: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(st, key=lambda kv: -kv[1])[:9]))
    src = page(rep, "source")
    h, data, nk = None, [], 0
    for r in src:
        if r and r[0] == "Kernel Name":
            nk += 1
            if nk == 2:
                break
            continue
        if r and r[0] == "Address":
            h = r
            continue
        if h and len(r) >= len(h) - 2:
            data.append(r)
    if h:
        ix = {k: i for i, k in enumerate(h)}
        op = collections.Counter()
        for r in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
            op[m.group(2) if m else "?"] += int(r[ix["Instructions Executed"]])
        tot = sum(op.values()) or 1
        print("\ndynamic instruction mix of the first captured launch (%d SASS lines, %d warp instructions):" % (len(data), tot))
        print(", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in op.most_common(16)))


if __name__ == "__main__":
    main()
