#!/usr/bin/env python
"""Condenses `ncu --page raw --csv` dumps of the step kernels into the few metrics DESIGN.md quotes and
refreshes profiles/traffic.json (read by bench.py for roofline.traffic).

  python tools/ncu_summary.py TAG      (reads profiles/TAG_step_kernel_*_ncu_full_raw.csv)
"""
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "lts__t_sectors_srcunit_ltcfabric.sum", "lts__t_sectors_op_write.sum"]


def main():
    tag = sys.argv[1]
    lines, traffic = [], {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"{tag}_*_ncu_full_raw.csv"))):
        name = os.path.basename(path)[len(tag) + 1:-len("_ncu_full_raw.csv")]
        if name.startswith("step_kernel_"):
            name = name[len("step_kernel_"):]
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        tot = []
        for d in data:
            lines.append(f"--- {name}: {d[col['Kernel Name']]}")
            for m in WANT:
                if m in col:
                    lines.append(f"{m:70s} {d[col[m]]:>18s} {units[col[m]]}")
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot.append(sum(float(d[col[m]]) * scale[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
        traffic[name] = sum(tot) / len(tot)
    with open(os.path.join(ROOT, "profiles", f"{tag}_step_kernel_ncu_summary.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    out = {"2048": traffic.get("single_2048", old.get("2048")), "2048_fused": traffic.get("fused_2048", old.get("2048_fused")),
           "4096x4096x512": traffic.get("single_4096w", old.get("4096x4096x512")),
           "4096x4096x512_fused": traffic.get("fused_4096w", old.get("4096x4096x512_fused")),
           "2048_fused4": traffic.get("fused4_2048", old.get("2048_fused4")),
           "1024_fused4": traffic.get("fused4_1024", old.get("1024_fused4")),
           "source": f"profiles/{tag}_*_ncu_full_raw.csv (entries this tag did not capture are kept from the earlier round's r01e_*): dram__bytes_read.sum + dram__bytes_write.sum per launch "
                     "(mean over the captured launches: both x-offsets, and both step parities for the single-step "
                     "kernels). 2048 = 2048^3 grid, 17.18 GB algorithmic per single-step launch; a fused launch "
                     "advances two steps (34.36 GB algorithmic by the 2 B/update definition) on the same traffic; a fused4 launch (step4_kernel) "
                     "advances FOUR steps; the first z-pair of every band of four is loaded twice (5/4 of the grid), but on big single slabs the second load hits L2 (grouped bands, staggered units: 25/24 from DRAM). "
                     "4096x4096x512 = one rank's slab of 4096^3 on 8 GPUs (warp-pair kernels), same voxel count"}
    with open(tp, "w") as f:
        json.dump(out, f, indent=1)
    print("\n".join(lines[:12]))
    print(json.dumps(out)[:300])


if __name__ == "__main__":
    main()
