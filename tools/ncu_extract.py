#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of metrics the profiles/ summaries quote.
usage: python tools/ncu_extract.py raw.csv [raw2.csv ...] > summary.md   (also writes per-kernel DRAM bytes as JSON with --json PATH)"""
import csv, json, re, sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
]
QUOTED = [
    ("gpu__time_duration.sum", "duration_us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_active_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_throughput_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_slots_busy_pct"),
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    args = sys.argv[1:]
    jpath = None
    if "--json" in args:
        i = args.index("--json"); jpath = args[i + 1]; del args[i:i + 2]
    mpath = None
    if "--metrics-json" in args:      # per kernel: the utilisation figures bench.py quotes next to its live timings
        i = args.index("--metrics-json"); mpath = args[i + 1]; del args[i:i + 2]
    quoted = {}
    traffic = {}
    seen = set()
    for path in args:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = re.sub(r"^void\s+", "", re.sub(r"<unnamed>::|\(.*", "", r[idx["Kernel Name"]]).strip())
            if name in seen:
                continue
            seen.add(name)
            print("### %s  (%s)" % (name, path.split("/")[-1]))
            print("| metric | value |\n|---|---|")
            for m, label in METRICS:
                if m in idx:
                    print("| %s | %s %s |" % (label, r[idx[m]], units[idx[m]]))
            rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT.get(units[idx["dram__bytes_read.sum"]], 1.0)
            wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT.get(units[idx["dram__bytes_write.sum"]], 1.0)
            traffic[name.split("<")[0]] = int(rd + wr)
            q = {}
            for m, label in QUOTED:
                if m in idx:
                    try:
                        q[label] = round(float(r[idx[m]]), 3)
                    except ValueError:
                        pass
            q["source"] = "ncu --set full --clock-control none, one launch, kernel alone (" + path.split("/")[-1].replace(".csv", ".ncu-rep") + ")"
            quoted[name.split("<")[0]] = q
            print()
    if jpath:
        json.dump(traffic, open(jpath, "w"), indent=1, sort_keys=True)
    if mpath:
        json.dump(quoted, open(mpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
