"""`ncu -i ... --page raw --csv` exports -> one summary CSV (+ DRAM bytes per launch for profiles/r02_traffic.json).
usage: python scratch/ncu_summary.py out.csv workload=raw.csv [workload=raw.csv ...]"""
import csv, sys, collections
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "lts__t_bytes.sum"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def main():
    out, pairs = sys.argv[1], [a.split("=", 1) for a in sys.argv[2:]]
    with open(out, "w", newline="") as f:
        w = None
        for wl, path in pairs:
            rows = list(csv.reader(open(path)))
            hdr, units = rows[0], rows[1]
            idx = {h: i for i, h in enumerate(hdr)}
            cols = [c for c in COLS if c in idx]
            if w is None:
                w = csv.writer(f)
                w.writerow(["workload", "kernel", "launches"] + [f"{c} [{'us' if c.startswith('gpu__time') else 'byte' if 'bytes' in c else units[idx[c]]}]" for c in cols])
            agg = collections.OrderedDict()
            for r in rows[2:]:
                name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
                vals = []
                for c in cols:
                    v = float(r[idx[c]].replace(",", "")) if r[idx[c]] not in ("", "n/a") else float("nan")
                    vals.append(v * SCALE.get(units[idx[c]], 1.0) if (c.startswith("gpu__time") or "bytes" in c) else v)
                a = agg.setdefault(name + "#" + r[idx["launch__grid_size"]], [0, [0.0] * len(cols)])
                a[0] += 1
                a[1] = [x + y for x, y in zip(a[1], vals)]
            for name, (n, sums) in agg.items():
                w.writerow([wl, name.split("#")[0], n] + [round(x / n, 4) for x in sums])


main()
