"""Summarise ncu reports (raw page) into a CSV for profiles/:  python tools/ncu_summary.py out.csv label=report.ncu-rep ..."""
import csv, io, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def main():
    out = sys.argv[1]
    rows_out = []
    for arg in sys.argv[2:]:
        label, rep = arg.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, unit = rows[0], rows[1]
        for data in rows[2:]:
            rec = {"label": label}
            for h, u, v in zip(head, unit, data):
                if h == "Kernel Name":
                    rec["kernel"] = v
                for w in WANT:
                    if h == w or h.endswith("." + w):
                        rec[w] = f"{v} {u}".strip()
            rows_out.append(rec)
    # labels of the form workload:op:pairs_per_launch also feed profiles/r1_traffic.json (bench.py's roofline.traffic)
    import json, os
    traffic = {}
    def num(v):
        x, unit = v.split()[0], (v.split() + [""])[1]
        return float(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    for r in rows_out:
        parts = r["label"].split(":")
        if len(parts) == 3 and "dram__bytes_read.sum" in r:
            tot = num(r["dram__bytes_read.sum"]) + num(r["dram__bytes_write.sum"])
            traffic.setdefault(parts[0], {})[parts[1]] = {
                "dram_bytes_per_pair": int(tot / int(parts[2])),
                "capture": f"{parts[2]} pairs per launch, {r['dram__bytes_read.sum']} read + {r['dram__bytes_write.sum']} written, {r.get('gpu__time_duration.sum')}"}
    if traffic:
        with open(os.path.join(os.path.dirname(out), "r1_traffic.json"), "w") as f:
            json.dump({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per image pair from the ncu --set full captures in "
                                   "r1_ncu_full_summary.csv (bytes per launch / pairs per launch); bench.py scales it to its launch size",
                       **traffic}, f, indent=1)
    cols = ["label", "kernel"] + WANT
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        for r in rows_out:
            w.writerow([r.get(c, "") for c in cols])
    for r in rows_out:
        try:
            busy = float(r["sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"].split()[0]) / float(r["sm__cycles_elapsed.avg"].split()[0])
        except (KeyError, ValueError, ZeroDivisionError):
            busy = float("nan")
        print(r["label"], r.get("gpu__time_duration.sum"), "dram R/W", r.get("dram__bytes_read.sum"), r.get("dram__bytes_write.sum"),
              "dram%", r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f"tensor-subpipe busy {busy:.3f}",
              "regs", r.get("launch__registers_per_thread"), "grid", r.get("launch__grid_size"))


if __name__ == "__main__":
    main()
