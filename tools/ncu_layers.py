"""Per-layer roofline table from one `ncu --metrics ... --csv` launch list + the op list tools/run_once.py dumps.

usage: ncu_layers.py <ncu csv> <ops json> <out prefix>
Writes <out>.csv and <out>.txt: per launch the device time, the tensor-pipe utilisation, DRAM bytes read / written, achieved DRAM GB/s
against the measured copy bandwidth (MEASURED_PEAKS.json), the algorithmic bytes and FLOPs of the op and the fractions.
Times under ncu are cold-cache and serialised (compare shares and fractions, not absolutes).

Tensor-pipe utilisation, normalised here (ncu's `sm__ops_path_tensor_op_hmma_*` counters read 0 for tcgen05.mma, and the
`sm__pipe_tensor_subpipe_hmma_cycles_active_realtime` counter is collected per TPC, which is how round 1's table got ratios > 1):
  mma_insts         = `sm__inst_executed_pipe_tensor_subpipe_hmma.sum`: this one does count tcgen05.mma (checked: SNUNet conv0_4.conv2,
                      64 pairs = 32768 tiles x 18 MMAs = 589 824, the counter reads 589 824);
  tensor_math_pct   = mma_insts x (128 x N x 16 MACs) / (elapsed SM cycles x 4096 MAC/clk/SM x 148 SMs): the share of the dense bf16
                      peak the kernel's MMAs amount to (includes MACs on padding / halo columns), in [0, 100];
  tensor_busy_pct   = mma_insts x cost(N) / (elapsed SM cycles x 148): the share of the cycles the tensor pipe is occupied, with
                      cost(N) the measured cycles one SS-mode M=128 K=16 MMA holds the pipe (tools/ubench/mma_n.cu: 45 for N <= 48,
                      48.5 / 56.5 / 64 / 96 / 128 for N = 64 / 96 / 128 / 192 / 256), in [0, 100]."""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = "sm__inst_executed_pipe_tensor_subpipe_hmma.sum"
N_SM = 148


def mma_cost(n):
    return max(45.0, 32.0 + n / 4.0, n / 2.0) if n else float("nan")


def main():
    ncu_csv, ops_json, out = sys.argv[1:4]
    peaks = {"hbm_gbs": 6549.1, "bf16_tflops": 1665.1}
    try:
        peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
    except OSError:
        pass
    rows = {}
    lines = [l for l in open(ncu_csv) if l.startswith('"')]
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        d = rows.setdefault(k, {"kernel": r["Kernel Name"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            d[r["Metric Name"]] = float("nan")
        d["unit:" + r["Metric Name"]] = r["Metric Unit"]
    meta = json.load(open(ops_json))
    labels = []
    for op in meta["ops"]:
        for j in range(op["kernels"]):
            labels.append((op["name"] + (f"#{j}" if op["kernels"] > 1 else ""), op if j == 0 else None))
    ids = sorted(rows)
    n_chunks = max(1, -(-meta["pairs"] // meta["chunk"]))
    pairs_per_launch = min(meta["pairs"], meta["chunk"])
    total_us = 0.0
    table = []
    for i, k in enumerate(ids):
        d = rows[k]
        # the plan's kernels come first, in op order (once per chunk); whatever follows (torch's own reductions on the result) is named by kernel
        name, op = labels[i % len(labels)] if i < len(labels) * n_chunks else (d["kernel"][:40], None)
        dur = d.get("gpu__time_duration.sum", float("nan"))
        if d.get("unit:gpu__time_duration.sum", "ns") in ("ns", "nsecond"):
            dur_us = dur / 1e3
        elif d.get("unit:gpu__time_duration.sum") in ("us", "usecond"):
            dur_us = dur
        else:
            dur_us = dur * 1e3      # ms
        rd, wr = d.get("dram__bytes_read.sum", float("nan")), d.get("dram__bytes_write.sum", float("nan"))
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            u = d.get("unit:" + key, "byte")
            f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            if key.endswith("read.sum"):
                rd *= f
            else:
                wr *= f
        gbs = (rd + wr) / (dur_us * 1e-6) / 1e9 if dur_us else float("nan")
        alg_b = op["bytes_per_pair"] * pairs_per_launch if op else float("nan")
        alg_f = 2.0 * op["macs_per_pair"] * pairs_per_launch if op else float("nan")
        total_us += dur_us
        nan = float("nan")
        n_mma = d.get(T, nan)
        n_tile = op["n_tile"] if op else 0
        cyc_avg = d.get("sm__cycles_elapsed.avg", nan)
        math_pct = 100.0 * n_mma * 128.0 * n_tile * 16.0 / (cyc_avg * 4096.0 * N_SM) if (cyc_avg and n_tile) else nan
        busy = 100.0 * n_mma * mma_cost(n_tile) / (cyc_avg * N_SM) if (cyc_avg and n_tile) else nan
        table.append({"id": k, "op": name, "kernel": d["kernel"].split("(")[0][-48:], "us": dur_us, "tensor_pct": math_pct,
                      "tensor_busy_pct": busy, "mma_insts": n_mma,
                      "dram_read_MB": rd / 1e6, "dram_write_MB": wr / 1e6, "dram_GBs": gbs, "dram_frac_of_measured": gbs / peaks["hbm_gbs"],
                      "dram_pct_ncu": d.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", float("nan")),
                      "alg_MB": alg_b / 1e6, "alg_GBs": alg_b / (dur_us * 1e-6) / 1e9 if dur_us else float("nan"),
                      "alg_hbm_frac": alg_b / (dur_us * 1e-6) / 1e9 / peaks["hbm_gbs"] if dur_us else float("nan"),
                      "alg_TFLOPs": alg_f / (dur_us * 1e-6) / 1e12 if dur_us else float("nan"),
                      "alg_tensor_frac": alg_f / (dur_us * 1e-6) / 1e12 / peaks["bf16_tflops"] if dur_us else float("nan"),
                      "regs": d.get("launch__registers_per_thread", float("nan")), "grid": d.get("launch__grid_size", float("nan")),
                      "n_tile": op["n_tile"] if op else 0, "xf": op["xf_cs"] if op else 0})
    keys = list(table[0])
    with open(out + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(table)
    with open(out + ".txt", "w") as f:
        f.write(f"# {meta['net']} {meta['pairs']} pairs of {meta['h']}x{meta['h']} (chunk {meta['chunk']}): one forward under ncu, --clock-control none; "
                f"times are cold-cache and serialised.\n# tensor% = tcgen05.mma count ({T}) x 128 x N x 16 MACs / (elapsed cycles x 4096 x {N_SM} SMs); busy% = count x measured cycles per MMA of that N / (elapsed cycles x {N_SM})\n# peaks: HBM {peaks['hbm_gbs']} GB/s, bf16 {peaks['bf16_tflops']} TFLOP/s (MEASURED_PEAKS.json)\n")
        f.write(f"{'op':28s} {'us':>8s} {'share':>6s} {'tensor%':>8s} {'busy%':>6s} {'dramGB/s':>9s} {'of HBM':>7s} {'alg GB/s':>9s} {'alg/HBM':>8s} {'algTF/s':>8s} {'N':>4s} {'xf':>3s} {'regs':>5s}\n")
        for r in table:
            f.write(f"{r['op'][:28]:28s} {r['us']:8.1f} {r['us'] / total_us:6.3f} {r['tensor_pct']:8.1f} {r['tensor_busy_pct']:6.1f} {r['dram_GBs']:9.0f} {r['dram_frac_of_measured']:7.2f} "
                    f"{r['alg_GBs']:9.0f} {r['alg_hbm_frac']:8.2f} {r['alg_TFLOPs']:8.1f} {r['n_tile']:4d} {r['xf']:3d} {r['regs']:5.0f}\n")
        f.write(f"total {total_us:.1f} us over {len(table)} launches\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
