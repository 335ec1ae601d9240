"""Measured time over floor per op, from a profiles/r2_layers_*.csv table (tools/ncu_layers.py).

floor = max(HBM floor, tensor-issue floor): the HBM floor is the op's algorithmic bytes (every source read once, every
output written once) at the measured copy bandwidth; the tensor-issue floor is its tcgen05.mma count x the measured cycles
per MMA of that N (profiles/r2_mma_cycles_by_n.txt) spread over 148 SMs.  Both are already in the table as fractions of the
measured time (alg_hbm_frac, tensor_busy_pct), so x_floor = 1 / max(of the two).  Times are ncu's (cold cache, serialised).
usage: floor_table.py profiles/r2_layers_c2_b64.csv > profiles/r2_floor_c2_b64.txt"""
import csv
import math
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
print(f"# {sys.argv[1]}: measured / max(HBM floor, tensor-issue floor) per op; round-1 verdict item 1 asked for <= 1.5 on *.conv2 and Up*")
print(f"{'op':30s} {'us':>8s} {'hbm floor us':>13s} {'mma floor us':>13s} {'x floor':>8s}  bound")
tot = tot_floor = 0.0
worst = []
for r in rows:
    us = float(r["us"])
    hb = float(r["alg_hbm_frac"]) if r["alg_hbm_frac"] not in ("", "nan") else float("nan")
    mm = float(r["tensor_busy_pct"]) / 100 if r["tensor_busy_pct"] not in ("", "nan") else float("nan")
    cands = [v for v in (hb, mm) if not math.isnan(v)]
    if not cands:
        continue
    f = max(cands)
    hb_us = us * hb if not math.isnan(hb) else float("nan")
    mm_us = us * mm if not math.isnan(mm) else float("nan")
    print(f"{r['op'][:30]:30s} {us:8.1f} {hb_us:13.1f} {mm_us:13.1f} {1 / f:8.2f}  {'hbm' if f == hb else 'tensor'}")
    tot += us
    tot_floor += us * f
    worst.append((1 / f, r["op"]))
print(f"{'total':30s} {tot:8.1f} {'':13s} {'':13s} {tot / tot_floor:8.2f}")
worst.sort(reverse=True)
print("# furthest from their floor: " + ", ".join(f"{n} {x:.2f}x" for x, n in worst[:8]))
