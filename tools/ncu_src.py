"""Summarise an ncu report's source page: python tools/ncu_src.py report.ncu-rep [min_samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][:2])
h = rows[1]; idx = {n: i for i, n in enumerate(h)}; data = rows[2:]
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
print("total samples", tot)
cum = 0
for r in data:
    s = int(r[idx['# Samples']] or 0); cum += s
    src = r[idx['Source']]
    key = any(k in src for k in ('UTCHMMA', 'UTMALDG', 'LDTM', 'UTCBAR', 'ELECT', 'BAR.SYNC', 'ACQBULK', 'DEPBAR'))
    if s >= thr or key:
        st = {n[6:]: int(r[idx[n]] or 0) for n in stalls}
        st = {k: v for k, v in st.items() if v > 2}
        print(r[idx['Address']][-5:], f"{s:5d} cum{cum:6d}", f"{r[idx['Instructions Executed']]:>8s}", src[:64], st)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg', 'smsp__average_warp_latency_per_inst', 'launch__grid_size', 'sm__warps_active.avg.pct']
for a, b, c in zip(rows[0], rows[1], rows[2]):
    if any(w in a + ' ' for w in want): print(a, b, c)
