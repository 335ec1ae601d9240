"""One line per conv op from a tools/trace_op.py log: grid, ring depths, cycles per tile and the share of the CTA's run time each
role spends WAITING (median over the CTAs).  A role that never waits is the pacing one.
usage: trace_table.py gpurun_out/trace_x.log > profiles/r2_roles_x.txt"""
import re
import sys

L = open(sys.argv[1]).read().split("\n")
print(f"# {sys.argv[1]}: per-op role waits (STCD_TRACE=1 clock stamps, median over CTAs).  wait shares are fractions of (first operand -> exit):")
print("# iss:a = issuer waiting for activations, iss:w = for streamed weights, iss:acc = for a free accumulator; epi = epilogue warp 3 waiting for an")
print("# accumulator; prod = activation producer waiting for a free stage.  gap = first MMA - end of the previous kernel (us), busy = first MMA -> end (us).")
print(f"{'op':28s} {'grid':>7s} {'aS':>2s} {'wS':>2s} {'Wres':>4s} {'tmem':>4s} {'t/cta':>5s} {'setup':>6s} {'cyc/tile':>8s} {'gap':>6s} {'busy':>7s}  {'iss:a':>5s} {'iss:w':>5s} {'iss:acc':>7s} {'epi':>5s} {'prod':>5s}")
for i, l in enumerate(L):
    m = re.match(r"(\S+)\s+grid=\((\d+),(\d+)\) smem=(\d+) aS=(\d+) wS=(\d+) res=(\d+) tmem=(\d+) tiles=(\d+)", l)
    if not m:
        continue
    med = dict(re.findall(r"(\S+)=(-?\d+)", L[i + 1]))
    sh = dict(re.findall(r"(\S+)=(-?[\d.]+)", L[i + 2]))
    tl = re.search(r"first-mma - prev end\s+([-\d.]+), busy\s+([\d.]+)", L[i - 1])
    g = m.groups()
    cyc = (int(med["exit"]) - int(med["mma:a0"])) / max(1, int(med["tiles"]))
    print(f"{g[0][:28]:28s} {g[1] + 'x' + g[2]:>7s} {g[4]:>2s} {g[5]:>2s} {g[6]:>4s} {g[7]:>4s} {med['tiles']:>5s} {med['setup']:>6s} {cyc:8.0f} {tl.group(1) if tl else '':>6s} {tl.group(2) if tl else '':>7s}  "
          f"{sh.get('mma:wait_a', ''):>5s} {sh.get('mma:wait_w', ''):>5s} {sh.get('mma:wait_acc', ''):>7s} {sh.get('epi:wait_acc', ''):>5s} {sh.get('prod:wait_empty', ''):>5s}")
