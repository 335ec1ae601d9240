"""profiles/r2_traffic.json: DRAM bytes (read + written) per image pair of every op, from the per-layer ncu tables
(tools/ncu_layers.py csv: dram__bytes_read.sum + dram__bytes_write.sum per launch / pairs per launch).  bench.py reports the
dominant kernel's entry as `roofline.traffic`."""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = {"snunet_256_b64": ("r2_layers_c2_b64", 64), "segcd_r34_1024_b16": ("r2_layers_c3_b4", 4), "siamunet_diff_256_b64": ("r2_layers_c1_b64", 64),
       "siamunet_diff_256": ("r2_layers_c1_b8", 8), "changegnn_v1_256_b32": ("r2_layers_c4_b8", 8), "changeformer_v6_256_b32": ("r2_layers_c5_b8", 8)}
out = {"_comment": "DRAM bytes per image pair (dram__bytes_read.sum + dram__bytes_write.sum of one launch / pairs in that launch), "
                   "ncu --clock-control none, round 2 kernels; source tables: profiles/r2_layers_*.csv"}
d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles")
for wl, (name, pairs) in SRC.items():
    path = os.path.join(d, name + ".csv")
    if not os.path.exists(path):
        continue
    out[wl] = {}
    for r in csv.DictReader(open(path)):
        try:
            b = (float(r["dram_read_MB"]) + float(r["dram_write_MB"])) * 1e6
        except ValueError:
            continue
        if b == b and not r["op"].startswith("void "):
            out[wl][r["op"]] = {"dram_bytes_per_pair": b / pairs, "pairs_in_capture": pairs, "us_under_ncu": float(r["us"])}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print({k: len(v) for k, v in out.items() if k != "_comment"})
