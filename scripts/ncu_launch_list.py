"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python scripts/ncu_launch_list.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ix = {h: i for i, h in enumerate(hdr)}
d = collections.defaultdict(list)
for r in rows[start:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    name = r[ix["Kernel Name"]].replace("dwb::", "").replace("void ", "")
    name = name.split("(")[0]
    d[(name, r[ix["Grid Size"]].replace(" ", ""))].append(v)
tot = sum(sum(v) for v in d.values())
n = sum(len(v) for v in d.values())
print(f"{n} launches, {tot / 1000:.2f} ms of kernel time (cold-cache, serialised under ncu)\n")
print("| kernel | grid | launches | avg us | total ms | share |\n|---|---|---:|---:|---:|---:|")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    if sum(v) / tot < 0.002:
        continue
    print(f"| `{k[0]}` | {k[1]} | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / 1000:.2f} | {100 * sum(v) / tot:.1f} % |")
by = collections.defaultdict(float)
for k, v in d.items():
    key = "tcgen05 GEMM (pair + single)" if "gemm_bf16" in k[0] else "tcgen05 attention fwd" if "attn_fwd_tc" in k[0] else \
        "tcgen05 attention bwd" if "attn_bwd" in k[0] else "LayerNorm fwd/bwd" if "layernorm" in k[0] else "other"
    by[key] += sum(v)
print("\n" + ", ".join(f"{k}: {100 * v / tot:.1f} %" for k, v in sorted(by.items(), key=lambda kv: -kv[1])))
