"""Top stall-sample SASS lines of an `ncu --page source --csv` dump:  python scripts/ncu_hot_sass.py dump.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(enumerate(body), key=lambda ir: -int(ir[1][ix["# Samples"]] or 0))[:n]
for i, r in sorted(top):
    st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0)}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {int(r[ix['Instructions Executed']] or 0):9d}  {r[ix['Source']][:90]:90s} {st}")
