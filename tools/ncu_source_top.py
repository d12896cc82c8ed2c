"""Top stall sites of one kernel from `ncu -i rep --page source --csv` (SASS view).
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME | python tools/ncu_source_top.py [N]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
# several kernels may be concatenated: take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
col = {h: i for i, h in enumerate(hdr)}
samp = col["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[samp] or 0) for r in body)
print("total samples", tot, " instructions", len(body))
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(body, key=lambda r: -int(r[samp] or 0))[:n]:
    st = {s[6:]: int(r[col[s]] or 0) for s in stalls if int(r[col[s]] or 0)}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%6d %5.1f%%  %-70s %s" % (int(r[samp]), 100.0 * int(r[samp]) / max(tot, 1), r[col["Source"]].strip()[:70], top))
