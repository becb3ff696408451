"""Summarise an `ncu --page source --csv` dump: top-N SASS lines by stall samples with the dominant stall reason."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
isamp = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "not_issued" not in h.lower()]
if not stall_cols:
    stall_cols = [i for i, h in enumerate(hdr) if h.lower().startswith("stall")]
data = []
for k, r in enumerate(rows[2:]):
    try:
        s = int(r[isamp])
    except Exception:
        continue
    data.append((s, k, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "stall columns:", [hdr[i] for i in stall_cols][:30])
for s, k, r in sorted(data, reverse=True)[:n]:
    reasons = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{s:7d} {100*s/tot:5.1f}%  line {k:5d}  {r[1].strip()[:70]:70s} {reasons}")
