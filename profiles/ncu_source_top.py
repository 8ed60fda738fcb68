"""Top stalled SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-id :::N`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
S = ix["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S]) for r in data)
print(rows[0][1][:120])
print("total samples", tot, "instructions", len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[S]))[:n]:
    why = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    why = " ".join(f"{w}:{c}" for c, w in why if c)
    print(f"{int(r[S]):7d} {100 * int(r[S]) / tot:5.1f}%  exec={r[ix['Instructions Executed']]:>8s}  {r[ix['Source']].strip()[:70]:70s} {why}")
