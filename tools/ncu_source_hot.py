"""Hot SASS instructions of one kernel from `ncu -i rep --page source --csv --launch-count 1 [--launch-skip n]` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
ix = {h: i for i, h in enumerate(hdr)}
S = lambda r: int(r[ix["# Samples"]] or 0)
tot = sum(S(r) for r in data)
print(rows[0][1][:100], "| instructions", len(data), "samples", tot)
for n, r in enumerate(data):
    r.append(n)
for r in sorted(data, key=lambda r: -S(r))[:n_top]:
    reasons = {k[6:]: int(r[ix[k]]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and r[ix[k]].isdigit() and int(r[ix[k]]) > 0}
    main = sorted(reasons.items(), key=lambda x: -x[1])[:2]
    print(f"{100 * S(r) / tot:5.1f}% #{r[-1]:5d} exec {r[ix['Instructions Executed']]:>9} {r[ix['Source']].strip()[:72]:72s} {main}")
loc = [r for r in data if "LDL" in r[ix["Source"]] or "STL" in r[ix["Source"]]]
print("local-memory instructions:", len(loc), "executed", sum(int(r[ix["Instructions Executed"]]) for r in loc), "samples", sum(S(r) for r in loc))
print("total executed", sum(int(r[ix["Instructions Executed"]]) for r in data))
