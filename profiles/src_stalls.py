"""Top stalled SASS instructions of a kernel from `ncu -i X.ncu-rep --page source --csv` (needs -lineinfo / --import-source).
    python profiles/src_stalls.py gpurun_out/dgpw_src.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(rows[0][1])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h[6:]: sum(int(r[ix[h]]) for r in data) for h in stalls}
print("samples", tot, "| SASS instructions", len(data), "| warp instructions executed", sum(int(r[ix["Instructions Executed"]]) for r in data))
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
    st = sorted(((h[6:], int(r[ix[h]])) for h in stalls if int(r[ix[h]]) > 0), key=lambda kv: -kv[1])[:3]
    print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(8), r[ix["Source"]].strip()[:72].ljust(72), st)
