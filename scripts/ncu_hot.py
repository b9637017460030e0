"""Top stall sites of one kernel from an .ncu-rep source page: python scripts/ncu_hot.py rep regex [N]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
lines = raw.splitlines()
# first kernel only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def f(x):
    try: return float(x)
    except: return 0.0
tot = sum(f(r[idx["# Samples"]]) for r in rows[1:])
print("total samples", tot)
agg = {h: sum(f(r[idx[h]]) for r in rows[1:]) for h in stall_cols}
print("stall totals:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0})
top = sorted(rows[1:], key=lambda r: -f(r[idx["# Samples"]]))[:N]
for r in top:
    st = {h[6:]: int(f(r[idx[h]])) for h in stall_cols if f(r[idx[h]]) > 0}
    print(f"{r[idx['Address']][-5:]} {f(r[idx['# Samples']]):7.0f} {100*f(r[idx['# Samples']])/tot:5.1f}%  {r[idx['Source']][:70]:70s} {st}")
