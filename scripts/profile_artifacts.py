"""Turn one gpu_full.sh run into the tracked artifacts: python scripts/profile_artifacts.py r01j
   gpurun_out/<tag>_prof.ncu-rep     -> profiles/<tag>_ncu_full_summary.md, profiles/traffic.json (dram bytes per launch)
   gpurun_out/<tag>_launches.csv     -> profiles/<tag>_launches_summary.md
   gpurun_out/bench_{default,reference}.json -> profiles/<tag>_bench_*.json"""
import collections, csv, io, json, shutil, subprocess, sys

tag = sys.argv[1]
rep = f"gpurun_out/{tag}_prof.ncu-rep"
with open(f"profiles/{tag}_ncu_full_summary.md", "w") as f:
    f.write(subprocess.run([sys.executable, "scripts/ncu_summary.py", rep], capture_output=True, text=True).stdout)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
names = {"hash_fwd_kernel": "hbr_hash_encode_fwd", "hash_bwd_kernel": "hbr_hash_encode_bwd", "mlp_fwd_tc_kernel": "hbr_mlp_fwd_tc",
         "mlp_bwd_tc_kernel": "hbr_mlp_bwd_tc", "composite_fwd_kernel": "hbr_composite_fwd", "composite_bwd_kernel": "hbr_composite_bwd"}
# the fused backward (scatter warps inside mlp_bwd_tc_kernel<.., SCAT = 11>) is the same kernel template: both keys get its traffic
ALIAS = {"hbr_mlp_bwd_tc": ["hbr_field_bwd_rays_tc"], "hbr_hash_encode_fwd": ["hbr_hash_encode_fwd_rays"]}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"_source": f"profiles/{tag}_ncu_full_summary.md (ncu --set full, one launch each, bench.py --steps 3 --warmup 3 --graph off, 524288 points)"}
for r in rows[2:]:
    for s, n in names.items():
        if s in r[idx["Kernel Name"]]:
            out[n] = int(round(sum(float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]]
                                   for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))))
for k, al in ALIAS.items():
    for a in al:
        if k in out:
            out[a] = out[k]
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
lines = [l for l in open(f"gpurun_out/{tag}_launches.csv").read().splitlines() if l.startswith('"')]
rd = csv.reader(lines)
h = next(rd)
ix = {x: i for i, x in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) < len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[ix["Metric Unit"]], 1)
    a = agg[r[ix["Kernel Name"]]]
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"profiles/{tag}_launches_summary.md", "w") as f:
    f.write(f"# {tag} launch list (ncu --metrics gpu__time_duration.sum --clock-control none; bench.py --steps 3 --warmup 3 --graph off)\n\n"
            "| kernel | launches | total ns | share |\n|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k[:70]} | {a[0]} | {int(a[1])} | {100 * a[1] / tot:.1f}% |\n")
for n in ("default", "reference"):
    shutil.copy(f"gpurun_out/bench_{n}.json", f"profiles/{tag}_bench_{n}.json")
print(json.dumps(out))
