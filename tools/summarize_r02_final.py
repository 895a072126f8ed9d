"""Launch list of the final round-2 benchmark stage (fused RK plan) from gpurun_out/r02ag_launches_cfg5.csv (tools/gpu/r02ag.sh):
    python tools/summarize_r02_final.py
writes profiles/r02_final_launches_cfg5.{csv,md}.  The stage boundary is the one lincomb launch per stage (the RK accumulator a u_tn + b u)."""
import collections, csv, os, re, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", "r02ag_launches_cfg5.csv")
dst = os.path.join(ROOT, "profiles", "r02_final_launches_cfg5")
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
h = rows[0]; ix = {c: i for i, c in enumerate(h)}
data = rows[1:]
names = [r[ix["Kernel Name"]] for r in data]
marks = [i for i, n in enumerate(names) if "lincomb_kernel" in n]
seg = data[marks[-2]:marks[-1]] if len(marks) >= 2 else data          # the last complete stage: from its accumulator initialisation to the next one
agg = collections.OrderedDict()
for r in seg:
    n = re.sub(r"\(amdg::.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[ix["Metric Value"]])
tot = sum(v[1] for v in agg.values())
shutil.copy(src, dst + ".csv")
with open(dst + ".md", "w") as f:
    f.write("# r02 (final) -- ncu launch list of the default benchmark stage (cfg5: d=6 k=1 m=2 NMAX=7, one nonlinear RK3SSP stage with the RK combination in the sweep epilogues, 1 x B200)\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02ag_launches_cfg5.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph`\n")
    f.write("(raw list: `r02_final_launches_cfg5.csv`; per-launch times are cold-cache and serialised under the profiler -- compare SHARES).\n\n")
    f.write("Last complete stage of the run: %d launches, %.2f ms serialised (no `rk_stage_kernel`, one `lincomb_kernel`: the plan of `stage.StagePlan(fuse_rk=True)`).\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % (len(seg), tot / 1e6))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (n, c, t / 1e3, 100 * t / tot))
    col32 = sum(t for n, (c, t) in agg.items() if "sweep_col_kernel<3, 2" in n)
    f.write("\n`sweep_col_kernel<3,2,*>` (the kernel `roofline` in bench.py reports): %.1f %% of the stage.\n" % (100 * col32 / tot))
print(open(dst + ".md").read())
