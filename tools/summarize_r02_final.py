"""Launch list of the final round-2 benchmark stage (fused RK plan) from gpurun_out/r02au_launches_cfg5.csv (tools/gpu/r02au.sh):
    python tools/summarize_r02_final.py
writes profiles/r02_final_launches_cfg5.{csv,md}.  The stage boundary is the one lincomb launch per stage (the RK accumulator a u_tn + b u)."""
import collections, csv, os, re, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", "r02au_launches_cfg5.csv")
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
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02au_launches_cfg5.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph`\n")
    f.write("(raw list: `r02_final_launches_cfg5.csv`; per-launch times are cold-cache and serialised under the profiler -- compare SHARES).\n\n")
    f.write("Last complete stage of the run: %d launches, %.2f ms serialised (no `rk_stage_kernel`, one `lincomb_kernel`: the plan of `stage.StagePlan(fuse_rk=True)`).\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % (len(seg), tot / 1e6))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (n, c, t / 1e3, 100 * t / tot))
    col32 = sum(t for n, (c, t) in agg.items() if "sweep_col_kernel<3, 2" in n)
    f.write("\n`sweep_col_kernel<3,2,*>` (the kernel `roofline` in bench.py reports): %.1f %% of the stage.\n" % (100 * col32 / tot))
print(open(dst + ".md").read())


# ---- full capture of the roofline kernel and DRAM bytes of the 3 -> 2 launches of a stage (final tree): profiles/r02_final_roofline_kernel_ncu.md, r02_traffic.json
import json, subprocess, sys
sys.path.insert(0, os.path.join(ROOT, "tools"))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
def to_bytes(v, unit):
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
rows = list(csv.reader(open(os.path.join(G, "r02au_col_raw.csv"))))
h, u, data = rows[0], rows[1], rows[2:]
kn = h.index("Kernel Name")
sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
with open(os.path.join(P, "r02_final_roofline_kernel_ncu.md"), "w") as f:
    f.write("# r02 (final) -- `ncu --set full --clock-control none` capture of the roofline kernel on the final tree (1 x B200, call AU)\n\n")
    f.write("cfg5 (d=6, k=1, m=2, NMAX=7): full 3 -> 2 sweep along dimension 0 (729 -> 486 doubles per element, 157 MB algorithmic), `tools/sweep_time.py --workload cfg5 --kernel 0 --lus 2 --dims 0 --shapes b>a`\n\n")
    f.write("| metric | unit | " + " | ".join("launch %d" % (i + 1) for i in range(len(data))) + " |\n|---|---|" + "---|" * len(data) + "\n")
    f.write("| kernel | | " + " | ".join("`" + re.sub(r"\(amdg::.*", "", r[kn]).replace("void amdg::", "") + "`" for r in data) + " |\n")
    for w in WANT:
        if w not in h: continue
        i = h.index(w)
        def num(x):
            try: return float(x.replace(",", ""))
            except ValueError: return None
        f.write("| %s | %s | " % (w, u[i]) + " | ".join(("%.4g" % num(r[i])) if num(r[i]) is not None else "-" for r in data) + " |\n")
# stage DRAM bytes
rows = [r for r in csv.reader(open(os.path.join(G, "r02au_stage_dram.csv"))) if len(r) > 10]
h = rows[0]; ix = {c: i for i, c in enumerate(h)}
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    d[r[ix["Metric Name"]]] = to_bytes(v, r[ix["Metric Unit"]]) if "bytes" in r[ix["Metric Name"]] else v
L = list(per.values())
per_stage = 264                              # sweep_col launches per stage
sel = [d for d in L[-per_stage:] if re.search(r"sweep_col_kernel<3, 2", d["name"])]
tot = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in sel)
tr = json.load(open(os.path.join(P, "r02_traffic.json")))
tr["sweep_col_kernel<3,2>"] = {"dram_bytes_per_launch": tot / len(sel), "launches": len(sel), "git": sha,
                                "source": "gpurun_out/r02au_stage_dram.csv: all 3 -> 2 sweep launches of one cfg5 stage on the final tree (tools/gpu/r02au.sh)"}
json.dump(tr, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
with open(os.path.join(P, "r02_final_roofline_kernel_ncu.md"), "a") as f:
    f.write("\n## the 3 -> 2 sweep launches of one cfg5 stage (what `roofline` in bench.py times)\n\n%d launches of `sweep_col_kernel<3,2,*>`: DRAM read + written %.2f GB = %.1f MB per launch "
            "(algorithmic B_sweep of the same launches: 170.9 MB per launch), serialised time %.2f ms.\n" % (len(sel), tot / 1e9, tot / len(sel) / 1e6, sum(d["gpu__time_duration.sum"] for d in sel) / 1e6))
print(open(os.path.join(P, "r02_final_roofline_kernel_ncu.md")).read())
