"""Tracked profile summaries of round 2 from the ncu pages brought back under gpurun_out/ (tools/gpu/r02u.sh, r02v.sh):
    python tools/summarize_r02.py
writes profiles/r02_launches_cfg5.{csv,md}, profiles/r02_roofline_kernels_ncu.md and profiles/r02_traffic.json (read by bench.py)."""
import collections, csv, json, os, re, shutil, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launch_list(src, dst_base, title, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]; ix = {c: i for i, c in enumerate(h)}
    data = rows[1:]
    names = [r[ix["Kernel Name"]] for r in data]
    rk = [i for i, n in enumerate(names) if "rk_stage" in n]
    seg = data[rk[-2] + 1:rk[-1] + 1] if len(rk) >= 2 else data          # the last complete stage
    agg = collections.OrderedDict()
    for r in seg:
        n = re.sub(r"\(amdg::.*", "", r[ix["Kernel Name"]]).replace("void ", "")
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[ix["Metric Value"]])
    tot = sum(v[1] for v in agg.values())
    shutil.copy(src, dst_base + ".csv")
    with open(dst_base + ".md", "w") as f:
        f.write("# %s\n\nCommand: `%s`\n(raw list: `%s.csv`; per-launch times are cold-cache and serialised under the profiler -- compare SHARES).\n\n" % (title, cmd, os.path.basename(dst_base)))
        f.write("Last complete stage of the run: %d launches, %.2f ms serialised.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % (len(seg), tot / 1e6))
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (n, c, t / 1e3, 100 * t / tot))
    return agg, tot


WANT = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def raw_table(path):
    rows = list(csv.reader(open(path)))
    h, u, data = rows[0], rows[1], rows[2:]
    kn = h.index("Kernel Name")
    out = ["| metric | unit | " + " | ".join("launch %d" % (i + 1) for i in range(len(data))) + " |", "|---|---|" + "---|" * len(data)]
    out.append("| kernel | | " + " | ".join("`" + re.sub(r"\(amdg::.*", "", r[kn]).replace("void amdg::", "") + "`" for r in data) + " |")
    vals = {}
    for w in WANT:
        if w not in h: continue
        i = h.index(w)
        def num(x):
            try: return float(x.replace(",", ""))
            except ValueError: return None
        vals[w] = [num(r[i]) for r in data]
        out.append("| %s | %s | " % (w, u[i]) + " | ".join(("%.4g" % v) if v is not None else "-" for v in vals[w]) + " |")
    return "\n".join(out), vals, u, h


def to_bytes(v, unit):
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main():
    os.makedirs(P, exist_ok=True)
    launch_list(os.path.join(G, "r02u_launches_cfg5.csv"), os.path.join(P, "r02_launches_cfg5"),
                "r02 -- ncu launch list of the default benchmark stage (cfg5: d=6 k=1 m=2 NMAX=7, one nonlinear RK3SSP stage, 1 x B200)",
                "ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02u_launches_cfg5.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph")
    traffic = {}
    sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    with open(os.path.join(P, "r02_roofline_kernels_ncu.md"), "w") as f:
        f.write("# r02 -- `ncu --set full --clock-control none` captures of the two roofline kernels (1 x B200)\n\n")
        for key, path, what in (("sweep_col_kernel<3,2>", "r02u_col_raw.csv", "cfg5 (d=6, k=1, m=2, NMAX=7): `sweep_col_kernel<3,2,4>`, full 3 -> 2 sweep along dimension 0 (729 -> 486 doubles per element, 157 MB algorithmic), `tools/sweep_time.py --workload cfg5 --kernel 0 --lus 2 --dims 0 --shapes b>a`"),
                                ("sweep_tc_kernel<4,4>", "r02u_tc_raw.csv", "cfg2 (d=4, k=3, m=3, NMAX=8): `sweep_tc_kernel<4,4,2>`, full 4 -> 4 sweep along dimension 1 (43 MB algorithmic), `tools/sweep_time.py --workload cfg2 --kernel 0 --lus 2 --dims 1`")):
            p = os.path.join(G, path)
            if not os.path.exists(p): continue
            tab, vals, u, h = raw_table(p)
            f.write("## %s\n\n%s\n\n%s\n\n" % (key, what, tab))
            rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
            if rd and wr:
                ur, uw = u[h.index("dram__bytes_read.sum")], u[h.index("dram__bytes_write.sum")]
                per = [to_bytes(a, ur) + to_bytes(b, uw) for a, b in zip(rd, wr) if a is not None and b is not None]
                traffic[key] = {"dram_bytes_per_launch": sum(per) / len(per), "launches": len(per), "git": sha, "source": "gpurun_out/" + path + " (tools/gpu/r02u.sh)"}
                f.write("`dram__bytes_read + dram__bytes_write` per launch: %.1f MB (mean of %d launches; the written half of a launch is still in L2 when the kernel ends).\n\n" % (traffic[key]["dram_bytes_per_launch"] / 1e6, len(per)))
    # DRAM bytes of the launches the benchmark's roofline times: every sweep_col_kernel<3,2,*> launch of one stage (tools/gpu/r02v.sh)
    sd = os.path.join(G, "r02v_stage_dram.csv")
    if os.path.exists(sd):
        rows = [r for r in csv.reader(open(sd)) if len(r) > 10]
        h = rows[0]; ix = {c: i for i, c in enumerate(h)}
        per = collections.OrderedDict()
        for r in rows[1:]:
            d = per.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
            d[r[ix["Metric Name"]]] = to_bytes(float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]) if "bytes" in r[ix["Metric Name"]] else float(r[ix["Metric Value"]].replace(",", ""))
        L = list(per.values())
        per_stage = 264                                                   # sweep launches per stage (273 launches - 1 point-wise - 7 lincomb - 1 RK)
        last = L[-per_stage:]
        sel = [d for d in last if re.search(r"sweep_col_kernel<3, 2", d["name"])]
        if sel:
            tot = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in sel)
            traffic["sweep_col_kernel<3,2>"] = {"dram_bytes_per_launch": tot / len(sel), "launches": len(sel), "git": sha,
                                                "source": "gpurun_out/r02v_stage_dram.csv: all 3 -> 2 sweep launches of one cfg5 stage (tools/gpu/r02v.sh)"}
            with open(os.path.join(P, "r02_roofline_kernels_ncu.md"), "a") as f:
                f.write("## the 3 -> 2 sweep launches of one cfg5 stage (what `roofline` in bench.py times)\n\n"
                        "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:sweep_col` over `bench.py --no-graph --steps 1`: "
                        "%d launches of `sweep_col_kernel<3,2,*>`, DRAM read + written %.2f GB = %.1f MB per launch (algorithmic B_sweep of the same launches: 36.9 GB = 170.9 MB per launch; "
                        "the written halves of most launches are still in L2 when they end and are read back from there by the next level), serialised time %.2f ms.\n"
                        % (len(sel), tot / 1e9, tot / len(sel) / 1e6, sum(d["gpu__time_duration.sum"] for d in sel) / 1e6))
    json.dump(traffic, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


main()
