"""Turn the ncu CSV pages brought back by tools/ncu_capture.sh (gpurun_out/) into the tracked summaries under profiles/:
    python tools/summarize_profiles.py r01"""
import collections, csv, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list
src = os.path.join(G, R + "_launches_bench.csv")
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    agg[r[kn]][0] += 1; agg[r[kn]][1] += v
tot = sum(v[1] for v in agg.values())
shutil.copy(src, os.path.join(P, R + "_launches_bench.csv"))
with open(os.path.join(P, R + "_launches_bench.md"), "w") as f:
    f.write("# %s — ncu launch list of `python bench.py --steps 2 --warmup 3 --no-graph` (cfg2 round trip, 1 x B200)\n\n" % R)
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/%s_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-graph`\n" % R)
    f.write("(raw list: `profiles/%s_launches_bench.csv`; per-launch times are cold-cache and serialised under the profiler — compare SHARES).\n\n" % R)
    f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (n[:150], c, t / 1e3, 100 * t / tot))
    f.write("\nThe 1D sweep kernel (`sweep_tc_kernel`, csrc/kernels_tc.cu) is the dominant kernel: it is what `roofline` in bench.py measures.\n")

# ---- full captures
WANT = ["launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def table(path, pick):
    rows = list(csv.reader(open(path)))
    h, u, data = rows[0], rows[1], rows[2:]
    data = [data[i] for i in pick if i < len(data)]
    out = ["| metric | unit | " + " | ".join("launch %d" % (i + 1) for i in range(len(data))) + " |", "|---|---|" + "---|" * len(data)]
    kn = h.index("Kernel Name")
    out.append("| kernel | | " + " | ".join(r[kn].split("(")[0][-40:] for r in data) + " |")
    for w in WANT:
        idx = [i for i, c in enumerate(h) if c == w]
        if not idx: continue
        i = idx[0]
        def fmt(x):
            try: return "%.4g" % float(x.replace(",", ""))
            except ValueError: return x
        out.append("| %s | %s | " % (w, u[i]) + " | ".join(fmt(r[i]) for r in data) + " |")
    return "\n".join(out), data, h

with open(os.path.join(P, R + "_sweep_tc_ncu.md"), "w") as f:
    f.write("# %s — `ncu --set full` captures of the dominant kernel (sweep_tc_kernel<4,4,*>), cfg2 size (d=4, k=3, NMAX=8), 1 x B200\n\n" % R)
    f.write("## single-job sweeps (tools/prof_sweep.py: full / L / U sweeps along dims 0..3; source resident in L2 between launches, so `dram__bytes_read` is the compulsory read of one 21.5 MB vector and the written half stays in L2)\n\n")
    f.write("Command: `ncu --set full --clock-control none --import-source on -k regex:sweep_tc -s 12 -c 12 -o /tmp/%s_single python tools/prof_sweep.py sweeps`\n\n" % R)
    t, data, h = table(os.path.join(G, R + "_sweep_tc_single_raw.csv"), range(0, 12))
    f.write(t + "\n\n")
    dr, dw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    f.write("## batched launches inside the bench step (shared-prefix schedule: 1, 2, 4, 8 jobs per launch)\n\n")
    f.write("Command: `ncu --set full --clock-control none -k regex:sweep_tc -s 150 -c 10 -o /tmp/%s_step python bench.py --steps 1 --warmup 3 --no-graph`\n\n" % R)
    t2, _, _ = table(os.path.join(G, R + "_sweep_tc_step_raw.csv"), range(0, 10))
    f.write(t2 + "\n")
print("written", P)
