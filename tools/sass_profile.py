"""Summarise an `ncu --page source --csv` dump (SASS view): executed warp instructions and stall samples by opcode and by code region.
    python tools/sass_profile.py gpurun_out/x_source.csv [--regions N] [--top N]"""
import csv, sys, collections, re

def load(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr): continue
        src = r[ix["Source"]].strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0] if src else ""
        out.append(dict(src=src, op=op, n=int(r[ix["Instructions Executed"]] or 0), samp=int(r[ix["# Samples"]] or 0),
                        stalls={k: int(r[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}))
    return out

def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    I = load(path)
    tot = sum(i["n"] for i in I); ts = sum(i["samp"] for i in I)
    print("SASS instructions %d, executed warp instructions %d, samples %d" % (len(I), tot, ts))
    by = collections.Counter(); bys = collections.Counter()
    for i in I:
        k = i["op"].split(".")[0]; by[k] += i["n"]; bys[k] += i["samp"]
    print("-- by opcode (executed %, sample %)")
    for k, v in by.most_common(top): print("  %-10s %9d %5.1f%%  samples %5.1f%%" % (k, v, 100.0 * v / tot, 100.0 * bys[k] / max(ts, 1)))
    st = collections.Counter()
    for i in I:
        for k, v in i["stalls"].items(): st[k] += v
    print("-- stall reasons (all samples)")
    for k, v in st.most_common(10): print("  %-26s %7d %5.1f%%" % (k, v, 100.0 * v / max(ts, 1)))
    # regions: split at barriers / big changes of the execution count
    print("-- regions (consecutive instructions with similar execution counts)")
    reg = []; cur = None
    for idx, i in enumerate(I):
        n = i["n"]
        if cur is None or not (0.5 * cur["ref"] <= n <= 2.0 * cur["ref"] or abs(n - cur["ref"]) < 64):
            cur = dict(start=idx, ref=max(n, 1), n=0, samp=0, cnt=0, ops=collections.Counter()); reg.append(cur)
        cur["n"] += n; cur["samp"] += i["samp"]; cur["cnt"] += 1; cur["ops"][i["op"].split(".")[0]] += n; cur["end"] = idx
    for r in reg:
        if r["n"] < 0.004 * tot and r["samp"] < 0.004 * ts: continue
        ops = ", ".join("%s %d" % (k, v // max(r["ref"], 1)) for k, v in r["ops"].most_common(6))
        print("  [%4d-%4d] per-instr %8d  executed %5.1f%%  samples %5.1f%%  | %s" % (r["start"], r["end"], r["ref"], 100.0 * r["n"] / tot, 100.0 * r["samp"] / max(ts, 1), ops))
    if "--list" in sys.argv:
        for idx, i in enumerate(I): print("%4d %9d %5d  %s" % (idx, i["n"], i["samp"], i["src"]))

main()
