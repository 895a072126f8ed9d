import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        r=json.loads(l); print("ms/step %.4f  sweep us %.2f  frac %.3f"%(r["ms_per_step"], r["roofline"]["us_per_launch"], r["roofline"]["frac"]))
