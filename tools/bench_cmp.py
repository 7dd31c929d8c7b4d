"""Per-layer step times of two bench.py JSON lines side by side (A/B runs on one box):
    python tools/bench_cmp.py a.json b.json"""
import json,sys
def load(p):
    return json.loads([l for l in open(p) if l.startswith("{")][-1])
a=load(sys.argv[1]); b=load(sys.argv[2])
print("value", a["value"], b["value"], "ms", a["ms_per_step"], b["ms_per_step"])
for k,v in sorted(a["breakdown"].items(), key=lambda kv:-kv[1]["ms_per_step"]):
    w=b["breakdown"].get(k,{"ms_per_step":0})
    print(f"  {k:18s} {v['ms_per_step']:8.2f} {w['ms_per_step']:8.2f}")
