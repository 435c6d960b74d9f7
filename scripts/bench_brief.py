"""One-screen digest of a bench.py JSON line."""
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("no JSON line in", sys.argv[1], e); sys.exit(0)
keys = ["metric", "value", "n_gpus", "ms_per_step", "scaling", "gpu_launches", "ms_per_step_eager", "ms_per_step_cuda_graph"]
print({k: (round(d[k], 4) if isinstance(d.get(k), float) else d.get(k)) for k in keys})
print(" e2e", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.get("e2e", {}).items() if k != "api"})
r = d.get("roofline", {})
print(" roofline", {k: (round(r[k], 4) if isinstance(r.get(k), float) else r.get(k)) for k in ("frac", "us_per_launch", "us_per_launch_eager", "us_per_launch_cuda_graph", "achieved", "in_timed_step")})
for e in d.get("roofline_step", []):
    print("  step kernel", e["kernel"][:50], {k: round(e[k], 4) for k in ("us_per_launch", "share_of_step", "frac", "hbm_frac") if k in e})
for k in ("dp_parity", "strong_scaling", "cpu_baseline", "reference_gpu", "clocks"):
    if k in d:
        v = d[k]
        print(" " + k, {kk: ((float(f"{vv:.4g}")) if isinstance(vv, float) else (vv if not isinstance(vv, str) else vv[:60])) for kk, vv in v.items()} if isinstance(v, dict) else v)
if "configs3" in d:
    c = d["configs3"]
    print(" configs3", {k: (round(c[k], 3) if isinstance(c.get(k), float) else c.get(k)) for k in ("value", "ms_per_step", "error")}, "roofline frac", c.get("roofline", {}).get("frac"))
if "train_epoch" in d:
    t = d["train_epoch"]
    print(" train_epoch", {k: (round(t[k], 3) if isinstance(t.get(k), float) else t.get(k)) for k in ("value", "ms_per_step", "error")}, "masked_l2", (t.get("masked_l2") or {}).get("ms_per_step"))
