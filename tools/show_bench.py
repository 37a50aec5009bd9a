"""Prints value / ms per step / e2e and the per-kernel table of bench.py JSON lines:  python tools/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    t = open(f).read()
    i = t.find('{"metric')
    if i < 0:
        print(f, "no JSON line:", t[-300:])
        continue
    d = json.loads(t[i:].splitlines()[0])
    print(f, d.get("value"), "samples/s", d.get("ms_per_step"), "ms/step  e2e", d.get("e2e", {}).get("value"), " launches/step", d.get("gpu_launches", 0) / max(d.get("steps", 1), 1))
    tot = 0.0
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["share"]):
        print(f"    {k:34s} {v['calls_per_step']:4.1f} x {v['ms_per_call'] * 1e3:7.1f} us   {v.get('algorithmic_GBs', '')}")
        tot += v["calls_per_step"] * v["ms_per_call"] * 1e3
    print("    sum of kernels (profile pass, serialised):", round(tot, 1), "us")
    for side in ("c4", "c5"):
        if side in d:
            print("   ", side, {k: v for k, v in d[side].items() if k in ("value", "unit", "ms_per_step", "ms_per_call", "n_gpus", "error")})
