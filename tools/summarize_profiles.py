"""Turns ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
  python tools/summarize_profiles.py launches <launches.csv> <out.md> [skip_launches]
  python tools/summarize_profiles.py full <file.ncu-rep> <out.csv>
The launch list is the `--metrics gpu__time_duration.sum --clock-control none` pass; per-launch times there are
cold-cache and serialised, so only each kernel's SHARE of the step is comparable with bench.py's live numbers."""
import csv
import collections
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.replace("void ", "").strip()


def launches(path, out, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    rows = rows[skip:]
    agg = collections.OrderedDict()
    for n, t, g, b in rows:
        a = agg.setdefault(n, [0, 0.0, g, b])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"launches: {len(rows)}  total {total / 1e3:.1f} us (ncu serialised, cold cache)\n\n")
        f.write("| kernel | launches | total us | avg us | share | grid | block |\n|---|---|---|---|---|---|---|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {n} | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.2f} | {a[1] / total:.3f} | {a[2]} | {a[3]} |\n")


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units = rd[0], rd[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        cols = [k for k in KEYS if k in idx]
        w.writerow(["kernel"] + [f"{c} [{units[idx[c]]}]" for c in cols])
        for r in rd[2:]:
            w.writerow([short(r[idx["Kernel Name"]])] + [r[idx[c]] for c in cols])


def traffic(rep, out):
    """DRAM bytes (read + write) per launch of the kernels bench.py reports, keyed by bench.py's kernel names."""
    import json
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units = rd[0], rd[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = collections.defaultdict(list)
    for r in rd[2:]:
        name = r[idx["Kernel Name"]]
        b = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]] + \
            float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
        key = None
        if "output_row_kernel" in name or "output_tile_kernel" in name: key = "output_pass"
        elif "sparse_z" in name: key = "sparse_z_bias_act"
        elif "sparse_wgrad_light" in name: key = "sparse_wgrad_update:light"
        elif "sparse_wgrad_heavy" in name: key = "sparse_wgrad_update:heavy"
        elif "transpose_scatter" in name: key = "sparse_transpose"
        elif re.search(r"gemm_tc(_ts|_reg)?_kernel<", name):
            m = re.search(r"gemm_tc(?:_ts|_reg)?_kernel<(?:\(bool\))?([01]), (?:\(bool\))?([01])", name)
            key = {("0", "1"): "gemm_fwd_bias_act_tc", ("1", "1"): "gemm_dw_tc", ("0", "0"): "gemm_dx_tc"}.get(m.groups()) if m else None
        elif "update_biases" in name: key = "update_biases:max"
        if key: per[key].append(b)
    res = {}
    for k, v in per.items():
        if k.endswith(":max"): res[k[:-4]] = max(v)
        else: res[k] = sum(v) / len(v)
    if "sparse_wgrad_update:light" in res:
        res["sparse_wgrad_update"] = res.pop("sparse_wgrad_update:light") + res.pop("sparse_wgrad_update:heavy", 0.0)
    res = {k: int(v) for k, v in res.items()}
    res["_source"] = f"ncu --set full --clock-control none, {rep.split('/')[-1]}, bench.py workload c2 on one B200; dram__bytes_read.sum + dram__bytes_write.sum per launch"
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3])
