#!/usr/bin/env python
"""Regenerates profiles/<name>.md (launch list of ONE step: per-kernel time, share, DRAM bytes) and the trunk's entry of
profiles/trunk_traffic.json (what bench.py reports as roofline.traffic) from an ncu CSV launch list:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c N --csv \\
        --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --sustain-s 0      (scripts/gpu_check.sh)
    python scripts/trunk_traffic.py gpurun_out/launches.csv profiles/r02_launches_fp16_step.md fp16_b256_c2 [--bench-line FILE]

One step = the launches between two consecutive `rng_schedule_kernel` launches that are followed by a sampler launch
(the key schedule opens every update_critics call); the LAST complete step of the capture is used (graph replay).
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict, defaultdict

TRUNK = re.compile(r"stem2?_tc_kernel|conv3x3_tc_kernel|conv3x3_res_kernel|conv3x3s2_res_kernel|conv_tc_kernel|stem_prep_kernel|pool_finish_kernel|affine_relu_kernel|"
                   r"block_combine_kernel|gn_finalize_kernel|maxpool_affine_kernel|conv_igemm_f32|groupnorm_f32|maxpool3x3s2_f32|FillFunctor<float>")


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"serl::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:70]


def load(path):
    rows = OrderedDict()
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        d = rows.setdefault(k, {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1e3)
        elif m.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["rd" if "read" in m else "wr"] = v * mult
    return list(rows.values())


def last_step(launches):
    # (the critic step's noise / dropout fills are prefetched right after the key schedule: the sampler follows within a few launches)
    starts = [i for i, l in enumerate(launches) if "rng_schedule_kernel" in l["name"]
              and any("sample_" in x["name"] for x in launches[i + 1:i + 6])]
    if len(starts) < 2:
        raise SystemExit(f"need two step boundaries, found {len(starts)}")
    return launches[starts[-2]:starts[-1]]


def main():
    src, out_md, key = sys.argv[1:4]
    bench = None
    if "--bench-line" in sys.argv:
        with open(sys.argv[sys.argv.index("--bench-line") + 1]) as f:
            for ln in f:
                if ln.startswith("{"):
                    bench = json.loads(ln)
    step = last_step(load(src))
    tot = sum(l.get("us", 0) for l in step)
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for l in step:
        a = agg[short(l["name"])]
        a[0] += 1; a[1] += l.get("us", 0); a[2] += l.get("rd", 0); a[3] += l.get("wr", 0)
    trunk = [l for l in step if TRUNK.search(l["name"])]
    t_us = sum(l.get("us", 0) for l in trunk)
    t_bytes = sum(l.get("rd", 0) + l.get("wr", 0) for l in trunk)
    md = [f"# Launch list of one critic step ({key})", "",
          f"Source: `{src}` (ncu `gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum`, `--clock-control none`; launches are serialised "
          "and caches flushed between them, so durations are cold-cache and side-stream overlap is absent: SHARES carry over, sums do not).",
          f"Regenerate: `python scripts/trunk_traffic.py {src} {out_md} {key}`.", "",
          f"Step under ncu: {tot:.1f} us over {len(step)} launches.  Frozen-trunk kernels: {t_us:.1f} us = {100 * t_us / tot:.1f} % of the serialised step, "
          f"{len(trunk)} launches, DRAM traffic {t_bytes / 1e6:.0f} MB per step."]
    if bench:
        r = bench.get("roofline", {})
        md += ["", f"Bench line of the same build: {bench['ms_per_step']:.3f} ms/step = {bench['value']:.1f} steps/s, e2e {bench['e2e']['value']:.1f}; "
                   f"trunk {r.get('ms_per_step', float('nan')):.3f} ms = {r.get('achieved', float('nan')):.0f} TFLOP/s = {100 * r.get('frac', float('nan')):.1f} % of "
                   f"{r.get('peak')} ({r.get('peak_source')})."]
    md += ["", "| kernel | launches | us | share | DRAM read MB | DRAM write MB |", "|---|---:|---:|---:|---:|---:|"]
    for name, (n, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{name}` | {n} | {us:.1f} | {100 * us / tot:.1f} % | {rd / 1e6:.1f} | {wr / 1e6:.1f} |")
    md += ["", "## In launch order", "", "| # | kernel | grid | us | DRAM MB |", "|---:|---|---|---:|---:|"]
    for i, l in enumerate(step):
        md.append(f"| {i} | `{short(l['name'])}` | {l['grid']} | {l.get('us', 0):.1f} | {(l.get('rd', 0) + l.get('wr', 0)) / 1e6:.1f} |")
    with open(out_md, "w") as f:
        f.write("\n".join(md) + "\n")
    tj = os.path.join(os.path.dirname(os.path.abspath(out_md)), "trunk_traffic.json")
    data = json.load(open(tj)) if os.path.exists(tj) else {}
    data[key] = {"dram_bytes_per_step": int(t_bytes), "launches": len(trunk), "trunk_us_serialised": round(t_us, 1),
                 "source": f"{os.path.relpath(out_md, os.path.dirname(os.path.dirname(os.path.abspath(out_md))))} (ncu launch list of one step, cold caches per launch)"}
    json.dump(data, open(tj, "w"), indent=1)
    print(f"{out_md}: {len(step)} launches, {tot:.1f} us; trunk {t_us:.1f} us, {t_bytes / 1e6:.0f} MB")


if __name__ == "__main__":
    main()
