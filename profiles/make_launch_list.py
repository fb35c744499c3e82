#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py --quick --steps 1` -> one train step grouped by
kernel and grid.

    python profiles/make_launch_list.py gpurun_out/r02_launches_step.csv > profiles/r02_launches_step.txt

A step starts at prompt_linear_fwd_kernel (the first launch of the fused step); the last COMPLETE step of the capture is
taken (every step of the run launches the same kernels).  Times under ncu are cold-cache and serialised: shares, not absolutes.
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    src = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        name = re.sub(r"^void ", "", r[ix["Kernel Name"]])
        name = re.sub(r"\(.*", "", name).replace("mudpt::", "")
        name = name.replace("(bool)", "").replace("(int)", "")
        launches.append((name, r[ix["Grid Size"]] if "Grid Size" in ix else "", us))
    starts = [i for i, l in enumerate(launches) if "prompt_linear_fwd" in l[0]]
    if len(starts) >= 2:
        step = launches[starts[-2]:starts[-1]]
    else:
        step = launches
    total = sum(l[2] for l in step)
    groups = OrderedDict()
    for name, grid, us in step:
        g = groups.setdefault((name, grid), [0, 0.0])
        g[0] += 1
        g[1] += us
    print("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --quick --steps 1   (BASELINE config 2, 1 GPU)")
    print(f"# last complete train step: {len(step)} launches, summed kernel time {total / 1e3:.3f} ms (cold-cache, serialised: shares, not "
          "absolutes); grouped by kernel and grid")
    for (name, grid), (n, us) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:58]:58s} grid {grid:16s} {n:4d} launches {us:10.1f} us total {us / n:8.1f} us avg {100 * us / total:5.1f}%")
    classes = OrderedDict([("GEMM (tcgen05)", "gemm_tn_tcgen05"), ("attention", "attn_"), ("LayerNorm", "ln_"), ("splice / gather / scatter", "splice|gather_rows|scatter_rows"),
                           ("heads", "head_|sgemm|splitk|gather_ln|scatter_ln|l2norm|ce_rows|normalize"), ("prompt algebra", "prompt_"), ("SGD", "sgd_")])
    print("# by class:")
    for label, pat in classes.items():
        t = sum(us for (name, _), (_, us) in groups.items() if re.search(pat, name))
        print(f"#   {label:28s} {t / 1e3:8.3f} ms  {100 * t / total:5.1f}%")


if __name__ == "__main__":
    main()
