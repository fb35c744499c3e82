#!/usr/bin/env python
"""DRAM traffic of the dominant kernel (the tcgen05 GEMM) over EVERY launch of the train step, for bench.py's
`roofline.traffic`.

On the GPU box (after the same command exited 0 without ncu):

    python bench.py --quick --steps 1 > gpurun_out/q.log 2>&1 && \
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:gemm_tn_tcgen05 -c 4000 --csv --log-file gpurun_out/r02_step_traffic.csv python bench.py --quick --steps 1

Here (no GPU needed):

    python profiles/make_step_traffic.py gpurun_out/r02_step_traffic.csv profiles/r02_step_traffic.json

Every step of `bench.py --quick` (warm-ups, timed, instrumented) launches the same GEMMs, so the mean over all captured
launches is the mean over the launches of one step -- the population bench.py averages the algorithmic bytes over.
The output names the build (sha256 of the GEMM-relevant sources: api.cu, common.cuh, gemm.cu, gemm.h) it was captured from; bench.py reports `traffic: null` for any other build.
"""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# the sources that decide what the GEMM launches of a step are and how they move data: the kernel, its helpers, and the
# tower orchestration that chooses the launches (attention / head / row kernels do not change GEMM traffic)
GEMM_SOURCES = ("api.cu", "common.cuh", "gemm.cu", "gemm.h")


def csrc_digest():
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "mudpt_b200", "csrc")
    for f in GEMM_SOURCES:
        h.update(f.encode())
        h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        key = (r[ix["ID"]], r[ix["Kernel Name"]])
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
                 "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1.0)
        per.setdefault(key, {})[r[ix["Metric Name"]]] = val * scale
    by_kernel = {}
    tot_b, tot_n, tot_us = 0.0, 0, 0.0
    for (_, name), m in per.items():
        b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        short = re.sub(r"\(.*", "", name).replace("void mudpt::", "")
        e = by_kernel.setdefault(short, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
        e["launches"] += 1
        e["dram_bytes"] += b
        e["us"] += m.get("gpu__time_duration.sum", 0.0)
        tot_b += b
        tot_n += 1
        tot_us += m.get("gpu__time_duration.sum", 0.0)
    for e in by_kernel.values():
        e["avg_dram_MB"] = round(e["dram_bytes"] / e["launches"] / 1e6, 2)
        e["avg_us"] = round(e["us"] / e["launches"], 2)
        del e["dram_bytes"], e["us"]
    res = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of every gemm_tn_tcgen05_kernel launch of `bench.py --quick --steps 1` "
                   "(BASELINE config 2 shapes, 1 GPU) under ncu --clock-control none; all steps launch the same GEMMs",
           "captured_launches": tot_n, "avg_traffic_bytes_per_gemm_launch": tot_b / max(tot_n, 1),
           "avg_us_per_gemm_launch_under_ncu": tot_us / max(tot_n, 1), "by_kernel": by_kernel, "csrc_sha256": csrc_digest()}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "by_kernel"}, indent=1))


if __name__ == "__main__":
    main()
