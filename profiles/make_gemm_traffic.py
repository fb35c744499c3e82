"""Summarise an `ncu --set full` capture of tests/gpu_gemm_prof.py (text-tower GEMMs with their fused
epilogues, BASELINE config 2 shapes) into the per-launch DRAM traffic table bench.py reads
(roofline.traffic).  Run here (no GPU needed):

    python profiles/make_gemm_traffic.py gpurun_out/prof_gemm_all.ncu-rep profiles/r01_gemm_traffic.json
"""
import csv
import json
import subprocess
import sys

TAGS = ["qkv", "out", "c_fc", "c_proj", "d_c_proj", "d_c_fc", "d_out", "d_qkv"]
NAMES = {"qkv": "QKV in-proj", "out": "out-proj+residual", "c_fc": "c_fc+QuickGELU", "c_proj": "c_proj+residual",
         "d_c_proj": "d c_proj * GELU'", "d_c_fc": "d c_fc", "d_out": "d out-proj", "d_qkv": "d QKV"}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    M, d = 77000, 512
    shapes = {"qkv": (3 * d, d, 0), "out": (d, d, 2), "c_fc": (4 * d, d, 3), "c_proj": (d, 4 * d, 2),
              "d_c_proj": (4 * d, d, 4), "d_c_fc": (d, 4 * d, 0), "d_out": (d, d, 0), "d_qkv": (d, 3 * d, 0)}
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]

    def to_mb(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]

    launches = []
    # gpu_gemm_prof.py launches each tag 3 times (2 warm-ups + 1 timed with --iters 1): keep the last of each
    for t, tag in enumerate(TAGS):
        r = data[3 * t + 2]
        N, K, mode = shapes[tag]
        f32 = mode in (1, 2)
        alg = 2 * (M * K + N * K) + M * N * ((4 if f32 else 2) + (2 if mode == 3 else 0) + (4 if mode == 2 else 0) + (2 if mode == 4 else 0))
        e = {"kernel": r[ix["Kernel Name"]][:60], "gemm": f"{NAMES[tag]}: M={M} N={N} K={K}"}
        for w in want:
            e[w] = float(r[ix[w]].replace(",", ""))
        rd = to_mb(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_mb(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        e["dram__bytes_read.sum"], e["dram__bytes_write.sum"] = rd, wr
        e["traffic_MB"] = rd + wr
        e["algorithmic_MB"] = alg / 1e6
        e["traffic_over_algorithmic"] = round((rd + wr) / (alg / 1e6), 3)
        launches.append(e)
    res = {"note": "DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, MB) of the 8 text-tower GEMM launches of one layer "
                   "(forward: QKV, out-proj, c_fc+GELU, c_proj; backward: d c_proj*GELU', d c_fc, d out-proj, d QKV) at BASELINE "
                   "config 2, 1 GPU, from one `ncu --set full` capture of tests/gpu_gemm_prof.py",
           "launches": launches,
           "avg_traffic_bytes_per_launch": sum(e["traffic_MB"] for e in launches) / len(launches) * 1e6,
           "avg_algorithmic_bytes_per_launch": sum(e["algorithmic_MB"] for e in launches) / len(launches) * 1e6}
    res["avg_traffic_over_algorithmic"] = round(res["avg_traffic_bytes_per_launch"] / res["avg_algorithmic_bytes_per_launch"], 3)
    json.dump(res, open(out, "w"), indent=1)
    for e in launches:
        print(f'{e["gemm"]:48s} {e["gpu__time_duration.sum"]:8.1f} us  tensor {e["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]:5.1f}%  '
              f'dram {e["traffic_MB"]:7.1f} MB ({e["traffic_over_algorithmic"]:.2f}x algorithmic)')


if __name__ == "__main__":
    main()
