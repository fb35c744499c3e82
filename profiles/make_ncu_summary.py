#!/usr/bin/env python
"""Compact per-launch summary of an `ncu --set full` report (read here, no GPU needed):

    python profiles/make_ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/r02_ncu_x.txt

One block per captured launch: duration, DRAM bytes and rate, pipe / issue utilisation, occupancy and its limiters, the
largest warp-stall reasons (cycles a warp waits per instruction it issues)."""
import csv
import subprocess
import sys

KEYS = [
    ("duration us", "gpu__time_duration.sum"),
    ("SM clock GHz", "sm__cycles_elapsed.max.per_second"),
    ("DRAM read MB", "dram__bytes_read.sum"),
    ("DRAM write MB", "dram__bytes_write.sum"),
    ("DRAM rate TB/s", "dram__bytes.sum.per_second"),
    ("DRAM throughput % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe active % (elapsed)", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("executed warp instructions", "smsp__inst_executed.sum"),
    ("registers / thread", "launch__registers_per_thread"),
    ("dynamic smem KB / block", "launch__shared_mem_per_block_dynamic"),
    ("waves per SM", "launch__waves_per_multiprocessor"),
    ("block limit: registers / smem / warps", None),
    ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle",
               "membar", "dispatch_stall", "branch_resolving", "no_instruction", "sleeping", "tex_throttle", "drain", "imc_miss",
               "not_selected", "selected"]


def fmt(v, unit):
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    u = unit.strip()
    scale = {"Kbyte": 1e-3, "byte": 1e-6, "Mbyte": 1.0, "Gbyte": 1e3}
    if u in scale:
        return f"{f * scale[u]:.1f}"
    if u.endswith("byte/s"):
        return f"{f * {'Tbyte/s': 1.0, 'Gbyte/s': 1e-3, 'Mbyte/s': 1e-6}.get(u, 1.0):.2f}"
    if u in ("ns", "nsecond"):
        return f"{f * 1e-3:.2f}"
    if u in ("Ghz", "Mhz"):
        return f"{f * (1.0 if u == 'Ghz' else 1e-3):.2f}"
    return f"{f:.2f}" if abs(f) < 1e6 else f"{f:.4g}"


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        print(f"# {rep.split('/')[-1]}  (ncu --set full --clock-control none; cold caches, serialised)")
        for r in rows[2:]:
            name = r[ix["Kernel Name"]]
            print(f"  {name[:110]}   grid {r[ix['Grid Size']]} x block {r[ix['Block Size']]}")
            for label, key in KEYS:
                if key is None:
                    lim = [r[ix[k]] for k in ("launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps") if k in ix]
                    print(f"    {label:44s} {' / '.join(str(int(float(x))) for x in lim)}")
                elif key in ix and r[ix[key]] not in ("", "n/a"):
                    print(f"    {label:44s} {fmt(r[ix[key]], units[ix[key]])}")
            st = []
            for s in STALL_NAMES:
                k = STALLS % s
                if k in ix:
                    try:
                        st.append((float(r[ix[k]]), s))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            print("    stall cycles per issued instruction:        " + ", ".join(f"{s} {v:.2f}" for v, s in st[:5]))
        print()


if __name__ == "__main__":
    main()
