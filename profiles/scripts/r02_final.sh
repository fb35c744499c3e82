# Final single-GPU evidence of round 2 (B200 box, repo root):  bash profiles/scripts/r02_final.sh
# Plain runs first; every ncu capture only after the same command has exited 0 without ncu.
mkdir -p gpurun_out
R=r02d
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log
: > gpurun_out/${R}_ln_bwd_variants.txt
MUDPT_LN_BWD_PIPE=0 timeout 120 python tests/gpu_ln_bwd_prof.py >> gpurun_out/${R}_ln_bwd_variants.txt 2>> gpurun_out/${R}_ln_bwd_variants.err
timeout 120 python tests/gpu_ln_bwd_prof.py >> gpurun_out/${R}_ln_bwd_variants.txt 2>> gpurun_out/${R}_ln_bwd_variants.err
python - <<'P'
import json
for l in open("gpurun_out/r02d_ln_bwd_variants.txt"):
    d = json.loads(l)
    print(d["MUDPT_LN_BWD_PIPE"], {k[:14] + k[-9:]: (v["us"], v["frac_of_hbm"]) for k, v in d["shapes"].items()})
P
timeout 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02d_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "launches", d["gpu_launches_per_step"], d["clocks"])
print({k: (v["ms_per_step"], v.get("gbs_algorithmic")) for k, v in d["kernels"].items() if k.startswith("ln")})
P
timeout 200 python bench.py --quick --classes 125 --steps 20 > gpurun_out/${R}_quick_classes125.json 2>/dev/null; echo "quick125 rc=$?"; cut -c1-120 gpurun_out/${R}_quick_classes125.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches_step.csv \
    python bench.py --quick --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_pipe --launch-skip 150 -c 3 -o gpurun_out/${R}_prof_lnbwd_pipe -f \
    python bench.py --quick --steps 1 > gpurun_out/ncu_lnbwd.log 2>&1; echo "ncu lnbwd rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_pp --launch-skip 40 -c 1 -o gpurun_out/${R}_prof_attn_bwd_pp -f \
    python bench.py --quick --steps 1 > gpurun_out/ncu_attn_pp.log 2>&1; echo "ncu attn pp rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_short_fwd --launch-skip 40 -c 1 -o gpurun_out/${R}_prof_attn_short_fwd -f \
    python bench.py --quick --steps 1 > gpurun_out/ncu_attn_sf.log 2>&1; echo "ncu attn short fwd rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:head_rows|head_cols|splice_bwd_fused|ln_fwd_kernel" --launch-skip 120 -c 8 -o gpurun_out/${R}_prof_small -f \
    python bench.py --quick --steps 1 > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
