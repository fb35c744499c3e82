# LayerNorm backward as a shared-memory pipeline (bulk copies + mbarriers): parity under every kernel selection, microbenchmark
# of the variants, then tests / A/B / bench line of the build with the better selection.
# On the B200 box from the repo root:  bash profiles/scripts/r02_ln_bwd_pipe.sh
mkdir -p gpurun_out
R=r02c
for m in 1 2 0; do
  MUDPT_LN_BWD_PIPE=$m timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "layernorm_splice or unfused_layernorm" > gpurun_out/${R}_pytest_pipe$m.log 2>&1
  echo "unit tests MUDPT_LN_BWD_PIPE=$m rc=$? $(tail -1 gpurun_out/${R}_pytest_pipe$m.log)"
done
: > gpurun_out/${R}_ln_bwd_variants.txt
run() { env "$@" timeout 120 python tests/gpu_ln_bwd_prof.py >> gpurun_out/${R}_ln_bwd_variants.txt 2>> gpurun_out/${R}_ln_bwd_variants.err; }
run MUDPT_LN_BWD_PIPE=0
run MUDPT_LN_BWD_PIPE=1
run MUDPT_LN_BWD_PIPE=2
run MUDPT_LN_BWD_PIPE=1 MUDPT_LN_PIPE_STAGES=2
run MUDPT_LN_BWD_PIPE=1 MUDPT_LN_PIPE_STAGES=3
run MUDPT_LN_BWD_PIPE=1 MUDPT_LN_PIPE_CTAS=1
run MUDPT_LN_BWD_PIPE=1 MUDPT_LN_PIPE_CTAS=3
python - <<'P'
import json
for l in open("gpurun_out/r02c_ln_bwd_variants.txt"):
    d = json.loads(l)
    print(d["MUDPT_LN_BWD_PIPE"], d["MUDPT_LN_PIPE_STAGES"], d["MUDPT_LN_PIPE_CTAS"], {k.split(" (")[0] + k[-8:]: (v["us"], v["frac_of_hbm"]) for k, v in d["shapes"].items()})
P
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log
for c in 1000 125; do for m in 0 1; do
  echo "classes $c pipe $m: $(MUDPT_LN_BWD_PIPE=$m timeout 200 python bench.py --quick --classes $c --steps 20 2>/dev/null | tail -1)" >> gpurun_out/${R}_ln_bwd_pipe_ab.txt
done; done
python - <<'P'
import json
for l in open("gpurun_out/r02c_ln_bwd_pipe_ab.txt"):
    h, js = l.split(": ", 1)
    try:
        d = json.loads(js)
        print(h, round(d["ms_per_step"], 3), "ln_bwd ms", d["kernels_ms_per_step"].get("ln_bwd"), "us", d["kernels_us_per_launch"].get("ln_bwd"))
    except Exception as e:
        print(h, "failed", js[:100])
P
timeout 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02c_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "launches", d["gpu_launches_per_step"], d["clocks"])
print({k: (v["ms_per_step"], v.get("gbs_algorithmic")) for k, v in d["kernels"].items() if k.startswith("ln")})
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches_step.csv \
    python bench.py --quick --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_bwd --launch-skip 150 -c 3 -o gpurun_out/${R}_prof_lnbwd -f \
    python bench.py --quick --steps 1 > gpurun_out/ncu_lnbwd.log 2>&1; echo "ncu lnbwd rc=$?"
