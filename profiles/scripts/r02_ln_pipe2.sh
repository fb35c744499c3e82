# Second pass on the LayerNorm pipelines (trimmed backward, forward on the same ring): tests, microbenchmark, step A/B.
# B200 box, repo root:  bash profiles/scripts/r02_ln_pipe2.sh
mkdir -p gpurun_out
R=r02f
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 100 python tests/gpu_ln_bwd_prof.py > gpurun_out/${R}_ln_bwd_variants.txt 2> gpurun_out/${R}_ln_bwd_variants.err
python - <<'P'
import json
for l in open("gpurun_out/r02f_ln_bwd_variants.txt"):
    d = json.loads(l)
    print(d["MUDPT_LN_BWD_PIPE"], {k[:14] + k[-9:]: (v["us"], v["frac_of_hbm"]) for k, v in d["shapes"].items()})
P
for m in 0 1; do
  echo "classes 125 ln_fwd pipe $m: $(MUDPT_LN_FWD_PIPE=$m timeout 150 python bench.py --quick --classes 125 --steps 20 2>/dev/null | tail -1)" >> gpurun_out/${R}_ln_fwd_pipe_ab.txt
done
echo "classes 1000 default: $(timeout 150 python bench.py --quick --steps 20 2>/dev/null | tail -1)" >> gpurun_out/${R}_ln_fwd_pipe_ab.txt
python - <<'P'
import json
for l in open("gpurun_out/r02f_ln_fwd_pipe_ab.txt"):
    h, js = l.split(": ", 1)
    try:
        d = json.loads(js)
        print(h, round(d["ms_per_step"], 3), "ln_fwd", d["kernels_ms_per_step"].get("ln_fwd"), d["kernels_us_per_launch"].get("ln_fwd"), "ln_bwd", d["kernels_ms_per_step"].get("ln_bwd"), d["kernels_us_per_launch"].get("ln_bwd"))
    except Exception as e:
        print(h, "failed", js[:100])
P
