# bf16 gradient stream (option grad_stream_bf16): tests, error metrics, A/B timing, then the final single-GPU evidence of the
# build (bench line, GEMM traffic for the new api.cu digest, launch list, ncu of the LayerNorm backward).
# On the B200 box from the repo root:  bash profiles/scripts/r02_grad_stream.sh
mkdir -p gpurun_out
R=r02b
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log
timeout 300 python tests/gpu_grad_stream_metrics.py > gpurun_out/${R}_grad_stream_bf16.txt 2> gpurun_out/${R}_grad_stream_bf16.err; echo "metrics rc=$?"
grep worst gpurun_out/${R}_grad_stream_bf16.txt
for c in 1000 125; do for g in 0 1 0 1; do
  echo "classes $c grad_bf16 $g: $(MUDPT_GRAD_BF16=$g timeout 200 python bench.py --quick --classes $c --steps 20 2>/dev/null | tail -1)" >> gpurun_out/${R}_grad_stream_ab.txt
done; done
cut -c1-200 gpurun_out/${R}_grad_stream_ab.txt
timeout 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02b_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches_per_step"], d["clocks"])
print({k: (v["ms_per_step"], v.get("gbs_algorithmic")) for k, v in d["kernels"].items() if k.startswith("ln")})
P
timeout 300 python bench.py --quick --steps 1 > gpurun_out/q.log 2>&1 && \
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tn_tcgen05 -c 4000 --csv \
    --log-file gpurun_out/${R}_step_traffic.csv python bench.py --quick --steps 1 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches_step.csv \
    python bench.py --quick --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel --launch-skip 150 -c 4 -o gpurun_out/${R}_prof_lnbwd -f \
    python bench.py --quick --steps 1 > gpurun_out/ncu_lnbwd.log 2>&1; echo "ncu lnbwd rc=$?"
