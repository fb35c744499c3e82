# Final single-GPU evidence run of a round (on the B200 box, from the repo root):  bash profiles/scripts/final_n1.sh r02
# Plain runs first; ncu captures only after the same command exited 0 without ncu.
R=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --config 3 > gpurun_out/${R}_bench_cfg3.json 2> gpurun_out/${R}_bench_cfg3.err; echo "cfg3 rc=$?"
timeout 300 python bench.py --config 4 > gpurun_out/${R}_bench_cfg4.json 2> gpurun_out/${R}_bench_cfg4.err; echo "cfg4 rc=$?"
timeout 300 python bench.py --quick --steps 1 > gpurun_out/q.log 2>&1 && \
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tn_tcgen05 -c 4000 --csv \
    --log-file gpurun_out/${R}_step_traffic.csv python bench.py --quick --steps 1 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic rc=$?"
timeout 300 python bench.py --quick --steps 1 > gpurun_out/q2.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches_step.csv \
    python bench.py --quick --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?"
tail -3 gpurun_out/${R}_pytest_gpu.log
