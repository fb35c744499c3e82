# bench line + GEMM traffic capture of the build in the tree (B200 box, repo root):  bash profiles/scripts/final_bench_traffic.sh r02
R=${1:-r02}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err; echo "ref rc=$?"
timeout 300 python bench.py --quick --steps 1 > gpurun_out/q.log 2>&1 && \
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tn_tcgen05 -c 4000 --csv \
    --log-file gpurun_out/${R}_step_traffic.csv python bench.py --quick --steps 1 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic rc=$?"
