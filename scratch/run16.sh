set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" 
timeout 600 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
timeout 300 python bench.py --config 3 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; echo "cfg3 rc=$?"
timeout 300 python bench.py --config 4 > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err; echo "cfg4 rc=$?"
timeout 300 python bench.py --quick --steps 1 > gpurun_out/q.log 2>&1 && timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tn_tcgen05 -c 4000 --csv --log-file gpurun_out/r02_step_traffic.csv python bench.py --quick --steps 1 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic rc=$?"
timeout 300 python bench.py --quick --steps 1 --classes 125 > gpurun_out/q125.log 2>&1; echo "q125 rc=$?"
tail -5 gpurun_out/r02_pytest_final.log
