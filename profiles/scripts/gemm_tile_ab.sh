for cfg in "0 0 0" "1 20000 100000" "1 20000 1024" "2 20000 100000" "2 20000 1024"; do
  set -- $cfg
  echo "== MUDPT_GEMM_TILE=$1 ROWS=$2 N=$3"
  MUDPT_GEMM_TILE=$1 MUDPT_GEMM_TILE_ROWS=$2 MUDPT_GEMM_TILE_N=$3 python bench.py --quick --steps 10 --classes 125 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], {k:v for k,v in d['kernels_us_per_launch'].items() if k.startswith('gemm')})"
done
