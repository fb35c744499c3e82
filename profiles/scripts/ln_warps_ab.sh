# LayerNorm kernels: rows (warps) per CTA A/B at the per-rank and N = 1 shapes (bench.py --quick)
for w in 8 4 2 1; do
  echo "== MUDPT_LN_WARPS=$w"
  for c in 125 1000; do
    MUDPT_LN_WARPS=$w python bench.py --quick --steps 10 --classes $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  classes', $c, 'ms/step', round(d['ms_per_step'],3), {k:v for k,v in d['kernels_us_per_launch'].items() if k.startswith('ln')})"
  done
done
