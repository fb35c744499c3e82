for p in 0 -1; do
  echo "== MUDPT_SIDE_PRIORITY=$p per-rank"; MUDPT_SIDE_PRIORITY=$p python bench.py --quick --steps 20 --classes 125 2>/dev/null | cut -c1-70
  echo "== MUDPT_SIDE_PRIORITY=$p N=1"; MUDPT_SIDE_PRIORITY=$p python bench.py --quick --steps 10 2>/dev/null | cut -c1-70
done
echo "== per-rank again prio 0"; python bench.py --quick --steps 20 --classes 125 2>/dev/null | cut -c1-70
echo "== per-rank again prio -1"; MUDPT_SIDE_PRIORITY=-1 python bench.py --quick --steps 20 --classes 125 2>/dev/null | cut -c1-70
