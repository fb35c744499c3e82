echo "== base per-rank"; python bench.py --quick --steps 20 --classes 125 2>/dev/null | cut -c1-70
echo "== main high per-rank"; MUDPT_MAIN_PRIORITY=-1 python bench.py --quick --steps 20 --classes 125 2>/dev/null | cut -c1-70
echo "== base N=1"; python bench.py --quick --steps 10 2>/dev/null | cut -c1-70
echo "== main high N=1"; MUDPT_MAIN_PRIORITY=-1 python bench.py --quick --steps 10 2>/dev/null | cut -c1-70
