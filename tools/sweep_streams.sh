for cfg in "2 8" "2 10" "2 12" "3 8" "3 10" "3 12" "4 10" "4 14"; do set -- $cfg; python bench.py --steps 100 --warmup 5 --no-cpu --no-aux --streams $1 --sm-reserve $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams $1 reserve $2:', round(d['value']), d['ms_per_step'])"; done
