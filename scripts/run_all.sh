W=${W:-"config3 config4 config5"}
for w in $W; do echo "== $w $*"; python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['config']['counters'], 'scene_ms', d['config']['scene_create_ms'])
except Exception as e: print('ERR', t[-2000:])"; done
