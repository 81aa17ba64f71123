one() { label=$1; w=$2; shift; shift
  env "$@" python bench.py --workload $w --others "" --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['measured']['counters']
print('$label $w', round(d['ms_per_step'],3), 'ms', {k: round(v,2) for k,v in (d['roofline'].get('stages_ms') or {}).items()}, 'nodes', c['node_visits'], 'filt', c['filter_tests'])"
}
for c in 0.5 1 4 8; do for w in config5 config4; do one isect$c $w SOFTRAY_SAH_ISECT=$c; done; done
