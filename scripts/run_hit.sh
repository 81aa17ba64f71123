one() { label=$1; w=$2; shift; shift
  env "$@" python bench.py --workload $w --others "" --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label $w', round(d['ms_per_step'],3), 'ms', {k: round(v,2) for k,v in (d['roofline'].get('stages_ms') or {}).items()})"
}
for o in 2 3 4; do one hit-occ$o config4 SOFTRAY_WAVE_HIT_OCC=$o; done
for o in 2 4; do one hit-occ$o config5 SOFTRAY_WAVE_HIT_OCC=$o; done
