python -m pytest tests/test_cuda_wave.py tests/test_cuda_parity.py -x -q -m gpu 2>&1 | tail -3
one() { label=$1; w=$2; shift; shift
  env "$@" python bench.py --workload $w --others "" --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['measured']['counters']
print('$label $w', round(d['ms_per_step'],3), 'ms stages', {k: round(v,2) for k,v in d['roofline']['stages_ms'].items()}, 'bundled', c['rays_bundled'], 'listed', c['rays_short_listed'], 'nodes', c['node_visits'])"
}
one default config3 X=1
one occ3 config3 SOFTRAY_WAVE_SHADOW_OCC=3 SOFTRAY_WAVE_WALK_OCC=3
one walk5 config3 SOFTRAY_WAVE_WALK_OCC=5
one default config5 X=1
one walk5 config5 SOFTRAY_WAVE_WALK_OCC=5
one walk3 config5 SOFTRAY_WAVE_WALK_OCC=3
