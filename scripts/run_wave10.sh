python -m pytest tests/test_cuda_wave.py tests/test_cuda_parity.py -x -q -m gpu 2>&1 | tail -3
one() { label=$1; w=$2; shift; shift
  env "$@" python bench.py --workload $w --others "" --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['measured']['counters']
print('$label $w', round(d['ms_per_step'],3), 'ms stages', {k: round(v,2) for k,v in d['roofline']['stages_ms'].items()}, 'beam', c['rays_beam'], 'of', c['rays_primary'], 'nodes', c['node_visits'], 'filt', c['filter_tests'], 'fb', c['rays_fallback'])"
}
one beams config4 X=1
one beams-occ3 config4 SOFTRAY_WAVE_SEARCH_OCC=3
one beams-occ2 config4 SOFTRAY_WAVE_SEARCH_OCC=2
one nobeams config4 SOFTRAY_BEAM_BUDGET=0
one beams-b256 config4 SOFTRAY_BEAM_BUDGET=256 SOFTRAY_WAVE_SEARCH_OCC=3
