python -m pytest tests/test_cuda_wave.py -x -q -m gpu 2>&1 | tail -2
one() { label=$1; w=$2; shift; shift
  env "$@" python bench.py --workload $w --others "" --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['measured']['counters']
print('$label $w', round(d['ms_per_step'],3), 'ms search', round(d['roofline']['stages_ms']['search'],2), 'beam', c['rays_beam'], 'of', c['rays_primary'], 'nodes', c['node_visits'], 'filt', c['filter_tests'], 'fb', c['rays_fallback'])"
}
for m in 2 8 32 128; do one margin$m config4 SOFTRAY_BEAM_MARGIN=$m SOFTRAY_WAVE_SEARCH_OCC=3; done
one margin32-b4096 config4 SOFTRAY_BEAM_MARGIN=32 SOFTRAY_BEAM_BUDGET=4096 SOFTRAY_WAVE_SEARCH_OCC=3
