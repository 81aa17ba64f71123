python -m pytest tests/test_cuda_wave.py -x -q 2>&1 | tail -3
one() { label=$1; shift
  env "$@" python bench.py --workload config3 --others "" --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['measured']['counters']
print('$label', round(d['ms_per_step'],3), 'ms shadow', round(d['roofline']['stages_ms']['shadow'],2), 'bundled', c['rays_bundled'], 'listed', c['rays_short_listed'], 'of', c['rays_shadow'], 'nodes', c['node_visits'], 'filt', c['filter_tests'])"
}
one b384-occ4 X=1
one b384-occ3 SOFTRAY_WAVE_SHADOW_OCC=3
one b1024-occ3 SOFTRAY_BUNDLE_BUDGET=1024 SOFTRAY_WAVE_SHADOW_OCC=3
one b128-occ3 SOFTRAY_BUNDLE_BUDGET=128 SOFTRAY_WAVE_SHADOW_OCC=3
one b0-occ4 SOFTRAY_BUNDLE_BUDGET=0
