# occupancy sweep of the stage kernels: per-stage CUDA-event times (SOFTRAY_WAVE_TIMING), last frame of a short bench
one() { # workload, env...
  w=$1; shift
  env SOFTRAY_PIPELINE=wave SOFTRAY_WAVE_TIMING=1 "$@" python bench.py --workload $w --steps 2 --warmup 2 --no-cpu --no-e2e 2>&1 >/dev/null | grep "wave stages" | tail -1 | sed "s/^/$w $* : /"
}
for o in 3 4 5 6; do one config5 SOFTRAY_WAVE_SEARCH_OCC=$o SOFTRAY_WAVE_SHADOW_OCC=$o SOFTRAY_WAVE_HIT_OCC=$((o>4?4:o)); done
for o in 2; do one config5 SOFTRAY_WAVE_SEARCH_OCC=$o SOFTRAY_WAVE_SHADOW_OCC=$o SOFTRAY_WAVE_HIT_OCC=$o; done
for o in 2 3 4 5 6; do one config3 SOFTRAY_WAVE_SHADOW_OCC=$o; done
for o in 3 4 5 6; do one config4 SOFTRAY_WAVE_SEARCH_OCC=$o SOFTRAY_WAVE_HIT_OCC=$((o>4?4:o)); done
python bench.py --workload config4 --steps 1 --warmup 1 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config4 counters', d['config']['counters'])"
