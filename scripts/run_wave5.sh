# A/B of builds and knobs: "label|env assignments" per line in $CASES; prints ms per frame for each workload
python -m pytest tests/test_cuda_wave.py -x -q 2>&1 | tail -3
run() { label=$1; shift
  for w in ${W:-config5 config3 config4}; do
    ms=$(env SOFTRAY_PIPELINE=wave "$@" python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3))
except Exception as e: print('ERR')")
    echo "$label $w $ms"
  done
}
run ww1-2streams X=1
run ww1-1stream SOFTRAY_WAVE_STREAMS=1
run ww0-2streams SOFTRAY_SO=variants/ww0.so
run ww0-1stream SOFTRAY_SO=variants/ww0.so SOFTRAY_WAVE_STREAMS=1
for w in config5 config3 config4; do SOFTRAY_PIPELINE=wave SOFTRAY_WAVE_STREAMS=1 SOFTRAY_WAVE_TIMING=1 python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --no-e2e 2>&1 >/dev/null | grep "wave stages" | tail -1 | sed "s/^/ww1 $w: /"; done
