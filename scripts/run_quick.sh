# parity suites that exercise the shadow path, then timings of the configs
python -m pytest tests/test_cuda_filter.py tests/test_cuda_lbvh.py -x -q 2>&1 | tail -8
python -m pytest tests/test_cuda_parity.py -x -q 2>&1 | tail -5
for w in ${W:-config2 config3 config5}; do echo "== $w"; python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(round(d['ms_per_step'],3), 'ms/frame', d['config']['counters'])
except Exception as e: print('ERR', t[-2000:])"; done
