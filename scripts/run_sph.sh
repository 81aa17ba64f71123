python -m pytest tests/test_cuda_parity.py tests/test_cuda_filter.py tests/test_cuda_wave.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do python bench.py --workload config2 --others "" --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config2', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['ms_per_step'],4), d['e2e']['matches_device_frame'])"; done
