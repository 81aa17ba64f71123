for v in 98304 0 98304 0; do SOFTRAY_STAGE_SPHERES_MAX=$v python bench.py --workload config2 --others "" --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stage_max $v config2', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['ms_per_step'],4), d['e2e']['matches_device_frame'], d['measured']['counters']['node_visits'])"; done
python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "config2 or spheres" 2>&1 | tail -2
