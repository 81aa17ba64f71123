python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in config3 config2; do for b in 384 0; do
SOFTRAY_BUNDLE_BUDGET=$b python bench.py --workload $w --others "" --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['measured']['counters']
print('$w SOFTRAY_BUNDLE_BUDGET=$b', round(d['ms_per_step'],3), 'ms', round(d['value']), 'Mrays/s; traced', d['measured']['rays_traced'], 'of', d['measured']['rays_per_step'], 'value_traced', round(d['measured']['value_traced']))"
done; done
