# one GPU, 1/8 of config5's pixels (what a rank renders at N = 8): one chunk on one stream against two in flight
for s in 2 1 2 1; do
SOFTRAY_WAVE_STREAMS=$s python bench.py --workload config5 --scale 0.3536 --others "" --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('streams $s:', round(d['ms_per_step'],3), 'ms, launches', d['measured']['launches_per_frame'])"
done
