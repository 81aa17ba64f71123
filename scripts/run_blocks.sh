for b in 2 3 4 5 6; do echo "== blocks/SM $b"; for w in ${W:-config4}; do SOFTRAY_BLOCKS_PER_SM=$b python bench.py --workload $w --steps 3 --warmup 2 --no-cpu --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], round(d['ms_per_step'],3))"; done; done
