for c in 0.5 1.0 2.0 4.0; do echo "== isect $c"; for w in ${W:-config2}; do SOFTRAY_SAH_ISECT=$c python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], round(d['ms_per_step'],3), d['config']['counters']['node_visits'], d['config']['counters']['filter_tests'])"; done; done
