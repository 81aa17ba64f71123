for c in 2.0 1.0 0.5 0.3; do echo "== isect $c"; SOFTRAY_SAH_ISECT=$c python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['counters'])"; done
