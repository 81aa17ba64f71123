fmt='import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(round(d["ms_per_step"],3), "nodes", d["config"]["counters"]["node_visits"], "filt", d["config"]["counters"]["filter_tests"])
except Exception as e: print("ERR", t[-600:])'
for c in ${COSTS:-1.0 3.0 5.0}; do for w in ${W:-config3 config4 config5}; do echo -n "isect $c $w: "; SOFTRAY_SAH_ISECT=$c timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "$fmt"; done; done
