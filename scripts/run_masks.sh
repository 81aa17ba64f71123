fmt='import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(round(d["ms_per_step"],4))
except Exception as e: print("ERR", t[-600:])'
for m in ${MASKS:-31 0 1 3 5 9 17 7 15 19 27 23 29}; do echo -n "mask $m: "; SOFTRAY_PHASE_SYNC=$m timeout 120 python bench.py --workload ${W:-config2} --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "$fmt"; done
