# usage: V="a b" W="config2 config3" run_vars.sh  -- time each variants/<v>.so on each workload (each run under timeout)
fmt='import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(round(d["ms_per_step"],3))
except Exception as e: print("ERR", t[-600:])'
cp softray_b200/libsoftray_cuda.so /tmp/orig.so
for v in $V; do
  cp variants/$v.so softray_b200/libsoftray_cuda.so
  if [ -n "$CHECK" ]; then echo "== $v parity"; timeout 300 python -m pytest tests/test_cuda_parity.py -x -q -k "goldens or config2_small or config3_small or known_answer or row_bands" 2>&1 | tail -2; fi
  for w in $W; do echo "== $v $w"; timeout 120 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "$fmt"; done
done
cp /tmp/orig.so softray_b200/libsoftray_cuda.so
