# baseline (HEAD tree) vs SR_OUTLINE levels, same box
WL=${W:-config2 config4 config3}
fmt='import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(round(d["ms_per_step"],3))
except Exception as e: print("ERR", t[-800:])'
for w in $WL; do echo "== old_tree $w"; (cd variants/old_tree && python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "$fmt"); done
cp softray_b200/libsoftray_cuda.so /tmp/orig.so
for v in ${V:-o0 o1 o2 o3}; do
  cp variants/$v.so softray_b200/libsoftray_cuda.so
  for w in $WL; do echo "== $v $w"; python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "$fmt"; done
done
cp /tmp/orig.so softray_b200/libsoftray_cuda.so
