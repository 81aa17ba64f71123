# usage: run_variants.sh "<variant .so files>" "<workloads>"
for v in $1; do
  cp softray_b200/libsoftray_cuda.so /tmp/orig.so
  cp $v softray_b200/libsoftray_cuda.so
  for w in $2; do echo "== $v $w"; python bench.py --workload $w --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value'],1))
except Exception as e: print('ERR', t[-1500:])"; done
  cp /tmp/orig.so softray_b200/libsoftray_cuda.so
done
