for v in main nopre nopow main; do
  if [ $v = main ]; then E="X=1"; else E="SOFTRAY_SO=variants/$v.so"; fi
  env $E python bench.py --workload config2 --others "" --steps 40 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v config2', round(d['ms_per_step'],4))"
done
