# A/B of experimental builds: VARIANTS="name1 name2" (variants/<name>.so; "main" = the in-tree library), W = workloads
one() { label=$1; w=$2; shift; shift
  env "$@" python bench.py --workload $w --others "" --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label $w', round(d['ms_per_step'],3), 'ms', {k: round(v,2) for k,v in (d['roofline'].get('stages_ms') or {}).items()})"
}
for v in ${VARIANTS:-main}; do for w in ${W:-config5 config3 config4}; do
  if [ $v = main ]; then one main $w X=1; else one $v $w SOFTRAY_SO=variants/$v.so; fi
done; done
