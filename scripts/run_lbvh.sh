# device-built tree: parity suite, then build time + trace time against the host SAH tree
SOFTRAY_BUILD_TIMING=1 python -m pytest tests/test_cuda_lbvh.py -x -q -s -k 'one_million' 2>&1 | tail -15
python -m pytest tests/test_cuda_lbvh.py -x -q 2>&1 | tail -5
for leaf in ${LEAVES:-1 2 4 8}; do for w in ${W:-config3 config5}; do
  echo "== lbvh leaf=$leaf $w"
  SOFTRAY_BUILD_TIMING=1 SOFTRAY_ACCEL=lbvh SOFTRAY_LBVH_LEAF=$leaf python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(round(d['ms_per_step'],3), 'ms/frame; scene_create_ms', round(d['config']['scene_create_ms'],1), d['config']['counters'])
except Exception as e: print('ERR', t[-2000:])"
done; done
for w in ${W:-config3 config5}; do
  echo "== host SAH $w"
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(round(d['ms_per_step'],3), 'ms/frame; scene_create_ms', round(d['config']['scene_create_ms'],1), d['config']['counters'])
except Exception as e: print('ERR', t[-2000:])"
done
