# stage kernels: parity first, then timings against the fused kernel
python -m pytest tests/test_cuda_wave.py -x -q 2>&1 | tail -15
for pl in fused wave; do for w in ${W:-config5 config3 config4}; do
  SOFTRAY_PIPELINE=$pl timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/w1_${pl}_$w.json 2> gpurun_out/w1_${pl}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/w1_${pl}_$w.json").read().strip().splitlines()[-1]); c=d["config"]["counters"]; print("$pl $w", round(d["ms_per_step"],3), "ms", round(d["value"]), "Mrays/s nodes", c["node_visits"], "prim", c["prim_tests"], "filt", c["filter_tests"], "unsure", c["filter_unsure"], "bundled", c["rays_bundled"])
except Exception as e: print("ERR $pl $w", open("gpurun_out/w1_${pl}_$w.err").read()[-1500:])
PY
done; done
