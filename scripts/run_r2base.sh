# round-2 baseline: timings of the big configs with the round-1 kernel, then ncu --set full on config5 (half size) and config3 (half size)
for w in config5 config3 config4; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2base_$w.json 2> gpurun_out/r2base_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2base_$w.json").read().strip().splitlines()[-1]); print("$w", round(d["ms_per_step"],3), "ms", round(d["value"]), "Mrays/s", d["config"]["counters"], "scene_ms", round(d["config"]["scene_create_ms"]))
except Exception as e: print("ERR $w", open("gpurun_out/r2base_$w.err").read()[-1500:])
PY
done
for w in config5 config3; do
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2base_$w \
    python bench.py --workload $w --scale 0.5 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r2base_$w.log 2>&1
tail -2 gpurun_out/ncu_r2base_$w.log | cut -c1-200
done
