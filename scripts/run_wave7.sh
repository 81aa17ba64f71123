python -m pytest tests/test_cuda_wave.py tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --no-cpu ${BENCH_ARGS} > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 300 gpurun_out/bench_r2c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2c.json").read().strip().splitlines()[-1])
print("headline", d["config"]["workload"], round(d["ms_per_step"],3), "ms", round(d["value"]), "Mrays/s; e2e", round(d["e2e"]["ms_per_step"],3), "ms ok", d["e2e"]["matches_device_frame"], "frac", round(d["roofline"]["frac"],4))
print(" stages", {k: round(v,2) for k,v in d["roofline"].get("stages_ms").items()})
for k,o in d["others"].items(): print(" ", k, round(o["ms_per_step"],3), "ms", "e2e", round(o["e2e"]["ms_per_step"],3), o["e2e"]["matches_device_frame"], "frac", round(o["roofline"]["frac"],4), o["pipeline"][:5], {k: round(v,2) for k,v in (o["roofline"].get("stages_ms") or {}).items()})
PY
