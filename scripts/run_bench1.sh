# the driver's default bench line + the reference arm (short)
python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 300 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2a.json").read().strip().splitlines()[-1])
print("headline", d["config"]["workload"], round(d["ms_per_step"],3), "ms", round(d["value"]), "Mrays/s; e2e", round(d["e2e"]["ms_per_step"],3), "ms ok", d["e2e"]["matches_device_frame"], "frac", round(d["roofline"]["frac"],4))
print(" stages", d["roofline"].get("stages_ms"), d["roofline"].get("dominant_kernel",{}).get("frac"))
print(" cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["threads_4"]["value"])
for k,o in d["others"].items(): print(" ", k, round(o["ms_per_step"],3), "ms", round(o["value"]), "Mrays/s traced", round(o["value_traced"]), "e2e", round(o["e2e"]["ms_per_step"],3), o["e2e"]["matches_device_frame"], "frac", round(o["roofline"]["frac"],4), o["pipeline"][:5], o["roofline"].get("stages_ms"))
PY
