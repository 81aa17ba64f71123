# per-kernel launch list of the stage kernels (cold-cache, serialised: shares, not absolutes)
for w in ${W:-config5 config3 config4}; do
SOFTRAY_PIPELINE=wave ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/w2_launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/w2_$w.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/w2_launches_$w.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ik].split("(")[0][-40:]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[iv].replace(",",""))
tot=sum(v[1] for v in agg.values())
print("$w")
for k,(n,t) in agg.items(): print(f"  {k:42s} n={n:4d} total={t/1e6:9.3f} ms  {100*t/tot:5.1f}%")
PY
done
