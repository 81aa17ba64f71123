# 1/2/4/8-GPU bench lines (the driver's scaling run), peer-mapped gather unless GATHER=nccl.  NS="1 2" limits the counts.
G=${GATHER:-peer}
for n in ${NS:-1 2 4 8}; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 --gather $G > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err || tail -c 1500 gpurun_out/scale_n$n.err
  fi
done
for n in ${NS:-1 2 4 8}; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print($n, d["config"]["workload"], round(d["ms_per_step"],3), "ms", round(d["value"]), "Mrays/s e2e", round(d["e2e"]["ms_per_step"],3), "single", d["measured"]["matches_single_gpu"], "e2e_ok", d["e2e"]["matches_device_frame"], {k: (round(o["ms_per_step"],3), o["matches_single_gpu"]) for k,o in d["others"].items()})
except Exception as e: print($n, "ERR", e)
PY
done
