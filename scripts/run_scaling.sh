# 1/2/4/8-GPU bench lines (the driver's scaling run), peer-mapped gather unless GATHER=nccl
G=${GATHER:-peer}
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 --gather $G > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err || tail -c 1500 gpurun_out/scale_n$n.err
done
for n in 1 2 4 8; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print($n, round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],4), round(d["e2e"]["value"]), d["clocks"])
except Exception as e: print($n, "ERR", e)
PY
done
