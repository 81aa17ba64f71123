N=${1:-2}
python -m pytest tests/test_abi.py tests/test_cuda_parity.py -m gpu -x -q -k "abi or handles or pinned or phase" 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/n$N.json 2> gpurun_out/n$N.err || tail -c 2000 gpurun_out/n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["ms_per_step"],4), round(d["value"]), "e2e", d["e2e"]["ms_per_step"], round(d["e2e"]["value"]), d["e2e"].get("matches_device_gather"))
PY
