N=${N:-8}
for bh in ${BHS:-32 16 8}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+bh)) bench.py --gpus $N --steps 10 --warmup 3 --others "" --band-height $bh --no-e2e > gpurun_out/bands_$bh.json 2> gpurun_out/bands_$bh.err || tail -c 800 gpurun_out/bands_$bh.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bands_$bh.json").read().strip().splitlines()[-1])
print("band_height $bh:", round(d["ms_per_step"],3), "ms", d["measured"]["matches_single_gpu"], {k: round(v,2) for k,v in d["roofline"]["stages_ms"].items()})
PY
done
