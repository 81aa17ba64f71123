# round-end evidence: GPU parity suite, smoke, default bench line (+ reference arm), launch list of the same command
TAG=${1:-r01v}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_${TAG}_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
grep -c render_kernel gpurun_out/launches_$TAG.csv
