# usage: run_prof.sh <tag> [workload]   -- bench line, then one ncu --set full capture of the render kernel
TAG=$1; W=${2:-config2}
python bench.py --workload $W --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
tail -c 600 gpurun_out/bench_$TAG.json
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --workload $W --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log | cut -c1-300
