# usage: run_prof2.sh <tag> <workload> [scale]  -- one ncu --set full capture of the render kernel (no bench line)
TAG=$1; W=${2:-config2}; S=${3:-1.0}
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --workload $W --scale $S --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log | cut -c1-1200
