# ncu --set full of one stage kernel: run_wave4.sh <tag> <workload> <kernel regex> [skip]
TAG=$1; W=$2; K=$3; S=${4:-0}
SOFTRAY_PIPELINE=wave ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --workload $W --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
