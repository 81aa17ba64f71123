# round-end evidence.  TAG=r02z
TAG=${1:-r02z}
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${TAG}_pytest_gpu.txt; cat gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 200 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
# launch list of the default command's frames (cold, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config5.csv python scripts/one_frame.py config5 3 > /dev/null 2>&1
# DRAM traffic of whole frames
for w in config5 config3 config4 config2; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${TAG}_dram_$w.csv python scripts/one_frame.py $w 2 > /dev/null 2>&1
done
# ncu --set full of the stage kernels (second frame, a middle chunk where the frame has several)
cuobjdump -xelf sr_wave softray_b200/libsoftray_cuda.so > /dev/null 2>&1; cuobjdump -xelf sr_render softray_b200/libsoftray_cuda.so > /dev/null 2>&1
cap() { # tag workload kernel-regex skip section cubin
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o /tmp/prof_$1 python scripts/one_frame.py $2 2 > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/prof_$1.ncu-rep gpurun_out/${TAG}_$1_ncu_full.txt "ncu --set full --clock-control none, $3 on $2 (python scripts/one_frame.py $2 2, launch $4 of that kernel)" > /dev/null 2>&1
  NCU_SECTION=$5 python scripts/ncu_lines.py /tmp/prof_$1.ncu-rep $6 softray_b200/csrc/sr_device.cuh 24 > gpurun_out/${TAG}_$1_by_line.txt 2>&1
  rm -f /tmp/prof_$1.ncu-rep
}
cap search_config5 config5 '^k_search$' 5 k_searchILi0ELi4 sr_wave.sm_100a.cubin
cap hit_config5 config5 '^k_hit$' 5 k_hitILi0ELi3 sr_wave.sm_100a.cubin
cap walk_config5 config5 '^k_shadow_walk$' 5 k_shadow_walkILi4 sr_wave.sm_100a.cubin
cap search_config4 config4 '^k_search$' 24 k_searchILi0ELi4 sr_wave.sm_100a.cubin
cap cone_config3 config3 '^k_shadow$' 2 k_shadowILi4 sr_wave.sm_100a.cubin
cap walk_config3 config3 '^k_shadow_walk$' 2 k_shadow_walkILi4 sr_wave.sm_100a.cubin
cap fused_config2 config2 'render_kernel' 1 render_kernel sr_render.sm_100a.cubin
ls gpurun_out/${TAG}_* | wc -l
