#!/bin/bash
# run on the GPU box (through gpurun): the artefacts profiles/ is made of.  usage: bash tools/make_profiles.sh <tag>
tag=${1:-r02}
out=gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err
for wl in cf_trimap_1080p green_4k replace_1080p bgstep_4k; do
  ncu --metrics $M --clock-control none --csv --log-file $out/${tag}_${wl}_launches.csv python tools/bench_configs.py --only $wl --steps 1 --warmup 1 --no-cpu --no-e2e --frames 120 > $out/${tag}_ncu_${wl}.log 2>&1
done
ncu --metrics $M --clock-control none --csv --log-file $out/${tag}_median_launches.csv python bench.py --steps 2 --warmup 3 --configs none --no-cpu > $out/${tag}_ncu_median.log 2>&1
ncu --set full --clock-control none -k regex:"cf_lowres2_wide|alpha_up_fuzzy|cross_march|trimap_bits_kernel|trimap_up_bits" -s 6 -c 6 -o $out/${tag}_cf_kernels python tools/bench_configs.py --only cf_trimap_1080p --steps 1 --warmup 1 --no-cpu --no-e2e > $out/${tag}_ncu_full_cf.log 2>&1
ncu --set full --clock-control none -k regex:"cf_lowres4_wide|alpha_up_fuzzy" -s 2 -c 2 -o $out/${tag}_green_kernels python tools/bench_configs.py --only green_4k --steps 1 --warmup 1 --no-cpu --no-e2e > $out/${tag}_ncu_full_green.log 2>&1
ncu --set full --clock-control none -k regex:"bgstep_frame|blend16" -s 2 -c 3 -o $out/${tag}_bgstep_blend_kernels python tools/bench_configs.py --only replace_1080p,bgstep_4k --steps 1 --warmup 1 --no-cpu --no-e2e --frames 96 > $out/${tag}_ncu_full_bgstep.log 2>&1
ls -la $out/${tag}_*
