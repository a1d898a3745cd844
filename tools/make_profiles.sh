#!/bin/bash
# run on the GPU box (through gpurun): the artefacts profiles/ is made of.  usage: bash tools/make_profiles.sh <tag>
# ncu is restricted to this library's kernels (-k): profiling torch's data-generation kernels of the benches takes minutes.
tag=${1:-r02}
out=gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
K='regex:cf_lowres|alpha_up_fuzzy|cross_march|trimap_|bgstep_frame|median|blend|ratio_flags|degenerate|get_fg|bgdiff'
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err
for wl in cf_trimap_1080p green_4k replace_1080p bgstep_4k; do
  C=""; [ $wl = bgstep_4k ] && C="-c 96"     # the 6 person-mask runs (16 launches each); the dense-mask runs that follow are left out
  ncu --metrics $M --clock-control none -k "$K" $C --csv --log-file $out/${tag}_${wl}_launches.csv python tools/bench_configs.py --only $wl --steps 1 --warmup 1 --no-cpu --no-e2e --frames 120 > $out/${tag}_ncu_${wl}.log 2>&1
done
# the bench line's own bgstep clip (500 frames: 64 launches per run, 6 person-mask runs) for traffic.json; SKIP500=1 leaves it out
[ "${SKIP500:-0}" = 1 ] || ncu --metrics $M --clock-control none -k "$K" -c 384 --csv --log-file $out/${tag}_bgstep_4k_500_launches.csv python tools/bench_configs.py --only bgstep_4k --steps 1 --warmup 1 --no-cpu --no-e2e > $out/${tag}_ncu_bgstep_4k_500.log 2>&1
ncu --metrics $M --clock-control none -k "$K" --csv --log-file $out/${tag}_median_launches.csv python bench.py --steps 2 --warmup 3 --configs none --no-cpu > $out/${tag}_ncu_median.log 2>&1
ls -la $out/${tag}_*
