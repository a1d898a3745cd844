#!/bin/bash
# the round's closing run on one B200: GPU tests, smoke, bench lines of both arms, ncu launch lists, ncu --set full of the two late kernels
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02z_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02z_pytest_gpu.log
tail -2 $O/r02z_pytest_gpu.log
python __graft_entry__.py smoke > $O/r02z_smoke.log 2>&1; tail -1 $O/r02z_smoke.log
SKIP500=1 bash tools/make_profiles.sh r02z > $O/r02z_make.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend16_int -c 2 -o $O/r02z_blend_int python tools/bench_configs.py --only replace_1080p --steps 1 --warmup 0 --no-cpu --no-e2e > $O/r02z_ncu_blend.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trimap_up_bits -c 1 -o $O/r02z_trimap_up python tools/bench_configs.py --only cf_trimap_1080p --steps 1 --warmup 0 --no-cpu --no-e2e > $O/r02z_ncu_trimap_up.log 2>&1
head -c 600 $O/r02z_bench.json; echo; tail -3 $O/r02z_bench.err
ls -la $O/r02z_* | head -30
