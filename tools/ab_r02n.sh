#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "blend or replace or composite or config4" > $O/r02n_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02n_pytest.log
tail -3 $O/r02n_pytest.log
for v in 0 1; do
  VU_BLEND_FP64=$v python tools/bench_configs.py --only replace_1080p --no-cpu --no-e2e --steps 20 > $O/r02n_replace_fp64_$v.json 2> $O/r02n_replace_fp64_$v.err
done
for v in 6 24 48; do
  VU_TM_WARPS=$v python tools/bench_configs.py --only cf_trimap_1080p --no-cpu --no-e2e --steps 20 > $O/r02n_tmwarps_$v.json 2> $O/r02n_tmwarps_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n_*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d.get('workload'), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d.get('bit_exact'), (d.get('realistic_matte') or {}).get('ms_per_step'))
PY
