#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "green or cf_ or config1 or config3 or fused or colorfilter" > $O/r02t_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02t_pytest.log
tail -3 $O/r02t_pytest.log
python tools/bench_configs.py --only cf_trimap_1080p,green_4k --no-cpu --no-e2e --steps 20 > $O/r02t_cfg.json 2> $O/r02t_cfg.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02t_*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d.get('workload'), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d.get('bit_exact'))
PY
