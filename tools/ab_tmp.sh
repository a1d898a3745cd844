#!/bin/bash
set -u
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > $O/r02v_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02v_pytest.log
tail -3 $O/r02v_pytest.log
timeout 200 python tools/bench_configs.py --only green_4k,bgstep_4k --no-cpu --no-e2e --steps 10 > $O/r02v_cfg.json 2> $O/r02v_cfg.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02v_*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d.get('workload'), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d.get('bit_exact'), (d.get('dense_masks') or {}).get('ms_per_step'))
PY
