#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "trimap or green or bgstep or config1 or config3 or config5" > $O/r02r_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02r_pytest.log
tail -3 $O/r02r_pytest.log
python tools/bench_configs.py --only cf_trimap_1080p,green_4k --no-cpu --no-e2e --steps 20 > $O/r02r_cfg.json 2> $O/r02r_cfg.err
python tools/bench_configs.py --only bgstep_4k --frames 120 --no-cpu --no-e2e --steps 10 > $O/r02r_bgstep.json 2> $O/r02r_bgstep.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02r_*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d.get('workload'), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d.get('bit_exact'), (d.get('dense_masks') or {}).get('ms_per_step'))
PY
