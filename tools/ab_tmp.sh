#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "blend or replace or composite or config4" > $O/r02s_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02s_pytest.log
tail -3 $O/r02s_pytest.log
python tools/bench_configs.py --only replace_1080p --no-cpu --no-e2e --steps 20 > $O/r02s_replace.json 2> $O/r02s_replace.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02s_*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d.get('workload'), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d.get('bit_exact'), (d.get('realistic_matte') or {}).get('ms_per_step'))
PY
