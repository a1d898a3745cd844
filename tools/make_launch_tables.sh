#!/bin/bash
# profiles/<round>_launch_tables.txt from the ncu launch lists tools/make_profiles.sh leaves in gpurun_out/.
# usage: bash tools/make_launch_tables.sh <tag of the csv files> > profiles/r02_launch_tables.txt
tag=${1:-r02}
d=gpurun_out
P1=2073600; P4=8294400
frames_of() {  # frames the csv covers: launches of the kernel that runs once per chunk x frames per chunk (grid.z or grid.y)
  python - "$1" "$2" <<'PY'
import csv, sys
rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")) if sys.argv[2] in r["Kernel Name"] and r["Metric Name"] == "gpu__time_duration.sum"]
print(sum(int(r["Grid Size"].strip("()").split(",")[-1 if "bgstep" in sys.argv[2] or "alpha_up" in sys.argv[2] else 1]) for r in rows))
PY
}
cat <<'TXT'
# per-kernel launch tables of the BASELINE configs (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
# -k <this library's kernels>; python tools/bench_configs.py --only <workload> --steps 1 --warmup 1 --no-cpu --no-e2e --frames 120 for the clip pipelines,
# python bench.py --steps 2 --warmup 3 --configs none --no-cpu for the headline; tools/make_profiles.sh + tools/make_launch_tables.sh).
# ncu times are cold-cache and serialised: compare shares and bytes, not absolutes.  P/frame = DRAM bytes per frame in units of the frame's pixel count.
TXT
n=$(frames_of $d/${tag}_cf_trimap_1080p_launches.csv alpha_up_fuzzy)
echo; echo "== cf_trimap_1080p (BASELINE configs[0]): $n frames in chunks of 50"
python tools/ncu_launch_table.py $d/${tag}_cf_trimap_1080p_launches.csv --pixels $P1 --frames $n
n=$(frames_of $d/${tag}_green_4k_launches.csv alpha_up_fuzzy)
echo; echo "== green_4k (BASELINE configs[2]): $n frames in chunks of 24"
python tools/ncu_launch_table.py $d/${tag}_green_4k_launches.csv --pixels $P4 --frames $n
echo; echo "== replace_1080p (BASELINE configs[3]): blend16 launches of 300 frames (uniform-random alpha, then the realistic matte)"
python tools/ncu_launch_table.py $d/${tag}_replace_1080p_launches.csv
n=$(frames_of $d/${tag}_bgstep_4k_launches.csv bgstep_frame)
echo; echo "== bgstep_4k (BASELINE configs[4], 120-frame clip on one GPU, person masks): $n frames in chunks of 24"
python tools/ncu_launch_table.py $d/${tag}_bgstep_4k_launches.csv --pixels $P4 --frames $n
echo; echo "== median_1080p (BASELINE configs[1], the bench headline)"
python tools/ncu_launch_table.py $d/${tag}_median_launches.csv
