#!/bin/bash
# static evidence that the library is hand-written sm_100a code: per kernel, the counts of the instructions that matter
# (TMA bulk-tensor loads, mbarrier waits, byte-SIMD arithmetic).  usage: tools/sass_summary.sh > profiles/rNN_sass_summary.txt
lib=${1:-video_unscreen_b200/libvu_b200.so}
echo "# cuobjdump -sass $lib : kernels and instruction counts (UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier, VABSDIFF4 / IDP.4A / VIMNMX = byte SIMD, REDUX = warp reduce)"
cuobjdump -sass "$lib" | awk '
/Function :/ { if (name != "") print_row(); name = $3; total = 0; delete c; next }
/^[[:space:]]+\/\*[0-9a-f]{4}\*\// {
  total++
  if ($0 ~ /UTMALDG/) c["UTMALDG"]++
  if ($0 ~ /SYNCS/) c["SYNCS"]++
  if ($0 ~ /VABSDIFF4/) c["VABSDIFF4"]++
  if ($0 ~ /IDP\.4A/) c["IDP4A"]++
  if ($0 ~ /VIMNMX/) c["VIMNMX"]++
  if ($0 ~ /REDUX/) c["REDUX"]++
  if ($0 ~ /ATOM|RED\./) c["ATOM"]++
  if ($0 ~ /DADD|DMUL|DFMA/) c["FP64"]++
}
function print_row() {
  printf "%6d instr  UTMALDG %2d  SYNCS %2d  VABSDIFF4 %4d  IDP.4A %4d  VIMNMX %4d  REDUX %3d  ATOM %3d  FP64 %4d  %s\n", total, c["UTMALDG"], c["SYNCS"], c["VABSDIFF4"], c["IDP4A"], c["VIMNMX"], c["REDUX"], c["ATOM"], c["FP64"], name
}
END { if (name != "") print_row() }' | sed -e 's/_ZN2vu[0-9]*_GLOBAL__N__[0-9a-f]*_[0-9]*_//' | sort -k14
