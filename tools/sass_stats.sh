#!/bin/bash
# usage: tools/sass_stats.sh <object> <function-substring> [first-lines]
# static SASS opcode mix of one kernel of an object file
obj=$1; pat=$2; n=${3:-0}
cuobjdump -sass "$obj" | awk -v pat="$pat" '/Function :/{f = index($0, pat) > 0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -e 's#/\* 0x[0-9a-f]* \*/##' > /tmp/sass_stats.$$
wc -l < /tmp/sass_stats.$$
awk '{ if ($2 ~ /^@/) print $3; else print $2}' /tmp/sass_stats.$$ | sed 's/\.[A-Z0-9a-z_.]*//' | sort | uniq -c | sort -rn | head -16
[ "$n" -gt 0 ] && sed -n "1,${n}p" /tmp/sass_stats.$$
cp /tmp/sass_stats.$$ /tmp/last.sass; rm -f /tmp/sass_stats.$$
