#!/bin/bash
# Development: run a bench script (default tools/alt_bench.py) against every variant library under
# pwstablenet_b200/var/*.so.xz (or those named). VB_SCRIPT, VB_TAIL configure it.
cd "$(dirname "$0")/.."
libs=${@:-$(ls pwstablenet_b200/var/*.so.xz)}
mkdir -p /tmp/pwsvar
for l in $libs; do
  b=$(basename $l .xz)
  xz -d -k -c $l > /tmp/pwsvar/$b
  echo "== $b"
  PWS_LIB_PATH=/tmp/pwsvar/$b timeout 120 python ${VB_SCRIPT:-tools/alt_bench.py} 2>&1 | tail -${VB_TAIL:-12}
done
