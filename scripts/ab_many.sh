#!/bin/bash
# A/B over library variants: timing (scripts/ab.py, 8-frame batch so the end-of-kernel tail does not blur the comparison)
# and executed warp instructions of one 1080p launch (ncu, 3 metrics).  usage: scripts/ab_many.sh <variant>...
for v in "$@"; do
  lib=cubiquity_b200/lib/$v.so
  CBQ_LIBRARY=$lib python scripts/ab.py --tag $v --frames 8 --skip lod,pt,replay 2>&1 | tail -1 | tee gpurun_out/ab8_$v.json
  CBQ_LIBRARY=$lib ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none \
      -k regex:tracePersistent --launch-skip 2 -c 1 --csv python scripts/one_frame.py 2>/dev/null | grep -v "^==" | tail -3 | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | tee gpurun_out/inst_$v.txt
done
