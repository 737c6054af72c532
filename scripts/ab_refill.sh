#!/bin/bash
# own-block refill experiment: ticket chunk x refill threshold on the 8-frame cold job (scripts/ab.py)
for v in ab_base ab_chunk64 ab_chunk128; do
  for t in 32 24 16 8; do
    CBQ_LIBRARY=cubiquity_b200/lib/$v.so python scripts/ab.py --tag ${v}_t$t --frames 8 --skip lod,pt,replay,random --opt refill_threshold=$t 2>&1 | tail -1
  done
done
