#!/bin/bash
# pipeline-granularity sweep of conv3_tc3 (slices per step x ring slots) at the bench shapes
for sb in 1 2 3 6 8; do for ring in 2 3 4 6 8; do
  echo "== TEM_TC3_SB=$sb TEM_TC3_RING=$ring"
  TEM_TC3_SB=$sb TEM_TC3_RING=$ring python tools/op_bench.py g1.fwd g1.dgrad g10.fwd g7.fwd 2>&1 | grep -v Warn
done; done
