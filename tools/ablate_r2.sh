#!/bin/bash
# stage ablation of the wf=8 hot kernels (needs the -DTEM_ABLATION library: python -m transfer_em_b200.build --ablation)
# usage: tools/ablate_r2.sh "0 1 29" [op names...]
export TEM_ABLATION_LIB=1
BITS=${1:-"0 1 2 4 8 12 16 28 29"}; shift
OPS=${@:-g7.wgrad g1.wgrad g10.wgrad g3.wgrad g8.wgrad g5.wgrad g1.fwd g10.fwd g7.fwd g1.dgrad g3.fwd g5.fwd}
for bits in $BITS; do
  echo "== TEM_S2_DBG=$bits (1 no epilogue, 2 no atomics, 4 no x/input loads, 8 no g loads, 16 no MMAs)"
  TEM_S2_DBG=$bits python tools/op_bench.py $OPS 2>&1 | grep -v Warn
done
