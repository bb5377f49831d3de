#!/bin/bash
# ncu --set full of the small-pass kernels inside a B = 1 forward (graphs off): fusion layer cluster kernel, graph head layer 1
O=gpurun_out/r02s; mkdir -p $O
B="python tools/b1_forward.py 5 4"
HMV_NO_GRAPH=1 $B > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
cap() {  # name, kernel regex, launches to skip
  HMV_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o $O/$1 $B > $O/$1.log 2>&1; echo "ncu $1 rc $?"
}
cap fusion_cluster_l0 fusion_block_cluster_kernel 10
cap fusion_cluster_l3 fusion_block_cluster_kernel 13
cap gcn_l1_small gcn_l1_small_kernel 2
ls -la $O
