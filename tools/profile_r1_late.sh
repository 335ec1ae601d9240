#!/bin/bash
# ncu evidence for the kernels rewritten at the end of round 1 (tensor-core attention and SegCD head, depth-wise conv, kNN with FFMA2,
# GELU epilogue); run under gpurun.  Each program is first run WITHOUT ncu; ncu only wraps a command line that has just exited 0.
set -x
O=gpurun_out
full() {  # full <out name> <kernel regex> <skip> -- <run_once args>
  local name=$1 rx=$2 skip=$3; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/$name python tools/run_once.py "$@" > $O/ncu_$name.log 2>&1
}
python tools/run_once.py ChangeFormerV6 8 256 8 > $O/plain_cf.log 2>&1 || exit 1
full prof_r1l_attn sr_attention_mma 0 -- ChangeFormerV6 8 256 8
full prof_r1l_dwconv dwconv3x3 0 -- ChangeFormerV6 8 256 8
python tools/run_once.py SegCD 2 1024 2 > $O/plain_seg.log 2>&1 || exit 1
full prof_r1l_head segcd_head_mma 0 -- SegCD 2 1024 2
python tools/run_once.py ChangeGNNV1 8 256 8 > $O/plain_gnn.log 2>&1 || exit 1
full prof_r1l_knn knn_graph 0 -- ChangeGNNV1 8 256 8
full prof_r1l_fc1 conv_ws 9 -- ChangeGNNV1 8 256 8            # encoder.backbone.0.1.fc1 (80 -> 320, GELU epilogue, N = 64 tiles)
python tools/ncu_summary.py $O/r1_late_ncu_full_summary.csv \
  changeformer_v6_256_b32:Tenc_x2.block1.0.attn.softmax:8=$O/prof_r1l_attn.ncu-rep \
  changeformer_v6_256_b32:Tenc_x2.block1.0.mlp.dwconv:8=$O/prof_r1l_dwconv.ncu-rep \
  segcd_r34_1024_b16:segmentation_head:2=$O/prof_r1l_head.ncu-rep \
  changegnn_v1_256_b32:encoder.backbone.0.0.graph:8=$O/prof_r1l_knn.ncu-rep \
  changegnn_v1_256_b32:conv9:8=$O/prof_r1l_fc1.ncu-rep > $O/r1_late_ncu_full_summary.txt 2>&1
rm -f $O/prof_r1l_*.ncu-rep
