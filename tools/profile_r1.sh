#!/bin/bash
# ncu evidence for round 1 (run under gpurun): launch lists + full captures of the dominant kernels of every configuration.
# Each program is first run WITHOUT ncu; ncu only wraps a command line that has just exited 0.
set -x
O=gpurun_out
run() { python tools/run_once.py "$@"; }
full() {  # full <out name> <kernel regex> <skip> -- <run_once args>
  local name=$1 rx=$2 skip=$3; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/$name python tools/run_once.py "$@" > $O/ncu_$name.log 2>&1
}
# C3: SegCD-ResNet34, 2 pairs of 1024x1024 per launch
run SegCD 2 1024 2 > $O/plain_segcd.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_r1_segcd_1024_p2.csv python tools/run_once.py SegCD 2 1024 2 > $O/ncu_l1.log 2>&1
full prof_r1_segcd_conv2 conv_ws 2 -- SegCD 2 1024 2       # encoder.layer1.0.conv2 (+identity, ReLU)
full prof_r1_segcd_conv18 conv_ws 18 -- SegCD 2 1024 2     # encoder.layer3.1.conv1
full prof_r1_segcd_conv42 conv_ws 42 -- SegCD 2 1024 2     # decoder.blocks.3.conv1 (phases folded into N)
full prof_r1_segcd_conv44 conv_ws 44 -- SegCD 2 1024 2     # decoder.blocks.4.conv1 (phases folded into N)
full prof_r1_segcd_head segcd_head 0 -- SegCD 2 1024 2
# C2: SNUNet-ECAM, 32 pairs of 256x256 per launch
run SNUNet_ECAM 32 256 32 > $O/plain_snunet.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_r1_snunet_256_p32.csv python tools/run_once.py SNUNet_ECAM 32 256 32 > $O/ncu_l2.log 2>&1
full prof_r1_snunet_conv38 conv_ws 38 -- SNUNet_ECAM 32 256 32   # conv0_4.conv1
# C4: ChangeGNNV1, 8 pairs of 256x256
run ChangeGNNV1 8 256 8 > $O/plain_gnn.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1_changegnn_256_p8.csv python tools/run_once.py ChangeGNNV1 8 256 8 > $O/ncu_l3.log 2>&1
full prof_r1_gnn_knn knn_graph 0 -- ChangeGNNV1 8 256 8
# C5: ChangeFormerV6, 8 pairs of 256x256
run ChangeFormerV6 8 256 8 > $O/plain_cf.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1_changeformer_256_p8.csv python tools/run_once.py ChangeFormerV6 8 256 8 > $O/ncu_l4.log 2>&1
full prof_r1_cf_dwconv dwconv3x3 0 -- ChangeFormerV6 8 256 8
full prof_r1_cf_attn sr_attention 0 -- ChangeFormerV6 8 256 8
# summarise on the box (the .ncu-rep files together exceed gpurun's 64 MiB pull limit) and keep only three reports
python tools/ncu_summary.py $O/r1_ncu_full_summary.csv \
  segcd_r34_1024_b16:encoder.layer1.0.conv2:2=$O/prof_r1_segcd_conv2.ncu-rep \
  segcd_r34_1024_b16:encoder.layer3.1.conv1:2=$O/prof_r1_segcd_conv18.ncu-rep \
  segcd_r34_1024_b16:decoder.blocks.3.conv1:2=$O/prof_r1_segcd_conv42.ncu-rep \
  segcd_r34_1024_b16:decoder.blocks.4.conv1:2=$O/prof_r1_segcd_conv44.ncu-rep \
  segcd_r34_1024_b16:segmentation_head:2=$O/prof_r1_segcd_head.ncu-rep \
  snunet_256_b64:conv0_4.conv1:32=$O/prof_r1_snunet_conv38.ncu-rep \
  changegnn_v1_256_b32:encoder.backbone.0.0.graph:8=$O/prof_r1_gnn_knn.ncu-rep \
  changeformer_v6_256_b32:Tenc_x2.block1.0.mlp.dwconv:8=$O/prof_r1_cf_dwconv.ncu-rep \
  changeformer_v6_256_b32:Tenc_x2.block1.0.attn.softmax:8=$O/prof_r1_cf_attn.ncu-rep > $O/r1_ncu_full_summary.txt 2>&1
for n in gnn_knn cf_dwconv cf_attn segcd_head; do python tools/ncu_src.py $O/prof_r1_$n.ncu-rep 40 > $O/r1_src_$n.txt 2>&1; done
rm -f $O/prof_r1_segcd_conv*.ncu-rep $O/prof_r1_snunet_conv38.ncu-rep $O/prof_r1_segcd_head.ncu-rep $O/prof_r1_cf_attn.ncu-rep $O/prof_r1_cf_dwconv.ncu-rep
