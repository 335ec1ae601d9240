#!/bin/bash
# ncu evidence for round 1 (run under gpurun): launch lists + full captures of the dominant conv kernels.
set -x
O=gpurun_out
python tools/run_once.py SegCD 2 1024 2 > $O/plain_segcd.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_r1_segcd_1024_p2.csv python tools/run_once.py SegCD 2 1024 2 > $O/ncu_l1.log 2>&1
for s in 2 18 42 44; do
  ncu --set full --clock-control none --import-source on -k regex:conv_ws -s $s -c 1 -f -o $O/prof_r1_segcd_conv$s python tools/run_once.py SegCD 2 1024 2 > $O/ncu_f$s.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:segcd_head -c 1 -f -o $O/prof_r1_segcd_head python tools/run_once.py SegCD 2 1024 2 > $O/ncu_fh.log 2>&1
python tools/run_once.py SNUNet_ECAM 32 256 32 > $O/plain_snunet.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_r1_snunet_256_p32.csv python tools/run_once.py SNUNet_ECAM 32 256 32 > $O/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_ws -s 38 -c 1 -f -o $O/prof_r1_snunet_conv38 python tools/run_once.py SNUNet_ECAM 32 256 32 > $O/ncu_f38.log 2>&1
