#!/bin/bash
# Source-level stall captures (ncu --set full --import-source on) of three C2 kernels; summaries by tools/ncu_src.py.
O=gpurun_out
python tools/run_once.py SNUNet_ECAM 64 256 64 > $O/plain_src.log 2>&1 || { echo plain run failed; exit 1; }
cap() {  # cap <name> <conv launch index>
  ncu --set full --clock-control none --import-source on -k regex:conv_ws -s $2 -c 1 -f -o $O/prof_r2_$1 python tools/run_once.py SNUNet_ECAM 64 256 64 > $O/ncu_src_$1.log 2>&1
  python tools/ncu_src.py $O/prof_r2_$1.ncu-rep 25 > $O/r2_src_$1.txt 2>&1
  head -3 $O/r2_src_$1.txt; tail -9 $O/r2_src_$1.txt
  rm -f $O/prof_r2_$1.ncu-rep
}
cap conv1_1_conv2 15
cap conv0_4_conv1 38
cap conv1_0_conv2 3
