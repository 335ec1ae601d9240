#!/bin/bash
# ncu evidence for the families added late in round 1 (DTCDSCN, BIT, IFNet); run under gpurun.
# Each program is first run WITHOUT ncu; ncu only wraps a command line that has just exited 0.
set -x
O=gpurun_out
full() {  # full <out name> <kernel regex> <skip> -- <run_once args>
  local name=$1 rx=$2 skip=$3; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/$name python tools/run_once.py "$@" > $O/ncu_$name.log 2>&1
}
python tools/run_once.py CDNet_model 8 256 8 > $O/plain_dt.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1_dtcdscn_256_p8.csv python tools/run_once.py CDNet_model 8 256 8 > $O/ncu_l5.log 2>&1
full prof_r1_dt_gate gate_apply 0 -- CDNet_model 8 256 8
python tools/run_once.py BASE_Transformer 8 256 8 > $O/plain_bit.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1_bit_dd8_256_p8.csv python tools/run_once.py BASE_Transformer 8 256 8 > $O/ncu_l6.log 2>&1
full prof_r1_bit_decoder bit_decoder 0 -- BASE_Transformer 8 256 8
full prof_r1_bit_tokenizer bit_tokenizer 0 -- BASE_Transformer 8 256 8
python tools/run_once.py DSIFN 8 256 8 > $O/plain_if.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1_ifnet_256_p8.csv python tools/run_once.py DSIFN 8 256 8 > $O/ncu_l7.log 2>&1
full prof_r1_if_conv2 conv_ws 1 -- DSIFN 8 256 8             # t1_base.features.2 (64 -> 64 at full resolution, + fused max-pool)
full prof_r1_if_ca_apply ca_apply 9 -- DSIFN 8 256 8         # ca5, first segment
python tools/ncu_summary.py $O/r1_f4_ncu_full_summary.csv \
  dtcdscn_256_b32:encoder1.0.se:8=$O/prof_r1_dt_gate.ncu-rep \
  bit_dd8_256_b32:bit.decoder:8=$O/prof_r1_bit_decoder.ncu-rep \
  bit_dd8_256_b32:bit.tokenizer:8=$O/prof_r1_bit_tokenizer.ncu-rep \
  ifnet_256_b16:t1_base.features.2:8=$O/prof_r1_if_conv2.ncu-rep \
  ifnet_256_b16:ca5.apply0:8=$O/prof_r1_if_ca_apply.ncu-rep > $O/r1_f4_ncu_full_summary.txt 2>&1
python tools/ncu_src.py $O/prof_r1_bit_decoder.ncu-rep > $O/r1_src_bit_decoder.txt 2>&1
rm -f $O/prof_r1_dt_gate.ncu-rep $O/prof_r1_bit_tokenizer.ncu-rep $O/prof_r1_if_conv2.ncu-rep $O/prof_r1_if_ca_apply.ncu-rep $O/prof_r1_bit_decoder.ncu-rep
