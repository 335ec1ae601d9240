set -x
O=gpurun_out
python tools/run_once.py SNUNet_ECAM 32 256 32 > $O/plain_snunet2.log 2>&1 || exit 1
for s in 12 15 11; do
ncu --set full --clock-control none --import-source on -k regex:conv_ws -s $s -c 1 -f -o $O/prof_snunet_$s python tools/run_once.py SNUNet_ECAM 32 256 32 > $O/ncu_snunet_$s.log 2>&1
python tools/ncu_src.py $O/prof_snunet_$s.ncu-rep 60 > $O/r1_src_snunet_conv$s.txt 2>&1
done
python tools/ncu_summary.py $O/r1_snunet_extra.csv snunet_256_b64:conv0_1.conv2:32=$O/prof_snunet_12.ncu-rep snunet_256_b64:conv1_1.conv2:32=$O/prof_snunet_15.ncu-rep snunet_256_b64:conv0_1.conv1:32=$O/prof_snunet_11.ncu-rep > $O/r1_snunet_extra.txt 2>&1
rm -f $O/prof_snunet_*.ncu-rep
