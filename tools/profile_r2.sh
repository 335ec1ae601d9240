#!/bin/bash
# Round-2 per-layer ncu tables (run under gpurun).  Each program is first run WITHOUT ncu; ncu wraps the same command line only after it exited 0.
O=gpurun_out
M=gpu__time_duration.sum,sm__mem_tensor_writes_op_utcmma.sum,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size
layers() {  # layers <tag> <run_once args...>
  local tag=$1; shift
  STCD_DUMP_OPS=$O/ops_$tag.json python tools/run_once.py "$@" > $O/plain_$tag.log 2>&1 || { echo "plain run failed: $tag"; tail -3 $O/plain_$tag.log; return 1; }
  STCD_PROFILE_RANGE=1 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/ncu_$tag.csv python tools/run_once.py "$@" > $O/ncu_$tag.log 2>&1 || { echo "ncu failed: $tag"; tail -3 $O/ncu_$tag.log; return 1; }
  python tools/ncu_layers.py $O/ncu_$tag.csv $O/ops_$tag.json $O/r2_layers_$tag > /dev/null && tail -1 $O/r2_layers_$tag.txt
}
for spec in "$@"; do
  layers $spec
done
