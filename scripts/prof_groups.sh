M="gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers"
for cfg in "32:group=32" "32:group=16" "32:group=8,carveout=100,slice=3584" "16:group=32" "16:group=16" "16:group=8"; do
  sz=${cfg%%:*}; op=${cfg#*:}
  LUX_SIZE=$sz LUX_OPTS=$op timeout 200 ncu --metrics $M --section WarpStateStats --clock-control none -k regex:lux_step_kernel -s 806 -c 2 --csv --log-file gpurun_out/g_${sz}_${op//[=,]/_}.csv python scripts/prof_target.py 2 > /dev/null 2>&1
done
ls gpurun_out
