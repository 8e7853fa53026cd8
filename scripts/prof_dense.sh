# Dense-batch profile capture (run on the GPU box from the repository root; writes gpurun_out/): the dense replay-state
# pool of scripts/sweep.py (LUX_SWEEP_DENSE=1), launch list + ncu --set full of the step-kernel launches of two steps.
T=${1:-r02f_dense}
LUX_SWEEP_DENSE=1 LUX_SWEEP_STEPS=10 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lux_ -s 100 -c 30 --csv \
    --log-file gpurun_out/${T}_launches.csv python scripts/sweep.py > gpurun_out/${T}_l.log 2>&1
LUX_SWEEP_DENSE=1 LUX_SWEEP_STEPS=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lux_step_kernel -s 60 -c 6 \
    -o gpurun_out/${T}_step python scripts/sweep.py > gpurun_out/${T}_s.log 2>&1
ls -la gpurun_out | tail -5
