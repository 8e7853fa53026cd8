# Round-2 profile capture (run on the GPU box from the repository root; writes gpurun_out/).
#   1. launch list of the bench step chain        2. ncu --set full of the two step-kernel classes
#   3. ncu --set full of the observation kernel inside the config-4 loop
# (five lux_ launches per step since r02f: classify, big class, small class, oversize launch, reset + action stream;
#  400 burn-in + 3 warm-up steps are skipped)
T=${1:-r02b}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lux_ -s 2015 -c 50 --csv \
    --log-file gpurun_out/${T}_launches.csv python scripts/prof_target.py 4 > gpurun_out/${T}_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lux_step_kernel -s 1209 -c 3 \
    -o gpurun_out/${T}_step python scripts/prof_target.py 4 > gpurun_out/${T}_s.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lux_observe -s 70 -c 1 \
    -o gpurun_out/${T}_obs python scripts/c4_breakdown.py > gpurun_out/${T}_o.log 2>&1
ls -la gpurun_out
