# plain run first (must exit 0), then the ncu passes of the same command line (one GPU, never multi-rank)
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --burnin 60 --no-cpu-baseline --no-other-configs"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_pre|k_dyn|k_scan|k_post|k_lidar|k_restore_bank" -s 310 -c 6 -f -o gpurun_out/prof_step $CMD > gpurun_out/ncu_step.log 2>&1
ls -la gpurun_out | head -20
