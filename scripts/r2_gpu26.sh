CMD="python bench.py --steps 3 --warmup 3 --burnin 60 --no-cpu-baseline --no-other-configs"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
