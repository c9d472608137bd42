mkdir -p gpurun_out/r01y
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r01y/pytest_gpu.log 2>&1; tail -4 gpurun_out/r01y/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qgemm_ -s 128 -c 128 --csv --log-file gpurun_out/r01y/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01y/ncu_launch.log 2>&1
tail -3 gpurun_out/r01y/launches.csv | cut -c1-300
