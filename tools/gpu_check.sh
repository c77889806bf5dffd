#!/bin/bash
# tools/gpu_check.sh TAG : GPU parity tests + single-frame bench + ncu launch list of three frames (run under gpurun)
TAG=$1; mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/${TAG}_tests.log 2>&1; tail -3 gpurun_out/r2/${TAG}_tests.log
python bench.py --no-cpu --stream 0 --batch 0 --steps 40 > gpurun_out/r2/${TAG}_bench.json 2> gpurun_out/r2/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/${TAG}_launches.csv python tools/one_frame.py turbine100k 4 > gpurun_out/r2/${TAG}_ncu.log 2>&1
