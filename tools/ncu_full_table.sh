#!/bin/bash
# tools/ncu_full_table.sh TAG : ncu --set full of every kernel of the third frame (plain launches), raw CSV page only
# (first half of tools/ncu_full.sh).  Run under gpurun AFTER the same command has exited 0 without ncu.
TAG=$1; mkdir -p gpurun_out/r2
OSEP_NO_GRAPH=1 python tools/one_frame.py turbine100k 3 > gpurun_out/r2/${TAG}_plain.log 2>&1 || exit 1
OSEP_NO_GRAPH=1 ncu --set full --clock-control none --launch-skip 86 -c 43 -f -o /tmp/${TAG}_full python tools/one_frame.py turbine100k 3 > gpurun_out/r2/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/r2/${TAG}_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2/${TAG}_*
