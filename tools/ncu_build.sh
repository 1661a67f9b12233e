#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_build_cell' -s 1 -c 1 \
    -o gpurun_out/${1:-ncu_build} -f python tools/tile_probe.py --steps 8 > gpurun_out/ncu_build.log 2>&1
tail -3 gpurun_out/ncu_build.log
ncu -i gpurun_out/${1:-ncu_build}.ncu-rep --page raw --csv > gpurun_out/${1:-ncu_build}_raw.csv 2>/dev/null
ncu -i gpurun_out/${1:-ncu_build}.ncu-rep --page source --csv > gpurun_out/${1:-ncu_build}_source.csv 2>/dev/null
ls -la gpurun_out/${1:-ncu_build}*
