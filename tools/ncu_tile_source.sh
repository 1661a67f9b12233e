#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_force_tile' -s 3 -c 1 \
    -o gpurun_out/${1:-ncu_tile_src} -f python tools/tile_probe.py --steps 10 > gpurun_out/ncu_tile_src.log 2>&1
tail -2 gpurun_out/ncu_tile_src.log
ncu -i gpurun_out/${1:-ncu_tile_src}.ncu-rep --page source --csv > gpurun_out/${1:-ncu_tile_src}_source.csv 2>/dev/null
ncu -i gpurun_out/${1:-ncu_tile_src}.ncu-rep --page raw --csv > gpurun_out/${1:-ncu_tile_src}_raw.csv 2>/dev/null
ls -la gpurun_out/${1:-ncu_tile_src}*
