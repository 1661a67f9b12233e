#!/bin/bash
# one ncu --set full capture of the cell-tile pair kernel and of the build kernel (N = 1e6 LJ lattice)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_force_tile|k_build_cell' -s 2 -c 4 \
    -o gpurun_out/${1:-ncu_tile} -f python tools/tile_probe.py --steps 12 > gpurun_out/ncu_tile.log 2>&1
tail -3 gpurun_out/ncu_tile.log
ncu -i gpurun_out/${1:-ncu_tile}.ncu-rep --page raw --csv > gpurun_out/${1:-ncu_tile}_raw.csv 2>/dev/null
ls -la gpurun_out/${1:-ncu_tile}*
