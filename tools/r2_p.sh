#!/bin/bash
# final ncu captures of round 2: pair kernel (raw + source), mask-mode build and localize kernels (raw)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_force_tile' -s 3 -c 2 \
    -o gpurun_out/r2p_tile -f python tools/tile_probe.py --steps 10 > gpurun_out/r2p_ncu1.log 2>&1
tail -1 gpurun_out/r2p_ncu1.log
ncu -i gpurun_out/r2p_tile.ncu-rep --page raw --csv > gpurun_out/r2p_force_tile_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r2p_tile.ncu-rep --page source --csv > gpurun_out/r2p_force_tile_ncu_source.csv 2>/dev/null
rm -f gpurun_out/r2p_tile.ncu-rep
ncu --set full --clock-control none -k regex:'k_build_cell|k_tile_localize_masks|k_verlet21' -s 3 -c 4 \
    -o gpurun_out/r2p_reb -f python tools/tile_probe.py --steps 20 > gpurun_out/r2p_ncu2.log 2>&1
tail -1 gpurun_out/r2p_ncu2.log
ncu -i gpurun_out/r2p_reb.ncu-rep --page raw --csv > gpurun_out/r2p_rebuild_kernels_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/r2p_reb.ncu-rep
