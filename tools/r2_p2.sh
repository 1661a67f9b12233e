#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'k_build_cell|k_tile_localize_masks' -s 1 -c 4 \
    -o gpurun_out/r2p_reb -f python tools/tile_probe.py --steps 20 > gpurun_out/r2p_ncu2.log 2>&1
tail -1 gpurun_out/r2p_ncu2.log
ncu -i gpurun_out/r2p_reb.ncu-rep --page raw --csv > gpurun_out/r2p_rebuild_kernels_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/r2p_reb.ncu-rep
