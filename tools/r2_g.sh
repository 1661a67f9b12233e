#!/bin/bash
mkdir -p gpurun_out
PARM_B200_BUILD_MASKS=1 ncu --set full --clock-control none --import-source on -k regex:'k_tile_localize_masks' -s 0 -c 1 \
    -o gpurun_out/r2g_loc -f python tools/tile_probe.py --steps 14 > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
ncu -i gpurun_out/r2g_loc.ncu-rep --page source --csv > gpurun_out/r2g_loc_source.csv 2>/dev/null
rm -f gpurun_out/r2g_loc.ncu-rep
