#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'^k_force$' -s 2 -c 1 \
    -o gpurun_out/r2o_cfg4f -f python tools/config4_probe.py --steps 6 > gpurun_out/r2o_ncu.log 2>&1
tail -2 gpurun_out/r2o_ncu.log
ncu -i gpurun_out/r2o_cfg4f.ncu-rep --page raw --csv > gpurun_out/r2o_cfg4f_raw.csv 2>/dev/null
ncu -i gpurun_out/r2o_cfg4f.ncu-rep --page source --csv > gpurun_out/r2o_cfg4f_source.csv 2>/dev/null
rm -f gpurun_out/r2o_cfg4f.ncu-rep
