#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:getrf_panel_cluster -s 3 -c 1 -o gpurun_out/r2_prof_getrf_panel -f \
    python probes/getrf_one.py 4096 f64 > gpurun_out/r2_ncu_getrf_panel.log 2>&1
tail -2 gpurun_out/r2_ncu_getrf_panel.log
python probes/ncu_summarise.py rep gpurun_out/r2_prof_getrf_panel.ncu-rep gpurun_out/r2_ncu_getrf_panel_summary.csv; head -40 gpurun_out/r2_ncu_getrf_panel_summary.csv | cut -c1-160
timeout 200 ncu --set full --clock-control none --import-source on -k regex:laswp_apply -s 0 -c 1 -o gpurun_out/r2_prof_laswp -f \
    python probes/getrf_one.py 16384 f64 > gpurun_out/r2_ncu_laswp.log 2>&1
python probes/ncu_summarise.py rep gpurun_out/r2_prof_laswp.ncu-rep gpurun_out/r2_ncu_laswp_summary.csv; head -40 gpurun_out/r2_ncu_laswp_summary.csv | cut -c1-160
