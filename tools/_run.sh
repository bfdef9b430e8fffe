ncu --set full --cache-control none --clock-control none --import-source on -k regex:step_group -s 3 -c 4 -f -o gpurun_out/prof_c5 python tools/prof_c5.py > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_ncu.log
