export PROF_SLOTS=128 PROF_REPS=2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:^sync_select_kernel -s 1 -c 1 -f -o gpurun_out/ncu_sel_r2j python tools/prof_sync.py > gpurun_out/ncu_sel_r2j.log 2>&1
python tools/ncu_lines.py gpurun_out/ncu_sel_r2j.ncu-rep 45 > gpurun_out/ncu_lines_sel_r2j.txt 2>&1
ncu -i gpurun_out/ncu_sel_r2j.ncu-rep --page raw --csv > gpurun_out/ncu_raw_sync_select_kernel_r2j.csv 2>/dev/null
rm -f gpurun_out/ncu_sel_r2j.ncu-rep
cat gpurun_out/ncu_lines_sel_r2j.txt
