mkdir -p gpurun_out
python profiles/prof_sep.py fwd 14 2 8 9 32 > gpurun_out/r2k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'sep_kernel' -s 2 -c 1 -o gpurun_out/prof_r2k_sep_fwd python profiles/prof_sep.py fwd 14 2 8 9 32 > gpurun_out/r2k_ncu.log 2>&1
tail -n 2 gpurun_out/r2k_ncu.log
