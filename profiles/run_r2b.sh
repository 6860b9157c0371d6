mkdir -p gpurun_out
python profiles/prof_os.py 14 2 16 16 2 3 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:car3d_grad_image_os -s 2 -c 1 -o gpurun_out/prof_r2e_dbg2 python profiles/prof_os.py 14 2 16 16 2 3 > gpurun_out/ncu1.log 2>&1
python profiles/prof_os.py 14 1 16 16 2 3 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:car3d_grad_image_os -s 2 -c 1 -o gpurun_out/prof_r2e_dbg1 python profiles/prof_os.py 14 1 16 16 2 3 > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log
