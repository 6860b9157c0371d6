mkdir -p gpurun_out
python profiles/prof_os.py 14 2 8 2 1 3 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:car3d_grad_image_os -s 2 -c 1 -o gpurun_out/prof_r2c_v2 python profiles/prof_os.py 14 2 8 2 1 3 > gpurun_out/ncu1.log 2>&1
tail -n 3 gpurun_out/ncu1.log
