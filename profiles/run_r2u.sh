mkdir -p gpurun_out
python profiles/bwd_occupancy_ab.py > gpurun_out/r2u_bwd_occ.txt 2>&1; cat gpurun_out/r2u_bwd_occ.txt
