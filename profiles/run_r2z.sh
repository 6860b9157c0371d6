mkdir -p gpurun_out
python profiles/e2e_order.py > gpurun_out/r2z_e2e_order.txt 2>&1; cat gpurun_out/r2z_e2e_order.txt
