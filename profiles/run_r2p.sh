mkdir -p gpurun_out
python profiles/bwd_stage_sweep.py > gpurun_out/r2p_bwd_stage.txt 2>&1; cat gpurun_out/r2p_bwd_stage.txt
