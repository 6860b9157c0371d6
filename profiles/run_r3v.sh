mkdir -p gpurun_out
timeout 1200 python benchmarks/sweep.py --reps 10 > gpurun_out/sweep_r2_final.md 2> gpurun_out/sweep_r2_final.jsonl; tail -n 9 gpurun_out/sweep_r2_final.md
