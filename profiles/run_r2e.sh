mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2e_tests.log; cat gpurun_out/r2e_tests.log
python tests/stress_nms.py > gpurun_out/r2e_stress.log 2>&1; tail -n 3 gpurun_out/r2e_stress.log
python profiles/nms_regimes.py > gpurun_out/r2e_nms_regimes.txt 2>&1; cat gpurun_out/r2e_nms_regimes.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 300 gpurun_out/r2e_bench.err
i=0
for cfg in "6000 1000 0" "6000 1000 1"; do
  i=$((i+1))
  python profiles/prof_nms.py $cfg 2 > gpurun_out/plain_nms.log 2>&1 &&
  ncu --set full --clock-control none -k regex:nms_ -c 12 -o gpurun_out/prof_r2_nms6k_$i python profiles/prof_nms.py $cfg 2 > gpurun_out/ncu_nms6k_$i.log 2>&1
  tail -n 1 gpurun_out/ncu_nms6k_$i.log
done
