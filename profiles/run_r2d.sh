mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2d_tests.log; cat gpurun_out/r2d_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 600 gpurun_out/r2d_bench.err
i=0
for cfg in "6000 1000 0" "6000 1000 1" "20000 2000 0" "20000 2000 1"; do
  i=$((i+1))
  python profiles/prof_nms.py $cfg 3 > gpurun_out/plain_nms.log 2>&1 &&
  ncu --set full --clock-control none -k regex:nms_ -s 12 -c 6 -o gpurun_out/prof_r2_nms_$i python profiles/prof_nms.py $cfg 3 > gpurun_out/ncu_nms_$i.log 2>&1
  tail -n 1 gpurun_out/ncu_nms_$i.log
done
