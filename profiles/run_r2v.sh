mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err ) 2> gpurun_out/r2v_time.txt; tail -n 3 gpurun_out/r2v_time.txt; tail -c 400 gpurun_out/r2v_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_ref.json 2> gpurun_out/r2v_ref.err ) 2> gpurun_out/r2v_ref_time.txt; tail -n 3 gpurun_out/r2v_ref_time.txt; tail -c 600 gpurun_out/r2v_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; tail -n 2 gpurun_out/r2v_smoke.log
