mkdir -p gpurun_out
python profiles/l2hint_ab.py 2>&1 | grep "hint on"
