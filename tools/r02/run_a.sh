set -x
python tools/r02/time_a.py pc_stagger_ns=5000 pc_stagger_ns=10000 pc_stagger_ns=15000 pc_stagger_ns=20000 pc_stagger_ns=25000 pc_stagger_ns=30000 pc_stagger_ns=40000 > gpurun_out/r02_time_a.txt 2>&1
