set -x
python tools/r02/refill_experiment.py 600 > gpurun_out/r02_refill_experiment.txt 2>&1; echo rc=$?
for rep in 1 2; do
python tools/r02/ab_timing.py 4096 "" "cf_debug=1024" "cf_debug=3072" "cf_debug=2048" >> gpurun_out/r02_refill_timing.txt 2>&1
done
