set -x
for i in 1 2 3; do
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_gputests_rep$i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests_rep$i.log
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.txt
