set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python scripts/fit_variants.py c3 full 5 > gpurun_out/s15_new.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s15_launches.csv python scripts/flop_probe.py c3 full gpurun_out/s15_counts.json > gpurun_out/s15_ncu.log 2>&1
head -n 1 gpurun_out/s15_new.log
