set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python scripts/fit_variants.py c3 full 5 > gpurun_out/s14_new.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s14_pytest.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s14_launches.csv python scripts/flop_probe.py c3 full gpurun_out/s14_counts.json > gpurun_out/s14_ncu.log 2>&1
tail -n 3 gpurun_out/s14_new.log gpurun_out/s14_pytest.log
