set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "caller_buffers or tiny" > gpurun_out/s17_pytest.log 2>&1
python bench.py > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fit_disp_kernel -c 4 -o gpurun_out/s17_fit_disp python scripts/flop_probe.py c3 full gpurun_out/s17_counts.json > gpurun_out/s17_ncu.log 2>&1
tail -n 3 gpurun_out/s17_pytest.log; tail -c 600 gpurun_out/s17_bench.err
