# last check of the head: the whole GPU suite, smoke(), the default bench line
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/fc_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fc_smoke.log 2>&1
python bench.py > gpurun_out/fc_bench.json 2> gpurun_out/fc_bench.err
tail -n 3 gpurun_out/fc_pytest.log gpurun_out/fc_smoke.log; tail -c 300 gpurun_out/fc_bench.err
