# last check of the head: the whole GPU suite and smoke()
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/fc_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fc_smoke.log 2>&1
tail -n 3 gpurun_out/fc_pytest.log gpurun_out/fc_smoke.log
