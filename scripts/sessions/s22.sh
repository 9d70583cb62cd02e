set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q -k "stand_alone or callback or error_paths or c1_shape" > gpurun_out/s22_pytest.log 2>&1
tail -n 30 gpurun_out/s22_pytest.log
