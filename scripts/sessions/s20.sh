set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "next_batch or caller_buffers" > gpurun_out/s20_pytest.log 2>&1
tail -n 30 gpurun_out/s20_pytest.log
