set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python scripts/fit_variants.py c3 full 5 > gpurun_out/s11_new.log 2>&1
CHICDIFF_B200_LIB=$PWD/chicdiff_b200/libchicdiff_b200_t512.so python scripts/fit_variants.py c3 full 5 > gpurun_out/s11_t512.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s11_pytest.log 2>&1
tail -n 3 gpurun_out/s11_new.log gpurun_out/s11_t512.log gpurun_out/s11_pytest.log
