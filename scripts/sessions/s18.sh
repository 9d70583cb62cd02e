set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CHICDIFF_B200_LIB=$PWD/chicdiff_b200/libchicdiff_b200_nostage.so timeout 600 python scripts/fit_variants.py c3 full 5 > gpurun_out/s18_nostage.log 2>&1
timeout 600 python scripts/fit_variants.py c3 full 5 > gpurun_out/s18_new.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1
tail -n 4 gpurun_out/s18_nostage.log gpurun_out/s18_new.log gpurun_out/s18_pytest.log
