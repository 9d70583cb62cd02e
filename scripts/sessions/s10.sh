set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python scripts/fit_variants.py c3 full 5 > gpurun_out/s10_t512.log 2>&1
CHICDIFF_B200_LIB=$PWD/chicdiff_b200/libchicdiff_b200_t128.so python scripts/fit_variants.py c3 full 5 > gpurun_out/s10_t128.log 2>&1
python scripts/fit_variants.py c3 full 5 > gpurun_out/s10_t512_b.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:fit_disp --csv --log-file gpurun_out/s10_flop_launches.csv python scripts/flop_probe.py c3 full gpurun_out/s10_flop_counts.json > gpurun_out/s10_flop.log 2>&1
python bench.py > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
tail -3 gpurun_out/s10_t512.log gpurun_out/s10_t128.log gpurun_out/s10_t512_b.log gpurun_out/s10_pytest.log
