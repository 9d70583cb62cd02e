set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python scripts/fit_variants.py c3 full 5 > gpurun_out/s12_new.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s12_pytest.log 2>&1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:fit_disp --csv --log-file gpurun_out/s12_flop_launches.csv python scripts/flop_probe.py c3 full gpurun_out/s12_flop_counts.json > gpurun_out/s12_flop.log 2>&1
python scripts/flop_per_eval.py gpurun_out/s12_flop_launches.csv gpurun_out/s12_flop_counts.json profiles/r02_fit_disp_flop_per_eval.json > gpurun_out/s12_flop_per_eval.log 2>&1
cp profiles/r02_fit_disp_flop_per_eval.json gpurun_out/s12_flop_per_eval.json
python bench.py > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err
tail -n 3 gpurun_out/s12_new.log gpurun_out/s12_pytest.log gpurun_out/s12_flop_per_eval.log
