# final 1-GPU evidence of the round: tests, flop table of the current kernel, bench (+ reference arm), launch list, full capture
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest.log 2>&1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:fit_disp --csv --log-file gpurun_out/f1_flop_launches.csv python scripts/flop_probe.py c3 full gpurun_out/f1_flop_counts.json > gpurun_out/f1_flop.log 2>&1
python scripts/flop_per_eval.py gpurun_out/f1_flop_launches.csv gpurun_out/f1_flop_counts.json profiles/r02_fit_disp_flop_per_eval.json > gpurun_out/f1_flop_per_eval.log 2>&1
cp profiles/r02_fit_disp_flop_per_eval.json gpurun_out/f1_flop_per_eval.json
python bench.py > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err
python bench.py --impl reference > gpurun_out/f1_bench_reference.json 2> gpurun_out/f1_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f1_launches.csv python scripts/flop_probe.py c3 full gpurun_out/f1_counts2.json > gpurun_out/f1_ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fit_disp_kernel -c 4 -o gpurun_out/f1_fit_disp python scripts/flop_probe.py c3 full gpurun_out/f1_counts3.json > gpurun_out/f1_ncu_full.log 2>&1
timeout 900 python bench.py --sweep > gpurun_out/f1_sweep.json 2> gpurun_out/f1_sweep.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f1_smoke.log 2>&1
tail -n 3 gpurun_out/f1_pytest.log gpurun_out/f1_smoke.log; tail -c 400 gpurun_out/f1_bench.err
