set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python bench.py > gpurun_out/fh_bench.json 2> gpurun_out/fh_bench.err
tail -c 300 gpurun_out/fh_bench.err
