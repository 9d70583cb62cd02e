set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gpu and (two or multi or shard)" > gpurun_out/s16_pytest.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s16_bench2.json 2> gpurun_out/s16_bench2.err
tail -n 3 gpurun_out/s16_pytest.log; tail -c 1500 gpurun_out/s16_bench2.err
