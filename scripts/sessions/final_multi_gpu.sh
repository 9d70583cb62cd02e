# final multi-GPU evidence of the round: usage  bash scripts/sessions/final_multi_gpu.sh N [sweep]
set -x
N=$1
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpu or multi_gpu" > gpurun_out/f${N}_pytest.log 2>&1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/multi_gpu_check.py c3 400000 > gpurun_out/f${N}_multi_gpu_check.log 2>&1
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/f${N}_bench.json 2> gpurun_out/f${N}_bench.err
if [ "$2" = "sweep" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --sweep > gpurun_out/f${N}_sweep.json 2> gpurun_out/f${N}_sweep.err
fi
tail -n 3 gpurun_out/f${N}_pytest.log gpurun_out/f${N}_multi_gpu_check.log 2>/dev/null; tail -c 600 gpurun_out/f${N}_bench.err
