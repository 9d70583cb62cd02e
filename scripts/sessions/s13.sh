set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s13_topo.txt 2>&1
numactl -H > gpurun_out/s13_numa.txt 2>&1 || lscpu | grep -i numa > gpurun_out/s13_numa.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s13_bench8.json 2> gpurun_out/s13_bench8.err
tail -c 3000 gpurun_out/s13_bench8.json
