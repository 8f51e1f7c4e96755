cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c29_bench2.json 2> gpurun_out/c29_bench2.err
echo "bench rc=$?" >> gpurun_out/c29_bench2.err
tail -2 gpurun_out/c29_bench2.err
