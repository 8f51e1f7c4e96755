cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c18_smi.log
timeout 900 python -m pytest tests/test_gpu_boundary.py -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/c18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c18_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/c18_dist.log 2>&1
echo "dist rc=$?" >> gpurun_out/c18_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c18_bench2.json 2> gpurun_out/c18_bench2.err
echo "bench rc=$?" >> gpurun_out/c18_bench2.err
tail -15 gpurun_out/c18_pytest.log; tail -12 gpurun_out/c18_dist.log; tail -3 gpurun_out/c18_bench2.err
