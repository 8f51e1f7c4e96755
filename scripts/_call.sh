cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 80 python bench.py --steps 5 --warmup 3 > gpurun_out/c30_bench1.json 2> gpurun_out/c30_bench1.err
echo "bench rc=$?" >> gpurun_out/c30_bench1.err
tail -2 gpurun_out/c30_bench1.err
