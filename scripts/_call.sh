cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/c16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c16_pytest.log
timeout 600 python scripts/gpu_variants.py 1000000 '{}' '{"reserve":2}' '{"reserve":8}' '{"reserve":16}' '{"persistent":0}' '{"lookahead":1}' > gpurun_out/c16_variants.log 2>&1
tail -3 gpurun_out/c16_pytest.log
cat gpurun_out/c16_variants.log
