cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/c14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
timeout 600 python scripts/gpu_variants.py 1000000 '{}' '{"p2_first":2048}' '{"p2_first":16384}' '{"p2_first":65536}'  '{"p2_first":32768, "p2_window":262144}' > gpurun_out/c14_variants.log 2>&1
tail -3 gpurun_out/c14_pytest.log
cat gpurun_out/c14_variants.log
