cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/gpu_case2.py s100k:100000:7:12:blosum62 '{}' '{"bucket_aux":0}' '{"min_iters":8}' '{"min_iters":8,"bucket_aux":0}' '{"min_iters":3}' '{"batch":256}' '{"lookahead":0}' > gpurun_out/c23_cases.log 2>&1
cat gpurun_out/c23_cases.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/c23_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c23_pytest.log
tail -4 gpurun_out/c23_pytest.log
