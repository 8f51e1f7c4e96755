cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/gpu_case2.py mix730:50000:7:30:blosum62 '{}' '{"batch":448}' '{"batch":96}' > gpurun_out/c25_cases.log 2>&1
timeout 600 python scripts/gpu_case2.py l30:50000:30:30:blosum62 '{}' >> gpurun_out/c25_cases.log 2>&1
cat gpurun_out/c25_cases.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/c25_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c25_pytest.log
tail -4 gpurun_out/c25_pytest.log
