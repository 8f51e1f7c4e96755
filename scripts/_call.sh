cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/c21_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c21_pytest.log
timeout 600 python scripts/gpu_case.py s100k:100000:7:12:blosum62 s100k_x:100000:9:12:pam250 b62_1m:1000000:12:12:blosum62 > gpurun_out/c21_cases.log 2>&1
tail -4 gpurun_out/c21_pytest.log
cat gpurun_out/c21_cases.log
