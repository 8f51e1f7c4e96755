cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
timeout 600 python scripts/gpu_variants.py 1000000 '{}' > gpurun_out/c11_variants.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hmk_p2_window -s 20 -c 1 -f -o gpurun_out/c11_window python scripts/gpu_prof.py 1000000 > gpurun_out/c11_ncu1.log 2>&1
tail -3 gpurun_out/c11_pytest.log
cat gpurun_out/c11_variants.log
