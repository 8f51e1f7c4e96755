cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/c10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
timeout 600 python scripts/gpu_variants.py 1000000 '{}' '{"p2_window":16384}' '{"p2_window":262144}' > gpurun_out/c10_variants.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c10_launches.csv python scripts/gpu_prof.py 1000000 > gpurun_out/c10_ncu.log 2>&1
tail -5 gpurun_out/c10_pytest.log
cat gpurun_out/c10_variants.log
