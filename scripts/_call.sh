cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c31_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/c31_smoke.log
timeout 25 python -m pytest tests/test_clinkage.py -m gpu -q -p no:cacheprovider -k "musi or status" > gpurun_out/c31_pytest.log 2>&1
tail -2 gpurun_out/c31_smoke.log; tail -2 gpurun_out/c31_pytest.log
