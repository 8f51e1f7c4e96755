mkdir -p gpurun_out
timeout 21 python -m pytest -q -x -p no:cacheprovider "tests/test_gpu_parity.py::test_cpp_host_driver_end_to_end" "tests/test_clinkage.py::test_clinkage_gpu_vs_oracle[case9]" "tests/test_clinkage.py::test_clinkage_gpu_vs_oracle[case10]" "tests/test_clinkage.py::test_clinkage_gpu_vs_oracle[case11]" "tests/test_clinkage.py::test_clinkage_gpu_vs_oracle[case12]" > gpurun_out/c32.log 2>&1
echo "exit $?" >> gpurun_out/c32.log
tail -5 gpurun_out/c32.log
