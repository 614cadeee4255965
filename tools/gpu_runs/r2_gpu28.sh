mkdir -p gpurun_out
# the tests added at the end of round 2, verbosely (which ran, which skipped), and the reference suite's own summary
timeout -s KILL 120 python -m pytest tests/test_gpu_reference_tests.py tests/test_gpu_reference_wrappers.py "tests/test_gpu_parity.py::test_axes_subsets" -m gpu -v -rs -p no:cacheprovider > gpurun_out/r2_pytest_reference_driven.log 2>&1
tail -15 gpurun_out/r2_pytest_reference_driven.log | cut -c1-200
tail -3 gpurun_out/reference_tests_on_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -2
