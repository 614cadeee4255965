mkdir -p gpurun_out
# final check of round 2: the whole -m gpu suite (incl. the axes= test and the tests that drive the kernels through
# the reference's own wrappers / test files, baseline/_ref*), no -x so that every failure shows
timeout -s KILL 175 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1
tail -5 gpurun_out/r2_pytest_final.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2_pytest_final.log | head -20
