mkdir -p gpurun_out
# last call of round 2: the whole -m gpu suite on the final tree (adds the zero-copy interop test)
timeout -s KILL 100 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/r2_pytest_final2.log 2>&1
tail -4 gpurun_out/r2_pytest_final2.log | cut -c1-300
grep -n "^FAILED\|^ERROR\|^SKIPPED" gpurun_out/r2_pytest_final2.log | cut -c1-260 | head -20
