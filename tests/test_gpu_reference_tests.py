"""The reference's OWN ndfilters test files, unmodified, on the GPU.

As tests/test_reference_suite_cpu.py, but with nothing patched: ``dask.array.from_array`` of the stand-in
(tests/_dask_stub) puts every test array into device chunks, the reference's wrappers and dispatch tables (imported
as they are from ``baseline/_ref``, the pip-installed copy ``__graft_entry__.build()`` makes) hand each chunk to this
package's chunk functions, and the sm_100a kernels do the arithmetic.  ``assert_eq`` holds float results to the
repository's tolerance (tests/_helpers.py) and everything else to equality.  321 of the reference's tests apply to a
non-numpy chunk type on the hot path (selection and reasons: tests/_reference_stub.py).
Skips when the reference's files are not there (they are git-ignored and built in the build container).
"""
import pytest

import _reference_stub

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(_reference_stub.unavailable_reason() is not None
                                 or _reference_stub.reference_tests_dir() is None,
                                 reason="no copy of the reference (baseline/_ref, baseline/_ref_tests) or a real dask")]


def test_reference_ndfilters_tests_pass_on_the_kernels(tmp_path):
    rc, passed, tail = _reference_stub.run_reference_tests(tmp_path)
    try:        # keep the reference suite's own summary with the session's records (copied to profiles/)
        import os

        out = os.path.join(_reference_stub.ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "reference_tests_on_gpu.log"), "w") as f:
            f.write(tail)
    except OSError:
        pass
    assert rc == 0 and passed == _reference_stub.REFERENCE_TESTS_COUNT, tail
