"""The reference's OWN ndfilters test files, unmodified, run against this repository's chunk type.

/root/reference/tests/test_dask_image/test_ndfilters/*.py are started in a pytest subprocess with
tests/_dask_stub on the path: a stand-in for the sliver of dask they touch (``dask.utils.Dispatch``,
``dask.array.from_array / stack / concatenate / utils.assert_eq``) whose arrays are ``HaloChunked`` objects with
``B200Array`` chunks.  Every ``dask_image.ndfilters.X(d, ...)`` call in those files therefore goes: reference wrapper
-> reference dispatch table -> the B200Array registration -> this repository's chunk function, and every
``assert_eq`` against scipy is exact (the CPU harness puts the oracle below the C ABI; see b200_refsuite_plugin.py).

What is selected: everything in test__gaussian / test__smooth / test__diff / test__edge / test__utils, the
minimum / maximum filter cases of test__order, the gaussian and argument cases of test__threshold.  What is not:
test__conv.py passes NUMPY weights next to the image, which the reference itself rejects for any non-numpy chunk type
(``check_arraytypes_compatible``, ref:dask_image/dispatch/_utils.py:23-25 -- its cupy test passes cupy weights,
ref:tests/.../test_cupy_ndfilters.py:27-30; tests/test_reference_wrappers_cpu.py does the same with B200Array weights);
median / rank / percentile / generic filters are out of scope.
Runs only where /root/reference exists (the build container).
"""
import pytest

import _reference_stub

pytestmark = pytest.mark.skipif(_reference_stub.unavailable_reason() is not None
                                or _reference_stub.reference_tests_dir() is None,
                                reason="needs the reference tree (build container) and no real dask")

SELECTION = [
    # (file, -k expression, tests that must pass: the counts of reference v2025.11.0 -- 321 in all)
    ("test__gaussian.py", None, 168),
    ("test__smooth.py", None, 15),
    ("test__diff.py", None, 2),
    ("test__edge.py", None, 20),
    ("test__utils.py", None, 46),
    ("test__order.py", "minimum_filter or maximum_filter", 64),
    ("test__threshold.py", "gaussian or invalid_threshold_method", 6),
]


@pytest.mark.parametrize("fname, expr, count", SELECTION, ids=[s[0] for s in SELECTION])
def test_reference_test_file_passes_with_b200_chunks(tmp_path, fname, expr, count):
    rc, passed, tail = _reference_stub.run_reference_tests(tmp_path, [fname], expr)
    assert rc == 0 and passed == count, tail


def test_selection_expression_counts():
    """The one-run form used on the GPU (tests/test_gpu_reference_tests.py) selects the same tests."""
    assert sum(s[2] for s in SELECTION) == _reference_stub.REFERENCE_TESTS_COUNT
