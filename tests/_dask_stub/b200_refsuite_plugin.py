"""pytest plugin for running the REFERENCE'S OWN ndfilters test files against this repository's chunk type
(tests/test_reference_suite_cpu.py starts pytest with ``-p b200_refsuite_plugin``) -- TEST INFRASTRUCTURE.

Without a GPU (the build container) every test runs under the CPU harness of tests/test_host_path_cpu.py: the launch
helpers below the C ABI are replaced by the oracle, everything above them runs for real.  With a GPU nothing is
patched: the chunks are device buffers and the CUDA kernels do the arithmetic.
"""
import pytest
import torch

import dask_image_b200  # noqa: F401  (registers B200Array in the reference's dispatch tables: dask_image is importable)
from dask_image_b200.dispatch._dispatch_ndfilters import register_with_dask_image
from test_host_path_cpu import cpu_path  # noqa: F401  (fixture)

assert register_with_dask_image(), "the reference (dask_image) must be importable next to the dask stand-in"


@pytest.fixture(autouse=True)
def _reset_data_scale():
    from dask.array.utils import DATA_SCALE

    DATA_SCALE[0] = 0.0
    yield


if not torch.cuda.is_available():
    # arrays created while the reference's test modules are COLLECTED (parametrize lists) need host buffers too
    from dask_image_b200 import _array

    def _host_init(self, tensor):
        self.tensor = tensor if tensor.is_contiguous() else tensor.contiguous()

    _array.B200Array.__init__ = _host_init
    _array.default_device = lambda: torch.device("cpu")

    @pytest.fixture(autouse=True)
    def _b200_cpu_harness(cpu_path):  # noqa: F811
        yield
