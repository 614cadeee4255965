"""Importing the UNMODIFIED reference (dask-image) without dask -- TEST INFRASTRUCTURE.

The only thing the reference's ndfilters path imports from dask is ``dask.utils.Dispatch``
(ref:dask_image/dispatch/_dispatcher.py:3); from a dask array it needs ``ndim`` / ``dtype`` / ``_meta`` and
``map_overlap`` (ref:dask_image/ndfilters/_gaussian.py:70-81).  ``reference_ndfilters()`` puts a stand-in for that one
class in ``sys.modules`` and imports ``dask_image.ndfilters`` from

* ``/root/reference`` (the build container), or
* ``baseline/_ref`` -- the reference pip-installed as it is by ``__graft_entry__.build()`` (git-ignored; it travels
  to the GPU box with the snapshot, where /root/reference does not exist),

then adds the ``B200Array`` registrations to the reference's dispatch tables (``register_with_dask_image()``: the lines
a maintainer would add).  Nothing of the reference is copied into the repository.
"""
import contextlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ["/root/reference", os.path.join(ROOT, "baseline", "_ref")]


def reference_root():
    for path in CANDIDATES:
        if os.path.isfile(os.path.join(path, "dask_image", "ndfilters", "__init__.py")):
            return path
    return None


def reference_tests_dir():
    """The reference's own ndfilters test files: in the reference tree, or the copy build() puts next to
    baseline/_ref (git-ignored)."""
    for path in ("/root/reference/tests", os.path.join(ROOT, "baseline", "_ref_tests")):
        d = os.path.join(path, "test_dask_image", "test_ndfilters")
        if os.path.isfile(os.path.join(d, "test__gaussian.py")):
            return d
    return None


# what of the reference's ndfilters suite applies to a non-numpy chunk type and the hot path (the rest: test__conv.py
# passes numpy weights, which the reference rejects for any other chunk type, ref:dask_image/dispatch/_utils.py:23-25;
# median / rank / percentile / generic filters are out of scope)
REFERENCE_TESTS_SELECTION = (
    "(not test__order or minimum_filter or maximum_filter) and "
    "(not test__threshold or gaussian or invalid_threshold_method) and "
    "not test__conv and not test__generic and not cupy")
REFERENCE_TESTS_COUNT = 321      # reference v2025.11.0


def run_reference_tests(tmp_path, files=None, expr=REFERENCE_TESTS_SELECTION, timeout=900):
    """Start the reference's test files in a pytest subprocess with the dask stand-in (tests/_dask_stub) and the
    plugin that keeps the arithmetic on the oracle when there is no GPU.  Returns (returncode, passed, output tail)."""
    import re
    import subprocess

    here = os.path.join(ROOT, "tests")
    tests_dir = reference_tests_dir()
    ini = os.path.join(str(tmp_path), "pytest.ini")
    with open(ini, "w") as f:
        f.write("[pytest]\nmarkers = cupy\n")        # not the reference's `--flake8` addopts
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(here, "_dask_stub"), reference_root(), ROOT, here])
    targets = [os.path.join(tests_dir, f) for f in files] if files else [tests_dir]
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-p", "b200_refsuite_plugin",
           "-c", ini, "--rootdir", str(tmp_path)] + targets
    if expr:
        cmd += ["-k", expr]
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=str(tmp_path))
    m = re.search(r"(\d+) passed", run.stdout)
    return run.returncode, int(m.group(1)) if m else 0, run.stdout[-3000:] + run.stderr[-2000:]


def unavailable_reason():
    """None when the reference can be driven through the stub, else why not."""
    if importlib.util.find_spec("dask") is not None:
        return "a real dask is installed: tests/test_gpu_dask.py covers this"
    if reference_root() is None:
        return "no reference tree (/root/reference or baseline/_ref)"
    return None


@contextlib.contextmanager
def reference_ndfilters():
    """Yields the reference's ``dask_image.ndfilters`` module with B200Array registered in its tables."""
    from dask_image_b200.dispatch import _dispatch_ndfilters as own
    from dask_image_b200.dispatch._dispatcher import _Base

    root = reference_root()
    dask = types.ModuleType("dask")
    utils = types.ModuleType("dask.utils")
    utils.Dispatch = _Base          # register / register_lazy / dispatch (MRO walk) / __call__, as dask's class
    dask.utils = utils
    saved = {k: sys.modules.get(k) for k in ("dask", "dask.utils")}
    sys.modules["dask"], sys.modules["dask.utils"] = dask, utils
    sys.path.insert(0, root)
    before = set(sys.modules)
    try:
        import dask_image.ndfilters as ref

        assert os.path.abspath(ref.__file__).startswith(os.path.abspath(root)), ref.__file__
        assert own.register_with_dask_image()
        yield ref
    finally:
        sys.path.remove(root)
        for name in set(sys.modules) - before:
            if name == "dask_image" or name.startswith("dask_image."):
                del sys.modules[name]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
