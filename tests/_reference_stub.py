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
