"""``dask.utils.Dispatch``: single dispatch on type with lazy registration (the documented contract of dask's class:
``register(type | tuple, func=None)``, ``register_lazy(toplevel_module, func=None)``, ``dispatch(cls)`` walking the
MRO, ``__call__(arg, ...)``).  One stand-in serves the tests and the CPU arm of bench.py (oracle/reference_arm.py)."""
from oracle.reference_arm import _Dispatch as Dispatch  # noqa: F401
