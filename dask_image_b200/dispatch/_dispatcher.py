"""Single dispatch on the chunk type of an array -- the reference's plug-in
mechanism (dask_image/dispatch/_dispatcher.py:6-24).

When ``dask`` is importable the class derives from ``dask.utils.Dispatch`` as
the reference's does; otherwise an equivalent type table with the same
``register`` / ``register_lazy`` / ``dispatch`` (MRO walk) contract is used, so
the package works on a box without dask.
"""
try:  # pragma: no cover - dask is not installed in the build image
    from dask.utils import Dispatch as _Base
except ImportError:

    class _Base:
        """Type -> implementation table with MRO lookup and lazy registration."""

        def __init__(self, name=None):
            self._lookup = {}
            self._lazy = {}
            if name:
                self.__name__ = name

        def register(self, type, func=None):
            def wrapper(func):
                if isinstance(type, tuple):
                    for t in type:
                        self.register(t, func)
                else:
                    self._lookup[type] = func
                return func

            return wrapper(func) if func is not None else wrapper

        def register_lazy(self, toplevel, func=None):
            def wrapper(func):
                self._lazy[toplevel] = func
                return func

            return wrapper(func) if func is not None else wrapper

        def dispatch(self, cls):
            lk = self._lookup
            for cls2 in cls.__mro__:
                try:
                    return lk[cls2]
                except KeyError:
                    pass
                toplevel, _, _ = cls2.__module__.partition(".")
                try:
                    register = self._lazy.pop(toplevel)
                except KeyError:
                    pass
                else:
                    register()
                    return self.dispatch(cls)
            raise TypeError("No dispatch for {0}".format(cls))

        def __call__(self, arg, *args, **kwargs):
            meth = self.dispatch(type(arg))
            return meth(arg, *args, **kwargs)


def get_type(array):
    """Chunk type behind ``array``: ``type(array._meta)`` for dask arrays,
    ``type(array)`` for anything else."""
    meta = getattr(array, "_meta", None)
    return type(array) if meta is None else type(meta)


class Dispatcher(_Base):
    """Dispatch on ``get_type(arg)`` and call the registered factory, which
    returns the per-chunk filter function."""

    def __call__(self, arg, *args, **kwargs):
        factory = self.dispatch(get_type(arg))
        return factory(arg, *args, **kwargs)
