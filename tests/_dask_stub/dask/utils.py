"""``dask.utils.Dispatch``: single dispatch on type with lazy registration (the documented contract of dask's class:
``register(type | tuple, func=None)``, ``register_lazy(toplevel_module, func=None)``, ``dispatch(cls)`` walking the
MRO, ``__call__(arg, ...)``)."""


class Dispatch:
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
        for cls2 in cls.__mro__:
            if cls2 in self._lookup:
                return self._lookup[cls2]
            toplevel = cls2.__module__.partition(".")[0]
            if toplevel in self._lazy:
                self._lazy.pop(toplevel)()
                return self.dispatch(cls)
        raise TypeError("No dispatch for {0}".format(cls))

    def __call__(self, arg, *args, **kwargs):
        return self.dispatch(type(arg))(arg, *args, **kwargs)
