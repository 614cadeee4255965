from ._dispatcher import Dispatcher, get_type  # noqa: F401
from ._utils import check_arraytypes_compatible  # noqa: F401
from ._dispatch_ndfilters import register_b200, register_with_dask_image  # noqa: F401
