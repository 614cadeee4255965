"""The entry points of ``dask_image.ndfilters`` that are NOT on the B200 correlation path (rank
statistics and arbitrary Python callables -- dask_image/ndfilters/_order.py:49-157, _generic.py:15-50).  They exist
here under the reference's names so that code written against ``dask_image.ndfilters`` fails with one clear error
instead of an AttributeError; there is deliberately no CPU fallback behind them."""

__all__ = ["generic_filter", "median_filter", "rank_filter", "percentile_filter", "OutOfScopeError"]


class OutOfScopeError(NotImplementedError):
    """A dask-image filter outside the B200 correlation path was called."""


def _out_of_scope(name, why):
    def stub(image, *args, **kwargs):
        raise OutOfScopeError(
            "dask_image_b200.ndfilters.%s: %s; it is outside the B200 correlation path and there is no CPU "
            "fallback.  Use dask_image.ndfilters.%s on numpy / cupy chunks." % (name, why, name))

    stub.__name__ = stub.__qualname__ = name
    stub.__doc__ = "Not available: %s (raises OutOfScopeError, a NotImplementedError)." % why
    return stub


generic_filter = _out_of_scope("generic_filter", "an arbitrary Python callable per window cannot run in a CUDA kernel")
median_filter = _out_of_scope("median_filter", "a rank statistic, not a correlation")
rank_filter = _out_of_scope("rank_filter", "a rank statistic, not a correlation")
percentile_filter = _out_of_scope("percentile_filter", "a rank statistic, not a correlation")
