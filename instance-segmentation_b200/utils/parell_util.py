"""Mirror of the reference's utils/parell_util.py:5-8: map `func` over the zipped arguments and transpose the
results into a tuple of lists.  (The batched decode in decode.decode_output does not need it; it is kept
because decode_single-style per-image callers do.)"""
from functools import partial


def multi_apply(func, *args, **kwargs):
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))
