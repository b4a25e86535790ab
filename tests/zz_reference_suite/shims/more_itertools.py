"""Four-function stand-in for the ``more_itertools`` names the reference's test-suite
uses (tests/test_adrt_iter.py:38-127, test_bdrt_iter.py, test_iadrt_fmg_iter.py): the
package is not installed in this image and there is no network.  Test infrastructure only."""
import collections

_marker = object()


def first(iterable, default=_marker):
    for item in iterable:
        return item
    if default is _marker:
        raise ValueError("first() was called on an empty iterable, and no default value was provided.")
    return default


def last(iterable, default=_marker):
    try:
        return collections.deque(iterable, maxlen=1)[-1]
    except IndexError:
        if default is _marker:
            raise ValueError("last() was called on an empty iterable, and no default was provided.") from None
        return default


def ilen(iterable):
    return sum(1 for _ in iterable)


def consume(iterator, n=None):
    if n is None:
        collections.deque(iterator, maxlen=0)
    else:
        for _ in zip(range(n), iterator):
            pass
