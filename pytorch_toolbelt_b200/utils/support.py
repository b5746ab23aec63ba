"""Deprecation decorator (mirrors the behaviour of reference ``pytorch_toolbelt/utils/support.py:10``)."""
import functools
import inspect
import warnings

__all__ = ["pytorch_toolbelt_deprecated"]


def _warn_and_call(target, message):
    @functools.wraps(target)
    def wrapper(*args, **kwargs):
        with warnings.catch_warnings():
            warnings.simplefilter("always", DeprecationWarning)
            warnings.warn(message, category=DeprecationWarning, stacklevel=2)
        return target(*args, **kwargs)

    return wrapper


def pytorch_toolbelt_deprecated(reason):
    """Usable as ``@pytorch_toolbelt_deprecated("why")`` or bare ``@pytorch_toolbelt_deprecated``."""
    if isinstance(reason, (str, bytes)):

        def decorate(target):
            kind = "class" if inspect.isclass(target) else "function"
            return _warn_and_call(target, f"Call to deprecated {kind} {target.__name__} ({reason}).")

        return decorate
    if inspect.isclass(reason) or inspect.isfunction(reason):
        kind = "class" if inspect.isclass(reason) else "function"
        return _warn_and_call(reason, f"Call to deprecated {kind} {reason.__name__}.")
    raise TypeError(repr(type(reason)))
