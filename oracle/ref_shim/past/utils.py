"""``past.utils.old_div`` restated from the published behaviour of the ``future`` package:
floor division when both operands are integral, true division otherwise."""
import numbers


def old_div(a, b):
    if isinstance(a, numbers.Integral) and isinstance(b, numbers.Integral):
        return a // b
    return a / b
