"""A number that lives on the device until somebody looks at it.

The reference computes `num_connect` with `torch.nonzero(...).shape[0]` (agent.py:1052-1056,1073-1077): a host
synchronisation in the middle of every evaluation step. The kernels here count the connections on the device; forward()
returns this DeviceScalar in place of the Python number. It behaves like one - `total += x`, `total / count`, `str()`,
comparisons, `float()` all work, so runningScore.update_bandW / get_avg_bandW (metrics.py:19-21,110-111) run on it
unchanged - but arithmetic stays on the device (tiny asynchronous kernels on the current stream) and the value is
copied to the host only when a host number is really needed (`float()`, printing, comparing): once per epoch in the
reference's evaluate loop instead of once per step.
"""
import numbers

import torch


class DeviceScalar(numbers.Real):
    __slots__ = ("_t", "_v")

    def __init__(self, tensor):
        """tensor: a 0-dim (or 1-element) float64 device tensor owned by this object (not a view of a static buffer)."""
        self._t = tensor.reshape(())
        self._v = None

    @classmethod
    def from_count(cls, count, denom):
        """count: 1-element integer device tensor (a static engine buffer: snapshotted here), value = count / denom."""
        return cls(count.reshape(()).to(torch.float64) / float(denom))

    def device_value(self):
        """The value as a 0-dim float64 device tensor (no synchronisation)."""
        return self._t

    def item(self):
        if self._v is None:
            self._v = float(self._t.item())   # the one host synchronisation
        return self._v

    # ---- host views
    def __float__(self):
        return self.item()

    def __int__(self):
        return int(self.item())

    def __trunc__(self):
        return int(self.item())

    def __floor__(self):
        import math
        return math.floor(self.item())

    def __ceil__(self):
        import math
        return math.ceil(self.item())

    def __round__(self, n=None):
        return round(self.item(), n)

    def __bool__(self):
        return self.item() != 0.0

    def __repr__(self):
        return repr(self.item())

    def __str__(self):
        return str(self.item())

    def __format__(self, spec):
        return format(self.item(), spec)

    def __hash__(self):
        return hash(self.item())

    # ---- comparisons resolve (anything that is not a plain number - e.g. pytest.approx - compares from its side)
    @staticmethod
    def _num(o):
        return float(o) if isinstance(o, (int, float, DeviceScalar)) else None

    def __eq__(self, o):
        v = self._num(o)
        return self.item() == v if v is not None else o == self.item()

    def __ne__(self, o):
        return not self.__eq__(o)

    def __lt__(self, o):
        v = self._num(o)
        return self.item() < v if v is not None else NotImplemented

    def __le__(self, o):
        v = self._num(o)
        return self.item() <= v if v is not None else NotImplemented

    def __gt__(self, o):
        v = self._num(o)
        return self.item() > v if v is not None else NotImplemented

    def __ge__(self, o):
        v = self._num(o)
        return self.item() >= v if v is not None else NotImplemented

    # ---- arithmetic stays on the device
    def _other(self, o):
        if isinstance(o, DeviceScalar):
            return o._t.to(self._t.device)
        if isinstance(o, (int, float)):
            return float(o)
        return NotImplemented

    def _bin(self, o, fn):
        v = self._other(o)
        if v is NotImplemented:
            return NotImplemented
        return DeviceScalar(fn(self._t, v))

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._bin(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, lambda a, b: a / b)

    def __rtruediv__(self, o):
        return self._bin(o, lambda a, b: b / a)

    def __floordiv__(self, o):
        return self.item() // float(o)

    def __rfloordiv__(self, o):
        return float(o) // self.item()

    def __mod__(self, o):
        return self.item() % float(o)

    def __rmod__(self, o):
        return float(o) % self.item()

    def __pow__(self, o):
        return self.item() ** float(o)

    def __rpow__(self, o):
        return float(o) ** self.item()

    def __neg__(self):
        return DeviceScalar(-self._t)

    def __pos__(self):
        return self

    def __abs__(self):
        return DeviceScalar(self._t.abs())
