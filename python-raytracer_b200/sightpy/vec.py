"""Host-side 3-vector used to *describe* scenes (reference: sightpy/utils/vector3.py:12-234).

The reference does all of its rendering arithmetic through this class.  Here it is only the
user-facing value type of the API (scene construction, camera set-up, caller supplied ray
bundles); rendering arithmetic happens in the CUDA library behind the C ABI.  Components may
be python scalars (real or complex) or numpy arrays.
"""
import numbers

import numpy as np

__all__ = ["vec3", "rgb", "extract", "array_to_vec3"]


def _is_scalar_like(v):
    return isinstance(v, (numbers.Number, np.ndarray, np.generic))


def extract(cond, x):
    """np.extract that passes plain numbers through (reference vector3.py:5-9)."""
    return x if isinstance(x, numbers.Number) else np.extract(cond, x)


class vec3:
    __slots__ = ("x", "y", "z")

    def __init__(self, x, y, z):
        self.x, self.y, self.z = x, y, z

    # -- helpers ---------------------------------------------------------------------------
    def _zip(self, other, op):
        if isinstance(other, vec3):
            return vec3(op(self.x, other.x), op(self.y, other.y), op(self.z, other.z))
        if _is_scalar_like(other):
            return vec3(op(self.x, other), op(self.y, other), op(self.z, other))
        return NotImplemented

    def _map(self, fn):
        return vec3(fn(self.x), fn(self.y), fn(self.z))

    def __repr__(self):
        return f"vec3({self.x}, {self.y}, {self.z})"

    __str__ = lambda self: f"({self.x}, {self.y}, {self.z})"

    # -- arithmetic (component-wise; scalars broadcast) --------------------------------------
    def __add__(self, v):
        return self._zip(v, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, v):
        return self._zip(v, lambda a, b: a - b)

    def __rsub__(self, v):
        return self._zip(v, lambda a, b: b - a)

    def __mul__(self, v):
        return self._zip(v, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, v):
        return self._zip(v, lambda a, b: a / b)

    def __rtruediv__(self, v):
        return self._zip(v, lambda a, b: b / a)

    def __neg__(self):
        return self._map(lambda a: -a)

    def __pow__(self, a):
        return self._map(lambda c: c ** a)

    def __abs__(self):
        return self._map(np.abs)

    def __eq__(self, other):
        return (self.x == other.x) & (self.y == other.y) & (self.z == other.z)

    __hash__ = None

    # numpy must never treat a vec3 as a sequence (np.abs(vec3) is used by reference-era scripts)
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if ufunc is np.absolute and method == "__call__":
            return abs(self)
        return NotImplemented

    # -- products / norms -------------------------------------------------------------------
    def dot(self, v):
        return self.x * v.x + self.y * v.y + self.z * v.z

    def cross(self, v):
        return vec3(
            self.y * v.z - self.z * v.y,
            self.z * v.x - self.x * v.z,
            self.x * v.y - self.y * v.x,
        )

    def square_length(self):
        return self.dot(self)

    def length(self):
        return np.sqrt(self.dot(self))

    def normalize(self):
        mag = self.length()
        return self * (1.0 / np.where(mag == 0, 1, mag))

    def average(self):
        return (self.x + self.y + self.z) / 3

    def matmul(self, matrix):
        """matrix (3x3) applied on the left: returns matrix @ self (reference vector3.py:93-97)."""
        m = np.asarray(matrix)
        a = self.to_array()
        out = np.dot(m, a) if a.ndim == 1 else np.tensordot(m, a, axes=([1], [0]))
        return vec3(out[0], out[1], out[2])

    def change_basis(self, new_basis):
        return vec3(self.dot(new_basis[0]), self.dot(new_basis[1]), self.dot(new_basis[2]))

    # -- complex helpers (refractive indices are complex per colour channel) -----------------
    @staticmethod
    def real(v):
        return v._map(np.real)

    @staticmethod
    def imag(v):
        return v._map(np.imag)

    @staticmethod
    def exp(v):
        return v._map(np.exp)

    @staticmethod
    def sqrt(v):
        return v._map(np.sqrt)

    # -- swizzles / containers ------------------------------------------------------------
    def xyz(self):
        return vec3(self.x, self.y, self.z)

    def yzx(self):
        return vec3(self.y, self.z, self.x)

    def zxy(self):
        return vec3(self.z, self.x, self.y)

    def components(self):
        return (self.x, self.y, self.z)

    def to_array(self):
        return np.array([self.x, self.y, self.z])

    def shape(self, *_):
        return self.x.shape if isinstance(self.x, np.ndarray) else 1

    def __len__(self):
        s = self.shape()
        return s if isinstance(s, int) else s[0]

    def __getitem__(self, ind):
        return vec3(self.x[ind], self.y[ind], self.z[ind])

    def broadcast_to(self, shape):
        return self._map(lambda c: np.broadcast_to(c, shape))

    def repeat(self, n):
        return self._map(lambda c: np.repeat(c, n))

    def reshape(self, *newshape):
        return self._map(lambda c: c.reshape(*newshape))

    def mean(self, axis):
        return self._map(lambda c: np.mean(c, axis=axis))

    def clip(self, lo, hi):
        return self._map(lambda c: np.clip(c, lo, hi))

    def extract(self, cond):
        return self._map(lambda c: extract(cond, c))

    def place(self, cond):
        out = []
        for c in self.components():
            buf = np.zeros(cond.shape)
            np.place(buf, cond, c)
            out.append(buf)
        return vec3(*out)

    @staticmethod
    def where(cond, out_true, out_false):
        return vec3(
            np.where(cond, out_true.x, out_false.x),
            np.where(cond, out_true.y, out_false.y),
            np.where(cond, out_true.z, out_false.z),
        )

    @staticmethod
    def select(mask_list, out_list):
        return vec3(
            np.select(mask_list, [o.x for o in out_list]),
            np.select(mask_list, [o.y for o in out_list]),
            np.select(mask_list, [o.z for o in out_list]),
        )

    @staticmethod
    def concatenate(vecs):
        return vec3(
            np.concatenate([v.x for v in vecs]),
            np.concatenate([v.y for v in vecs]),
            np.concatenate([v.z for v in vecs]),
        )


def array_to_vec3(array):
    return vec3(array[0], array[1], array[2])


rgb = vec3
