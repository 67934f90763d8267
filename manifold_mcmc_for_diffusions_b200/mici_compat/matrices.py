"""mici.matrices: the identity metric is the only one the CHMC scripts construct (scripts/utils.py:255-270)."""
import numpy as np


class IdentityMatrix:
    # NumPy must defer `ndarray @ IdentityMatrix` to __rmatmul__ (Mici's Matrix classes do the same): the reference's
    # h2 is `0.5 * state.mom @ self.metric.inv @ state.mom` (mici_extensions.py:1202)
    __array_ufunc__ = None

    def __init__(self, size=None, scalar=1.0):
        self.size = size
        self.scalar = scalar

    @property
    def inv(self):
        return IdentityMatrix(self.size, 1.0 / self.scalar)

    @property
    def sqrt(self):
        return IdentityMatrix(self.size, np.sqrt(self.scalar))

    @property
    def log_abs_det(self):
        return 0.0 if self.scalar == 1.0 else (self.size or 0) * np.log(abs(self.scalar))

    def __matmul__(self, other):
        return other if self.scalar == 1.0 else self.scalar * other

    def __rmatmul__(self, other):
        return other if self.scalar == 1.0 else self.scalar * other

    def __rmul__(self, scalar):
        return IdentityMatrix(self.size, self.scalar * scalar)

    __mul__ = __rmul__


class DensePositiveDefiniteMatrix:
    __array_ufunc__ = None

    def __init__(self, array):
        self.array = np.asarray(array)

    @property
    def inv(self):
        return DensePositiveDefiniteMatrix(np.linalg.inv(self.array))

    @property
    def sqrt(self):
        return np.linalg.cholesky(self.array)

    @property
    def log_abs_det(self):
        return np.linalg.slogdet(self.array)[1]

    def __matmul__(self, other):
        return self.array @ other

    def __rmatmul__(self, other):
        return other @ self.array


class PositiveDefiniteBlockDiagonalMatrix:
    def __init__(self, blocks):
        self.blocks = tuple(blocks)
