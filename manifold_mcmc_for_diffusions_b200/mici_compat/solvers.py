"""mici.solvers (only the norm the reference imports, mici_extensions.py:23)"""
import numpy as np


def maximum_norm(vct):
    return np.max(np.abs(vct))
