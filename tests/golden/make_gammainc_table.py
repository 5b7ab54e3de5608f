"""Generates tests/golden/gammainc_table.npz: scipy.special.gammainc -- the function the reference's gamma SAS family calls
(roger/core/sas.py:153) -- on a seeded sample of the (a, x) domain the backend accepts, plus the argument rules.

    python tests/golden/make_gammainc_table.py
"""
import os

import numpy as np
import scipy
import scipy.special as sp

HERE = os.path.dirname(os.path.abspath(__file__))

rng = np.random.default_rng(2024)
a = np.concatenate([rng.uniform(0.05, 5, 3000), rng.uniform(5, 30, 2000), rng.uniform(30, 170, 1000),
                    [0.5, 1, 2, 20, 21, 100, 170, 1.5, 0.8, 2.0]])
x = np.concatenate([10 ** rng.uniform(-6, np.log10(50), 6000), [0.3, 1, 2, 20, 21, 30, 49, 4.0, 2.5, 3.0]])
edge_a = np.array([0.0, 0.0, -1.0, 1.0, np.inf, 2.0, np.inf, 2.0, np.nan, 1.0])
edge_x = np.array([0.0, 1.0, 1.0, -1.0, 1.0, np.inf, np.inf, 0.0, 1.0, np.nan])
with np.errstate(all="ignore"):
    np.savez_compressed(os.path.join(HERE, "gammainc_table.npz"), a=a, x=x, p=sp.gammainc(a, x), edge_a=edge_a, edge_x=edge_x,
                        edge_p=sp.gammainc(edge_a, edge_x), scipy_version=np.array(scipy.__version__))
print("gammainc_table.npz", a.size, "points, scipy", scipy.__version__)
