"""Structural test of gpry_b200.integration.patch_gpry against the REAL reference (build
container only: skipped where /root/reference is absent).  No device call is made; the
GPU-side behaviour of the very same methods is covered by tests/test_gpu_gpr.py."""
import dill as pickle
from copy import deepcopy

import numpy as np
import pytest

from oracle.ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


def test_patch_and_unpatch():
    gpry = import_reference()
    from gpry.preprocessing import Normalize_bounds, Normalize_y
    from gpry_b200 import integration
    cls = gpry.gpr.GaussianProcessRegressor
    orig_predict, orig_inv = cls.predict, cls._kernel_inverse
    integration.patch_gpry(gpry)
    try:
        assert cls.predict is not orig_predict and cls._kernel_inverse is not orig_inv
        bounds = np.array([[0.0, 1.0]] * 3)
        g = cls(kernel={"Matern": {"nu": 2.5}}, bounds=bounds, account_for_inf=None,
                preprocessing_X=Normalize_bounds(bounds), preprocessing_y=Normalize_y(), verbose=0)
        from sklearn.base import clone
        g.kernel_ = clone(g.kernel)
        g.kernel_.theta = np.log([2.0, 0.3, 0.4, 0.5])
        kind, c, ell = g._kernel_spec()
        assert kind == "matern25" and c == pytest.approx(2.0) and np.allclose(ell, [0.3, 0.4, 0.5])
        g._dev = object()                       # stands for a live device handle
        state = pickle.loads(pickle.dumps(g)).__dict__
        assert state["_dev"] is None            # never pickled, rebuilt lazily
        c2 = deepcopy(g)
        assert "_dev" not in c2.__dict__ or c2.__dict__["_dev"] is None
        with pytest.raises(ValueError):         # same argument checks as the reference
            g.predict(np.zeros((2, 3)), return_mean_grad=True)
    finally:
        integration.unpatch_gpry(gpry)
    assert cls.predict is orig_predict and cls._kernel_inverse is orig_inv
