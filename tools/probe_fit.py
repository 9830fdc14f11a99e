"""Prints what the hyper-parameter fit of the golden case ends with (theta, LML, evaluations)
for the lock-step and the one-by-one drivers, next to the reference's numbers."""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from gpry_b200.gpr import GaussianProcessRegressor
from gpry_b200.preprocessing import Normalize_bounds, Normalize_y
for name in ("fit_rbf_d2_n40", "fit_rbf_d8_n300"):
    path = os.path.join("tests", "golden", name + ".npz")
    if not os.path.exists(path):
        continue
    z = np.load(path)
    print(name, "reference: theta", z["theta_opt"], "lml", float(z["lml_opt"]), "n_eval", int(z["n_eval_loglike"]))
    for lockstep in (True, False):
        gpr = GaussianProcessRegressor(
            kernel="RBF", bounds=z["bounds"], noise_level=1e-2, n_restarts_optimizer=int(z["n_restarts"]) if "n_restarts" in z.files else 4,
            preprocessing_X=Normalize_bounds(z["bounds"]), preprocessing_y=Normalize_y(),
            account_for_inf=None, random_state=7, verbose=0)
        gpr.append_to_data(z["X_train"], z["y_train"], fit_gpr={"lockstep": lockstep})
        print("  lockstep", lockstep, "theta", gpr.kernel_.theta, "dtheta", np.max(np.abs(gpr.kernel_.theta - z["theta_opt"])),
              "lml", gpr.log_marginal_likelihood_value_, "rel", abs(gpr.log_marginal_likelihood_value_ - float(z["lml_opt"])) / abs(float(z["lml_opt"])),
              "n_eval", gpr.n_eval_loglike)
