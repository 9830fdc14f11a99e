"""
Lock-step multi-start L-BFGS-B.

The reference runs its restarts one after the other, each a scipy L-BFGS-B whose objective
evaluates ONE point (hyper-parameters: gpr.py:944-989; acquisition: gp_acquisition.py:
280-390, 503-511).  On a GPU one evaluation costs about as much as a few hundred, so here every
restart runs the same scipy optimiser in its own thread and the objective calls of all
still-active restarts meet at a barrier: one batched device call evaluates them together.
The per-restart iterates are those scipy would produce on its own from the same function
values (same line search, same stopping rules): only the evaluation is shared.  (The values
themselves can differ in the last bits between a batched and a one-at-a-time evaluation, e.g.
where the training-side GEMMs split their k range for small batches; the optimiser then stops
at points that agree to its own tolerance.)
"""
import threading
import warnings

import numpy as np
import scipy.optimize


def lockstep_minimize(batch_func, x0s, bounds, method="L-BFGS-B", options=None):
    """Minimise from every start in ``x0s`` (n, p) at once.

    ``batch_func(X)`` takes an (m, p) array (m <= n, the restarts that are waiting) and returns
    ``(values (m,), grads (m, p))``.  Returns a list of ``(x_opt, f_opt)`` in start order.
    """
    x0s = [np.array(x0, dtype=float) for x0 in x0s]
    n = len(x0s)
    cond = threading.Condition()
    pending = {}          # restart -> point waiting for evaluation
    results = {}          # restart -> (f, g)
    active = set(range(n))
    optima = [None] * n
    errors = []

    def flush_locked():
        idx = sorted(pending)
        X = np.array([pending[i] for i in idx])
        try:
            vals, grads = batch_func(X)
        except Exception as excpt:     # wake everybody up: nobody must wait forever
            errors.append(excpt)
            pending.clear()
            cond.notify_all()
            raise
        for j, i in enumerate(idx):
            results[i] = (float(vals[j]), np.array(grads[j], dtype=float))
        pending.clear()
        cond.notify_all()

    def objective(i):
        def f(x):
            with cond:
                if errors:
                    raise RuntimeError("another restart failed") from errors[0]
                pending[i] = np.array(x, dtype=float)
                if len(pending) == len(active):
                    flush_locked()
                else:
                    while i not in results:
                        if errors:
                            raise RuntimeError("another restart failed") from errors[0]
                        cond.wait()
                return results.pop(i)
        return f

    def run(i):
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = scipy.optimize.minimize(objective(i), x0s[i], method=method, jac=True,
                                              bounds=bounds, options=options)
            optima[i] = (res.x, res.fun)
        except Exception as excpt:
            with cond:
                if not errors:
                    errors.append(excpt)
        finally:
            with cond:
                active.discard(i)
                if pending and len(pending) == len(active) and not errors:
                    try:
                        flush_locked()
                    except Exception:
                        pass
                cond.notify_all()

    threads = [threading.Thread(target=run, args=(i,)) for i in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return optima
