"""Optimisers for the device ``Objective`` -- the caller on the other side of the hot path.

Mirror of the reference's ``OptimizationUtils.py`` (file:line citations below) with the same
function names, argument orders and return values.  The reference drives ``scipy.optimize``
with ``fun_free`` / ``fun_free_grad`` / ``fun_free_hvp``; that still works here, every call is a
CUDA evaluation.  ``minimize_objective_newton`` is new: Newton's method whose linear solve is the
direct arrowhead solve of the cached Hessian (Schur complement on the FP64 tensor cores), so an
iteration is one order-2 evaluation + one solve + a few order-0 evaluations of the line search,
and the iterate never leaves the device.
"""
import numpy as np
import scipy as sp
import scipy.optimize   # noqa: F401  (sp.optimize)

from ._tensors import is_torch


def get_sym_matrix_inv_sqrt(hessian, ev_min=None, ev_max=None):
    """Inverse square root of a symmetric matrix with eigenvalue thresholding
    (OptimizationUtils.py:6-20).  Dense: meant for small problems, as in the reference."""
    if sp.sparse.issparse(hessian):
        hessian = hessian.toarray()
    hessian = np.asarray(hessian, dtype=np.float64)
    hessian_sym = 0.5 * (hessian + hessian.T)
    eig_val, eig_vec = np.linalg.eigh(hessian_sym)
    if ev_min is not None:
        eig_val[eig_val <= ev_min] = ev_min
    if ev_max is not None:
        eig_val[eig_val >= ev_max] = ev_max
    hess_corrected = (eig_vec * eig_val) @ eig_vec.T
    hess_inv_sqrt = (eig_vec / np.sqrt(eig_val)) @ eig_vec.T
    return np.array(hess_inv_sqrt), np.array(hess_corrected)


def set_objective_preconditioner(objective, free_par=None, hessian=None, ev_min=None, ev_max=None):
    """objective.preconditioner = H^{-1/2} at ``free_par`` (OptimizationUtils.py:25-42)."""
    if free_par is None and hessian is None:
        raise ValueError(
            'You must specify either a Hessian or the free_par at which ' +
            'the objective\'s Hessian is to be evaluated.')
    if hessian is None:
        hessian = objective.fun_free_hessian(free_par)
    inv_hess_sqrt, hessian_corrected = get_sym_matrix_inv_sqrt(hessian, ev_min=ev_min, ev_max=ev_max)
    objective.preconditioner = inv_hess_sqrt
    return hessian, inv_hess_sqrt, hessian_corrected


def _prepare_logger(objective, print_every, init_logger):
    if init_logger:
        objective.logger.initialize()
    if print_every is not None:
        objective.logger.print_every = print_every


def _scipy_minimize(objective, init_x, precondition, method, options, disp, with_hessp):
    """One scipy.optimize.minimize run over the plain or the preconditioned objective; the
    preconditioned run starts from P^{-1} x0 and maps the optimum back with ``uncondition_x``."""
    objective.preconditioning = precondition
    if precondition:
        if objective.preconditioner is None:
            raise AssertionError("precondition=True needs objective.preconditioner")
        funs = (objective.fun_free_cond, objective.fun_free_grad_cond, objective.fun_free_hvp_cond)
        start = np.linalg.solve(objective.preconditioner, init_x)
    else:
        funs = (objective.fun_free, objective.fun_free_grad, objective.fun_free_hvp)
        start = init_x
    value, grad, hvp = funs
    kwargs = dict(hessp=hvp) if with_hessp else {}
    result = sp.optimize.minimize(lambda par: value(par, verbose=disp), x0=start, jac=grad,
                                  method=method, options=options, **kwargs)
    opt_x = objective.uncondition_x(result.x) if precondition else result.x
    return opt_x, result


def minimize_objective_trust_ncg(objective, init_x, precondition, maxiter=50, gtol=1e-6, disp=True,
                                 print_every=None, init_logger=True):
    """scipy trust-ncg over the device objective: value, gradient and Hessian-vector products are
    CUDA evaluations (OptimizationUtils.py:45-77; same arguments and ``(opt_x, result)`` return)."""
    _prepare_logger(objective, print_every, init_logger)
    return _scipy_minimize(objective, init_x, precondition, 'trust-ncg',
                           {'maxiter': maxiter, 'gtol': gtol, 'disp': disp}, disp, True)


def minimize_objective_bfgs(objective, init_x, precondition=False, maxiter=500, disp=True,
                            print_every=None, init_logger=True):
    """scipy BFGS over the device objective (OptimizationUtils.py:80-108)."""
    _prepare_logger(objective, print_every, init_logger)
    return _scipy_minimize(objective, init_x, precondition, 'BFGS',
                           {'maxiter': maxiter, 'disp': disp}, disp, False)


def minimize_objective_newton(objective, init_x, maxiter=50, gtol=1e-8, disp=False,
                              max_backtracks=30, armijo=1e-4):
    """Newton's method with the direct arrowhead solve, entirely on the device.

    Each iteration: one order-2 evaluation (KL, gradient, Hessian blocks), ``step = -H^{-1} g`` by
    the Schur-complement solve of the cached Hessian, Armijo backtracking with order-0
    evaluations.  Where the Hessian is not positive definite (the Schur complement has a
    non-positive pivot, or the step is not a finite descent direction) the step is recomputed with
    Levenberg damping ``(H + lambda I)^{-1} g``, lambda growing tenfold from 1e-3 of the largest
    diagonal entry until the damped matrix is positive definite (``model.add_diagonal_``).  Returns
    ``(opt_x, scipy.optimize.OptimizeResult)`` like the scipy-driven optimisers above; ``opt_x`` is
    numpy for numpy input, a CUDA tensor for tensor input.
    """
    import torch
    model = getattr(objective, "model", None)
    if not getattr(model, "_lrvb_device_model", False):
        raise TypeError("minimize_objective_newton needs a device Objective")
    want_torch = is_torch(init_x)
    x = (init_x.detach().to(model.device, torch.float64) if want_torch
         else torch.as_tensor(np.asarray(init_x, dtype=np.float64), device=model.device)).reshape(-1).clone()
    nfev = njev = nhev = 0
    converged, message = False, "maximum number of iterations reached"
    kl = gmax = float("nan")
    it = 0
    for it in range(maxiter + 1):
        model.evaluate(x, 2, "free")
        nfev, njev, nhev = nfev + 1, njev + 1, nhev + 1
        kl = float(model.kl_tensor().item())
        g = model.grad_tensor()
        gmax = float(g.abs().max().item())
        if disp:
            print("Newton iter {}: value {:.12g}  |grad|_inf {:.3e}".format(it, kl, gmax))
        if gmax <= gtol:
            converged, message = True, "gradient tolerance reached"
            break
        if it == maxiter:
            break
        step, slope, lam, lam_total = None, 0.0, 0.0, 0.0
        for _try in range(24):
            try:
                step = -model.solve(g).reshape(-1)
                slope = float(torch.dot(g, step).item())
            except np.linalg.LinAlgError:
                slope = float("nan")
            if np.isfinite(slope) and slope < 0.0:
                break
            # not positive definite here: damp the cached blocks (the next iteration re-evaluates them)
            lam = 1e-3 * model.max_abs_diagonal() if lam == 0.0 else 10.0 * lam
            model.add_diagonal_(lam - lam_total)
            lam_total = lam
            step = None
        if step is None:
            message = "no descent direction found"
            break
        t, accepted = 1.0, False
        for _ in range(max_backtracks):
            xn = x + t * step
            model.evaluate(xn, 0, "free")
            nfev += 1
            kn = float(model.kl_tensor().item())
            if np.isfinite(kn) and kn <= kl + armijo * t * slope:
                accepted = True
                break
            t *= 0.5
        if not accepted:
            message = "line search failed"
            break
        x = xn
    objective._set_par(x, "free")
    objective.par        # an optimiser leaves the optimum in par (read here once: the only host copy of the run)
    res = sp.optimize.OptimizeResult(
        x=x if want_torch else x.cpu().numpy(), fun=kl, success=converged, message=message, nit=it,
        nfev=nfev, njev=njev, nhev=nhev, grad_inf_norm=gmax)
    return res.x, res


def repeatedly_optimize(objective, optimization_fun, init_x, initial_optimization_fun=None,
                        max_iter=100, gtol=1e-8, ftol=1e-8, xtol=1e-8, disp=False,
                        keep_intermediate_optimizations=False):
    """Repeat ``optimization_fun`` (start point -> (x, result)) until x, f or the gradient stops
    moving (OptimizationUtils.py:114-167); same return tuple."""
    opt_results = []
    if initial_optimization_fun is not None:
        if disp:
            print('Running intitial optimization.')
        init_x, init_opt = initial_optimization_fun(init_x)
        if keep_intermediate_optimizations:
            opt_results.append(init_opt)
    f_val = objective.fun_free(init_x)
    converged = x_conv = f_conv = grad_conv = False
    i = 0
    x = new_x = init_x
    obj_opt = None
    while i < max_iter and (not converged):
        if disp:
            print('\n---------------------------------\n' + 'Repeated optimization iteration ', i)
        i += 1
        new_x, obj_opt = optimization_fun(x)
        if keep_intermediate_optimizations:
            opt_results.append(obj_opt)
        new_f_val = objective.fun_free(new_x)
        grad_val = objective.fun_free_grad(new_x)
        x_diff = float(abs(new_x - x).sum())
        f_diff = float(abs(new_f_val - f_val))
        grad_l1 = float(abs(grad_val).sum())
        x_conv, f_conv, grad_conv = x_diff < xtol, f_diff < ftol, grad_l1 < gtol
        x, f_val = new_x, new_f_val
        converged = x_conv or f_conv or grad_conv
        if disp:
            print('Iter {}: x_diff = {}, f_diff = {}, grad_l1 = {}'.format(i, x_diff, f_diff, grad_l1))
    return new_x, converged, x_conv, f_conv, grad_conv, obj_opt, opt_results
