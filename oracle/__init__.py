"""CPU oracle for the LRVB logistic-GLMM hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker (or as the thing timed on the
host cores) -- never on the CUDA product path.

Parity status
-------------
* Forward values (GH logistic term, exponential-family terms, parameter packing,
  the composed GLMM KL) are PINNED: ``tests/golden/make_golden.py`` runs the
  reference's own unmodified ``Modeling.py`` / ``ExponentialFamilies.py`` /
  ``Parameters.py`` / ``ParameterDictionary.py`` / ``NormalParams.py`` /
  ``GammaParams.py`` (imported from /root/reference under the ``sys.modules`` shim
  in ``oracle/ref_shim.py``) on seeded inputs and commits the outputs under
  ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
  against them.
* Derivatives: the reference differentiates with the third-party package
  ``autograd>=1.3,<1.4`` (setup.py:39), which is absent from /root/reference and
  not installable here.  Gradient / Hessian parity is therefore anchored three
  ways: (i) the analytic numpy formulas in ``glmm_oracle.py``, (ii) torch-fp64
  autodiff of the restated KL (``glmm_torch.py``), and (iii) high-order finite
  differences of the *reference's own forward code* committed as golden vectors
  (gradient to ~1e-10, directional second derivatives to ~1e-8).  No GLMM
  gradient/Hessian appears in any reference test, so for derivatives the
  reference's own tests leave **parity unpinned**; (iii) is the strongest pin
  available.
* Matrix / simplex parameter packing (``packing_oracle.py``): value maps and the simplex
  Jacobian / Hessian are PINNED to outputs of the reference's own functions
  (``tests/golden/packing.npz``); the log-Cholesky Jacobian / Hessian, which the reference obtains
  from autograd, are pinned to Richardson-extrapolated differences of the reference's forward map
  (1e-8 / 1e-6) and cross-checked against torch-fp64 autodiff (1e-13).
"""
