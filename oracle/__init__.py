"""CPU oracle for the efficient-tlbo surrogate-and-acquisition hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the timed CPU baseline) -- never on the product path.  The product
(``efficient-tlbo_b200/``) fails loudly when its CUDA library is missing; it has
no CPU fallback and never imports ``oracle``.

What it is: a numpy/scipy restatement of the reference's algorithm for the hot
path named in SURVEY.md section 8 (rows a1-a10 and f-1: ``gp.py``, ``facade.py``, ``acquisition.py``,
``gp_mcmc.py`` -- GaussianProcessMCMC + emcee's stretch move --, ``ei_optimization.py`` -- the online
scenario's maximisers --), with a plain ``float64[N, d]`` +
``types[d]`` input contract instead of ConfigSpace objects.  Every function
cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).

Why a restatement instead of the reference itself: the reference cannot be
imported in this image as it stands (scikit-optimize, ConfigSpace, emcee, pyrfr are
absent; numpy 2.x dropped ``np.int`` / ``np.float`` / ``np.mat`` that the reference
uses) and does not exist on the GPU box at all.  The dense numerics live in
un-vendored third-party code: ``scikit-optimize`` (unpinned in
``requirements.txt:18``) wrapping ``scikit-learn==0.21.3``
``GaussianProcessRegressor`` (``requirements.txt:13``), and scipy's L-BFGS-B /
SLSQP.  The GPR algorithm restated here is Rasmussen & Williams Alg. 2.1 as
implemented by sklearn ``_gpr.py`` plus skopt's variance form
``diag - einsum("ki,kj,ij->k", K*, K*, K_inv_)`` with ``K_inv_ = L^-T L^-1``.

Pinning.  The reference ships no golden vectors, KATs or asserting tests for this
path.  The oracle is pinned two ways:

1. AGAINST THE REFERENCE'S OWN CODE, RUN HERE: ``tests/golden/make_reference_golden.py``
   imports the unmodified reference modules from ``/root/reference`` in the build
   container -- with the three removed numpy aliases restored and stand-ins under
   ``tests/golden/refshims/`` for the ABSENT third-party imports only (scikit-learn's
   kernels / regressor under scikit-optimize's names plus its ``K_inv_`` variance
   form; value classes for ConfigSpace; an ``emcee.EnsembleSampler`` surface over the
   restated stretch move) -- and records what the reference computes into
   ``tests/golden/reference_v1.npz``: ``gp_kernels.py`` K / dK / K*, ``GaussianProcess``
   ``_nll`` / ``_optimize`` (start points, per-restart optima, theta*) / ``predict``,
   ``EI``, ``scipy_solve``, ``GaussianProcessMCMC`` (``_ll``, ensemble, prediction), and
   whole ``SMBO_OFFLINE`` loops over ``RGPE``, ``TransBO_RGPE``, ``OBTLV``, ``TST`` and
   ``POGPE`` (selected rows, weights, dilution flags, the integer ranking-loss tables,
   every fitted theta, the global ``np.random`` state after each iteration).
   ``tests/test_reference_golden.py`` (CPU) holds the oracle to those numbers --
   floats to rounding, integers / rows / stream states exactly --;
   ``tests/test_gpu_reference_golden.py`` holds the CUDA path to them.
2. Against the independent implementation available offline: scikit-learn 1.9's
   ``GaussianProcessRegressor`` / ``log_marginal_likelihood(eval_gradient=True)``
   for the continuous part of the kernel stack (``tests/test_oracle_sklearn.py``),
   finite differences for the Hamming part, scipy's L-BFGS-B for the device-portable
   optimiser.  ``tests/golden/hotpath_v1.npz`` / ``mcmc_v1.npz`` freeze the oracle
   itself (``make_golden.py`` / ``make_golden_mcmc.py``).

What stays UNPINNED (stated again in DESIGN.md): the three things that live in
third-party packages absent from this machine and therefore run through a
restatement on BOTH sides of the fixtures -- scikit-optimize's ``K_inv_`` variance
form (mathematically the triangular form sklearn 1.9 uses; they agree to 1e-13),
``emcee``'s stretch-move sampler (``gp_mcmc.py``) and ``ConfigSpace``'s neighbourhood /
sampling streams of the online scenario (``ei_optimization.py``).
"""
