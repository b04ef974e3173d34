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
imported in this image (scikit-optimize, ConfigSpace, emcee, pyrfr are absent;
numpy 2.x dropped ``np.int`` / ``np.float`` / ``np.mat`` that the reference
uses).  The dense numerics live in un-vendored third-party code:
``scikit-optimize`` (unpinned in ``requirements.txt:18``) wrapping
``scikit-learn==0.21.3`` ``GaussianProcessRegressor`` (``requirements.txt:13``),
and scipy's L-BFGS-B / SLSQP.  The GPR algorithm restated here is Rasmussen &
Williams Alg. 2.1 as implemented by sklearn ``_gpr.py`` plus skopt's variance
form ``diag - einsum("ki,kj,ij->k", K*, K*, K_inv_)`` with
``K_inv_ = L^-T L^-1``.

Pinning: the reference ships no golden vectors, KATs or asserting tests for
this path, and it cannot be executed here, so the
oracle is pinned against the one independent implementation that IS available:
scikit-learn 1.9's ``GaussianProcessRegressor`` /
``log_marginal_likelihood(eval_gradient=True)`` for the continuous part of the
kernel stack (``tests/test_oracle_sklearn.py``), finite differences for the
Hamming part, and first-principles identities.  Against the *reference's own*
outputs parity is therefore UNPINNED (stated again in DESIGN.md).  Golden
vectors under ``tests/golden/`` are minted from this oracle with fixed seeds by
``tests/golden/make_golden.py`` / ``make_golden_mcmc.py``.  The same holds for the two further un-vendored
dependencies restated for rows a10 / f-1 -- ``emcee`` (sampler) and ``ConfigSpace`` (neighbourhood /
sampling streams): parity UNPINNED, see the headers of ``gp_mcmc.py`` and ``ei_optimization.py``.
"""
