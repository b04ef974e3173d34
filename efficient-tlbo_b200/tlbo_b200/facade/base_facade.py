"""Ensemble-facade base (reference tlbo/facade/base_facade.py:15-208) on the device path.

All K source surrogates, the target surrogate and the (up to 5) cross-validation surrogates
of one facade share ONE device-resident ``SurrogatePack``:
    slots [0, K)      source surrogates (fitted once, in one batched launch)
    slot  K           target surrogate
    slots (K, K + 5]  cross-validation surrogates of the current iteration
so that an iteration costs one batched fit launch, one small-query predict launch and one
fused predict + EI + arg-max launch.

With ``surrogate_type='gp_mcmc'`` (BASELINE config 5) every surrogate is a GROUP of W walker GPs
(``GaussianProcessMCMC``): logical slot j occupies pack slots ``[j * W, (j + 1) * W)``, per-walker
posteriors are marginalised on the device (``tlbo_mcmc_marginalize``) and the ensemble
combination + EI + arg-max run over the marginalised per-surrogate posteriors."""
import abc

import numpy as np

from tlbo_b200.config_space import convert_configurations_to_array, get_types
from tlbo_b200.model.gp import fit_models_async, fit_prepared_async
from tlbo_b200.model.gp_mcmc import fit_mcmc_models_async
from tlbo_b200.model.model_builder import build_model
from tlbo_b200.utils.constants import VERY_SMALL_NUMBER
from tlbo_b200.utils.normalization import zero_mean_unit_var_normalization, zero_one_normalization

N_CV_SLOTS = 5
SMALL_QUERY_ROWS = 2048      # at or below: per-(row, surrogate) warps instead of the pool-tile kernel


def _history_to_arrays(hist):
    """One source history: an ``OrderedDict{Configuration -> perf}`` (the reference's
    pickle format) or an ``(X, y)`` pair of arrays."""
    if isinstance(hist, (tuple, list)) and len(hist) == 2 and isinstance(hist[0], np.ndarray):
        return np.asarray(hist[0], dtype=np.float64), np.array(hist[1], dtype=np.float64)
    _X, _y = list(), list()
    for _config, _perf in hist.items():
        _X.append(_config)
        _y.append(_perf)
    return convert_configurations_to_array(_X), np.array(_y, dtype=np.float64)


class BaseFacade(object, metaclass=abc.ABCMeta):
    def __init__(self, config_space, source_hpo_data, seed, target_hp_configs=None,
                 history_dataset_features=None, num_src_hpo_trial=50, surrogate_type='gp', max_target_rows=96,
                 mcmc_steps=None):
        self.method_id = None
        self.config_space = config_space
        self.random_seed = seed
        self.K = len(source_hpo_data)
        self.num_src_hpo_trial = num_src_hpo_trial
        self.source_hpo_data = source_hpo_data
        self.target_hp_configs = target_hp_configs
        self.source_surrogates = None
        self.target_surrogate = None
        self.history_dataset_features = history_dataset_features
        if history_dataset_features is not None:
            assert len(history_dataset_features) == self.K
        if surrogate_type == 'gp_mcmc' and getattr(get_types(config_space)[0], 'has_conditions', False):
            raise ValueError('gp_mcmc surrogates over a space with conditions are not supported on the device path')
        if surrogate_type not in ('gp', 'gp_b200', 'gp_mcmc'):
            raise ValueError("surrogate_type %r is not part of the B200 hot path (use 'gp' or 'gp_mcmc')"
                             % surrogate_type)
        self.surrogate_type = surrogate_type
        self.types, self.bounds = get_types(config_space)
        # walker GPs per surrogate (1 for the MAP GP): model_builder.py:74-77
        self._G = 1
        if surrogate_type == 'gp_mcmc':
            self._G = 3 * (len(self.types) + 2)
            self._G += self._G % 2
        self._mcmc_steps = mcmc_steps       # (burn-in, chain) override of the reference's (250, 250)
        self._mcmc_shared = {}              # initial ensemble + random tables shared by identically seeded models
        self._logical_pack = None           # K + 1 placeholder slots for the all-resident fused call
        self._small = None                  # pinned buffers of the small-query acquisition path
        self.instance_features = None
        self.var_threshold = VERY_SMALL_NUMBER
        self.w = None
        self.eta_list = list()
        self.target_weight = []
        self._pack = None
        self._cap_hint = max_target_rows
        # accumulation order of the fused ensemble predict (pack slot ids) and TST mode flag
        self._fused_mode = 0
        self._stage = None            # pinned staging buffers (created with the pack)
        self._overlap_hook = None     # host callback run while the batched fit is in flight
        self._p0_cache = None         # restart start points shared by every model of this facade
        self._proto = None            # prototype of a freshly built target-side surrogate
        self._ens_cache = None        # (pack, device weights / skip / order, pack view) of acq_small
        self._pool_cache = {}         # (data_ptr, N) -> resident source posteriors over a fixed pool
        self._pool_keepalive = None   # the registered pool tensor: while it is alive its address cannot be reused

    # ---- data-set meta-features (base_facade.py:193-208; scikit-learn on the host as the reference) --------
    def scale_fit_meta_features(self, meta_features):
        from sklearn.impute import SimpleImputer
        from sklearn.preprocessing import MinMaxScaler
        meta_features = np.array(meta_features)
        self.meta_feature_imputer = SimpleImputer(missing_values=np.nan, strategy='mean')
        self.meta_feature_imputer.fit(meta_features)
        meta_features = self.meta_feature_imputer.transform(meta_features)
        self.meta_feature_scaler = MinMaxScaler()
        self.meta_feature_scaler.fit(meta_features)
        return self.meta_feature_scaler.transform(meta_features)

    def scale_transform_meta_features(self, meta_feature):
        _meta_features = np.array([meta_feature])
        _meta_feature = self.meta_feature_imputer.transform(_meta_features)
        _meta_feature = self.meta_feature_scaler.transform(_meta_feature)
        return np.clip(_meta_feature, 0, 1)

    # ---- device pack ------------------------------------------------------------------
    def _ensure_pack(self, rows):
        from tlbo_b200 import ops
        need = ops.batch_cap(max(rows, self.num_src_hpo_trial, 8))
        if self._pack is not None and self._pack.ncap >= need:
            return self._pack
        cap = ops.round_up(max(need, self._cap_hint), 8)
        if self._pack is not None:
            cap = max(cap, ops.round_up(int(self._pack.ncap * 1.5), 8))
        if self._stage is None:
            self._stage = ops.PinnedStage()
        new = ops.SurrogatePack((self.K + 1 + N_CV_SLOTS) * self._G, cap, len(self.types))
        if self._pack is not None:
            for s in range(self._pack.M):
                new.copy_slot_from(s, self._pack, s)
            for mdl in (self.source_surrogates or []):
                mdl.bind(new, mdl._slot)
        self._pack = new
        return new

    @abc.abstractmethod
    def train(self, X, y):
        pass

    @abc.abstractmethod
    def predict(self, X):
        pass

    # ---- surrogate construction ---------------------------------------------------------
    def _new_model(self):
        """``build_model(surrogate_type, config_space, RandomState(seed))`` (base_facade.py:64-65,95)."""
        model = build_model(self.surrogate_type, self.config_space, np.random.RandomState(self.random_seed))
        if self._G > 1:
            assert model.group_size == self._G
            if self._mcmc_steps is not None:
                model.burnin_steps, model.chain_length = int(self._mcmc_steps[0]), int(self._mcmc_steps[1])
        return model

    def _fit_async(self, models, Xs, ys):
        """One batched device fit of ``models`` (MAP restarts or MCMC chains); returns ``finish``."""
        self._ens_cache = None
        if self._G > 1:
            return fit_mcmc_models_async(models, Xs, ys)
        return fit_models_async(models, Xs, ys)

    def build_source_surrogates(self, normalize):
        """base_facade.py:58-91 -- the K source fits run as ONE batched launch."""
        self.source_surrogates = list()
        pack = self._ensure_pack(self.num_src_hpo_trial)
        Xs, ys = [], []
        for k, hist in enumerate(self.source_hpo_data):
            model = self._new_model()
            X, y = _history_to_arrays(hist)
            X = X[:self.num_src_hpo_trial]
            y = y[:self.num_src_hpo_trial]
            if normalize == 'standardize':
                if (y == y[0]).all():
                    y[0] += 1e-4
                y, _, _ = zero_mean_unit_var_normalization(y)
            elif normalize == 'scale':
                if (y == y[0]).all():
                    y[0] += 1e-4
                y, _, _ = zero_one_normalization(y)
                y = 2 * y - 1.
            else:
                raise ValueError('Invalid parameter in norm.')
            self.eta_list.append(np.min(y))
            self._share_start_points(model)
            model.bind(pack, k)
            Xs.append(X)
            ys.append(y)
            self.source_surrogates.append(model)
        if self.K:
            self._fit_async(self.source_surrogates, Xs, ys)()

    def _share_start_points(self, model):
        """Every model of a facade is built from a fresh ``RandomState(self.random_seed)``
        (base_facade.py:64-65,95), so the 10 sampled L-BFGS-B start points are identical:
        draw them once and share the array (also lets the batched fit see one ``p0``)."""
        if self._G > 1:
            model._shared = self._mcmc_shared        # same seeds -> same initial ensemble and random tables
            return
        if self._p0_cache is None:
            self._p0_cache = model._start_points()
        model._p0_override = self._p0_cache

    def _prepare_single_surrogate(self, X, y, normalize, slot):
        """The host half of ``build_single_surrogate`` (base_facade.py:93-108): build the
        model, apply the all-equal guard (mutating the caller's ``y``) and the facade-level
        normalisation.  Returns ``(model, X, y_normalised)`` for a later batched fit."""
        assert normalize in ['standardize', 'scale', 'none']
        if self._proto is None:
            self._proto = self._new_model()
            self._share_start_points(self._proto)
        # every build_model(type, cs, RandomState(seed)) call yields the same fresh model
        model = self._proto.fresh_copy()
        if normalize == 'standardize':
            if (y == y[0]).all():
                y[0] += 1e-4
            y, _, _ = zero_mean_unit_var_normalization(y)
        elif normalize == 'scale':
            if (y == y[0]).all():
                y[0] += 1e-4
            y, _, _ = zero_one_normalization(y)
        model.bind(self._ensure_pack(X.shape[0]), slot)
        return model, X, y

    def _native_prep_ok(self, X, y):
        """The native host-prep path covers MAP GPs fed with a plain float64 vector."""
        return (self._G == 1 and isinstance(y, np.ndarray) and y.dtype == np.float64 and y.ndim == 1
                and y.flags.c_contiguous and isinstance(X, np.ndarray) and X.ndim == 2 and X.shape[0] == y.shape[0])

    def _prepare_cv_jobs(self, X, y, ranges, outer_guard, normalize):
        """``[_prepare_single_surrogate(X minus ranges[j], ...) for j]`` with the host half (guards, the two
        normalisations, imputation, padding) done by ONE native call (``tlbo_host_prepare_jobs``,
        bit-identical to the numpy statement above; ``y`` is mutated in place where the reference mutates
        it).  Job j goes to slot ``K + j``.  Returns ``(models, launch)``; ``launch()`` enqueues the batched
        fit and returns its ``finish``."""
        from tlbo_b200 import ops
        if self._proto is None:
            self._proto = self._new_model()
            self._share_start_points(self._proto)
        self._ens_cache = None
        pack = self._ensure_pack(X.shape[0])
        prep = ops.prepare_jobs(X, y, ranges, outer_guard, normalize)
        models = [self._proto.fresh_copy().bind(pack, self.K + j) for j in range(len(ranges))]
        return models, (lambda: fit_prepared_async(models, prep))

    def build_single_surrogate(self, X, y, normalize, slot=None):
        job = self._prepare_single_surrogate(X, y, normalize, self.K if slot is None else slot)
        finish = self._fit_async([job[0]], [job[1]], [job[2]])
        if self._overlap_hook is not None:
            self._overlap_hook()
        finish()
        return job[0]

    # ---- prediction -----------------------------------------------------------------------
    def _ensemble_args(self):
        """(w[M], skip[M] or None, order[M] or None) over pack slots 0..K for the fused
        kernel; subclasses override."""
        return np.asarray(self.w, dtype=np.float64), None, None

    def _ensemble_args_dev(self):
        """``_ensemble_args`` as device tensors (host arrays go through the small-constant upload cache; a facade
        whose weights were formed on the device hands its tensors through)."""
        from tlbo_b200 import ops

        def dev(a, dt):
            if a is None or hasattr(a, 'data_ptr'):
                return a
            return ops.cached_upload(np.ascontiguousarray(a, dtype=dt))
        w, skip, order = self._ensemble_args()
        return dev(w, np.float64), dev(skip, np.int32), dev(order, np.int32)

    def _fused(self, X_dev, eta=0.0, par=0.0, var_floor=0.0, keys=None, excluded=None, idx_offset=0,
               want_rows=False, ws=None, sync=True):
        from tlbo_b200 import ops
        dw, dskip, dorder = self._ensemble_args_dev()
        N = int(X_dev.shape[0])
        if getattr(self.types, 'has_conditions', False) and self._G == 1:
            # a space with conditional hyper-parameters: every kernel value carries the activity mask of
            # gp_kernels.py:21-29.  The masked cross-covariance lives in the per-(row, surrogate) posterior
            # kernel (tlbo_gp_predict); the fused kernel then combines, scores and reduces the resident rows.
            mu, var = self._predict_slots(0, self.K + 1, X_dev)
            return ops.predict_ei_fused(self._logical(), self.types, dw, dskip, X_dev, eta, par=par,
                                        var_floor=var_floor, keys=keys, excluded=excluded, idx_offset=idx_offset,
                                        mode=self._fused_mode, order=dorder, want_rows=want_rows, ws=ws, sync=sync,
                                        nmax=8, cache=(mu, var))
        if N <= SMALL_QUERY_ROWS and (X_dev.data_ptr(), N) not in self._pool_cache:
            # small ad-hoc query sets (the online scenario's neighbourhoods and sampled configurations):
            # a pool tile is scored by ONE CTA walking the surrogates in turn, which leaves the device
            # idle here -- instead one warp per (row, surrogate) evaluates every posterior in parallel
            # (tlbo_gp_predict [+ tlbo_mcmc_marginalize]) and the fused kernel only combines, scores
            # and reduces the resident posteriors
            mu, var = self._predict_slots(0, self.K + 1, X_dev)
            return ops.predict_ei_fused(self._logical(), self.types, dw, dskip, X_dev, eta, par=par,
                                        var_floor=var_floor, keys=keys, excluded=excluded, idx_offset=idx_offset,
                                        mode=self._fused_mode, order=dorder, want_rows=want_rows, ws=ws, sync=sync,
                                        nmax=8, cache=(mu, var))
        if self._G > 1:
            # marginalised per-surrogate posteriors, all resident: sources from the pool cache (or
            # computed now for an ad-hoc query), the target's W walker GPs evaluated and
            # marginalised into row K; the fused kernel then only combines, scores and reduces
            cmu, cvar = self._group_posteriors(X_dev)
            return ops.predict_ei_fused(self._logical(), self.types, dw, dskip, X_dev, eta, par=par,
                                        var_floor=var_floor, keys=keys, excluded=excluded, idx_offset=idx_offset,
                                        mode=self._fused_mode, order=dorder, want_rows=want_rows, ws=ws, sync=sync,
                                        nmax=8, cache=(cmu, cvar))
        view = ops.pack_view(self._pack, 0, self.K + 1)
        nmax = self._nmax()
        from tlbo_b200.model import gp_large
        if nmax > gp_large.LARGE_N:
            # a surrogate beyond the fused kernel's shared-memory plan (SCoT's stacked GP): its pool posterior by
            # the two-step large-n path, then the fused kernel in all-resident mode (combine + EI + arg-max)
            if self.K != 0:
                raise ValueError('large-n surrogates are supported for single-surrogate facades only')
            mu, var = gp_large.posterior_pool(self._pack, self.K, self.types, X_dev)
            return ops.predict_ei_fused(self._logical(), self.types, dw, dskip, X_dev, eta, par=par,
                                        var_floor=var_floor, keys=keys, excluded=excluded, idx_offset=idx_offset,
                                        mode=self._fused_mode, order=dorder, want_rows=want_rows, ws=ws, sync=sync,
                                        nmax=8, cache=(mu, var))
        cache = self._pool_cache.get((X_dev.data_ptr(), int(X_dev.shape[0])))
        return ops.predict_ei_fused(view, self.types, dw, dskip, X_dev, eta, par=par, var_floor=var_floor,
                                    keys=keys, excluded=excluded, idx_offset=idx_offset, mode=self._fused_mode,
                                    order=dorder, want_rows=want_rows, ws=ws, sync=sync, nmax=nmax, cache=cache)

    def _nmax(self):
        """Largest training-set size among the surrogates in use (the kernels' shared-memory plans are sized by it)."""
        return max([self.num_src_hpo_trial if self.K else 0, getattr(self.target_surrogate, 'n_train_', 0)] +
                   [m.n_train_ for m in (self.source_surrogates or [])])

    def acq_small(self, X, eta, par, var_floor):
        """EI of a small host query set ``X [N, d]`` (N <= SMALL_QUERY_ROWS) through mapped pinned
        buffers: returns ``(ei float64[N], n_negative)``, or None when this facade has no fast path
        (walker groups).  Same kernels, same values as ``_fused(want_rows=True)``."""
        import torch
        from tlbo_b200 import ops
        N = int(X.shape[0])
        if self._G > 1 or N == 0 or N > SMALL_QUERY_ROWS:
            return None
        if self.target_surrogate is None:
            raise ValueError('Target surrogate is none.')
        sq = self._small
        if sq is None or sq.cap < N:
            sq = self._small = ops.SmallQuery(max(256, 2 * N), len(self.types), self.K + 1)
        Xp = sq.X_np[:N]
        Xp[...] = X
        Xp[~np.isfinite(Xp)] = -1
        # the ensemble weights / flags and the pack view only change with a fit: built once per training
        # (this path runs a few hundred times per online BO iteration)
        ens = self._ens_cache
        if ens is None or ens[0] is not self._pack:
            dw, dskip, dorder = self._ensemble_args_dev()
            ens = self._ens_cache = (self._pack, dw, dskip, dorder, ops.pack_view(self._pack, 0, self.K + 1))
        _, dw, dskip, dorder, view = ens
        view.c.nmax = self._nmax()
        ops.ensemble_ei_small(view, self.types, dw, dskip, dorder, self._fused_mode, sq, N, eta, par, var_floor)
        ops.stream_synchronize()
        ops.STATS.h2d += N * len(self.types) * 8
        ops.STATS.d2h += N * 8 + 32
        ei = sq.ei_np[:N].copy()
        n_neg = int(np.count_nonzero(ei < 0.0))
        ei[np.isnan(ei)] = -np.finfo(np.float64).max              # acquisition.py:72-76
        return ei, n_neg

    def climb(self, x0, acq0, rng, eta, par, var_floor, max_steps, const_mask, num_neighbors, stdev, stats):
        """One hill-climb of the online local search (ei_optimization.py:239-294) through the native driver
        ``tlbo_local_search_climb``: same values, same accept rule, same stream consumption as the per-step Python
        loop, one launch per step (``rng.randint(0, 10000)`` per step: a block of seeds is drawn ahead and the
        generator is rewound to exactly the number the climb consumed).  Returns ``(x, acq)`` or None when this
        facade has no native path (walker groups, spaces with conditions, large-n surrogates)."""
        import ctypes
        from tlbo_b200 import ops
        from tlbo_b200._cabi import check, lib, ptr
        from tlbo_b200.model import gp_large
        if (self._G > 1 or getattr(self.types, 'has_conditions', False) or self.target_surrogate is None
                or getattr(self.target_surrogate, 'n_train_', 0) > gp_large.LARGE_N):
            return None
        d = len(self.types)
        tarr = np.asarray(self.types, dtype=np.int64)
        cap = int(np.sum(np.where(tarr > 0, tarr - 1, num_neighbors))) + 1
        sq = self._small
        if sq is None or sq.cap < cap:
            sq = self._small = ops.SmallQuery(max(256, cap), d, self.K + 1)
        ens = self._ens_cache
        if ens is None or ens[0] is not self._pack:
            dw, dskip, dorder = self._ensemble_args_dev()
            ens = self._ens_cache = (self._pack, dw, dskip, dorder, ops.pack_view(self._pack, 0, self.K + 1))
        _, dw, dskip, dorder, view = ens
        view.c.nmax = self._nmax()
        t = ops._types_arr(self.types)
        cm = np.ascontiguousarray(const_mask, dtype=np.uint8)
        x = np.array(x0, dtype=np.float64)
        acq = ctypes.c_double(float(acq0))
        counters = np.zeros(5, dtype=np.int32)
        stream = ops._stream()
        while True:
            state = rng.get_state()
            seeds = rng.randint(0, 10000, size=64).astype(np.uint32)
            before = int(counters[0])
            rc = lib.tlbo_local_search_climb(view.ref(), ptr(t), ptr(cm), ptr(dw), ptr(dskip), ptr(dorder),
                                             int(self._fused_mode), float(eta), float(par), float(var_floor),
                                             int(num_neighbors), float(stdev), ptr(seeds), 64,
                                             -1 if max_steps is None else int(max_steps), ptr(x), ctypes.byref(acq),
                                             ptr(sq.X), ptr(sq.mu), ptr(sq.var), ptr(sq.ei), ptr(sq.counter),
                                             sq.cap, ptr(counters), stream)
            used = int(counters[0]) - before
            rng.set_state(state)                    # rewind to exactly the draws the climb consumed
            if used:
                rng.randint(0, 10000, size=used)
            if rc not in (0, 1):
                check(rc)
            if rc == 0:
                break
        stats['steps'] += int(counters[0])
        stats['neighbours'] += int(counters[1])
        stats['launches'] += int(counters[2])
        ops.STATS.launches['ensemble_ei_small'] = ops.STATS.launches.get('ensemble_ei_small', 0) + int(counters[2])
        if counters[3] > 0:
            raise ValueError("Expected Improvement is smaller than 0 for at least one sample.")
        # ConfigSpace's lazy generator: one np.random.random() from the GLOBAL stream per inspected neighbour
        np.random.random(int(counters[1]))
        return x, float(acq.value)

    # ---- MCMC-marginalised surrogates: group-level posteriors ----------------------------------
    def _logical(self):
        """K + 1 placeholder slots (n = 1) standing for the marginalised surrogates in the fused
        kernel's all-resident mode (it only reads their posteriors from HBM)."""
        from tlbo_b200 import ops
        if self._logical_pack is None:
            self._logical_pack = ops.SurrogatePack(self.K + 1, 8, len(self.types))
            self._logical_pack.n.fill_(1)
        return self._logical_pack

    def _walker_posteriors_into(self, slot, X_dev, out_mu, out_var):
        """Marginalised posterior of logical surrogate ``slot`` over ``X_dev`` into row ``slot`` of
        ``out_mu / out_var`` (gp_mcmc.py:418-440): W walker posteriors, then one reduction."""
        from tlbo_b200 import ops
        G = self._G
        view = ops.pack_view(self._pack, slot * G, G)
        mdl = self.target_surrogate if slot == self.K else self.source_surrogates[slot]
        mu_w, var_w = ops.predict_pool_posteriors(view, self.types, X_dev, nmax=mdl.n_train_)
        ops.mcmc_marginalize(mu_w, var_w, 1, G, out=(out_mu[slot:slot + 1], out_var[slot:slot + 1]))

    def _group_posteriors(self, X_dev):
        import torch
        key = (X_dev.data_ptr(), int(X_dev.shape[0]))
        hit = self._pool_cache.get(key)
        if hit is None:
            N = int(X_dev.shape[0])
            cmu = torch.zeros(self.K + 1, N, dtype=torch.float64, device=X_dev.device)
            cvar = torch.zeros(self.K + 1, N, dtype=torch.float64, device=X_dev.device)
            # facades without source surrogates (NoTL) skip those slots in the combination
            for k in range(self.K if self.source_surrogates is not None else 0):
                self._walker_posteriors_into(k, X_dev, cmu, cvar)
        else:
            cmu, cvar = hit
        self._walker_posteriors_into(self.K, X_dev, cmu, cvar)
        return cmu, cvar

    def _predict_slots(self, slot0, M, X_dev):
        """Device ``mu, var [M, n]`` of logical surrogates ``slot0 .. slot0 + M - 1`` at a small
        query set: ONE predict launch (+ one marginalisation for walker groups)."""
        from tlbo_b200 import ops
        G = self._G
        view = ops.pack_view(self._pack, slot0 * G, M * G)
        mu, var = ops.gp_predict(view, self.types, X_dev)
        if G > 1:
            mu, var = ops.mcmc_marginalize(mu, var, M, G)
        return mu, var

    def register_static_pool(self, X_dev):
        """Declare ``X_dev`` (device tensor [N, d]) an immutable candidate pool: the source
        surrogates' posteriors over it -- independent of the target observations, recomputed by
        the reference in every ``predict`` -- are evaluated once and kept resident in HBM
        (2 * K * N * 8 bytes).  Later fused calls on this very tensor read them back instead
        of re-evaluating K surrogates; results are bit-identical."""
        from tlbo_b200 import ops
        if self.K == 0 or self.source_surrogates is None:
            return
        if getattr(self.types, 'has_conditions', False):
            return                      # masked posteriors are evaluated per call (see _fused)
        key = (X_dev.data_ptr(), int(X_dev.shape[0]))
        if key in self._pool_cache:
            return
        # the cache is keyed by address and length: hold the tensor so that the caching allocator cannot hand
        # the same address to a different query tensor of the same length while the entry exists
        self._pool_keepalive = X_dev
        if self._G > 1:
            import torch
            N = int(X_dev.shape[0])
            cmu = torch.empty(self.K + 1, N, dtype=torch.float64, device=X_dev.device)
            cvar = torch.empty(self.K + 1, N, dtype=torch.float64, device=X_dev.device)
            for k in range(self.K):            # row K is the target's, refreshed by every fused call
                self._walker_posteriors_into(k, X_dev, cmu, cvar)
            self._pool_cache = {key: (cmu, cvar)}
            return
        view = ops.pack_view(self._pack, 0, self.K)
        nmax = max(m.n_train_ for m in self.source_surrogates)
        self._pool_cache = {key: ops.predict_pool_posteriors(view, self.types, X_dev, nmax=nmax)}

    def _to_device(self, X):
        import torch
        from tlbo_b200 import ops
        if torch.is_tensor(X):
            return X
        X = np.array(X, dtype=np.float64)
        X[~np.isfinite(X)] = -1
        return ops.h2d(X)

    def _predict_rows(self, X, var_floor):
        if len(X.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X.shape))
        if X.shape[1] != len(self.types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.types), X.shape[1]))
        if self.target_surrogate is None:
            raise ValueError('Target surrogate is none.')
        from tlbo_b200 import ops
        _, mu, var, _ = self._fused(self._to_device(X), var_floor=var_floor, want_rows=True)
        return ops.d2h(mu).reshape((-1, 1)), ops.d2h(var).reshape((-1, 1))

    def predict_marginalized_over_instances(self, X):
        """base_facade.py:110-143: ensemble predict with the 1e-10 variance floor."""
        return self._predict_rows(X, self.var_threshold)

    def combine_predictions(self, X, combination_method='idp_lc'):
        """base_facade.py:145-183.  Only 'idp_lc' (independent linear combination) is on
        the benchmark path; it is what the fused kernel's mode 0 evaluates."""
        if combination_method != 'idp_lc':
            raise ValueError('Invalid combination method %s.' % combination_method)
        return self._predict_rows(X, 0.0)

    def _source_predictions(self, X, n_extra=0):
        """mu, var [K + 1 + n_extra, n] of every pack slot at the target points -- the K
        source predicts of rgpe.py:27-30 / topo_variant1.py:27-35 plus the CV surrogates,
        in ONE launch."""
        from tlbo_b200 import ops
        mu, var = self._predict_slots(0, self.K + 1 + n_extra, self._to_device(X))
        return ops.d2h(mu), ops.d2h(var)
