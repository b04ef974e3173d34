"""Acquisition maximiser over a fixed candidate table
(reference tlbo/optimizer/ei_offline_optimizer.py:10-29 ``OfflineSearch``,
tlbo/optimizer/ei_optimization.py:50-132).

The table is uploaded once and stays resident in HBM.  ``maximize`` consumes one
``rng.rand(N)`` per call exactly like ``_sort_configs_by_acq_value`` and returns the
candidates best-first; by default only the head of that order is materialised (the caller
-- smbo_offline.py:261-263 -- takes the first unevaluated entry), via the fused arg-max."""
import numpy as np

from tlbo_b200.config_space import Configuration, convert_configurations_to_array


class AcquisitionFunctionMaximizer(object):
    def __init__(self, acquisition_function, config_space, rng=None):
        self.acquisition_function = acquisition_function
        self.config_space = config_space
        if rng is None:
            self.rng = np.random.RandomState(seed=1)
        else:
            self.rng = rng

    def maximize(self, runhistory, num_points, **kwargs):
        return [t[1] for t in self._maximize(runhistory, num_points, **kwargs)]


class OfflineSearch(AcquisitionFunctionMaximizer):
    def __init__(self, configuration_list, acquisition_function, config_space, rng=None, shard=None,
                 prefetch_keys=False, device_keys=None):
        super().__init__(acquisition_function, config_space, rng)
        import torch
        from tlbo_b200 import ops
        if isinstance(configuration_list, np.ndarray):
            self.X_table = np.ascontiguousarray(configuration_list, dtype=np.float64)
            self.configuration_list = None
        else:
            self.configuration_list = configuration_list
            self.X_table = convert_configurations_to_array(configuration_list)
        X = self.X_table.copy()
        X[~np.isfinite(X)] = -1
        self.N = X.shape[0]
        # candidate-pool sharding (tlbo_b200.distributed.PoolShard): this rank's row range
        self.shard = shard
        lo, hi = (0, self.N) if shard is None else shard.row_range(self.N)
        self.lo, self.hi = lo, hi
        ops._dev()        # no CPU fallback: raises without a CUDA device
        self._host_rows = torch.from_numpy(X[lo:hi]).pin_memory()
        self.X_dev = self._host_rows.to('cuda', non_blocking=True)
        self._ws = ops.FusedWorkspace()
        # two pinned staging buffers for the tie-break keys (same memory as numpy views)
        self._keys_host = [torch.empty(hi - lo, dtype=torch.float64).pin_memory() for _ in range(2)]
        self._keys_host_np = [t.numpy() for t in self._keys_host]
        self._keys_flip = 0
        # prefetch_keys: the keys of the NEXT maximisation -- rng.rand(N), independent of the BO
        # state -- are drawn by a worker thread (numpy releases the GIL) while the current
        # iteration runs.  Only valid when `rng` is owned by this maximiser (SMBO_OFFLINE's is).
        self._prefetch = bool(prefetch_keys)
        self._pool = None
        self._future = None
        self._keys_dev = torch.empty(hi - lo, dtype=torch.float64, device='cuda')
        # device_keys (default: whenever this maximiser owns its rng, i.e. with prefetch_keys): the MT19937
        # state of `rng` lives on the device and `rng.rand(N)` is produced there, bit for bit, on a side
        # stream under the fit kernel -- no host generation (5 ms at N = 1e6), no 8 MB upload per iteration
        self._device_keys = bool(prefetch_keys) if device_keys is None else bool(device_keys)
        if self._device_keys:
            self._prefetch = False
            self._dev_rng = ops.DeviceRandomState(self.rng, n_hint=self.N)
            # two key buffers: the keys of the NEXT maximisation (independent of the BO state) are generated
            # right after the current one has been enqueued, so the generator never sits on the critical path
            self._keys_full = [torch.empty(self.N, dtype=torch.float64, device='cuda') for _ in range(2)]
            self._key_cur = 0
            self._keys_after = None
            self._key_stream = torch.cuda.Stream()
            self._state_before = torch.empty_like(self._dev_rng.state)
            self._generate_keys(0)
            self._keys_dev = self._keys_full[0][lo:hi]
        # the table is fixed for the lifetime of this maximiser: let an ensemble facade keep its
        # source surrogates' posteriors over it resident
        model = getattr(acquisition_function, 'model', None)
        if hasattr(model, 'register_static_pool'):
            model.register_static_pool(self.X_dev)
        self.last_best = None
        self._prepared = None
        self._copy_stream = None
        self._copy_pending = False
        self._stage = ops.PinnedStage()

    def reupload(self):
        """Host -> device copy of the candidate rows (the end-to-end benchmark path, where
        the caller hands host configurations on every call like the reference does)."""
        import torch
        from tlbo_b200 import ops
        # on a copy stream: the candidate rows are first read by the fused kernel at the END of the
        # iteration, so the transfer runs under the fit kernel instead of in front of it (the previous
        # reader of X_dev, the last maximisation's kernel, has completed -- its result was read back)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        with torch.cuda.stream(self._copy_stream):
            self.X_dev.copy_(self._host_rows, non_blocking=True)
        self._copy_pending = True
        ops.STATS.h2d += self._host_rows.numel() * 8

    def _config(self, idx):
        if self.configuration_list is not None:
            return self.configuration_list[idx]
        return Configuration(self.config_space, self.X_table[idx], idx)

    def prepare(self, evaluated_rows):
        """Host half of one maximisation -- the ``rng.rand(N)`` tie-break keys of
        ``_sort_configs_by_acq_value`` and the sorted evaluated-row list -- staged through
        pinned memory with asynchronous uploads, so that it can run while the facade's fit
        kernel is in flight."""
        import torch
        from tlbo_b200 import ops
        if self._device_keys:
            # the keys of this maximisation were generated ahead of time on the side stream
            # (`_generate_keys`); the main stream waits for them right before the fused kernel
            self._keys_dev = self._keys_full[self._key_cur][self.lo:self.hi]
            ex = None
            if len(evaluated_rows):
                ex = self._stage.upload('excluded', np.sort(np.asarray(evaluated_rows, dtype=np.int64)))
            self._prepared = (ex,)
            return
        slot = self._draw_keys() if self._future is None else self._future.result()
        self._future = None
        self._keys_dev.copy_(self._keys_host[slot], non_blocking=True)
        ops.STATS.h2d += self._keys_host[slot].numel() * 8
        if self._prefetch:
            if self._pool is None:
                from concurrent.futures import ThreadPoolExecutor
                self._pool = ThreadPoolExecutor(max_workers=1)
            self._future = self._pool.submit(self._draw_keys)
        ex = None
        if len(evaluated_rows):
            ex = self._stage.upload('excluded', np.sort(np.asarray(evaluated_rows, dtype=np.int64)))
        self._prepared = (ex,)

    def _generate_keys(self, buf):
        """Enqueue ``rng.rand(N)`` into key buffer ``buf`` on the side stream.  The buffer's last reader (the
        fused kernel of the maximisation before the previous one) completed long ago -- its result was read
        back --, so the side stream never waits for the main stream (where a fit may be running)."""
        import torch
        with torch.cuda.stream(self._key_stream):
            self._state_before.copy_(self._dev_rng.state)          # lets sync_host_rng() undo the look-ahead
            self._dev_rng.rand_into(self._keys_full[buf])

    def _consume_keys(self):
        """Called once the current maximisation's kernel is enqueued: switch buffers and start generating the
        next maximisation's keys -- BEHIND that kernel: the host enqueues it while the fit is still running, and a
        generator that started right away shared the SMs with the post-fit kernels (the fused kernel's second
        CTA per SM then waited for a free slot: 0.09 -> 0.17 ms in some runs).  Behind it the generator runs in
        the host gap before the next fit and has two milliseconds to spare."""
        import torch
        self._key_cur ^= 1
        if self._keys_after is None:
            self._keys_after = torch.cuda.Event()
        self._keys_after.record()
        self._key_stream.wait_event(self._keys_after)
        self._generate_keys(self._key_cur)

    def sync_host_rng(self):
        """Device-keys mode: bring the host ``rng`` object to the state the reference's maximiser would have
        now (the device generator is one ``rand(N)`` ahead because of the look-ahead buffer)."""
        import torch
        if not self._device_keys:
            return self.rng
        self._key_stream.synchronize()
        h = self._state_before.cpu().numpy().view(np.uint32)
        st = self.rng.get_state()
        self.rng.set_state(('MT19937', h[:624].copy(), int(h[624]), st[3], st[4]))
        return self.rng

    def _draw_keys(self):
        """One ``rng.rand(N)`` (the reference's tie-break draw) into the next pinned buffer.
        numpy memcpy, not torch copy_: torch's multi-threaded CPU copy stalls for milliseconds
        when it contends with the BLAS threads of scipy's SLSQP."""
        keys = self.rng.rand(self.N)
        self._keys_flip ^= 1
        self._keys_host_np[self._keys_flip][:] = keys[self.lo:self.hi]
        return self._keys_flip

    def best_unevaluated(self, evaluated_rows):
        """Row index the reference's loop would return: first entry of the descending
        (EI, random key) order that is not in ``evaluated_rows``."""
        if self._prepared is None:
            self.prepare(evaluated_rows)
        ex, = self._prepared
        self._prepared = None
        if self._copy_pending:
            import torch
            torch.cuda.current_stream().wait_stream(self._copy_stream)
            self._copy_pending = False
        if self._device_keys:
            import torch
            torch.cuda.current_stream().wait_stream(self._key_stream)
        if self.shard is None:
            best_dev = self.acquisition_function.best_unevaluated(self.X_dev, keys=self._keys_dev, excluded=ex,
                                                                  ws=self._ws, sync=False)[0]
            if self._device_keys:
                self._consume_keys()              # next maximisation's keys: generated under everything that follows
            from tlbo_b200 import ops
            b = ops.d2h(best_dev)
            res = (float(b[0]), float(b[1]), int(b[2]), int(b[3]))
            if res[3] > 0:
                raise ValueError("Expected Improvement is smaller than 0 for at least one sample.")
        else:
            best_dev = self.acquisition_function.best_unevaluated(self.X_dev, keys=self._keys_dev, excluded=ex,
                                                                  idx_offset=self.lo, ws=self._ws, sync=False)[0]
            if self._device_keys:
                self._consume_keys()
            res = self.shard.reduce_best(best_dev)
            if res[3] > 0:
                raise ValueError("Expected Improvement is smaller than 0 for at least one sample.")
        self.last_best = res
        if res[2] < 0:
            raise ValueError('The configuration in the SET (%d) is over' % self.N)
        return res[2]

    def _maximize(self, runhistory, num_points, evaluated_rows=(), full_order=False, **kwargs):
        if not full_order:
            idx = self.best_unevaluated(evaluated_rows)
            return [(self.last_best[0], self._config(idx))]
        # full descending order (reference semantics, O(N log N) on the host)
        acq = self.acquisition_function(self.X_table)
        if self._device_keys:
            # the look-ahead buffer already holds this call's rng.rand(N)
            self._key_stream.synchronize()
            random = self._keys_full[self._key_cur].cpu().numpy()
            self._consume_keys()
        else:
            random = self.rng.rand(self.N)
        indices = np.lexsort((random.flatten(), acq.flatten()))
        return [(acq[ind][0], self._config(ind)) for ind in indices[::-1]]
