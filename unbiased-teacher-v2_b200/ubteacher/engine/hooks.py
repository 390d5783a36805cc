"""Training-loop hooks with the roles of the [D2] hooks the reference trainer installs
(ubteacher/engine/trainer.py:503-552 ``build_hooks``): PeriodicCheckpointer, two EvalHooks (student, teacher),
PeriodicWriter. ([D2] IterationTimer / LRScheduler are folded into the loop; PreciseBN does not apply: FrozenBN.)"""
import json
import logging
import os
import time

logger = logging.getLogger(__name__)


class HookBase:
    trainer = None

    def before_train(self):
        pass

    def after_train(self):
        pass

    def before_step(self):
        pass

    def after_step(self):
        pass


class PeriodicCheckpointer(HookBase):
    """[D2] hooks.PeriodicCheckpointer: ``model_{iter:07d}`` every ``period`` iterations, ``model_final`` after the last."""

    def __init__(self, checkpointer, period, max_iter=None):
        self.checkpointer, self.period, self.max_iter = checkpointer, int(period), max_iter

    def before_train(self):
        self.max_iter = self.trainer.max_iter

    def after_step(self):
        it = self.trainer.iter
        if self.period > 0 and (it + 1) % self.period == 0 and it + 1 < self.max_iter:
            self.checkpointer.save(f"model_{it:07d}", iteration=it)
        if it + 1 >= self.max_iter:
            self.checkpointer.save("model_final", iteration=it)


class EvalHook(HookBase):
    """[D2] hooks.EvalHook: run ``fn`` every ``period`` iterations and after the last one; flattened scalar results go to
    the event storage."""

    def __init__(self, period, fn):
        self.period, self.fn = int(period), fn

    def _do(self):
        res = self.fn() or {}
        flat = {}
        for k, v in res.items():
            if isinstance(v, dict):
                flat.update({f"{k}/{kk}": vv for kk, vv in v.items()})
            else:
                flat[k] = v
        st = self.trainer.storage
        for k, v in flat.items():
            try:
                st.put_scalar(k, float(v), smoothing_hint=False)
            except (TypeError, ValueError):
                pass

    def after_step(self):
        nxt = self.trainer.iter + 1
        if self.period > 0 and nxt % self.period == 0 and nxt != self.trainer.max_iter:
            self._do()

    def after_train(self):
        if self.period > 0 and self.trainer.iter + 1 >= self.trainer.max_iter:
            self._do()


class JSONWriter:
    """[D2] JSONWriter: one JSON object per flush with the latest scalars (``metrics.json``)."""

    def __init__(self, path):
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        self._f = open(path, "a")
        self._last = -1

    def write(self, storage):
        rec = {"iteration": storage.iter}
        for k, (v, it) in storage.latest().items():
            if it > self._last:
                rec[k] = v
        self._last = storage.iter
        self._f.write(json.dumps(rec, sort_keys=True) + "\n")
        self._f.flush()

    def close(self):
        self._f.close()


class CommonMetricPrinter:
    """[D2] CommonMetricPrinter, reduced: iteration, total loss, the loss terms, lr, seconds per iteration."""

    def __init__(self, max_iter):
        self.max_iter, self._t, self._it = max_iter, time.perf_counter(), None

    def write(self, storage):
        lat = storage.latest()
        now = time.perf_counter()
        per_it = (now - self._t) / max(storage.iter - self._it, 1) if self._it is not None else float("nan")
        self._t, self._it = now, storage.iter
        losses = "  ".join(f"{k}: {v:.4g}" for k, (v, _) in lat.items() if "loss" in k)
        lr = lat.get("lr", (float("nan"), 0))[0]
        logger.info("iter: %d/%d  %s  lr: %.5g  time: %.4f s/it", storage.iter, self.max_iter, losses, lr, per_it)

    def close(self):
        pass


class PeriodicWriter(HookBase):
    def __init__(self, writers, period=20):
        self.writers, self.period = writers, int(period)

    def after_step(self):
        it = self.trainer.iter
        if (it + 1) % self.period == 0 or it == self.trainer.max_iter - 1:
            for w in self.writers:
                w.write(self.trainer.storage)

    def after_train(self):
        for w in self.writers:
            w.close()
