"""``inference_on_dataset`` / ``inference_context`` with the reference's role (ubteacher/evaluation/evaluator.py:14-104,
SURVEY.md §8(f) rank 3): put the model in eval mode, run it over a fixed-length loader — FCOS models are called with
``nms_method=cfg.MODEL.FCOS.NMS_CRITERIA_TEST`` like the reference does — feed (inputs, outputs) to the evaluator and return
``evaluator.evaluate()``. The [D2] ``DatasetEvaluator`` protocol (reset / process / evaluate) is all an evaluator needs."""
import inspect
import logging
import time
from contextlib import contextmanager

import torch

logger = logging.getLogger(__name__)


class DatasetEvaluators:
    """[D2] DatasetEvaluators: fan out to several evaluators and merge their result dicts."""

    def __init__(self, evaluators):
        self._evaluators = list(evaluators)

    def reset(self):
        for e in self._evaluators:
            e.reset()

    def process(self, inputs, outputs):
        for e in self._evaluators:
            e.process(inputs, outputs)

    def evaluate(self):
        merged = {}
        for e in self._evaluators:
            res = e.evaluate()
            if res:
                overlap = set(merged) & set(res)
                assert not overlap, f"different evaluators produce results with the same key {overlap}"
                merged.update(res)
        return merged


@contextmanager
def inference_context(model):
    """Temporarily eval(); the previous training flag is restored afterwards."""
    was_training = model.training
    model.eval()
    try:
        yield
    finally:
        model.train(was_training)


def inference_on_dataset(model, data_loader, evaluator, cfg):
    total = len(data_loader)            # the inference loader must have a fixed length
    evaluator = DatasetEvaluators([]) if evaluator is None else evaluator
    evaluator.reset()
    takes_nms = "nms_method" in inspect.signature(model.forward).parameters
    kwargs = {"nms_method": cfg.MODEL.FCOS.NMS_CRITERIA_TEST} if takes_nms else {}
    warmup = min(5, total - 1)
    compute = 0.0
    t_start = time.perf_counter()
    with inference_context(model), torch.no_grad():
        for idx, inputs in enumerate(data_loader):
            if idx == warmup:
                t_start, compute = time.perf_counter(), 0.0
            t0 = time.perf_counter()
            outputs = model(inputs, **kwargs)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            compute += time.perf_counter() - t0
            evaluator.process(inputs, outputs)
    n = max(total - warmup, 1)
    logger.info("inference: %.4f s / batch (%.4f s pure compute) over %d batches", (time.perf_counter() - t_start) / n,
                compute / n, n)
    results = evaluator.evaluate()
    return {} if results is None else results
