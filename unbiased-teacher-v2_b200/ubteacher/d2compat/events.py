"""Tiny EventStorage (detectron2.utils.events) — scalars per iteration, a context stack, nothing else."""
from collections import defaultdict

_STACK = []


def get_event_storage():
    assert len(_STACK), "get_event_storage() has to be called inside a 'with EventStorage(...)' context!"
    return _STACK[-1]


class EventStorage:
    def __init__(self, start_iter=0):
        self._history = defaultdict(list)
        self._latest = {}
        self._iter = start_iter

    def put_scalar(self, name, value, smoothing_hint=True):
        value = float(value)
        self._history[name].append((value, self._iter))
        self._latest[name] = (value, self._iter)

    def put_scalars(self, *, smoothing_hint=True, **kwargs):
        for k, v in kwargs.items():
            self.put_scalar(k, v, smoothing_hint=smoothing_hint)

    def history(self, name):
        return self._history[name]

    def latest(self):
        return self._latest

    def step(self):
        self._iter += 1

    @property
    def iter(self):
        return self._iter

    @iter.setter
    def iter(self, val):
        self._iter = int(val)

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        assert _STACK[-1] == self
        _STACK.pop()
