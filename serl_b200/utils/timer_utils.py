"""Wall-clock section timer with the reference's interface (utils/timer_utils.py: `Timer.tick/tock/context/
get_average_times`), used by the learner loops (`with timer.context("train_critics"): ...`).  Like the reference's
numbers these are DISPATCH times of asynchronous device work unless the section synchronises."""
from __future__ import annotations

import contextlib
import time


class Timer:
    def __init__(self):
        self.reset()

    def reset(self):
        self._sum, self._n, self._open = {}, {}, {}

    def tick(self, key):
        if key in self._open:
            raise ValueError(f"Timer is already ticking for key: {key}")
        self._open[key] = time.time()

    def tock(self, key):
        if key not in self._open:
            raise ValueError(f"Timer is not ticking for key: {key}")
        dt = time.time() - self._open.pop(key)
        self._sum[key] = self._sum.get(key, 0.0) + dt
        self._n[key] = self._n.get(key, 0) + 1

    @contextlib.contextmanager
    def context(self, key):
        self.tick(key)
        try:
            yield
        finally:
            self.tock(key)

    def get_average_times(self, reset=True):
        out = {k: self._sum[k] / self._n[k] for k in self._n}
        if reset:
            self.reset()
        return out
