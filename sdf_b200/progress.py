"""Minimal console progress line for generate(verbose=True).

The reference prints a ``\\r``-rewritten bar with percentage, counts and ETA
(sdf/progress.py); it is cosmetic and off the hot path.  This is an independent,
smaller implementation with the same ``Bar(n, enabled=...)`` /
``increment()`` / ``done()`` calls that core.generate makes.
"""
import sys
import time


def _hms(t):
    t = int(t + 0.5)
    return '%d:%02d:%02d' % (t // 3600, t // 60 % 60, t % 60)


class Bar:
    WIDTH = 30

    def __init__(self, total, enabled=True, out=None):
        self.total = max(int(total), 1)
        self.count = 0
        self.enabled = enabled
        self.out = out or sys.stdout
        self.t0 = time.time()

    def _draw(self):
        if not self.enabled:
            return
        frac = min(self.count / self.total, 1.0)
        fill = int(frac * self.WIDTH + 0.5)
        dt = time.time() - self.t0
        left = dt * (1 - frac) / frac if frac > 0 else 0
        self.out.write('  %3.0f%% (%d of %d) [%s%s] %s %s    \r' % (
            frac * 100, self.count, self.total, '#' * fill, '-' * (self.WIDTH - fill),
            _hms(dt), _hms(left)))
        self.out.flush()

    def increment(self, n=1):
        self.count += n
        self._draw()

    def done(self):
        self.count = self.total
        self._draw()
        if self.enabled:
            self.out.write('\n')
            self.out.flush()
