"""`custma.Timer` / `custma.utils.TimerError`: the wall-clock helper of the reference (custma/utils.py:6-87),
re-implemented with the same observable behaviour:

  * `Timer(print_tmpl=None, start=True)`; a template without a `{:.Nf}` field gets " {:.3f}" appended (:38-42),
    the default template is "{:.3f}";
  * as a context manager it (re)starts on entry and on exit prints the template formatted with the seconds since
    the last check, then stops (:51-57);
  * `since_start()` / `since_last_check()` raise TimerError when the timer is not running (:66-87); both count as
    a "check".

Like the reference it reads the host clock only and does NOT synchronise CUDA: around an asynchronous kernel launch
it measures launch latency.  Use `cuda_sync=True` (an addition, off by default) to synchronise the current device
at start and at every check; `bench.py` uses CUDA events instead.
"""
import re
import time

_FLOAT_FIELD = re.compile(r"({:.*\df})")


class TimerError(Exception):
    def __init__(self, message):
        super().__init__(message)
        self.message = message


class Timer:
    def __init__(self, print_tmpl=None, start=True, cuda_sync=False):
        self._running = False
        self._cuda_sync = bool(cuda_sync)
        if print_tmpl is not None and not _FLOAT_FIELD.findall(print_tmpl):
            print_tmpl = print_tmpl + " {:.3f}"
        self.print_tmpl = print_tmpl or "{:.3f}"
        if start:
            self.start()

    def _now(self):
        if self._cuda_sync:
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()
        return time.time()

    @property
    def is_running(self):
        return self._running

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, exc_type, exc_value, tb):
        print(self.print_tmpl.format(self.since_last_check()))
        self._running = False

    def start(self):
        now = self._now()
        if not self._running:
            self._t_start = now
            self._running = True
        self._t_last = now

    def _require_running(self):
        if not self._running:
            raise TimerError("timer is not running")

    def since_start(self):
        self._require_running()
        self._t_last = self._now()
        return self._t_last - self._t_start

    def since_last_check(self):
        self._require_running()
        now = self._now()
        elapsed = now - self._t_last
        self._t_last = now
        return elapsed
