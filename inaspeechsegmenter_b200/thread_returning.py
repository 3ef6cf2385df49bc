"""Thread whose ``join()`` hands back the target's return value -- the helper
the reference's prefetch loop relies on (thread_returning.py:11-25)."""
from threading import Thread


class ThreadReturning(Thread):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._result = None

    def run(self):
        target, args, kwargs = self._target, self._args, self._kwargs
        if target is not None:
            self._result = target(*args, **kwargs)

    def join(self, timeout=None):
        super().join(timeout)
        return self._result
