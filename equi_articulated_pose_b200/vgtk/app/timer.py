import time


class Timer():
    """Wall-clock stopwatch (reference: vgtk/vgtk/app/timer.py:3-16)."""

    def __init__(self):
        self.start_time = None

    def set_point(self):
        self.start_time = time.time()

    def reset(self):
        self.start_time = None

    def elapsed(self):
        return 0.0 if self.start_time is None else time.time() - self.start_time

    def reading(self):
        return self.elapsed()
