import os
import time


class Logger():
    """stdout + file logger (reference: vgtk/vgtk/app/logger.py:11-37)."""

    def __init__(self, log_file=None):
        self.log_file = log_file
        if log_file is not None:
            os.makedirs(os.path.dirname(os.path.abspath(log_file)), exist_ok=True)

    def log(self, tag, msg):
        line = f"[{time.strftime('%Y-%m-%d %H:%M:%S')}] [{tag}] {msg}"
        print(line, flush=True)
        if self.log_file is not None:
            with open(self.log_file, 'a') as fh:
                fh.write(line + "\n")
