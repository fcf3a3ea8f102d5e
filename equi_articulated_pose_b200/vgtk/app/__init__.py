"""vgtk.app -- host-side plumbing the reference's run script imports (Trainer base, logger, flags).

Out of the hot-path scope: kept minimal so that `import vgtk` exposes the
same names as the reference package (vgtk/vgtk/app/__init__.py)."""
from .trainer import Trainer
from .parse_config import HierarchyArgmentParser, dump_args
from .logger import Logger
from .summary import Summary
from .timer import Timer
