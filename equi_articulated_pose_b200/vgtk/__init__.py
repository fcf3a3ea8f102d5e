"""B200-native `vgtk`: same module paths and operator names as the reference package
(vgtk/vgtk/__init__.py:1-11), backed by libvgtkb200.so.

Make it importable under the reference's name with
    import equi_articulated_pose_b200 as eap; eap.install()      # then: import vgtk, chamfer, extensions.chamfer_dist
"""
from . import functional
from . import point3d
from . import pc
from . import spconv
from . import so3conv

from .app import *
from .loss import *
from .utils import batch_gather, batch_zip, LearningRateScheduler
