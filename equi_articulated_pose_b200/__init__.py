"""equi_articulated_pose_b200 -- B200 (sm_100a) implementation of the vgtk SE(3)-equivariant
point-convolution hot path of Meowuu7/equi-articulated-pose.

    import equi_articulated_pose_b200 as eap
    eap.install()            # `import vgtk`, `import chamfer`, `import extensions.chamfer_dist` now
                             # resolve to this package (same module paths as the reference)

There is no CPU fallback: the operators raise if libvgtkb200.so (csrc/, built by
__graft_entry__.build()) is missing or the tensors are not on a B200.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
__version__ = "0.1.0"


def install():
    """Expose the drop-in packages under the reference's import names."""
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)       # contains vgtk/, extensions/, chamfer.py
    for name in ("vgtk", "chamfer", "extensions"):
        mod = sys.modules.get(name)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(_HERE):
            raise RuntimeError(f"another `{name}` is already imported from {getattr(mod, '__file__', None)}")
    shims = os.path.join(_HERE, "shims")          # torch_cluster.fps on the new FPS kernel; a real installation wins
    if shims not in sys.path:
        sys.path.append(shims)
    return importlib.import_module("vgtk")
