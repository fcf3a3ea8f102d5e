"""Build the reference's OWN CUDA extensions (test infrastructure only).

This is the "oracle/_ref" recipe: it compiles the reference kernels from the
sources where they lie under /root/reference, for sm_100a, into
``oracle/_ref/`` (git-ignored, travels to the GPU box with the snapshot).
Nothing of the reference is committed: the only edit is the mechanical
``AT_DISPATCH_FLOATING_TYPES(x.type()`` -> ``x.scalar_type()`` substitution that
torch >= 2.x needs, applied in memory and written to the
build scratch directory.

The resulting pybind modules are used ONLY by ``tests/`` (bit-exact index
parity against the real reference kernels on the B200) and by
``bench.py --impl reference`` / ``profiles`` as the recompiled-reference
baseline.  The product never imports them.

    python oracle/build_ref.py            # builds vgtk_ref_grouping, vgtk_ref_gathering, chamfer_ref
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("VGTK_REFERENCE_ROOT", "/root/reference")

TARGETS = {
    # module name -> (source dir, [files])
    "vgtk_ref_grouping": ("vgtk/vgtk/cuda", ["grouping_cuda.cpp", "grouping_cuda_kernel.cu"]),
    "vgtk_ref_gathering": ("vgtk/vgtk/cuda", ["gathering_cuda.cpp", "gathering_cuda_kernel.cu"]),
    "chamfer_ref": ("extensions/chamfer_dist", ["chamfer_cuda.cpp", "chamfer.cu"]),
}


def _patched(text: str) -> str:
    # torch 2.x: DeprecatedTypeProperties no longer converts to ScalarType.
    text = re.sub(r"(AT_DISPATCH_FLOATING_TYPES\(\s*[A-Za-z_0-9\.]+?)\.type\(\)", r"\1.scalar_type()", text)
    text = text.replace(".type().is_cuda()", ".is_cuda()")
    return text


def build(names=None, verbose=False):
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree not present at {REF}; oracle/_ref can only be built where it is mounted")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    built = {}
    for name, (sub, files) in TARGETS.items():
        if names and name not in names:
            continue
        so = os.path.join(OUT, name + ".so")
        if os.path.exists(so):
            built[name] = so
            continue
        scratch = os.path.join(OUT, "build_" + name)
        os.makedirs(scratch, exist_ok=True)
        srcs = []
        for f in files:
            with open(os.path.join(REF, sub, f)) as fh:
                txt = _patched(fh.read())
            dst = os.path.join(scratch, f)
            with open(dst, "w") as fh:
                fh.write(txt)
            srcs.append(dst)
        load(name=name, sources=srcs, build_directory=scratch, verbose=verbose,
             extra_cuda_cflags=["-O3", "-lineinfo"], is_python_module=False)
        os.replace(os.path.join(scratch, name + ".so"), so)
        shutil.rmtree(scratch, ignore_errors=True)      # patched source copies and objects do not stay around
        built[name] = so
    return built


def load_ref(name):
    """Import a previously built reference module (tests only)."""
    import importlib.util
    import torch  # noqa: F401  (must be imported before the extension)
    so = os.path.join(OUT, name + ".so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(sys.argv[1:] or None, verbose=True))
