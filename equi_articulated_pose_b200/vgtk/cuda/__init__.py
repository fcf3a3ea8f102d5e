"""vgtk.cuda: the extension-module surface of the reference (pybind modules `grouping`,
`gathering`, `zpconv`), served by libvgtkb200.so through ctypes."""
from . import gathering, grouping, zpconv
