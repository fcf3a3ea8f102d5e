from vgtk.spconv import SphericalPointCloud, SphericalPointCloudPose
from .functional import *
from .modules import *
