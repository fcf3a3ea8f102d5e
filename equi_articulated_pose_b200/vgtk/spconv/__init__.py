from .base import SphericalPointCloud, SphericalPointCloudPose
from .functional import *
from .modules import *
