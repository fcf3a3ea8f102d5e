from .base import PointSet
