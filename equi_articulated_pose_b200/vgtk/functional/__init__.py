from .rotation import *
