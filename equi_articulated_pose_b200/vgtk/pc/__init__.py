from .io import load_ply, save_ply
from .sample import *
