from . import rasterize
