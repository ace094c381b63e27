"""2D grid factory -- drop-in for contourist/field2d.py:8-9."""
from . import grid_field


def Function2DGrid(xmin, ymin, xmax, ymax, dx, dy, function, materialize=False, cache=False):
    return grid_field.FunctionGrid((xmin, ymin), (xmax, ymax), (dx, dy), function, materialize, cache)
