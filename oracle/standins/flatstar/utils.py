"""``flatstar.utils`` stand-in."""
import numpy as np


def cylindrical_r(grid, reference=(0, 0)):
    ny, nx = np.shape(grid)
    yy, xx = np.mgrid[0:ny, 0:nx]
    return np.sqrt((xx - reference[0]) ** 2 + (yy - reference[1]) ** 2)
