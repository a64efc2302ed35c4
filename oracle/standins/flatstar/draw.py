"""``flatstar.draw`` stand-in: Pillow-ellipse disks + flux-conserving box rebin."""
import numpy as np
from PIL import Image, ImageDraw


def _disk(center, radius, shape):
    top_left = (center[0] - radius, center[1] - radius)
    bottom_right = (center[0] + radius, center[1] + radius)
    image = Image.new('1', (shape[1], shape[0]))
    ImageDraw.Draw(image).ellipse([top_left, bottom_right], outline=1, fill=1)
    return np.array(image).astype(float)


def _limb_darkening(mu, law, c):
    if law is None:
        return np.ones_like(mu)
    if law == 'linear':
        return 1 - c * (1 - mu)
    c = np.asarray(c, dtype=float)
    if law == 'quadratic':
        return 1 - c[0] * (1 - mu) - c[1] * (1 - mu) ** 2
    if law == 'square-root':
        return 1 - c[0] * (1 - mu) - c[1] * (1 - mu ** 0.5)
    if law == 'log':
        with np.errstate(all='ignore'):
            t = np.where(mu > 0, mu * np.log(np.where(mu > 0, mu, 1.0)), 0.0)
        return 1 - c[0] * (1 - mu) - c[1] * t
    if law == 'exp':
        return 1 - c[0] * (1 - mu) - c[1] / (1 - np.exp(mu))
    if law == 's3':
        return 1 - c[0] * (1 - mu) - c[1] * (1 - mu ** 1.5) - c[2] * (1 - mu ** 2)
    if law == 'c4':
        return 1 - c[0] * (1 - mu ** 0.5) - c[1] * (1 - mu) \
            - c[2] * (1 - mu ** 1.5) - c[3] * (1 - mu ** 2)
    raise ValueError('unknown limb-darkening law')


class _Grid:
    pass


def star(grid_size, radius=0.5, limb_darkening_law=None, ld_coefficient=None):
    shape = (grid_size, grid_size)
    r_px = grid_size * radius
    cen = grid_size / 2
    mask = _disk((cen, cen), r_px, shape)
    yy, xx = np.mgrid[0:grid_size, 0:grid_size]
    rho2 = ((xx - cen) ** 2 + (yy - cen) ** 2) / r_px ** 2
    mu = np.sqrt(np.clip(1 - rho2, 0.0, 1.0))
    inten = mask * _limb_darkening(mu, limb_darkening_law, ld_coefficient)
    g = _Grid()
    g.intensity = inten / np.sum(inten)
    g.radius_px = r_px
    return g


def planet_transit(star_grid, planet_to_star_ratio, impact_parameter=0.0,
                   phase=0.0, rescaling_factor=None, resample_method=None):
    inten = star_grid.intensity
    n = inten.shape[0]
    r_star = star_grid.radius_px
    r_pl = r_star * planet_to_star_ratio
    x_p = n / 2 + 2 * phase * r_star
    y_p = n / 2 + impact_parameter * r_star
    planet = _disk((x_p, y_p), r_pl, inten.shape)
    transit = inten * (1 - planet)
    depth = 1 - np.sum(transit) / np.sum(inten)
    g = _Grid()
    if rescaling_factor is not None:
        if resample_method not in (None, 'box'):
            raise NotImplementedError('stand-in implements box resampling only')
        m = int(round(n * rescaling_factor))
        ss = n // m
        transit = transit[:m * ss, :m * ss].reshape(m, ss, m, ss).sum(axis=(1, 3))
        x_p, y_p, r_pl = x_p * rescaling_factor, y_p * rescaling_factor, \
            r_pl * rescaling_factor
    g.intensity = transit
    g.transit_depth = depth
    g.planet_px_coordinates = (x_p, y_p)
    g.planet_radius_px = r_pl
    return g
