"""Import-only stand-in for ``astropy.io.fits`` (the hot path never opens FITS)."""


def open(*a, **k):  # noqa: A001
    raise NotImplementedError('astropy.io.fits is not available in this image')


def getdata(*a, **k):
    raise NotImplementedError('astropy.io.fits is not available in this image')
