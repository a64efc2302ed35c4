"""Import the UNMODIFIED reference (``/root/reference/p_winds``) with the two
stand-in packages on ``sys.path``.  TEST INFRASTRUCTURE ONLY: used by
``oracle/make_golden.py`` and by CPU tests that are skipped when the reference
tree is absent (it does not exist on the GPU box)."""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get('PWINDS_REFERENCE_ROOT', '/root/reference')
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'standins')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'p_winds'))


def load():
    """Return the reference ``p_winds`` package (modules parker, hydrogen, helium,
    transit, microphysics, lines, tools are imported eagerly)."""
    if not available():
        raise ImportError('reference tree not found at %s' % REFERENCE_ROOT)
    for p in (_STANDINS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import p_winds
        from p_winds import (parker, hydrogen, helium, transit,  # noqa: F401
                             microphysics, lines, tools)
    return p_winds
