"""Import the UNMODIFIED reference with the two stand-in packages on ``sys.path``.
TEST / BENCH INFRASTRUCTURE ONLY: used by ``oracle/make_golden.py``,
``oracle/make_anchors.py``, by CPU tests that are skipped when no reference tree is
present, and by the reference arm / ``cpu_baseline`` leg of ``bench.py``.

Where the reference is looked for, in this order: ``$PWINDS_REFERENCE_ROOT``,
``/root/reference`` (the build container), ``oracle/_ref`` (the byte-for-byte copy that
``oracle/build_ref.py`` makes so that the reference travels to the GPU box)."""
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_STANDINS = os.path.join(_HERE, 'standins')


def _find():
    for p in (os.environ.get('PWINDS_REFERENCE_ROOT'), '/root/reference', os.path.join(_HERE, '_ref')):
        if p and os.path.isfile(os.path.join(p, 'p_winds', 'helium.py')):
            return p
    return os.environ.get('PWINDS_REFERENCE_ROOT', '/root/reference')


REFERENCE_ROOT = _find()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'p_winds', 'helium.py'))


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
