"""Make the UNMODIFIED reference travel to the GPU box.

TEST / BENCH INFRASTRUCTURE.  Copies ``/root/reference/p_winds`` (the ten module files,
byte for byte) and the two data files the hot path reads into ``oracle/_ref/`` -- a
directory that is git-ignored (reference sources never enter the history) but not
gpurun-ignored, so that ``bench.py --impl reference`` and the ``cpu_baseline`` leg can time
the reference's own code on the GPU box's host cores, where ``/root/reference`` does not
exist.  The reference is pure Python: there is nothing to compile.  Its two absent
dependencies are served by ``oracle/standins`` (astropy: units and constants only;
flatstar: the disk rasteriser restated from Pillow's ellipse fill).

    python oracle/build_ref.py          # no-op when /root/reference is absent

Writes ``oracle/_ref/MANIFEST.json`` with the sha256 of every file copied.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('PWINDS_REFERENCE_SRC', '/root/reference')
DST = os.path.join(HERE, '_ref')
DATA = ('solar_spectrum_scaled_lambda.dat', 'he_3_profile.dat')


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC, 'p_winds')):
        if verbose:
            print('build_ref: %s not present, nothing to do' % SRC)
        return False
    man = {}
    for sub, names in (('p_winds', sorted(f for f in os.listdir(os.path.join(SRC, 'p_winds')) if f.endswith('.py'))),
                       ('data', DATA)):
        os.makedirs(os.path.join(DST, sub), exist_ok=True)
        for f in names:
            a, b = os.path.join(SRC, sub, f), os.path.join(DST, sub, f)
            blob = open(a, 'rb').read()
            man['%s/%s' % (sub, f)] = hashlib.sha256(blob).hexdigest()
            if not (os.path.exists(b) and open(b, 'rb').read() == blob):
                if os.path.exists(b):
                    os.chmod(b, 0o644)
                shutil.copyfile(a, b)
    ver = ''
    try:
        ver = open(os.path.join(SRC, 'setup.py')).read().split('version=')[1].split(',')[0].strip('"\' ')
    except Exception:
        pass
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': SRC, 'version': ver, 'sha256': man}, fh, indent=1, sort_keys=True)
    if verbose:
        print('build_ref: %d files -> %s' % (len(man), DST))
    return True


if __name__ == '__main__':
    build()
