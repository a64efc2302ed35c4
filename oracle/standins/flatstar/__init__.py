"""Stand-in for the third-party ``flatstar`` package (test infrastructure only).

``flatstar`` (PyPI, same author as p-winds, unpinned in the reference's
requirements.txt:4) is not installed in this image and is not vendored under
/root/reference.  Its rasteriser was recovered from the reference's golden
numbers: a star/planet disk drawn with Pillow's ``ImageDraw.ellipse`` on the
supersampled grid reproduces ``tests/test_transit.py:37`` (transit depth
0.015012130070238605) to 5e-15.  Everything not pinned by a golden number
(limb-darkening pixel convention, sign of the impact parameter) is documented as
"parity unpinned" in DESIGN.md.
"""
