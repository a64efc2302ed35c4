"""Minimal stand-in for the parts of ``astropy`` the p-winds hot path touches.

TEST INFRASTRUCTURE ONLY.  astropy is not installed in this image (no network).
This package exists so that the *unmodified* reference under /root/reference can
be imported to generate golden vectors (``oracle/make_golden.py``).  It is never
imported by the product package ``p_winds_b200``.
"""
