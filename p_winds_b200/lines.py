"""Spectral-line constants (reference: p_winds/lines.py).  NIST ASD values."""

__all__ = ['he_3_properties', 'c_ii_properties', 'o_i_properties', 'lyman_alpha_properties']


def he_3_properties():
    """He I 2^3S-2^3P triplet near 1.083 micron (lines.py:14-55).

    Returns lambda_0, lambda_1, lambda_2 [m], f_0, f_1, f_2, a_ij [1/s]."""
    return (1.082909114 * 1E-6, 1.083025010 * 1E-6, 1.083033977 * 1E-6,
            5.9902e-02, 1.7974e-01, 2.9958e-01, 1.0216e+07)


def c_ii_properties():
    """C II 1334/1335 A lines (lines.py:60-110).

    Returns lambda_0, lambda_1, lambda_2 [m], f_0..f_2, a_0..a_2 [1/s]."""
    return (1334.5323 * 1E-10, 1335.6628 * 1E-10, 1335.7079 * 1E-10,
            1.29e-01, 1.27e-02, 1.15e-01, 2.41e+08, 4.76e+07, 2.88e+08)


def o_i_properties():
    """O I 1302/1304/1306 A lines (lines.py:114-163).

    Returns lambda_0, lambda_1, lambda_2 [m], f_0..f_2, a_0..a_2 [1/s]."""
    return (1302.168 * 1E-10, 1304.858 * 1E-10, 1306.029 * 1E-10,
            5.20e-02, 5.18e-02, 5.19e-02, 3.41e+08, 2.03e+08, 6.76e+07)


def lyman_alpha_properties():
    """H I Ly-alpha.  Not in the reference's line list (the caller
    supplies it to radiative_transfer_2d); values of BASELINE config 2."""
    return 1215.67e-10, 0.4164, 6.265e8
