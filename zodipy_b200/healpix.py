"""HEALPix pixel centres (RING scheme) as unit vectors, NumPy only.

Used to generate the synthetic full-sky pointings of the BASELINE map configurations without
healpy / astropy-healpix (neither is in the image).  Standard HEALPix geometry (Gorski et al.
2005): 12 nside^2 equal-area pixels on iso-latitude rings; north polar cap, equatorial belt, south
polar cap.  The reference's own map examples obtain the same centres from healpy
(``docs/examples/healpy_map.py:14-18``).
"""
from __future__ import annotations

import numpy as np


def nside2npix(nside: int) -> int:
    return 12 * nside * nside


def _isqrt(v):
    r = np.floor(np.sqrt(v.astype(np.float64))).astype(np.int64)
    r -= (r * r > v)
    r += ((r + 1) * (r + 1) <= v)
    return r


def pix2vec_ring(nside: int, ipix) -> np.ndarray:
    """Unit vectors (3, n) of RING-ordered pixel centres."""
    ipix = np.asarray(ipix, dtype=np.int64)
    npix = nside2npix(nside)
    ncap = 2 * nside * (nside - 1)
    fact2 = 4.0 / npix
    fact1 = 2 * nside * fact2
    z = np.empty(ipix.shape, dtype=np.float64)
    sth = np.empty_like(z)
    phi = np.empty_like(z)

    north = ipix < ncap
    south = ipix >= npix - ncap
    belt = ~(north | south)

    p = ipix[north]
    iring = (1 + _isqrt(1 + 2 * p)) >> 1
    iphi = p + 1 - 2 * iring * (iring - 1)
    tmp = iring.astype(np.float64) ** 2 * fact2
    z[north] = 1.0 - tmp
    sth[north] = np.sqrt(tmp * (2.0 - tmp))
    phi[north] = (iphi - 0.5) * (np.pi / 2) / iring

    p = ipix[belt] - ncap
    iring = p // (4 * nside) + nside
    iphi = p % (4 * nside) + 1
    fodd = np.where(((iring + nside) & 1) == 1, 1.0, 0.5)
    zb = (2 * nside - iring) * fact1
    z[belt] = zb
    sth[belt] = np.sqrt((1.0 - zb) * (1.0 + zb))
    phi[belt] = (iphi - fodd) * (np.pi / 2) / nside

    p = npix - ipix[south]
    iring = (1 + _isqrt(2 * p - 1)) >> 1
    iphi = 4 * iring + 1 - (p - 2 * iring * (iring - 1))
    tmp = iring.astype(np.float64) ** 2 * fact2
    z[south] = tmp - 1.0
    sth[south] = np.sqrt(tmp * (2.0 - tmp))
    phi[south] = (iphi - 0.5) * (np.pi / 2) / iring

    return np.stack([sth * np.cos(phi), sth * np.sin(phi), z])


def full_sky_vectors(nside: int, start: int = 0, stop: int | None = None, chunk: int = 1 << 22,
                     out: np.ndarray | None = None) -> np.ndarray:
    """(3, stop-start) pixel-centre unit vectors of pixels [start, stop), built in chunks."""
    stop = nside2npix(nside) if stop is None else stop
    n = stop - start
    if out is None:
        out = np.empty((3, n), dtype=np.float64)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        out[:, lo:hi] = pix2vec_ring(nside, np.arange(start + lo, start + hi, dtype=np.int64))
    return out


# ---- NESTED scheme --------------------------------------------------------------------------
_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)


def _compact_bits(v):
    """De-interleave: keep the even-position bits of v (x index of a Morton code)."""
    v = v & 0x5555555555555555
    v = (v | (v >> 1)) & 0x3333333333333333
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFF
    v = (v | (v >> 16)) & 0x00000000FFFFFFFF
    return v


def nest2ring(nside: int, ipix) -> np.ndarray:
    """NESTED -> RING pixel index (nside must be a power of two)."""
    if nside & (nside - 1):
        raise ValueError("NESTED ordering needs nside to be a power of two")
    ipix = np.asarray(ipix, dtype=np.int64)
    npface = nside * nside
    face = ipix // npface
    ipf = ipix & (npface - 1)
    ix, iy = _compact_bits(ipf), _compact_bits(ipf >> 1)
    jr = _JRLL[face] * nside - ix - iy - 1  # ring number in 1 .. 4 nside - 1
    nr = np.where(jr < nside, jr, np.where(jr > 3 * nside, 4 * nside - jr, nside))
    n_before = np.where(jr < nside, 2 * nr * (nr - 1),
                        np.where(jr > 3 * nside, 12 * npface - 2 * (nr + 1) * nr,
                                 2 * nside * (nside - 1) + (jr - nside) * 4 * nside))
    kshift = np.where((jr < nside) | (jr > 3 * nside), 0, (jr - nside) & 1)
    jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
    jp = np.where(jp > 4 * nr, jp - 4 * nr, jp)
    jp = np.where(jp < 1, jp + 4 * nr, jp)
    return n_before + jp - 1


def pix2vec_nest(nside: int, ipix) -> np.ndarray:
    """Unit vectors (3, n) of NESTED-ordered pixel centres."""
    return pix2vec_ring(nside, nest2ring(nside, ipix))
