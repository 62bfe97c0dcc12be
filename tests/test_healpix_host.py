"""Host HEALPix helpers (NumPy): known values and structural properties."""
import numpy as np
import pytest

from zodipy_b200 import healpix as hp


def test_ring_known_centres():
    # (theta, phi) of a few nside=2 RING pixels, as tabulated by healpy.pix2ang(2, ...)
    v = hp.pix2vec_ring(2, np.array([0, 4, 12, 20, 44, 47]))
    theta = np.degrees(np.arccos(v[2]))
    phi = np.degrees(np.arctan2(v[1], v[0])) % 360
    np.testing.assert_allclose(theta, [23.55646431, 48.1896851, 70.52877937, 90.0, 156.44353569, 156.44353569],
                               atol=1e-7)
    np.testing.assert_allclose(phi, [45.0, 22.5, 0.0, 22.5, 45.0, 315.0], atol=1e-9)
    v1 = hp.pix2vec_ring(1, np.arange(12))
    np.testing.assert_allclose(v1[:, 4], [1.0, 0.0, 0.0], atol=1e-15)
    np.testing.assert_allclose(v1[2, :4], 2.0 / 3.0)


@pytest.mark.parametrize("nside", [1, 2, 8, 64, 512])
def test_ring_unit_norm_and_symmetry(nside):
    npix = hp.nside2npix(nside)
    idx = np.arange(npix) if npix <= 49152 else np.random.default_rng(0).choice(npix, 40000, replace=False)
    v = hp.pix2vec_ring(nside, idx)
    np.testing.assert_allclose(np.linalg.norm(v, axis=0), 1.0, atol=3e-16)
    # north/south mirror symmetry of the pixelisation: pixel npix-1-i is the antipode in z of ring mate
    vs = hp.pix2vec_ring(nside, npix - 1 - idx)
    np.testing.assert_allclose(vs[2], -v[2], atol=3e-16)


def test_nest2ring_known_and_bijective():
    # healpy documentation example: nest2ring(2, arange(10))
    assert hp.nest2ring(2, np.arange(10)).tolist() == [13, 5, 4, 0, 15, 7, 6, 1, 17, 9]
    assert hp.nest2ring(1, np.arange(12)).tolist() == list(range(12))
    for nside in (2, 4, 32, 128):
        n = hp.nside2npix(nside)
        assert np.array_equal(np.sort(hp.nest2ring(nside, np.arange(n))), np.arange(n))
    with pytest.raises(ValueError):
        hp.nest2ring(3, np.arange(4))


@pytest.mark.parametrize("nside", [2, 16, 64])
def test_nested_hierarchy(nside):
    """Children 4p..4p+3 of NESTED pixel p cluster around the parent's centre."""
    n = hp.nside2npix(nside)
    child = hp.pix2vec_nest(nside, np.arange(n)).reshape(3, n // 4, 4)
    parent = hp.pix2vec_nest(nside // 2, np.arange(n // 4))
    mean = child.mean(axis=2)
    mean /= np.linalg.norm(mean, axis=0)
    ang = np.arccos(np.clip((mean * parent).sum(axis=0), -1, 1))
    pix_size = np.sqrt(4 * np.pi / (n // 4))
    assert ang.max() < 0.1 * pix_size  # polar-cap pixels are the most distorted (~7 %)
    # every child is inside a disc of one parent pixel size around the parent centre
    sep = np.arccos(np.clip((child * parent[:, :, None]).sum(axis=0), -1, 1))
    assert sep.max() < 1.0 * pix_size
