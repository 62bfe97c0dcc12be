"""Parity of the CUDA path (through the C ABI) with the reference / oracle.  Needs a GPU."""
import numpy as np
import pytest

import zodi_oracle as oracle
import zodipy_b200 as zp
from helpers import (COMP_FLOOR_FP32, COMP_FLOOR_FP64, EARTH_20220114, TOL_FP32, TOL_FP64, case_ids, comp_floor,
                     fibonacci_sphere, golden_case, max_rel_comps, max_rel_total)
from zodipy_b200 import engine

pytestmark = pytest.mark.gpu

TOL = {"fp64": (TOL_FP64, COMP_FLOOR_FP64), "fp32": (TOL_FP32, COMP_FLOOR_FP32)}


def device_model(spec, generic=False, no_x2=False):
    """generic=True forces the generic kernel for models the fused Kelsall kernel would take;
    no_x2=True the scalar fused kernel where the packed (2 lines of sight per thread) one would run."""
    import os

    knobs = {"ZODI_FORCE_GENERIC": "1" if generic else "0", "ZODI_NO_X2": "1" if no_x2 else "0"}
    old = {k: os.environ.get(k) for k in knobs}
    os.environ.update(knobs)
    try:
        return engine.DeviceModel(spec, device=0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_kernel_selection():
    """Kelsall-family and RRM layouts take their fused kernels; user-edited layouts the generic one."""
    expect = {"planck18_857": "kelsall", "dirbe_25um_rand": "kelsall", "dirbe_1p25um": "kelsall",
              "planck13_545": "kelsall", "rrm_60um": "rrm", "dirbe_25um_mutated": "generic"}
    for case_id, which in expect.items():
        assert device_model(golden_case(case_id)[0]["spec"]).kernel_name == f"zodi_los_{which}_kernel"
    assert device_model(golden_case("planck18_857")[0]["spec"], generic=True).kernel_name == \
        "zodi_los_generic_kernel"


@pytest.mark.parametrize("generic", [False, True], ids=["auto", "generic"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("case_id", case_ids())
def test_golden_host_memory(case_id, precision, generic):
    """Committed reference outputs, C-ABI call with HOST buffers (H2D + kernel + D2H); every case
    through the kernel the library picks and through the generic kernel."""
    case, a = golden_case(case_id)
    dm = device_model(case["spec"], generic=generic)
    if generic and device_model(case["spec"]).kernel_name == "zodi_los_generic_kernel":
        pytest.skip("already covered by the auto run")
    launches = engine.kernel_launch_count()
    em = dm.evaluate(a["u"], a["obs"], a["earth"], return_comps=True, precision=precision)
    assert engine.kernel_launch_count() > launches  # the CUDA kernels ran
    tol, floor = TOL[precision][0], comp_floor(precision, case["spec"]["kind"])
    assert em.shape == a["emission"].shape and em.dtype == np.float64
    assert max_rel_total(em, a["emission"]) <= tol
    assert max_rel_comps(em, a["emission"], floor=floor) <= tol
    total = dm.evaluate(a["u"], a["obs"], a["earth"], return_comps=False, precision=precision)
    assert total.shape == (a["u"].shape[1],)
    # summed output = component sum in model order (comps.sum(axis=0), tests/test_evaluate.py:198-212)
    np.testing.assert_allclose(total, em.sum(axis=0), rtol=1e-6 if precision == "fp32" else 1e-15)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("case_id", ["dirbe_25um_rand", "dirbe_1p25um_tod", "rrm_60um", "planck18_tod",
                                     "dirbe_25um_tod_straddle"])
def test_golden_device_memory(case_id, precision):
    """Same through device-resident torch tensors (async launch on the current stream)."""
    import torch

    case, a = golden_case(case_id)
    dm = device_model(case["spec"])
    dev = torch.device("cuda:0")
    u, obs, earth = (torch.as_tensor(a[k], device=dev) for k in ("u", "obs", "earth"))
    em = dm.evaluate(u, obs, earth, return_comps=True, precision=precision)
    assert em.is_cuda and tuple(em.shape) == a["emission"].shape
    em = em.cpu().numpy()
    tol, floor = TOL[precision]
    assert max_rel_total(em, a["emission"]) <= tol
    assert max_rel_comps(em, a["emission"], floor=floor) <= tol
    out32 = dm.evaluate(u, obs, earth, return_comps=False, precision=precision, out_dtype=np.float32)
    assert out32.dtype == torch.float32
    np.testing.assert_allclose(out32.cpu().numpy(), a["emission"].sum(axis=0), rtol=2e-6 + tol)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name,x,unit,n", [("dirbe", 25.0, "um", 49152), ("planck18", 857.0, "GHz", 20000),
                                           ("planck13", 545.0, "GHz", 20000),
                                           ("rrm-experimental", 25.0, "um", 6000)])
def test_against_oracle_on_fresh_inputs(name, x, unit, n, precision):
    """CUDA vs oracle on seeded inputs that are not in the fixtures (BASELINE config 1 is the
    full nside=64-sized dirbe 25 um case)."""
    model = zp.Model(zp.Quantity(x, unit), name=name, precision=precision)
    u = fibonacci_sphere(n)
    em = model.evaluate_xyz(u, EARTH_20220114, return_comps=True)
    sel = np.random.default_rng(0).choice(n, size=3000, replace=False)
    ref = oracle.evaluate(model.spec, u[:, sel], EARTH_20220114, EARTH_20220114)
    tol, floor = TOL[precision][0], comp_floor(precision, model.spec["kind"])
    assert max_rel_total(em[:, sel], ref) <= tol
    assert max_rel_comps(em[:, sel], ref, floor=floor) <= tol


def test_edge_shapes_and_chunking():
    """Empty, single and ragged sizes; sizes that straddle the lane-count and chunk thresholds."""
    model = zp.Model(zp.Quantity(25.0, "um"))
    dm = model.device_model
    assert dm.evaluate(np.zeros((3, 0)), EARTH_20220114).shape == (0,)
    assert dm.evaluate(np.zeros((3, 0)), EARTH_20220114, return_comps=True).shape == (6, 0)
    u_all = fibonacci_sphere(400000)
    full = dm.evaluate(u_all, EARTH_20220114, return_comps=True)
    for n in (1, 2, 31, 33, 255, 257, 9472, 9473, 303105):
        part = dm.evaluate(u_all[:, :n], EARTH_20220114, return_comps=True)
        # different lane splits change the summation order only
        np.testing.assert_allclose(part, full[:, :n], rtol=1e-13)
    # strided view (a shard of a larger (3, N) array) without copying
    shard = u_all[:, 1000:5000]
    assert not shard.flags.c_contiguous
    np.testing.assert_allclose(dm.evaluate(shard, EARTH_20220114, return_comps=True), full[:, 1000:5000],
                               rtol=1e-13)
    # host path chunking (> 1 Mi lines of sight -> several pipeline chunks, ragged tail)
    n_big = (1 << 21) + 12345
    u_big = fibonacci_sphere(n_big)
    tot = dm.evaluate(u_big, EARTH_20220114, precision="fp32", out_dtype=np.float32)
    idx = np.array([0, 1, (1 << 20) - 1, 1 << 20, (1 << 21) - 1, 1 << 21, n_big - 1])
    ref = oracle.evaluate(model.spec, u_big[:, idx], EARTH_20220114, EARTH_20220114).sum(axis=0)
    np.testing.assert_allclose(tot[idx], ref, rtol=TOL_FP32)
    assert np.all(np.isfinite(tot))


def test_sharding_invariance_bitwise():
    """Any contiguous split evaluated separately equals the one-shot result bit for bit when the
    lane count is the same (reference: nprocesses equality, tests/test_evaluate.py:215-262)."""
    case, a = golden_case("dirbe_25um_tod_straddle")
    dm = device_model(case["spec"])
    n = a["u"].shape[1]
    flags = dm.outside_flags(a["obs"])  # GLOBAL flags (quirk Q1)
    one = dm.evaluate(a["u"], a["obs"], a["earth"], return_comps=True, outside_flags=flags)
    parts = [dm.evaluate(a["u"][:, s], a["obs"][:, s], a["earth"][:, s], return_comps=True,
                         outside_flags=flags)
             for s in (slice(0, 100), slice(100, 217), slice(217, n))]
    np.testing.assert_array_equal(np.concatenate(parts, axis=1), one)
    # without global flags a shard whose observers stay inside a cutoff differs (documented Q1)
    local = dm.evaluate(a["u"][:, :100], a["obs"][:, :100], a["earth"][:, :100], return_comps=True)
    assert local.shape == (6, 100)


def test_update_parameters_reuploads():
    model = zp.Model(zp.Quantity(25.0, "um"))
    u = fibonacci_sphere(512)
    before = model.evaluate_xyz(u, EARTH_20220114, return_comps=True)
    p = model.get_parameters()
    p["comps"]["cloud"]["n_0"] *= 2.0
    p["comps"]["band2"]["p"] = 3.5
    model.update_parameters(p)
    after = model.evaluate_xyz(u, EARTH_20220114, return_comps=True)
    np.testing.assert_allclose(after[0], 2.0 * before[0], rtol=1e-13)
    ref = oracle.evaluate(model.spec, u, EARTH_20220114, EARTH_20220114)
    assert max_rel_comps(after, ref, floor=COMP_FLOOR_FP64) <= TOL_FP64


def test_max_observer_radius_device_and_host():
    import torch

    case, a = golden_case("dirbe_25um_tod_straddle")
    dm = device_model(case["spec"])
    r_host = dm.max_observer_radius(a["obs"])
    r_dev = dm.max_observer_radius(torch.as_tensor(a["obs"], device="cuda:0"))
    expect = np.sqrt((a["obs"] ** 2).sum(axis=0)).max()
    assert r_host == pytest.approx(expect, rel=1e-15) and r_dev == pytest.approx(expect, rel=1e-15)
    np.testing.assert_array_equal(dm.outside_flags(a["obs"]), oracle.outside_flags(case["spec"], a["obs"]))


def test_linearity_in_emissivity_full_size():
    """Size-independent property at a BASELINE-sized input (nside 512 = 3.1 M lines of sight):
    scaling every emissivity by k scales the map by k; doubling a component density doubles it."""
    import torch

    n = 12 * 512 * 512
    dev = torch.device("cuda:0")
    u = torch.as_tensor(fibonacci_sphere(n), device=dev)
    obs = torch.as_tensor(EARTH_20220114, device=dev)
    model = zp.Model(zp.Quantity(857.0, "GHz"), name="planck18", precision="fp32")
    base = model.evaluate_xyz(u, obs, return_comps=True, out_dtype=np.float32)
    p = model.get_parameters()
    for label in p["emissivities"]:
        p["emissivities"][label] = tuple(3.0 * v for v in p["emissivities"][label])
    model.update_parameters(p)
    scaled = model.evaluate_xyz(u, obs, return_comps=True, out_dtype=np.float32)
    assert torch.allclose(scaled, 3.0 * base, rtol=2e-6, atol=0.0)
    assert bool(torch.isfinite(base).all())


def test_healpix_vectors_match_host_pix2vec():
    from zodipy_b200 import healpix

    for nside in (1, 2, 8, 64, 1024):
        npix = healpix.nside2npix(nside)
        rng = (0, npix) if npix <= 49152 else (npix // 2 - 3000, npix // 2 + 3000)
        got = engine.healpix_vectors(nside, rng)
        ref = healpix.pix2vec_ring(nside, np.arange(*rng))
        np.testing.assert_allclose(got, ref, rtol=0, atol=3e-16)
    # caps of a large map (first / last pixels) and a rotation
    nside = 2048
    npix = healpix.nside2npix(nside)
    for rng in ((0, 5000), (npix - 5000, npix)):
        np.testing.assert_allclose(engine.healpix_vectors(nside, rng),
                                   healpix.pix2vec_ring(nside, np.arange(*rng)), rtol=0, atol=3e-16)
    for nside in (1, 2, 64, 2048):  # NESTED ordering
        npix = healpix.nside2npix(nside)
        rng = (0, npix) if npix <= 49152 else (npix // 3, npix // 3 + 5000)
        np.testing.assert_allclose(engine.healpix_vectors(nside, rng, nest=True),
                                   healpix.pix2vec_nest(nside, np.arange(*rng)), rtol=0, atol=3e-16)
    with pytest.raises(engine._cabi.ZodiError):
        engine.healpix_vectors(12, nest=True)  # NESTED needs a power of two
    c, s_ = np.cos(0.4), np.sin(0.4)
    rot = np.array([[1, 0, 0], [0, c, s_], [0, -s_, c]])
    np.testing.assert_allclose(engine.healpix_vectors(16, rot=rot),
                               rot @ healpix.pix2vec_ring(16, np.arange(12 * 256)), rtol=0, atol=4e-16)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_evaluate_healpix_equals_array_seam(precision):
    """On-device directions give the same map as uploading pix2vec() through the array seam, for
    host and device outputs, sub-ranges, rotation and return_comps."""
    import torch

    from zodipy_b200 import healpix

    model = zp.Model(zp.Quantity(25.0, "um"), precision=precision)
    nside = 64
    u = healpix.full_sky_vectors(nside)
    ref = model.evaluate_xyz(u, EARTH_20220114, return_comps=True)
    tol = 1e-12 if precision == "fp64" else 2e-6
    got = model.evaluate_healpix(nside, EARTH_20220114, return_comps=True)
    np.testing.assert_allclose(got, ref, rtol=tol)
    np.testing.assert_allclose(model.evaluate_healpix(nside, EARTH_20220114), ref.sum(axis=0), rtol=tol)
    part = model.evaluate_healpix(nside, EARTH_20220114, pix_range=(1000, 20001), out_dtype=np.float32)
    assert part.dtype == np.float32 and part.shape == (19001,)
    np.testing.assert_allclose(part, ref.sum(axis=0)[1000:20001], rtol=tol + 1e-6)
    dev = model.evaluate_healpix(nside, EARTH_20220114, device_out=True)
    assert dev.is_cuda and torch.allclose(dev.cpu(), torch.from_numpy(ref.sum(axis=0)), rtol=tol, atol=0)
    c, s_ = np.cos(1.1), np.sin(1.1)
    rot = np.array([[c, s_, 0], [-s_, c, 0], [0, 0, 1.0]])
    np.testing.assert_allclose(model.evaluate_healpix(nside, EARTH_20220114, frame_rotation=rot),
                               model.evaluate_xyz(rot @ u, EARTH_20220114), rtol=tol)
    # oracle check on a subset
    sel = np.arange(0, u.shape[1], 97)
    ref_o = oracle.evaluate(model.spec, u[:, sel], EARTH_20220114, EARTH_20220114)
    assert max_rel_total(got[:, sel], ref_o) <= TOL[precision][0]
    nested = model.evaluate_healpix(nside, EARTH_20220114, nest=True)
    np.testing.assert_allclose(nested, ref.sum(axis=0)[healpix.nest2ring(nside, np.arange(u.shape[1]))], rtol=tol)
    with pytest.raises(ValueError):
        model.evaluate_healpix(4, EARTH_20220114, pix_range=(0, 12 * 16 + 1))
    assert model.evaluate_healpix(4, EARTH_20220114, pix_range=(7, 7)).shape == (0,)


def _sph2cart(lon, lat):
    return np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])


def _random_rotation(seed):
    q, r = np.linalg.qr(np.random.default_rng(seed).normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    return q * np.sign(np.linalg.det(q))


def test_lonlat_vectors_match_host_trigonometry():
    rng = np.random.default_rng(5)
    lon = np.concatenate([rng.uniform(-2 * np.pi, 4 * np.pi, 20000), [0.0, np.pi, 2 * np.pi, -np.pi, 1e-300, 7.0, 0.3]])
    lat = np.concatenate([rng.uniform(-np.pi / 2, np.pi / 2, 20000), [np.pi / 2, -np.pi / 2, 0.0, 1e-17, 0.5, -0.5, 0.0]])
    np.testing.assert_allclose(engine.lonlat_vectors(lon, lat), _sph2cart(lon, lat), rtol=0, atol=3e-16)
    rot = _random_rotation(1)
    np.testing.assert_allclose(engine.lonlat_vectors(lon, lat, rot=rot), rot @ _sph2cart(lon, lat), rtol=0, atol=5e-16)
    assert engine.lonlat_vectors(np.empty(0), np.empty(0)).shape == (3, 0)
    with pytest.raises(ValueError):
        engine.lonlat_vectors(lon, lat[:-1])


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name,x,unit", [("dirbe", 25.0, "um"), ("planck18", 857.0, "GHz"),
                                         ("rrm-experimental", 60.0, "um")])
def test_evaluate_lonlat_equals_array_seam(name, x, unit, precision):
    """Spherical-coordinate entry == array seam fed with the host-built unit vectors: host and device
    memory, with / without rotation, return_comps, per-sample observers, small and large N
    (scalar, lanes-per-line-of-sight and packed kernels), and the oracle on a subset."""
    import torch

    model = zp.Model(zp.Quantity(x, unit), name=name, precision=precision)
    tol = 1e-12 if precision == "fp64" else 2e-6
    rng = np.random.default_rng(11)
    rot = _random_rotation(3)
    for n in (1, 7, 5000, 700_001):
        lon, lat = rng.uniform(0, 2 * np.pi, n), np.arcsin(rng.uniform(-1, 1, n))
        u = rot @ _sph2cart(lon, lat)
        ref = model.evaluate_xyz(u, EARTH_20220114, return_comps=True)
        got = model.evaluate_lonlat(lon, lat, EARTH_20220114, frame_rotation=rot, return_comps=True)
        assert got.shape == ref.shape
        np.testing.assert_allclose(got.sum(axis=0), ref.sum(axis=0), rtol=tol)
        np.testing.assert_allclose(model.evaluate_lonlat(lon, lat, EARTH_20220114, frame_rotation=rot),
                                   ref.sum(axis=0), rtol=tol)
        if n == 5000:
            ref_o = oracle.evaluate(model.spec, u, EARTH_20220114, EARTH_20220114)
            assert max_rel_total(got, ref_o) <= TOL[precision][0]
    # no rotation (ecliptic angles), float32 output, device memory
    lon, lat = rng.uniform(-np.pi, np.pi, 3000), rng.uniform(-1.5, 1.5, 3000)
    ref = model.evaluate_xyz(_sph2cart(lon, lat), EARTH_20220114)
    out32 = model.evaluate_lonlat(lon, lat, EARTH_20220114, out_dtype=np.float32)
    assert out32.dtype == np.float32
    np.testing.assert_allclose(out32, ref, rtol=tol + 1e-6)
    dev = model.evaluate_lonlat(torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda(), EARTH_20220114)
    assert dev.is_cuda
    np.testing.assert_allclose(dev.cpu().numpy(), ref, rtol=tol)
    # per-sample observer / Earth positions (time-ordered data through the array seam)
    ang = rng.uniform(0, 2 * np.pi, 3000)
    earth = np.array([np.cos(ang), np.sin(ang), np.zeros_like(ang)]) * rng.uniform(0.98, 1.02, 3000)
    obs = earth * 1.01
    np.testing.assert_allclose(model.evaluate_lonlat(lon, lat, obs, earth, frame_rotation=rot),
                               model.evaluate_xyz(rot @ _sph2cart(lon, lat), obs, earth), rtol=tol)
    assert model.evaluate_lonlat(np.empty(0), np.empty(0), EARTH_20220114).shape == (0,)
    with pytest.raises(ValueError):
        model.evaluate_lonlat(lon, lat[:-1], EARTH_20220114)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_multi_device_host_sharding_bitwise(precision, monkeypatch):
    """Model(devices=[...]) splits host arrays over several device handles driven by threads of one
    process (here: all on GPU 0, or spread over the GPUs present); bit-identical to one handle,
    including time-ordered data whose observers straddle a cutoff sphere (global early-out flags)."""
    import torch

    monkeypatch.setattr(engine.MultiDeviceModel, "MIN_LOS_PER_DEVICE", 1)  # split the small golden case too
    devices = [i % torch.cuda.device_count() for i in range(3)]
    one = zp.Model(zp.Quantity(25.0, "um"), precision=precision, device=0)
    many = zp.Model(zp.Quantity(25.0, "um"), precision=precision, devices=devices)
    case, a = golden_case("dirbe_25um_tod_straddle")
    ref = one.evaluate_xyz(a["u"], a["obs"], a["earth"], return_comps=True)
    got = many.evaluate_xyz(a["u"], a["obs"], a["earth"], return_comps=True)
    np.testing.assert_array_equal(got, ref)
    rng = np.random.default_rng(2)
    n = 1_500_001
    lon, lat = rng.uniform(0, 2 * np.pi, n), np.arcsin(rng.uniform(-1, 1, n))
    rot = _random_rotation(4)
    ref = one.evaluate_lonlat(lon, lat, EARTH_20220114, frame_rotation=rot, out_dtype=np.float32)
    got = many.evaluate_lonlat(lon, lat, EARTH_20220114, frame_rotation=rot, out_dtype=np.float32)
    np.testing.assert_array_equal(got, ref)
    u = rot @ _sph2cart(lon, lat)
    np.testing.assert_array_equal(many.evaluate_xyz(u, EARTH_20220114), one.evaluate_xyz(u, EARTH_20220114))
    # parameter updates reach every device
    params = many.get_parameters()
    params["comps"]["cloud"]["n_0"] *= 1.5
    many.update_parameters(params)
    one.update_parameters(params)
    np.testing.assert_array_equal(many.evaluate_xyz(u[:, :5000], EARTH_20220114, return_comps=True),
                                  one.evaluate_xyz(u[:, :5000], EARTH_20220114, return_comps=True))


def test_evaluate_lonlat_with_device_ephemeris_and_multiband():
    n = 20000
    rng = np.random.default_rng(4)
    knots = 59000.0 + np.arange(0, 24 * 12 + 1) / 24.0
    ang = 2 * np.pi * (knots - 59000.0) / 365.25
    earth_knots = np.array([np.cos(ang), np.sin(ang), 1e-3 * np.sin(3 * ang)])
    eph = engine.DeviceEphemeris(float(knots[0]), float(knots[1] - knots[0]), earth_knots, device=0)
    t = np.sort(rng.uniform(knots[0], knots[-1], n))
    lon, lat = rng.uniform(0, 2 * np.pi, n), np.arcsin(rng.uniform(-1, 1, n))
    rot = _random_rotation(9)
    u = rot @ _sph2cart(lon, lat)
    for precision, tol in (("fp64", 1e-12), ("fp32", 2e-6)):
        model = zp.Model(zp.Quantity(25.0, "um"), precision=precision)
        for observer in ("earth", "semb-l2"):
            ref = model.evaluate_tod_xyz(u, t, eph, observer=observer)
            got = model.evaluate_lonlat(lon, lat, frame_rotation=rot, ephemeris=eph, obstime=t, observer=observer)
            np.testing.assert_allclose(got, ref, rtol=tol)
    mb = zp.MultiBandModel([zp.Quantity(v, "um") for v in (12.0, 25.0, 60.0)], name="dirbe", precision="fp32")
    ref = mb.evaluate_xyz(u, EARTH_20220114)
    got = mb.device_model.evaluate_lonlat(lon, lat, EARTH_20220114, rot=rot, precision="fp32")
    assert got.shape == (3, n)
    np.testing.assert_allclose(got, ref, rtol=2e-6)


@pytest.mark.parametrize("name,x,unit", [("planck18", 857.0, "GHz"), ("dirbe", 25.0, "um"),
                                         ("planck13", 545.0, "GHz"), ("dirbe", 1.25, "um"),
                                         ("dirbe", 3.5, "um")])  # the last two: scattering
def test_packed_kernel_equals_scalar_fused_kernel(name, x, unit, monkeypatch):
    """The packed-fp32 kernel (FFMA2, two lines of sight per thread) performs the same operations
    as the scalar fused kernel: results must be bit-identical (incl. ragged tails, per-sample
    observers), and within tolerance of the oracle."""
    model = zp.Model(zp.Quantity(x, unit), name=name, precision="fp32")
    n = 148 * 2048 * 2 + 777  # large enough for the packed kernel, ragged tail
    u = fibonacci_sphere(n)
    rng = np.random.default_rng(5)
    obs = EARTH_20220114 * (1.0 + 0.02 * rng.standard_normal((1, n)))
    packed, scalar = device_model(model.spec), device_model(model.spec, no_x2=True)
    for kwargs in ({"return_comps": True}, {"return_comps": False}):
        a = packed.evaluate(u, EARTH_20220114, precision="fp32", **kwargs)
        b = scalar.evaluate(u, EARTH_20220114, precision="fp32", **kwargs)
        np.testing.assert_array_equal(a, b)
    a = packed.evaluate(u, obs, EARTH_20220114, precision="fp32", return_comps=True)
    b = scalar.evaluate(u, obs, EARTH_20220114, precision="fp32", return_comps=True)
    np.testing.assert_array_equal(a, b)
    sel = rng.choice(n, 2000, replace=False)
    ref = oracle.evaluate(model.spec, u[:, sel], obs[:, sel], EARTH_20220114)
    assert max_rel_total(a[:, sel], ref) <= TOL_FP32
    assert max_rel_comps(a[:, sel], ref, floor=COMP_FLOOR_FP32) <= TOL_FP32
    # mid-size input: both kernels split the nodes of a line of sight over 8 lanes (same shuffle tree)
    m = 30001
    monkeypatch.setenv("ZODI_X2_LANES", "8")
    a = packed.evaluate(u[:, :m], obs[:, :m], EARTH_20220114, precision="fp32", return_comps=True)
    b = scalar.evaluate(u[:, :m], obs[:, :m], EARTH_20220114, precision="fp32", return_comps=True)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("threads", [128, 256])
@pytest.mark.parametrize("lanes", [1, 2, 4, 8])
def test_packed_kernel_lane_splits(lanes, threads, monkeypatch):
    """Every shape of the packed kernel (lanes per pair of lines of sight x CTA size): same results up to
    the summation order of the lane partials, ragged sizes, per-sample observers, component output."""
    model = zp.Model(zp.Quantity(25.0, "um"), precision="fp32")
    dm = model.device_model
    n = 70001
    u = fibonacci_sphere(n)
    rng = np.random.default_rng(6)
    obs = EARTH_20220114 * (1.0 + 0.02 * rng.standard_normal((1, n)))
    monkeypatch.setenv("ZODI_X2_LANES", "1")
    monkeypatch.setenv("ZODI_X2_THREADS", "256")
    base = dm.evaluate(u, obs, EARTH_20220114, precision="fp32", return_comps=True)
    monkeypatch.setenv("ZODI_X2_LANES", str(lanes))
    monkeypatch.setenv("ZODI_X2_THREADS", str(threads))
    for m in (n, 1, 2, 3, 255, 513, 4097):
        got = dm.evaluate(u[:, :m], obs[:, :m], EARTH_20220114, precision="fp32", return_comps=True)
        np.testing.assert_allclose(got, base[:, :m], rtol=3e-6, atol=1e-12)
        tot = dm.evaluate(u[:, :m], obs[:, :m], EARTH_20220114, precision="fp32")
        np.testing.assert_allclose(tot, got.sum(axis=0), rtol=1e-6)
    if lanes == 1:
        np.testing.assert_array_equal(dm.evaluate(u, obs, EARTH_20220114, precision="fp32", return_comps=True), base)


def _tod_inputs(n, seed=0):
    """Hourly Earth knots over ~40 days (analytic orbit), random sample times and pointings."""
    rng = np.random.default_rng(seed)
    t0, dt, n_knots = 59215.0, 1.0 / 24.0, 40 * 24 + 1
    tk = t0 + dt * np.arange(n_knots)
    lon = 2 * np.pi * (tk - t0) / 365.25 + 1.7
    r = 1.0 - 0.0167 * np.cos(lon - 1.8)
    earth_knots = np.array([r * np.cos(lon), r * np.sin(lon), 1e-5 * np.sin(3 * lon)])
    t = np.sort(rng.uniform(t0, tk[-1], n))
    u = rng.normal(size=(3, n))
    u /= np.linalg.norm(u, axis=0)
    return t0, dt, earth_knots, tk, t, np.ascontiguousarray(u)


def test_device_ephemeris_matches_scipy_cubic_spline():
    from scipy.interpolate import CubicSpline

    t0, dt, earth_knots, tk, t, _ = _tod_inputs(5000)
    eph = engine.DeviceEphemeris(t0, dt, earth_knots)
    spline = CubicSpline(tk, earth_knots, axis=-1)  # bodies.py:34
    c_ref = spline.c  # (4, n-1, 3)
    c = eph.coefficients()
    # coefficient differences weighted by their largest contribution to a position, dt^(3-k)
    weight = dt ** np.arange(3, -1, -1).reshape(4, 1, 1)
    assert np.max(np.abs(c - c_ref) * weight) < 1e-14
    earth, obs = eph.positions(t)
    np.testing.assert_allclose(earth, spline(t), rtol=0, atol=2e-15)
    np.testing.assert_allclose(obs, earth, rtol=0, atol=0)  # default observer = Earth
    sum_r2, max_e, max_o = eph.stats(t)
    ref = spline(t)
    assert sum_r2 == pytest.approx((ref**2).sum(), rel=1e-13)
    assert max_e == pytest.approx(np.sqrt((ref**2).sum(axis=0)).max(), rel=1e-14)
    # observer knots (e.g. another body): own spline
    obs_knots = earth_knots * 1.3 + 0.01
    eph2 = engine.DeviceEphemeris(t0, dt, earth_knots, obs_knots)
    _, obs2 = eph2.positions(t)
    np.testing.assert_allclose(obs2, CubicSpline(tk, obs_knots, axis=-1)(t), rtol=0, atol=3e-15)
    assert eph2.prepare(t, "knots") == pytest.approx(np.sqrt((obs2**2).sum(axis=0)).max(), rel=1e-14)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("observer", ["earth", "semb-l2", "knots"])
def test_evaluate_tod_on_device_equals_host_interpolation(observer, precision):
    """On-device ephemeris == the reference's host interpolation (scipy CubicSpline per sample,
    get_semb_l2_pos incl. its whole-array norm) fed through the array seam; host and device memory."""
    import torch
    from scipy.interpolate import CubicSpline

    n = 20000
    t0, dt, earth_knots, tk, t, u = _tod_inputs(n, seed=3)
    earth = CubicSpline(tk, earth_knots, axis=-1)(t)  # bodies.py:22-35
    obs_knots = None
    if observer == "earth":
        obs = earth
    elif observer == "semb-l2":  # bodies.py:38-50 (np.linalg.norm over the whole (3, N) array)
        norm = np.linalg.norm(earth)
        obs = earth / norm * (norm + engine.MEAN_DIST_TO_L2)
    else:
        obs_knots = earth_knots * 1.01 + np.array([[0.0], [0.0], [0.002]])
        obs = CubicSpline(tk, obs_knots, axis=-1)(t)
    model = zp.Model(zp.Quantity(25.0, "um"), precision=precision)
    eph = engine.DeviceEphemeris(t0, dt, earth_knots, obs_knots)
    ref = model.evaluate_xyz(u, obs, earth, return_comps=True)
    got = model.evaluate_tod_xyz(u, t, eph, observer=observer, return_comps=True)
    tol = 1e-11 if precision == "fp64" else 3e-6
    np.testing.assert_allclose(got, ref, rtol=tol, atol=1e-30)
    dev = torch.device("cuda:0")
    got_dev = model.evaluate_tod_xyz(torch.as_tensor(u, device=dev), torch.as_tensor(t, device=dev), eph,
                                     observer=observer, return_comps=True)
    np.testing.assert_array_equal(got_dev.cpu().numpy(), got)
    sel = np.arange(0, n, 40)
    ref_o = oracle.evaluate(model.spec, u[:, sel], obs[:, sel], earth[:, sel])
    assert max_rel_total(got[:, sel], ref_o) <= TOL[precision][0]


def test_device_math_routines_on_gpu():
    """The transcendental routines of csrc/zodi_device.cuh evaluated on the GPU itself (MUFU seeds,
    shared-memory tables) against NumPy in double precision."""
    rng = np.random.default_rng(11)
    x = np.exp2(rng.uniform(-60, 60, 300_000))
    assert np.abs(engine.device_math("log2_f64", x) - np.log2(x)).max() <= 4e-16 * 60
    near_one = 1.0 + rng.uniform(-1e-3, 1e-3, 100_000)
    assert np.abs(engine.device_math("log2_f64", near_one) - np.log2(near_one)).max() <= 3e-16
    sp = engine.device_math("log2_f64", np.array([0.0, -1.0, np.inf, np.nan, 5e-324]))
    assert sp[0] == -np.inf and np.isnan(sp[1]) and sp[2] == np.inf and np.isnan(sp[3]) and sp[4] == -1074.0

    e = np.concatenate([rng.uniform(-1019.9, 1019.9, 300_000), rng.uniform(-3, 3, 300_000)])
    got, ref = engine.device_math("exp2_f64", e), np.exp2(e)
    assert (np.abs(got - ref) / ref).max() <= 5e-16
    np.testing.assert_array_equal(engine.device_math("exp2_f64", np.array([-1020.0, -1075.0, -1e9, -np.inf])), 0.0)
    assert np.isnan(engine.device_math("exp2_f64", np.array([np.nan]))).all()

    r = np.exp(rng.uniform(-20, 20, 200_000))
    got = engine.device_math("rsqrt_f64", r)
    assert (np.abs(got * np.sqrt(r) - 1.0)).max() <= 4.5e-16

    for ax in (1.0, -1.0, 0.37, -2.5, 1e-3):
        yv = np.concatenate([rng.uniform(-3, 3, 100_000), rng.uniform(-1e-3, 1e-3, 20_000), [0.0, ax, -ax]])
        got = engine.device_math("atan2_abs_f64", yv, aux=ax)
        assert np.abs(got - np.abs(np.arctan2(yv, ax))).max() <= 9e-16, ax
        got32 = engine.device_math("atan2_abs_f32", yv, aux=ax)
        ref32 = np.abs(np.arctan2(yv.astype(np.float32).astype(np.float64), np.float64(np.float32(ax))))
        assert np.abs(got32 - ref32).max() <= 6e-7, ax

    c = np.linspace(-1, 1, 200_001)
    assert np.abs(engine.device_math("asin_f32", c) - np.arcsin(c.astype(np.float32).astype(np.float64))).max() <= 2.5e-7

    yy = np.concatenate([rng.uniform(0, 0.1, 100_000), rng.uniform(0, 30, 100_000), [0.0, 25.05, 26.0, 200.0]])
    y32 = yy.astype(np.float32).astype(np.float64)
    got = engine.device_math("one_minus_exp2_neg_f32", yy)
    ref = -np.expm1(-y32 * np.log(2.0))
    assert (np.abs(got - ref) <= 4e-6 * ref + 1e-12).all()  # relative, also for tiny arguments
    assert (got[y32 >= 25.05] == 1.0).all()  # the window the kernels skip (kRadialOne) is exactly 1


def test_staged_obstime_protocol():
    """zodi_ephemeris_stats(host times) stages them on the device; zodi_evaluate(obstime=NULL) uses
    that copy (same result as passing the times again) and refuses when nothing matching is staged."""
    import ctypes as C

    from zodipy_b200 import _cabi
    from zodipy_b200.spec import outside_flags

    n = 3 * (1 << 20) + 17  # several pipeline chunks
    t0, dt, earth_knots, tk, t, u = _tod_inputs(n, seed=9)
    model = zp.Model(zp.Quantity(25.0, "um"), precision="fp32")
    dm, eph = model.device_model, engine.DeviceEphemeris(t0, dt, earth_knots)
    r_max = eph.prepare(t, "earth")
    flags = outside_flags(model.spec, r_max)

    def call(obstime_ptr, count):
        out = np.empty(count, dtype=np.float32)
        a = _cabi.EvalArgs()
        a.n, a.u, a.u_stride = count, u.ctypes.data, n
        a.outside_flags = flags.ctypes.data_as(_cabi.c_uint8_p)
        a.precision, a.out_dtype, a.memory = 1, _cabi.OUT_F32, _cabi.MEM_HOST
        a.out, a.out_stride = out.ctypes.data, count
        a.ephemeris, a.obstime = eph._handle, obstime_ptr
        return dm._lib.zodi_evaluate(dm._handle, C.byref(a)), out

    rc, explicit = call(t.ctypes.data, n)
    assert rc == 0
    rc, staged = call(None, n)
    assert rc == 0
    np.testing.assert_array_equal(staged, explicit)
    assert call(None, n - 1)[0] != 0  # staged copy is for n samples
    eph.release_times()
    assert call(None, n)[0] != 0
    assert b"staged" in dm._lib.zodi_last_error()


@pytest.mark.parametrize("name", ["dirbe", "planck18", "rrm-experimental"])
def test_grid_number_density_matches_reference_functions(name):
    """grid_number_density (reference tests/test_model.py:94-123: shape; here also values against
    the oracle's restatement of the 11 density functions)."""
    x = np.linspace(-5, 5, 40)
    y = np.linspace(-5, 5, 30)
    z = np.linspace(-2, 2, 20)
    earth = EARTH_20220114[:, 0]
    grid = zp.grid_number_density_xyz(x, y, z, earth, model=name)
    model = zp.Model(zp.Quantity(25.0, "um") if name != "planck18" else zp.Quantity(857.0, "GHz"), name=name)
    assert grid.shape == (model.ncomps, 30, 40, 20)
    pts = np.asarray(np.meshgrid(x, y, z)).reshape(3, -1)
    for ci, comp in enumerate(model.spec["comps"]):
        with np.errstate(all="ignore"):
            ref = np.broadcast_to(oracle.DENSITY[comp["type"]](pts, comp["params"], EARTH_20220114), pts.shape[1:])
        got = grid[ci].reshape(-1)
        scale = np.nanmax(np.abs(ref))
        np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-13 * scale, err_msg=comp["label"])
    with pytest.raises(TypeError):
        zp.grid_number_density_xyz(x, y, z, earth, model=3)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name,xs,unit", [
    ("dirbe", [1.25, 2.2, 3.5, 4.9, 12.0, 25.0, 60.0, 100.0, 140.0, 240.0], "um"),  # incl. scattering bands
    ("planck18", [100.0, 143.0, 217.0, 353.0, 545.0, 857.0], "GHz"),
    ("dirbe", [25.0, 60.0, 100.0], "um"),                                             # thermal only, NB = 4
])
def test_multiband_equals_per_band_models(name, xs, unit, precision):
    """MultiBandModel == the per-band loop over single-band Models (same kernels' arithmetic up to
    summation order), and within tolerance of the oracle; host and device memory, healpix entry."""
    import torch

    n = 6000
    u = fibonacci_sphere(n)
    mb = zp.MultiBandModel([zp.Quantity(x, unit) for x in xs], name=name, precision=precision)
    got = mb.evaluate_xyz(u, EARTH_20220114)
    assert got.shape == (len(xs), n)
    tol, _ = TOL[precision]
    sel = np.arange(0, n, 7)
    for b, band in enumerate(mb.bands):
        single = band.evaluate_xyz(u, EARTH_20220114)
        np.testing.assert_allclose(got[b], single, rtol=1e-12 if precision == "fp64" else 3e-6)
        ref = oracle.evaluate(band.spec, u[:, sel], EARTH_20220114, EARTH_20220114).sum(axis=0)
        assert np.max(np.abs(got[b, sel] - ref) / np.abs(ref)) <= tol, (xs[b], precision)
    dev = torch.device("cuda:0")
    got_dev = mb.evaluate_xyz(torch.as_tensor(u, device=dev), torch.as_tensor(EARTH_20220114, device=dev))
    np.testing.assert_array_equal(got_dev.cpu().numpy(), got)
    hp = mb.evaluate_healpix(16, EARTH_20220114)
    assert hp.shape == (len(xs), 12 * 256)
    np.testing.assert_allclose(hp[-1], mb.bands[-1].evaluate_healpix(16, EARTH_20220114),
                               rtol=1e-12 if precision == "fp64" else 3e-6)


@pytest.mark.parametrize("name,xs,unit", [
    ("dirbe", [1.25, 2.2, 3.5, 4.9, 12.0, 25.0, 60.0, 100.0, 140.0, 240.0], "um"),  # NB = 16, scattering in 3 bands
    ("planck18", [100.0, 143.0, 217.0, 353.0, 545.0, 857.0], "GHz"),                # NB = 8, four components
    ("dirbe", [25.0, 60.0, 100.0], "um"),                                             # NB = 4
])
def test_multiband_packed_kernel(name, xs, unit, monkeypatch):
    """Packed multi-band kernel (two lines of sight per thread; taken once the pairs fill the machine): equals
    the per-band single-band evaluations and the oracle, odd count (padded last pair), host == device memory,
    and the scalar multi-band kernel (ZODI_NO_X2) within fp32 rounding."""
    import torch

    n = 60001
    u = fibonacci_sphere(n)
    mb = zp.MultiBandModel([zp.Quantity(x, unit) for x in xs], name=name, precision="fp32")
    assert mb.device_model.kernel_name_for(n, "fp32") == "zodi_los_multiband_x2_kernel"
    assert mb.device_model.kernel_name_for(6000, "fp32") == "zodi_los_multiband_kernel"
    assert mb.device_model.kernel_name_for(n, "fp64") == "zodi_los_multiband_kernel"
    got = mb.evaluate_xyz(u, EARTH_20220114)
    assert got.shape == (len(xs), n) and np.isfinite(got).all()
    sel = np.r_[np.arange(0, n, 61), n - 1]
    for b, band in enumerate(mb.bands):
        single = band.evaluate_xyz(u, EARTH_20220114)
        np.testing.assert_allclose(got[b], single, rtol=3e-6)
        ref = oracle.evaluate(band.spec, u[:, sel], EARTH_20220114, EARTH_20220114).sum(axis=0)
        assert np.max(np.abs(got[b, sel] - ref) / np.abs(ref)) <= TOL["fp32"][0], xs[b]
    dev = torch.device("cuda:0")
    got_dev = mb.evaluate_xyz(torch.as_tensor(u, device=dev), torch.as_tensor(EARTH_20220114, device=dev))
    np.testing.assert_array_equal(got_dev.cpu().numpy(), got)
    monkeypatch.setenv("ZODI_NO_X2", "1")
    scalar = zp.MultiBandModel([zp.Quantity(x, unit) for x in xs], name=name, precision="fp32")
    assert scalar.device_model.kernel_name_for(n, "fp32") == "zodi_los_multiband_kernel"
    np.testing.assert_allclose(got, scalar.evaluate_xyz(u, EARTH_20220114), rtol=3e-6)


@pytest.mark.parametrize("name,x,unit", [("planck18", 857.0, "GHz"), ("dirbe", 25.0, "um"), ("dirbe", 1.25, "um")])
def test_persistent_tiles_equal_one_tile_per_cta(name, x, unit, monkeypatch):
    """ZODI_X2_PERSIST=2: the packed kernel as a machine-sized grid whose CTAs claim tiles from a counter (the
    form taken with peer stores) == the default one-tile-per-CTA launch, bit for bit; the counters reset
    themselves, so repeated and interleaved calls keep working; small maps fall back to the plain grid."""
    nside = 256  # 3072 tiles > 1480 resident CTAs
    plain = zp.Model(zp.Quantity(x, unit), name=name, precision="fp32")
    want = plain.evaluate_healpix(nside, EARTH_20220114, out_dtype=np.float32)
    want_c = plain.device_model.evaluate_healpix(nside, EARTH_20220114, return_comps=True, precision="fp32",
                                                 out_dtype=np.float32)
    small = plain.evaluate_healpix(32, EARTH_20220114, out_dtype=np.float32)
    monkeypatch.setenv("ZODI_X2_PERSIST", "2")
    pers = zp.Model(zp.Quantity(x, unit), name=name, precision="fp32")
    for _ in range(3):
        np.testing.assert_array_equal(pers.evaluate_healpix(nside, EARTH_20220114, out_dtype=np.float32), want)
        np.testing.assert_array_equal(
            pers.device_model.evaluate_healpix(nside, EARTH_20220114, return_comps=True, precision="fp32",
                                               out_dtype=np.float32), want_c)
        np.testing.assert_array_equal(pers.evaluate_healpix(32, EARTH_20220114, out_dtype=np.float32), small)
    # host-memory arrays (chunked pipeline: several launches in flight on different streams)
    u = fibonacci_sphere(700001)
    np.testing.assert_array_equal(pers.evaluate_xyz(u, EARTH_20220114), plain.evaluate_xyz(u, EARTH_20220114))


def test_multiband_rejects_unsupported():
    with pytest.raises(engine._cabi.ZodiError):
        zp.MultiBandModel([zp.Quantity(25.0, "um"), zp.Quantity(60.0, "um")], name="rrm-experimental").device_model
    with pytest.raises(ValueError):
        zp.MultiBandModel([zp.Quantity(25.0, "um")] * 17).device_model
    with pytest.raises(ValueError):
        zp.MultiBandModel([zp.Quantity(25.0, "um")], weights=[None, None])


def _year_ephemeris_inputs(n, seed=0):
    """BASELINE config 4 at reduced size: samples spread uniformly over one year (t_i = t0 + i * 365.25 / n),
    hourly knots on np.arange's grid, uniform random pointings."""
    t0, dt = 59215.0, 1.0 / 24.0
    t = t0 + np.arange(n, dtype=np.float64) * (365.25 / n)
    tk = np.arange(t[0], t[-1] + dt, dt)  # arrange_obstimes (zodipy/bodies.py:16-19)
    lon = 2 * np.pi * (tk - t0) / 365.25 + 1.7
    r = 1.0 - 0.0167 * np.cos(lon - 1.8)
    earth_knots = np.array([r * np.cos(lon), r * np.sin(lon), 1e-5 * np.sin(3 * lon)])
    rng = np.random.default_rng(seed)
    u = rng.normal(size=(3, n))
    u /= np.linalg.norm(u, axis=0)
    return tk, earth_knots, t, np.ascontiguousarray(u)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_tod_one_year_semb_l2_device_ephemeris(precision):
    """BASELINE config 4's path at 1e6 samples: per-sample obstimes over one year, hourly knots -> device
    spline, observer = semb-l2 with the reference's whole-array norm (zodipy/bodies.py:38-50), against the
    oracle fed with scipy CubicSpline positions on the host; host-memory and device-memory entries agree."""
    import torch
    from scipy.interpolate import CubicSpline

    n = 1_000_000
    tk, earth_knots, t, u = _year_ephemeris_inputs(n)
    model = zp.Model(zp.Quantity(25.0, "um"), precision=precision)
    eph = model.ephemeris(float(tk[0]), float(tk[1] - tk[0]), earth_knots)
    got = model.evaluate_tod_xyz(u, t, eph, observer="semb-l2")
    earth = CubicSpline(tk, earth_knots, axis=-1)(t)
    norm = np.linalg.norm(earth)  # un-axised: Frobenius norm of the (3, n) array
    scale = (norm + engine.MEAN_DIST_TO_L2) / norm
    sel = np.sort(np.random.default_rng(1).choice(n, 3000, replace=False))
    ref = oracle.evaluate(model.spec, u[:, sel], scale * earth[:, sel], earth[:, sel]).sum(axis=0)
    tol = TOL_FP64 if precision == "fp64" else TOL_FP32
    assert np.max(np.abs(got[sel] - ref) / np.abs(ref)) <= tol
    dev = torch.device("cuda:0")
    got_dev = model.evaluate_tod_xyz(torch.as_tensor(u, device=dev), torch.as_tensor(t, device=dev), eph,
                                     observer="semb-l2")
    np.testing.assert_array_equal(got_dev.cpu().numpy(), got)
    lon, lat = np.arctan2(u[1], u[0]), np.arcsin(np.clip(u[2], -1, 1))
    got_ll = model.evaluate_lonlat(lon, lat, ephemeris=eph, obstime=t, observer="semb-l2")
    np.testing.assert_allclose(got_ll, got, rtol=1e-12 if precision == "fp64" else 3e-6)


@pytest.mark.parametrize("observer", ["earth", "semb-l2"])
def test_multi_device_tod_and_healpix_bitwise(observer, monkeypatch):
    """Model(devices=[...]): time-ordered data with on-device ephemerides split over several handles (global
    semb-l2 norm and early-out flags combined on the host) and HEALPix maps split by pixel range are
    bit-identical to one handle."""
    import torch

    monkeypatch.setattr(engine.MultiDeviceModel, "MIN_LOS_PER_DEVICE", 1)
    devices = [i % torch.cuda.device_count() for i in range(3)]
    one = zp.Model(zp.Quantity(25.0, "um"), precision="fp32", device=0)
    many = zp.Model(zp.Quantity(25.0, "um"), precision="fp32", devices=devices)
    n = 200_003
    tk, earth_knots, t, u = _year_ephemeris_inputs(n, seed=5)
    eph1 = one.ephemeris(float(tk[0]), float(tk[1] - tk[0]), earth_knots)
    ephn = many.ephemeris(float(tk[0]), float(tk[1] - tk[0]), earth_knots)
    assert hasattr(ephn, "parts") and len(ephn.parts) == 3
    ref = one.evaluate_tod_xyz(u, t, eph1, observer=observer, return_comps=True)
    got = many.evaluate_tod_xyz(u, t, ephn, observer=observer, return_comps=True)
    # shards of this size take another lane split than the whole (summation order of the lane partials), and
    # the global sum |earth|^2 of semb-l2 is added in another order (scale equal to ~1 ulp)
    np.testing.assert_allclose(got, ref, rtol=2e-6)
    if observer == "earth":  # same launch shape everywhere -> bit-identical
        monkeypatch.setenv("ZODI_X2_LANES", "2")
        np.testing.assert_array_equal(many.evaluate_tod_xyz(u, t, ephn, observer=observer, return_comps=True),
                                      one.evaluate_tod_xyz(u, t, eph1, observer=observer, return_comps=True))
    lon, lat = np.arctan2(u[1], u[0]), np.arcsin(np.clip(u[2], -1, 1))
    np.testing.assert_allclose(many.evaluate_lonlat(lon, lat, ephemeris=ephn, obstime=t, observer=observer),
                               one.evaluate_lonlat(lon, lat, ephemeris=eph1, obstime=t, observer=observer), rtol=2e-6)
    monkeypatch.setenv("ZODI_X2_LANES", "2")
    for kwargs in ({}, {"return_comps": True}, {"pix_range": (1000, 150_001), "nest": True}):
        np.testing.assert_array_equal(many.evaluate_healpix(128, EARTH_20220114, out_dtype=np.float32, **kwargs),
                                      one.evaluate_healpix(128, EARTH_20220114, out_dtype=np.float32, **kwargs))


def test_nprocesses_maps_to_devices(monkeypatch):
    """Model.evaluate(..., nprocesses=k) -> min(k, visible GPUs) devices when the placement was left to the
    library (zodipy/model.py:182-198: k workers); explicit device= / devices= and LOCAL_RANK are respected."""
    import torch

    monkeypatch.delenv("LOCAL_RANK", raising=False)
    auto = zp.Model(zp.Quantity(25.0, "um"))
    auto._use_processes(1)
    assert auto._devices == [0]
    auto._use_processes(64)
    assert auto._devices == list(range(torch.cuda.device_count()))
    pinned = zp.Model(zp.Quantity(25.0, "um"), device=0)
    pinned._use_processes(64)
    assert pinned._devices == [0]
    monkeypatch.setenv("LOCAL_RANK", "0")
    ranked = zp.Model(zp.Quantity(25.0, "um"))
    ranked._use_processes(64)
    assert ranked._devices == [0]


def test_multiband_outside_flags_are_per_model_component():
    """A multi-band handle returns one row per BAND but reads 2 flag bytes per MODEL component: supplied
    global flags must have shape (n_model_comps, 2) whatever the number of bands (3 bands, 6 components)."""
    mb = zp.MultiBandModel([zp.Quantity(v, "um") for v in (12.0, 25.0, 60.0)], name="dirbe", precision="fp64")
    dm = mb.device_model
    assert dm.ncomps == 3 and dm.n_model_comps == 6
    u = fibonacci_sphere(5000)
    obs = EARTH_20220114 * 1.25  # outside the ring's outer cutoff sphere: early-out flags matter
    flags = zp.Model(zp.Quantity(25.0, "um")).device_model.outside_flags(obs)
    assert flags.shape == (6, 2) and flags.any()
    auto = dm.evaluate(u, obs, EARTH_20220114)
    given = dm.evaluate(u, obs, EARTH_20220114, outside_flags=flags)
    np.testing.assert_array_equal(given, auto)
    with pytest.raises(ValueError):
        dm.evaluate(u, obs, EARTH_20220114, outside_flags=flags[:3])


def test_generic_kernel_with_the_largest_tables_the_abi_admits():
    """1024 table knots + 600 quadrature nodes need more dynamic shared memory than the 48 KB default next
    to the static fp64 math tables: the generic kernel opts in instead of failing at launch."""
    model = zp.Model(zp.Quantity(25.0, "um"), gauss_quad_degree=600)
    spec = dict(model.spec)
    t_old = np.asarray(spec["table"][0])
    t_new = np.linspace(t_old[0], t_old[-1], 1024)
    spec["table"] = np.array([t_new, np.interp(t_new, t_old, np.asarray(spec["table"][1]))])
    dm = engine.DeviceModel(spec, 0)
    assert dm.kernel_name == "zodi_los_generic_kernel"
    u = fibonacci_sphere(300)
    for precision, tol in (("fp64", TOL_FP64), ("fp32", TOL_FP32)):
        got = dm.evaluate(u, EARTH_20220114, return_comps=True, precision=precision)
        ref = oracle.evaluate(spec, u, EARTH_20220114, EARTH_20220114)
        assert max_rel_total(got, ref) <= tol
