/*
 * zodi_b200.h - C ABI of the B200-native line-of-sight brightness integrator.
 *
 * This is the drop-in boundary for ZodiPy's array-only hot path.  The reference
 * (Cosmoglobe/zodipy v1.1.3, pure Python/NumPy) has no FFI; the seam this ABI replaces is the
 * array tail of `Model._evaluate`, /root/reference/zodipy/model.py:253-279, i.e.
 *
 *   get_line_of_sight_range()          zodipy/line_of_sight.py:88-105   (+ :64-85)
 *   integrate_leggauss()               zodipy/line_of_sight.py:55-61
 *   kelsall_brightness_at_step()       zodipy/brightness.py:21-56
 *   rrm_brightness_at_step()           zodipy/brightness.py:59-83
 *   get_dust_grain_temperature()       zodipy/blackbody.py:16-30
 *   np.interp(T, *bp_interpolation_table)   zodipy/brightness.py:48,81
 *   get_scattering_angle()/get_phase_function()   zodipy/scattering.py:11-59
 *   DENSITY_FUNCS[...]                 zodipy/number_density.py:47-420
 *
 * Everything is plain C: pointers, sizes, POD structs.  No torch / C++ types cross the ABI.
 * All functions return 0 (ZODI_OK) on success or a negative zodi_status; the message of the
 * last failure on the calling thread is available from zodi_last_error().  Nothing throws.
 *
 * The library is CUDA-only (sm_100a).  There is no CPU fallback: if no CUDA device is usable
 * every compute entry point fails with ZODI_ERR_CUDA.
 */
#ifndef ZODI_B200_H
#define ZODI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZODI_ABI_VERSION 2
#define ZODI_MAX_COMPS 16   /* reference models have 4, 6 or 8 components */
#define ZODI_MAX_NODES 1024 /* gauss_quad_degree upper bound (reference default: 50) */
#define ZODI_MAX_TEMPS 1024 /* blackbody table knots (reference: 100, zodipy/blackbody.py:11) */
#define ZODI_N_SHAPE 8
#define ZODI_MAX_PEERS 8        /* GPUs of one NVSwitch box */
#define ZODI_IPC_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */

typedef enum {
    ZODI_OK = 0,
    ZODI_ERR_INVALID = -1, /* bad argument / descriptor */
    ZODI_ERR_CUDA = -2,    /* CUDA runtime error or no device */
    ZODI_ERR_NOMEM = -3,
    ZODI_ERR_UNSUPPORTED = -4
} zodi_status;

/* Component density types = keys of DENSITY_FUNCS, zodipy/number_density.py:408-420.
 * shape[] holds the RAW reference parameters (dataclass fields, zodipy/component.py:47-204) in
 * the order given below; angles already in the *_rad form the reference's density functions
 * consume (component.py:85-89,130-136).  The library derives its own device-side constants. */
typedef enum {
    ZODI_CLOUD = 0,        /* n_0, alpha, beta, gamma, mu                         :47-73   */
    ZODI_BAND = 1,         /* n_0, delta_zeta_rad, v, p, delta_r                  :76-110  */
    ZODI_RING = 2,         /* n_0, R, sigma_r, sigma_z                            :113-139 */
    ZODI_FEATURE = 3,      /* n_0, R, sigma_r, sigma_z, theta_rad, sigma_theta_rad :142-181 */
    ZODI_FAN = 4,          /* Q, P, gamma, Z_0, R_outer                           :184-218 */
    ZODI_COMET = 5,        /* gamma, Z_0, P, amp, R_inner, R_outer                :221-256 */
    ZODI_INTERSTELLAR = 6, /* amp                                                 :259-264 */
    ZODI_NARROW_BAND = 7,  /* beta_nb, G, gamma, A, R_inner, R_outer              :267-304 */
    ZODI_BROAD_BAND = 8,   /* beta_bb, sigma_bb, gamma, A, R_inner, R_outer       :307-342 */
    ZODI_RING_RRM = 9,     /* n_0, R, sigma_r, sigma_z, A                         :345-370 */
    ZODI_FEATURE_RRM = 10, /* n_0, R, sigma_r, sigma_z, theta_rad, sigma_theta_rad, A :373-404 */
    ZODI_N_COMP_TYPES = 11
} zodi_comp_type;

typedef enum { ZODI_KELSALL = 0, ZODI_RRM = 1 } zodi_model_kind; /* zodiacal_light_model.py:72-102 */
typedef enum { ZODI_FP64 = 0, ZODI_FP32 = 1 } zodi_precision;
typedef enum { ZODI_OUT_F64 = 0, ZODI_OUT_F32 = 1 } zodi_out_dtype;
typedef enum { ZODI_MEM_HOST = 0, ZODI_MEM_DEVICE = 1 } zodi_memory;

/* One dust component: geometry (zodipy/component.py:27-44), shape parameters, the two
 * heliocentric cutoff radii of its line-of-sight range (COMPONENT_CUTOFFS,
 * zodipy/line_of_sight.py:19-52) and its source-function scalars
 * (zodipy/unpack_model.py:34-57 Kelsall: emissivity, albedo; :120-126 RRM: T_0, delta). */
typedef struct {
    int32_t type; /* zodi_comp_type */
    int32_t reserved;
    double x0[3];
    double sin_Omega, cos_Omega, sin_i, cos_i;
    double shape[ZODI_N_SHAPE];
    double cutoff_inner, cutoff_outer;
    double emissivity, albedo; /* Kelsall kind */
    double T_0, delta;         /* RRM kind (per component) */
} zodi_component_desc;

/* Everything `Model.__init__` prepares for the hot path (zodipy/model.py:101-108,281-301). */
typedef struct {
    int32_t abi_version; /* must be ZODI_ABI_VERSION */
    int32_t kind;        /* zodi_model_kind */
    int32_t n_comps;
    int32_t n_nodes; /* gauss_quad_degree */
    int32_t n_temps; /* knots of the blackbody table */
    int32_t reserved;
    /* Kelsall shared source parameters (zodipy/unpack_model.py:34-105) */
    double T_0, delta, C1, C2, C3, solar_irradiance;
    /* RRM shared source parameter (zodipy/unpack_model.py:128-136) */
    double calibration;
    /* bp_interpolation_table rows (zodipy/blackbody.py:44-49): temperatures must be uniformly
     * spaced and ascending (the reference's are linspace(40, 550, 100)); bnu in MJy/sr. */
    const double* temps;
    const double* bnu;
    /* np.polynomial.legendre.leggauss(n_nodes), zodipy/model.py:103 */
    const double* nodes;
    const double* weights;
    zodi_component_desc comps[ZODI_MAX_COMPS];
} zodi_model_desc;

typedef struct zodi_model_s* zodi_model_t;
typedef struct zodi_ephemeris_s* zodi_ephemeris_t; /* device-resident ephemeris spline, see below */

/* Arguments of one evaluation = the arrays at the seam zodipy/model.py:253-279.
 * Layout is the reference's: structure-of-arrays, row k of a (3, n) array starts at
 * ptr + k*stride (stride >= n, in elements), so a contiguous shard of a larger (3, N) array can
 * be passed without copying (np.array_split on the last axis, zodipy/model.py:184-188). */
typedef struct {
    int64_t n;          /* number of lines of sight in this call */
    const double* u;    /* (3, n) ecliptic unit vectors, model.py:249-251 */
    int64_t u_stride;
    const double* obs;  /* (3, n_obs) observer position [AU], model.py:232-239 */
    int64_t n_obs;      /* 1 (instantaneous) or n (time-ordered) */
    int64_t obs_stride;
    const double* earth; /* (3, n_earth) Earth position [AU], model.py:213-220,240-241 */
    int64_t n_earth;     /* 1 or n */
    int64_t earth_stride;
    /* (n_comps, 2) bytes, ALWAYS host memory: [c][0] != 0 <=> any observer of the WHOLE job is
     * outside component c's inner cutoff sphere, [c][1] likewise for the outer one (the global
     * `.any()` early-out of get_sphere_intersection, line_of_sight.py:72-73).
     * NULL: the library derives the flags from the observers of THIS call. */
    const uint8_t* outside_flags;
    int32_t return_comps; /* 0: out is (n,) sum over components; 1: out is (n_comps, n) */
    int32_t precision;    /* zodi_precision */
    int32_t out_dtype;    /* zodi_out_dtype */
    int32_t memory;       /* zodi_memory: where u/obs/earth/out live */
    void* out;
    int64_t out_stride; /* row stride of out in elements when return_comps (>= n) */
    void* stream;       /* cudaStream_t for ZODI_MEM_DEVICE (NULL = default stream); async.
                           ZODI_MEM_HOST calls are synchronous and ignore it. */
    /* Fused all-gather (ZODI_MEM_DEVICE only): when n_peers > 0 the kernel epilogue stores element
     * j of this call (component row c) to peer_out[p][c * peer_stride + peer_offset + j] for every
     * p < n_peers - the map buffers of all GPUs of the box (own buffer included), mapped with
     * zodi_peer_buffer_open - so the slice a rank computes lands in every rank's full map while
     * the kernel is still running; `out` is then ignored and may be NULL.  This replaces the
     * np.concatenate of the reference's worker results (zodipy/model.py:198). */
    int32_t n_peers;
    int32_t reserved;
    void* peer_out[ZODI_MAX_PEERS];
    int64_t peer_offset;
    int64_t peer_stride;
    /* Block-cyclic shard layout (load balance: contiguous shards of a RING map are latitude bands
     * of unequal cost).  When cyclic_block > 0, line of sight j of this call is element
     *     g(j) = ((j / cyclic_block) * cyclic_parts + cyclic_rank) * cyclic_block + j % cyclic_block
     * of the global job: peer stores go to peer_offset + g(j) and zodi_evaluate_healpix integrates
     * pixel ipix_start + g(j).  Input arrays and `out` stay indexed by the local j. */
    int64_t cyclic_block;
    int32_t cyclic_parts;
    int32_t cyclic_rank;
    /* On-device ephemeris: when `ephemeris` is non-NULL, obs / earth (and their n_*, strides) are
     * ignored and sample j uses the spline positions at obstime[j] (obstime: n doubles, same
     * memory kind as u).  outside_flags must then be supplied (see zodi_ephemeris_stats).
     * ZODI_MEM_HOST only: obstime may be NULL when the preceding zodi_ephemeris_stats call was given
     * the same n host times - the copy it staged on the device is integrated from, so the times
     * cross the bus once (32 instead of 40 B per sample in total). */
    zodi_ephemeris_t ephemeris;
    const double* obstime;
} zodi_eval_args;

/* Time-ordered data with Earth / observer positions evaluated ON THE DEVICE from a cubic spline
 * through uniformly spaced ephemeris knots.  Replaces get_interp_bodypos (zodipy/bodies.py:22-35:
 * hourly knots from arrange_obstimes :16-19 + scipy.interpolate.CubicSpline, default not-a-knot
 * ends, evaluated at every sample) and the SEMB-L2 scaling (:38-50), and removes the 48 B per
 * sample of position arrays from the host -> device traffic. */
typedef struct {
    int64_t n_knots;           /* >= 4 */
    double t0, dt;             /* knot k is at time t0 + k * dt (same unit as the obstime array) */
    const double* earth_knots; /* (3, n_knots) row-major, host: Earth position at the knots [AU] */
    const double* obs_knots;   /* (3, n_knots) host, or NULL: observer = obs_scale * Earth */
    double obs_scale;          /* 1 for obspos="earth"; 1 + L2 / ||earth|| for "semb-l2" */
} zodi_ephemeris_desc;

/* HEALPix map evaluation with directions generated on the device (no (3, N) upload):
 * line of sight j of the call is the centre of pixel ipix_start + j (RING or NESTED) of resolution nside,
 * optionally rotated by `rot` (row-major 3x3: pixel frame -> mean ecliptic, i.e. the constant
 * matrix behind skycoord.transform_to(BarycentricMeanEcliptic), zodipy/model.py:247).
 * `base.u`/`base.u_stride` are ignored; everything else in `base` keeps its meaning
 * (base.memory describes obs / earth / out).  Replaces, for map-making callers, healpy.pix2ang +
 * SkyCoord construction (docs/examples/healpy_map.py:14-18) in front of the same hot path. */
typedef struct {
    zodi_eval_args base;
    int64_t nside;
    int64_t ipix_start;
    int32_t nest;      /* 0 = RING, 1 = NESTED (nside must then be a power of two) */
    int32_t has_rot;
    double rot[9];
} zodi_healpix_args;

/* Evaluation with directions given as spherical sky coordinates, the form a SkyCoord holds them in:
 * line of sight j points at longitude lon[j], latitude lat[j] [rad] of a frame whose constant
 * rotation to the mean ecliptic is `rot` (row-major 3x3; identity when has_rot == 0).  The kernel
 * prologue forms (cos lat cos lon, cos lat sin lon, sin lat) and rotates it, i.e. it replaces
 * skycoord.transform_to(BarycentricMeanEcliptic).cartesian.xyz (zodipy/model.py:247-251) for every
 * frame that differs from the mean ecliptic by a fixed rotation (ICRS, Galactic, FK5, ecliptic).
 * lon / lat: n doubles each in the memory kind of `base` (base.u / base.u_stride are ignored);
 * everything else in `base` keeps its meaning (per-sample obs / earth, ephemeris + obstime,
 * peer output, ...).  Host -> device traffic is 16 instead of 24 B per line of sight. */
typedef struct {
    zodi_eval_args base;
    const double* lon;
    const double* lat;
    int32_t has_rot;
    int32_t reserved;
    double rot[9];
} zodi_lonlat_args;

/* ---- library ---------------------------------------------------------------------------- */
int zodi_abi_version(void);
const char* zodi_last_error(void);
int zodi_device_count(int* count);

/* ---- model handle: device copy of the parameter block on one GPU -------------------------- */
int zodi_model_create(const zodi_model_desc* desc, int device, zodi_model_t* out);
int zodi_model_update(zodi_model_t model, const zodi_model_desc* desc); /* Model.update_parameters */
int zodi_model_destroy(zodi_model_t model);

/* ---- on-device ephemeris (time-ordered data) ------------------------------------------------- */
int zodi_ephemeris_create(int device, const zodi_ephemeris_desc* desc, zodi_ephemeris_t* out);
int zodi_ephemeris_set_obs_scale(zodi_ephemeris_t eph, double obs_scale);
int zodi_ephemeris_destroy(zodi_ephemeris_t eph);
/* Piecewise-cubic coefficients of the Earth spline as scipy's CubicSpline.c lays them out:
 * c[(k * (n_knots - 1) + i) * 3 + axis], k = 0..3 (highest power first); host output, for tests. */
int zodi_ephemeris_coefficients(zodi_ephemeris_t eph, double* c);
/* Spline positions at n times: earth_out / obs_out (3, n) row-major (either may be NULL). */
int zodi_ephemeris_positions(zodi_ephemeris_t eph, const double* t, int64_t n, int32_t memory,
                             double* earth_out, double* obs_out, void* stream);
/* Reductions over the samples needed by the reference's semantics: sum of |earth|^2 (the
 * un-axised np.linalg.norm of get_semb_l2_pos, bodies.py:47), max |earth|^2 and
 * max |observer|^2 (the global early-out flags, line_of_sight.py:72-73).  stats = 3 doubles (host). */
int zodi_ephemeris_stats(zodi_ephemeris_t eph, const double* t, int64_t n, int32_t memory, void* stream,
                         double* stats);
/* With ZODI_MEM_HOST, zodi_ephemeris_stats keeps its device copy of the n times in the handle (8 B
 * per sample; the buffer is reused by later calls) for a following zodi_evaluate(obstime = NULL);
 * this frees it (destroy does too). */
int zodi_ephemeris_release_times(zodi_ephemeris_t eph);

/* ---- the hot path ------------------------------------------------------------------------ */
int zodi_evaluate(zodi_model_t model, const zodi_eval_args* args);

/* Name of the kernel family zodi_evaluate launches for this model: "zodi_los_kelsall_kernel"
 * (fused Kelsall-family kernel) or "zodi_los_generic_kernel" (any component list). */
const char* zodi_model_kernel_name(zodi_model_t model);
/* Exact kernel zodi_evaluate would launch for n lines of sight in the given precision
 * ("zodi_los_kelsall_x2_kernel" = packed fp32 variant, ...). */
const char* zodi_model_kernel_for(zodi_model_t model, int64_t n, int32_t precision);

int zodi_evaluate_healpix(zodi_model_t model, const zodi_healpix_args* args);
/* Pixel-centre unit vectors (3, n) of pixels [ipix_start, ipix_start + n) (RING or NESTED) into
 * device or host memory (`memory`), rotated by rot if non-NULL: the directions
 * zodi_evaluate_healpix integrates. */
int zodi_healpix_vectors(int device, int64_t nside, int32_t nest, int64_t ipix_start, int64_t n,
                         const double* rot, double* out, int64_t out_stride, int32_t memory, void* stream);

int zodi_evaluate_lonlat(zodi_model_t model, const zodi_lonlat_args* args);
/* The unit vectors zodi_evaluate_lonlat integrates along: (3, n) into `out` (row stride out_stride),
 * lon / lat / out in `memory`; rot may be NULL. */
int zodi_lonlat_vectors(int device, const double* lon, const double* lat, int64_t n, const double* rot,
                        double* out, int64_t out_stride, int32_t memory, void* stream);

/* Largest heliocentric observer distance sqrt(x^2+y^2+z^2) over (3, n_obs) observers and the
 * resulting early-out flags; used to form the GLOBAL flags when a job is sharded over GPUs
 * (max-reduce r_max over ranks, then zodi_flags_from_radius). */
int zodi_max_observer_radius(zodi_model_t model, const double* obs, int64_t n_obs,
                             int64_t obs_stride, int32_t memory, void* stream, double* r_max);
int zodi_flags_from_radius(zodi_model_t model, double r_max, uint8_t* flags /* (n_comps,2) */);

/* ---- peer-visible map buffers (one process per GPU; CUDA IPC over NVLink/NVSwitch) ----------
 * alloc: cudaMalloc on `device` + export handle; open: map another process's buffer into this
 * process (peer access is enabled lazily); close/free undo them. */
int zodi_peer_buffer_alloc(int device, int64_t bytes, void** ptr, uint8_t handle[ZODI_IPC_HANDLE_BYTES]);
int zodi_peer_buffer_open(int device, const uint8_t handle[ZODI_IPC_HANDLE_BYTES], void** ptr);
int zodi_peer_buffer_close(int device, void* ptr);
int zodi_peer_buffer_free(int device, void* ptr);

/* Completion rendezvous of the fused all-gather, without NCCL: every rank owns a flag array of
 * ZODI_MAX_PEERS + 1 uint32 (allocated zeroed with zodi_peer_buffer_alloc, mapped by all peers with
 * zodi_peer_buffer_open); peer_flags[p] is rank p's array as mapped in THIS process.  The call enqueues
 * one tiny kernel on `stream` behind the integrator kernel: it publishes `epoch` into word `rank` of
 * every peer's array (system-scope release after the kernel's peer stores) and waits until words
 * 0..n_peers-1 of the own array have reached `epoch` (acquire), i.e. until every rank's slice has landed
 * in this rank's map.  Epochs must increase by one per rendezvous.  A peer that never arrives is
 * reported after ~10 s in word ZODI_MAX_PEERS of the own array (1 = timed out) instead of hanging.
 * Replaces the 4-byte NCCL all-reduce (23 us) of round 1 by ~3 us. */
int zodi_peer_rendezvous(int device, void* const* peer_flags, int32_t n_peers, int32_t rank, uint32_t epoch,
                         void* stream);

/* ---- several bands of one model in a single pass ----------------------------------------------
 * descs[0..n_bands-1] describe the SAME Kelsall-family model (identical components, cutoffs, T_0,
 * delta, quadrature) at different wavelengths / bandpasses: they may differ only in the blackbody
 * table values and in emissivity / albedo / C1-C3 / solar_irradiance (what Model.__init__ derives
 * from x, zodipy/model.py:101, zodipy/unpack_model.py:34-105).  zodi_multiband_evaluate takes the
 * same zodi_eval_args as zodi_evaluate (u, obs, earth, flags, ephemeris, ...; return_comps is
 * ignored) and writes the component-summed emission of band b to row b of `out`
 * (row stride out_stride >= n): out is (n_bands, n).  Works for host and device memory. */
#define ZODI_MAX_BANDS 16
typedef struct zodi_multiband_s* zodi_multiband_t;
int zodi_multiband_create(const zodi_model_desc* descs, int32_t n_bands, int device, zodi_multiband_t* out);
int zodi_multiband_evaluate(zodi_multiband_t mb, const zodi_eval_args* args);
int zodi_multiband_evaluate_healpix(zodi_multiband_t mb, const zodi_healpix_args* args);
int zodi_multiband_evaluate_lonlat(zodi_multiband_t mb, const zodi_lonlat_args* args);
int zodi_multiband_destroy(zodi_multiband_t mb);

/* ---- component densities on a set of points -------------------------------------------------
 * Replaces the array part of grid_number_density (zodipy/number_density.py:482-536): the number
 * density of every component at n heliocentric ecliptic points xyz (3, n) [AU], for one Earth
 * position earth[3] (host memory; used by the Earth-trailing feature).  out: (n_comps, n) float64.
 * xyz / out live in `memory`. */
int zodi_number_density(zodi_model_t model, const double* xyz, int64_t n, int64_t xyz_stride,
                        const double* earth, double* out, int64_t out_stride, int32_t memory, void* stream);

/* ---- measurement support ------------------------------------------------------------------ */
typedef enum {
    ZODI_PEAK_FP32_FMA = 0, /* FFMA  : flop/s (2 per FMA)      */
    ZODI_PEAK_FP64_FMA = 1, /* DFMA  : flop/s (2 per FMA)      */
    ZODI_PEAK_MUFU_EX2 = 2, /* MUFU.EX2 : op/s                 */
    ZODI_PEAK_HBM_COPY = 3  /* device copy: bytes/s (read+write) */
} zodi_peak_kind;
int zodi_peak_probe(int device, int32_t kind, double* per_second);
/* Element-wise evaluation of the device math routines the integrators are built from, ON the
 * GPU (for tests: the host cannot reproduce MUFU seeds / shared-memory tables).  x, y: n doubles in
 * host memory; single-precision routines convert x (and aux) to float first.  aux: second argument
 * of the two-argument routines (atan2_abs(y = x[i], x = aux)). */
typedef enum {
    ZODI_MATH_LOG2_F64 = 0,
    ZODI_MATH_EXP2_F64 = 1,
    ZODI_MATH_RSQRT_F64 = 2,
    ZODI_MATH_ATAN2_ABS_F64 = 3,
    ZODI_MATH_ASIN_F32 = 4,
    ZODI_MATH_ATAN2_ABS_F32 = 5,
    ZODI_MATH_ONE_MINUS_EXP2_NEG_F32 = 6,
    ZODI_MATH_EXP2_F32 = 7,
    ZODI_MATH_LOG2_F32 = 8
} zodi_math_op;
int zodi_device_math(int device, int32_t op, int64_t n, const double* x, double aux, double* y);
/* Number of kernels this library launched on behalf of the calling process (all threads). */
int64_t zodi_kernel_launch_count(void);
/* Device time [ms] of the LAST zodi_evaluate kernel(s) issued with ZODI_MEM_HOST memory
 * (CUDA events around the kernels only, excluding copies); for reporting. */
double zodi_last_kernel_ms(zodi_model_t model);

#ifdef __cplusplus
}
#endif
#endif /* ZODI_B200_H */
