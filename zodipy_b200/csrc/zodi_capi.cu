// zodi_capi.cu - implementation of the C ABI declared in include/zodi_b200.h.
//
// Host side of the library: validates the model descriptor, derives the device-form constants
// (in double, then narrows for the fp32 kernels), owns the small device buffers of a model
// handle (blackbody table, quadrature nodes, a pipelined staging workspace for host-memory
// calls) and launches the kernels of zodi_kernels.cuh.  Only the CUDA runtime is used.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "zodi_launch.hpp"
#include "zodi_misc_kernels.cuh"
#include "zodi_model_build.hpp"

using namespace zodi;

namespace zodi {

std::atomic<int64_t> g_launches{0};

int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load();
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n);
    }
    return n;
}

int pick_lanes(int64_t n, int n_nodes) {
    // enough threads to fill every SM's 2048 resident threads; never more lanes than nodes
    const int64_t fill = (int64_t)sm_count() * 2048;
    int L = 1;
    while (L < 32 && n * L < fill && 2 * L <= n_nodes) L *= 2;
    return L;
}

// Packed kernels: a thread works on a PAIR of lines of sight, 1280 threads are resident per SM (10 CTAs
// of 128).  L lanes per pair split the quadrature nodes; the count minimises a cost model fitted to the
// measured shape sweep (profiles/r2_packed_shape_sweep.jsonl):
//     time ~ (P + loops * ceil(n_nodes / L)) * L / throughput(f),   f = pairs * L / resident thread slots,
// P = 6 node-iterations' worth of per-thread prologue / reduction, loops = 1 (cloud + bands) or 2 (ring and
// feature loops weigh about as much as the first), throughput(f) = min(1, 1.25 f / (f + 0.25)): half-filled
// SMs already reach 84 % of the full rate.  ZODI_X2_LANES / ZODI_X2_THREADS force a shape (measurements,
// tests).
PackedShape pick_packed_shape(int64_t n, int n_nodes, int n_comps) {
    const char* el = std::getenv("ZODI_X2_LANES");
    const char* et = std::getenv("ZODI_X2_THREADS");
    const int env_lanes = el ? std::atoi(el) : 0, env_threads = et ? std::atoi(et) : 0;
    PackedShape s;
    s.threads = (env_threads == 128 || env_threads == 256) ? env_threads : kPackedDefaultThreads;
    if (env_lanes == 1 || env_lanes == 2 || env_lanes == 4 || env_lanes == 8) {
        s.lanes = env_lanes;
        return s;
    }
    const double pairs = 0.5 * (double)(n + 1), slots = (double)sm_count() * 1280.0;
    const double loops = n_comps == 6 ? 2.0 : 1.0;
    double best = 0.0;
    s.lanes = 1;
    for (int L = 1; L <= 8 && (L == 1 || 2 * L <= 2 * n_nodes); L *= 2) {
        const double f = pairs * L / slots;
        const double thr = std::fmin(1.0, 1.25 * f / (f + 0.25));
        const double cost = (6.0 + loops * ((n_nodes + L - 1) / L)) * L / thr;
        if (L == 1 || cost < best) { best = cost; s.lanes = L; }
    }
    return s;
}

cudaError_t launch_kelsall_packed_l1(const KelsallModel<float>&, const LaunchArgs&, const Pair<float>*,
                                     const Pair<float>*, int, cudaStream_t);
cudaError_t launch_kelsall_packed_l2(const KelsallModel<float>&, const LaunchArgs&, const Pair<float>*,
                                     const Pair<float>*, int, cudaStream_t);
cudaError_t launch_kelsall_packed_l4(const KelsallModel<float>&, const LaunchArgs&, const Pair<float>*,
                                     const Pair<float>*, int, cudaStream_t);
cudaError_t launch_kelsall_packed_l8(const KelsallModel<float>&, const LaunchArgs&, const Pair<float>*,
                                     const Pair<float>*, int, cudaStream_t);

cudaError_t launch_kelsall_packed(const KelsallModel<float>& K, const LaunchArgs& a, const Pair<float>* tab,
                                  const Pair<float>* nodes, PackedShape shape, cudaStream_t stream) {
    switch (shape.lanes) {
        case 1: return launch_kelsall_packed_l1(K, a, tab, nodes, shape.threads, stream);
        case 2: return launch_kelsall_packed_l2(K, a, tab, nodes, shape.threads, stream);
        case 4: return launch_kelsall_packed_l4(K, a, tab, nodes, shape.threads, stream);
        default: return launch_kelsall_packed_l8(K, a, tab, nodes, shape.threads, stream);
    }
}

}  // namespace zodi

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU_CHECK(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(ZODI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr int kSlots = 3;  // pipeline depth of the host-memory path
constexpr int kTileSlots = 64;

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t k0 = nullptr, k1 = nullptr;
    double* d_in = nullptr;   // u (3 rows) [+ obs (3 rows)] [+ earth (3 rows)], row pitch = chunk
    void* d_out = nullptr;    // up to ZODI_MAX_COMPS rows
    bool used = false;
};

}  // namespace

struct zodi_model_s {
    int device = 0;
    zodi_model_desc desc;  // raw copy (pointers nulled)
    DevModel<double> m64;
    DevModel<float> m32;
    bool kelsall_ok = false;  // model fits the fused Kelsall-family kernel
    KelsallModel<double> k64;
    KelsallModel<float> k32;
    bool rrm_ok = false;      // model fits the fused RRM kernel
    RrmModel<double> r64;
    RrmModel<float> r32;
    bool rrm_x2_ok = false;   // ... and its packed fp32 form
    RrmModelX2 rx2;
    // multi-band extension (zodi_multiband_*): this handle then carries band 0 and the shared parts
    int mb_bands = 0;
    MultiBandModel<double> mb64;
    MultiBandModel<float> mb32;
    Pair<double>* d_mbtab64 = nullptr;  // [n_bands][n_temps]
    Pair<float>* d_mbtab32 = nullptr;
    Pair<float>* d_mbrows32 = nullptr;  // the same tables knot-major, n_bands_padded + 2 pairs per row (packed kernel)
    int force_generic = 0;    // testing knob (ZODI_FORCE_GENERIC=1): always use the generic kernel
    int no_x2 = 0;            // testing knob (ZODI_NO_X2=1): scalar fused kernel instead of packed
    Pair<double>* d_table64 = nullptr;
    Pair<double>* d_nodes64 = nullptr;
    Pair<float>* d_table32 = nullptr;
    Pair<float>* d_nodes32 = nullptr;
    unsigned long long* d_scratch = nullptr;  // 8 bytes for the max-radius reduction
    // persistent-tile counters of the packed kernel (LaunchArgs::tile_counter): kTileSlots pairs handed out
    // round robin, so launches of one model that overlap on different streams do not share a pair
    unsigned int* d_tiles = nullptr;
    std::atomic<unsigned> tile_slot{0};
    // ZODI_X2_PERSIST: 0 one tile per CTA (default), 1 persistent tiles for launches that store to peers,
    // 2 for every large launch.  Opt-in: measured -0.5 % at 2 GPUs in one pass and +2 ... +3.6 % at 2 - 8 GPUs
    // in another, so the plain grid stays the default.
    int persist = 0;
    // host-memory path workspace
    std::mutex ws_mutex;
    Slot slots[kSlots];
    int64_t ws_chunk = 0;
    double last_kernel_ms = 0.0;
};

struct zodi_ephemeris_s {
    int device = 0;
    int64_t n_knots = 0;
    double t0 = 0.0, dt = 1.0, obs_scale = 1.0;
    std::vector<double> earth_c;  // host copy, device layout [seg][axis][4]
    double* d_earth = nullptr;
    double* d_obs = nullptr;      // NULL: observer = obs_scale * Earth
    double* d_stats = nullptr;    // 3 doubles: sum, max bits x2
    double* d_times = nullptr;    // sample times staged by the last host-memory zodi_ephemeris_stats
    int64_t n_times = 0;          // staged samples (0: nothing valid)
    int64_t cap_times = 0;        // allocated samples (the buffer is reused across calls)
};

namespace {

int validate_desc(const zodi_model_desc* d) {
    if (!d) return fail(ZODI_ERR_INVALID, "model descriptor is NULL");
    if (d->abi_version != ZODI_ABI_VERSION)
        return fail(ZODI_ERR_INVALID, "descriptor abi_version %d != library %d", d->abi_version,
                    ZODI_ABI_VERSION);
    if (d->kind != ZODI_KELSALL && d->kind != ZODI_RRM)
        return fail(ZODI_ERR_INVALID, "unknown model kind %d", d->kind);
    if (d->n_comps < 1 || d->n_comps > ZODI_MAX_COMPS)
        return fail(ZODI_ERR_INVALID, "n_comps=%d outside [1, %d]", d->n_comps, ZODI_MAX_COMPS);
    if (d->n_nodes < 1 || d->n_nodes > ZODI_MAX_NODES)
        return fail(ZODI_ERR_INVALID, "n_nodes=%d outside [1, %d]", d->n_nodes, ZODI_MAX_NODES);
    if (d->n_temps < 2 || d->n_temps > ZODI_MAX_TEMPS)
        return fail(ZODI_ERR_INVALID, "n_temps=%d outside [2, %d]", d->n_temps, ZODI_MAX_TEMPS);
    if (!d->temps || !d->bnu || !d->nodes || !d->weights)
        return fail(ZODI_ERR_INVALID, "temps/bnu/nodes/weights must be non-NULL");
    const double dt = (d->temps[d->n_temps - 1] - d->temps[0]) / (d->n_temps - 1);
    if (!(dt > 0.0)) return fail(ZODI_ERR_INVALID, "table temperatures must be ascending");
    for (int i = 0; i < d->n_temps; ++i) {
        const double expect = d->temps[0] + dt * i;
        if (std::fabs(d->temps[i] - expect) > 1e-9 * std::fabs(d->temps[d->n_temps - 1]))
            return fail(ZODI_ERR_UNSUPPORTED,
                        "table temperatures must be uniformly spaced (knot %d = %.17g, expected %.17g)",
                        i, d->temps[i], expect);
    }
    for (int i = 0; i < d->n_comps; ++i)
        if (n_shape_params(d->comps[i].type) < 0)
            return fail(ZODI_ERR_INVALID, "component %d has unknown type %d", i, d->comps[i].type);
    return ZODI_OK;
}

int upload_model(zodi_model_s* m, const zodi_model_desc* d) {
    build_dev_model(*d, m->m64);
    narrow_model(m->m64, m->m32);
    m->kelsall_ok = d->n_temps <= kFastMaxTemps && d->n_nodes <= kFastMaxNodes &&
                    build_kelsall_model(*d, m->k64);
    m->rrm_ok = d->n_temps <= kFastMaxTemps && d->n_nodes <= kFastMaxNodes && build_rrm_model(*d, m->m64, m->r64);
    if (m->rrm_ok) narrow_rrm(m->r64, m->m32, m->r32);
    m->rrm_x2_ok = m->rrm_ok && build_rrm_x2(m->r64, m->r32, m->rx2);
    if (m->kelsall_ok) narrow_kelsall(m->k64, m->k32);
    const char* fg = std::getenv("ZODI_FORCE_GENERIC");
    m->force_generic = (fg && fg[0] == '1');
    const char* nx = std::getenv("ZODI_NO_X2");
    m->no_x2 = (nx && nx[0] == '1');
    const char* ps = std::getenv("ZODI_X2_PERSIST");
    if (ps && ps[0] >= '0' && ps[0] <= '2') m->persist = ps[0] - '0';

    // ---- table as (B_i, B_{i+1}-B_i) pairs, nodes as (x_k, w_k) pairs ----
    std::vector<Pair<double>> t64, n64;
    std::vector<Pair<float>> t32, n32;
    build_pairs(*d, t64, n64, t32, n32);
    auto put = [](void** dst, const void* src, size_t bytes) -> cudaError_t {
        if (*dst) { cudaFree(*dst); *dst = nullptr; }
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    CU_CHECK(put((void**)&m->d_table64, t64.data(), t64.size() * sizeof(Pair<double>)));
    CU_CHECK(put((void**)&m->d_nodes64, n64.data(), n64.size() * sizeof(Pair<double>)));
    CU_CHECK(put((void**)&m->d_table32, t32.data(), t32.size() * sizeof(Pair<float>)));
    CU_CHECK(put((void**)&m->d_nodes32, n32.data(), n32.size() * sizeof(Pair<float>)));
    if (!m->d_scratch) CU_CHECK(cudaMalloc((void**)&m->d_scratch, sizeof(unsigned long long)));
    if (!m->d_tiles) {
        CU_CHECK(cudaMalloc((void**)&m->d_tiles, kTileSlots * 2 * sizeof(unsigned int)));
        CU_CHECK(cudaMemset(m->d_tiles, 0, kTileSlots * 2 * sizeof(unsigned int)));
    }

    m->desc = *d;
    m->desc.temps = m->desc.bnu = m->desc.nodes = m->desc.weights = nullptr;
    return ZODI_OK;
}

void free_workspace(zodi_model_s* m) {
    for (Slot& s : m->slots) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.k0) cudaEventDestroy(s.k0);
        if (s.k1) cudaEventDestroy(s.k1);
        if (s.stream) cudaStreamDestroy(s.stream);
        s = Slot();
    }
    m->ws_chunk = 0;
}

uint32_t flags_to_mask(const uint8_t* flags, int n_comps) {
    uint32_t mask = 0;
    for (int c = 0; c < n_comps; ++c) {
        if (flags[2 * c]) mask |= 1u << (2 * c);
        if (flags[2 * c + 1]) mask |= 1u << (2 * c + 1);
    }
    return mask;
}

void flags_from_r(const zodi_model_s* m, double r_max, uint8_t* flags) {
    for (int c = 0; c < m->desc.n_comps; ++c) {  // r_obs > cutoff, line_of_sight.py:72
        flags[2 * c] = r_max > m->desc.comps[c].cutoff_inner;
        flags[2 * c + 1] = r_max > m->desc.comps[c].cutoff_outer;
    }
}

// Packed RRM kernel (one thread per PAIR of lines of sight, no lane split): once the pairs half-fill the
// machine; smaller inputs take the scalar fused kernel with several lanes per line of sight.
bool rrm_takes_packed(const zodi_model_s* m, int64_t n) {
    return m->rrm_x2_ok && !m->no_x2 && n >= (int64_t)sm_count() * 1024;
}

// Packed multi-band kernel (one thread per PAIR of lines of sight): once the pairs give every SM a CTA;
// smaller inputs keep one line of sight per thread.
bool multiband_takes_packed(const zodi_model_s* m, int64_t n) {
    return !m->no_x2 && n >= (int64_t)sm_count() * 2 * kPackedDefaultThreads;
}

cudaError_t launch_eval(zodi_model_s* m, const LaunchArgs& a, int precision, cudaStream_t stream) {
    if (a.n <= 0) return cudaSuccess;
    if (m->mb_bands > 0) {
        if (precision == ZODI_FP32) {
            if (multiband_takes_packed(m, a.shape_n > 0 ? a.shape_n : a.n))
                return launch_multiband_packed(m->mb32, a, m->d_mbrows32, m->d_nodes32, stream);
            return launch_multiband_f32(m->mb32, a, m->d_mbtab32, m->d_nodes32, stream);
        }
        return launch_multiband_f64(m->mb64, a, m->d_mbtab64, m->d_nodes64, stream);
    }
    if (m->kelsall_ok && !m->force_generic) {
        // packed-fp32 kernels: every fp32 evaluation of a Kelsall-family model
        if (precision == ZODI_FP32 && !m->no_x2) {
            const PackedShape shape = pick_packed_shape(a.shape_n > 0 ? a.shape_n : a.n, m->k32.n_nodes, m->k32.n_comps);
            if (m->persist == 2 || (m->persist == 1 && a.n_peers > 0)) {
                // persistent tiles (the launcher drops the counter again when the grid fits the machine anyway)
                LaunchArgs b = a;
                b.tile_counter = m->d_tiles + 2 * (m->tile_slot.fetch_add(1) % kTileSlots);
                return launch_kelsall_packed(m->k32, b, m->d_table32, m->d_nodes32, shape, stream);
            }
            return launch_kelsall_packed(m->k32, a, m->d_table32, m->d_nodes32, shape, stream);
        }
        if (precision == ZODI_FP32) return launch_kelsall_f32(m->k32, a, m->d_table32, m->d_nodes32, stream);
        return launch_kelsall_f64(m->k64, a, m->d_table64, m->d_nodes64, stream);
    }
    if (m->rrm_ok && !m->force_generic) {
        if (precision == ZODI_FP32 && rrm_takes_packed(m, a.shape_n > 0 ? a.shape_n : a.n))
            return launch_rrm_packed(m->rx2, a, m->d_table32, m->d_nodes32, stream);
        if (precision == ZODI_FP32) return launch_rrm_f32(m->r32, a, m->d_table32, m->d_nodes32, stream);
        return launch_rrm_f64(m->r64, a, m->d_table64, m->d_nodes64, stream);
    }
    if (precision == ZODI_FP32) return launch_generic_f32(m->m32, a, m->d_table32, m->d_nodes32, stream);
    return launch_generic_f64(m->m64, a, m->d_table64, m->d_nodes64, stream);
}

int max_r_device(zodi_model_s* m, const double* d_obs, int64_t n_obs, int64_t stride,
                 cudaStream_t stream, double* r_max) {
    CU_CHECK(cudaMemsetAsync(m->d_scratch, 0, sizeof(unsigned long long), stream));
    const int grid = (int)std::min<int64_t>((n_obs + 255) / 256, (int64_t)sm_count() * 8);
    zodi_max_r2_kernel<<<grid, 256, 0, stream>>>(d_obs, n_obs, stride, m->d_scratch);
    g_launches.fetch_add(1);
    CU_CHECK(cudaGetLastError());
    unsigned long long bits = 0;
    CU_CHECK(cudaMemcpyAsync(&bits, m->d_scratch, sizeof(bits), cudaMemcpyDeviceToHost, stream));
    CU_CHECK(cudaStreamSynchronize(stream));
    double r2;
    std::memcpy(&r2, &bits, sizeof(r2));
    *r_max = std::sqrt(r2);
    return ZODI_OK;
}

double max_r_host(const double* obs, int64_t n_obs, int64_t stride) {
    auto range_max = [&](int64_t lo, int64_t hi) {
        double m = 0.0;
        for (int64_t i = lo; i < hi; ++i) {
            const double x = obs[i], y = obs[stride + i], z = obs[2 * stride + i];
            m = std::fmax(m, x * x + y * y + z * z);
        }
        return m;
    };
    if (n_obs < (1 << 20)) return std::sqrt(range_max(0, n_obs));
    // time-ordered data: one pass over 24 B per sample, split over the host cores
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = (int)std::max(1u, std::min(16u, hw ? hw : 4u));
    std::vector<double> part(nt, 0.0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&, t] { part[t] = range_max(n_obs * t / nt, n_obs * (t + 1) / nt); });
    for (auto& th : pool) th.join();
    double m = 0.0;
    for (double v : part) m = std::fmax(m, v);
    return std::sqrt(m);
}

int check_args(const zodi_model_s* m, const zodi_eval_args* a, bool need_u = true) {
    if (!m) return fail(ZODI_ERR_INVALID, "model handle is NULL");
    if (!a) return fail(ZODI_ERR_INVALID, "eval args are NULL");
    if (a->n < 0) return fail(ZODI_ERR_INVALID, "n=%lld is negative", (long long)a->n);
    if (a->n == 0) return ZODI_OK;
    if (a->cyclic_block < 0 || (a->cyclic_block > 0 && (a->cyclic_parts < 1 || a->cyclic_rank < 0 ||
                                                        a->cyclic_rank >= a->cyclic_parts)))
        return fail(ZODI_ERR_INVALID, "bad block-cyclic layout");
    if (a->cyclic_block > 0 && a->memory != ZODI_MEM_DEVICE)
        return fail(ZODI_ERR_INVALID, "block-cyclic layout needs ZODI_MEM_DEVICE");
    if (a->n_peers < 0 || a->n_peers > ZODI_MAX_PEERS)
        return fail(ZODI_ERR_INVALID, "n_peers=%d outside [0, %d]", a->n_peers, ZODI_MAX_PEERS);
    if (a->n_peers > 0) {
        if (a->memory != ZODI_MEM_DEVICE)
            return fail(ZODI_ERR_INVALID, "peer output needs ZODI_MEM_DEVICE inputs");
        for (int p = 0; p < a->n_peers; ++p)
            if (!a->peer_out[p]) return fail(ZODI_ERR_INVALID, "peer_out[%d] is NULL", p);
        // the kernel stores to peer_offset + g(j), j < n, of every peer's map (row stride peer_stride):
        // the largest global index must lie inside a row whatever the layout and the number of rows
        int64_t last = a->n - 1;
        if (a->cyclic_block > 0 && a->cyclic_parts >= 1) {
            const int64_t lb = last / a->cyclic_block;
            last = (lb * a->cyclic_parts + a->cyclic_rank) * a->cyclic_block + (last - lb * a->cyclic_block);
        }
        if (a->peer_offset < 0 || a->peer_stride < a->peer_offset + last + 1)
            return fail(ZODI_ERR_INVALID, "peer slice [%lld, %lld] does not fit a map row of %lld elements",
                        (long long)a->peer_offset, (long long)(a->peer_offset + last), (long long)a->peer_stride);
    }
    if (a->ephemeris) {
        if (!a->obstime && (a->memory != ZODI_MEM_HOST || !a->ephemeris->d_times || a->ephemeris->n_times != a->n))
            return fail(ZODI_ERR_INVALID, "obstime is NULL and the ephemeris holds no staged times for n=%lld samples "
                                          "(zodi_ephemeris_stats with host memory stages them)", (long long)a->n);
        if (!a->outside_flags)
            return fail(ZODI_ERR_INVALID, "outside_flags must be supplied with an ephemeris (zodi_ephemeris_stats)");
        if (a->ephemeris->device != m->device) return fail(ZODI_ERR_INVALID, "ephemeris lives on another device");
        if ((need_u && !a->u) || (!a->out && a->n_peers == 0)) return fail(ZODI_ERR_INVALID, "u/out must be non-NULL");
        if (need_u && a->u_stride < a->n) return fail(ZODI_ERR_INVALID, "row strides must be >= row lengths");
        goto positions_checked;
    }
    if ((need_u && !a->u) || !a->obs || !a->earth || (!a->out && a->n_peers == 0))
        return fail(ZODI_ERR_INVALID, "u/obs/earth/out must be non-NULL");
    if (a->n_obs != 1 && a->n_obs != a->n)
        return fail(ZODI_ERR_INVALID, "n_obs=%lld must be 1 or n=%lld", (long long)a->n_obs,
                    (long long)a->n);
    if (a->n_earth != 1 && a->n_earth != a->n)
        return fail(ZODI_ERR_INVALID, "n_earth=%lld must be 1 or n=%lld", (long long)a->n_earth,
                    (long long)a->n);
    if ((need_u && a->u_stride < a->n) || a->obs_stride < a->n_obs || a->earth_stride < a->n_earth)
        return fail(ZODI_ERR_INVALID, "row strides must be >= row lengths");
positions_checked:
    if ((a->return_comps || m->mb_bands > 0) && a->n_peers == 0 && a->out_stride < a->n)
        return fail(ZODI_ERR_INVALID, "out_stride=%lld < n", (long long)a->out_stride);
    if (a->precision != ZODI_FP64 && a->precision != ZODI_FP32)
        return fail(ZODI_ERR_INVALID, "unknown precision %d", a->precision);
    if (a->out_dtype != ZODI_OUT_F64 && a->out_dtype != ZODI_OUT_F32)
        return fail(ZODI_ERR_INVALID, "unknown out_dtype %d", a->out_dtype);
    if (a->memory != ZODI_MEM_HOST && a->memory != ZODI_MEM_DEVICE)
        return fail(ZODI_ERR_INVALID, "unknown memory kind %d", a->memory);
    return ZODI_OK;
}

void set_ephemeris(LaunchArgs& la, const zodi_ephemeris_s* e, const double* obstime) {
    la.eph_coef = nullptr; la.eph_obs_coef = nullptr; la.obstime = nullptr;
    la.eph_nseg = 0; la.eph_t0 = 0.0; la.eph_dt = 1.0; la.eph_scale = 1.0;
    if (!e) return;
    la.eph_coef = e->d_earth; la.eph_obs_coef = e->d_obs; la.obstime = obstime;
    la.eph_nseg = e->n_knots - 1; la.eph_t0 = e->t0; la.eph_dt = e->dt; la.eph_scale = e->obs_scale;
}

void set_healpix(LaunchArgs& la, const zodi_healpix_args* hp, int64_t offset) {
    la.cyc_block = 0; la.cyc_parts = 1; la.cyc_rank = 0; la.cyc_shift = -1;
    set_ephemeris(la, nullptr, nullptr);
    la.hp_nside = 0; la.hp_start = 0; la.hp_rotate = 0; la.hp_nest = 0;
    la.lon = nullptr; la.lat = nullptr;
    for (int i = 0; i < 9; ++i) la.hp_rot[i] = 0.0;
    if (!hp) return;
    la.hp_nside = hp->nside;
    la.hp_start = hp->ipix_start + offset;
    la.hp_rotate = hp->has_rot != 0;
    la.hp_nest = hp->nest != 0;
    for (int i = 0; i < 9; ++i) la.hp_rot[i] = hp->rot[i];
}

// Directions from spherical coordinates (after set_healpix(la, nullptr, ..)): device pointers.
void set_lonlat(LaunchArgs& la, const zodi_lonlat_args* ll, const double* d_lon, const double* d_lat) {
    if (!ll) return;
    la.lon = d_lon; la.lat = d_lat;
    la.hp_rotate = ll->has_rot != 0;
    for (int i = 0; i < 9; ++i) la.hp_rot[i] = ll->rot[i];
}

// ---- host-memory path: chunked, 3-deep pipeline H2D | kernel | D2H on private streams ------
int ensure_workspace(zodi_model_s* m, int64_t chunk) {
    if (m->ws_chunk >= chunk) return ZODI_OK;
    free_workspace(m);
    for (Slot& s : m->slots) {
        CU_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU_CHECK(cudaEventCreate(&s.k0));
        CU_CHECK(cudaEventCreate(&s.k1));
        CU_CHECK(cudaMalloc((void**)&s.d_in, (size_t)chunk * 9 * sizeof(double)));
        CU_CHECK(cudaMalloc(&s.d_out, (size_t)chunk * ZODI_MAX_COMPS * sizeof(double)));
    }
    m->ws_chunk = chunk;
    return ZODI_OK;
}

int evaluate_host(zodi_model_s* m, const zodi_eval_args* a, uint32_t mask, const zodi_healpix_args* hp,
                  const zodi_lonlat_args* ll) {
    std::lock_guard<std::mutex> lock(m->ws_mutex);
    const int64_t n = a->n;
    // 1 Mi-line chunks for large inputs; mid-size inputs (one device's shard of a map split over several
    // GPUs) are cut into >= 12 chunks so that the first kernel does not wait for a long first upload and
    // the last download is short (pipeline fill / drain)
    int64_t chunk = 1 << 20;
    if (n < 12 * chunk) chunk = std::max<int64_t>(1 << 17, ((n + 11) / 12 + 0xFFFF) & ~(int64_t)0xFFFF);
    if (n < chunk) chunk = n;
    int rc = ensure_workspace(m, chunk);
    if (rc) return rc;
    chunk = std::min<int64_t>(m->ws_chunk, std::max<int64_t>(chunk, 1));
    const size_t osz = a->out_dtype == ZODI_OUT_F32 ? sizeof(float) : sizeof(double);
    const int out_rows = m->mb_bands > 0 ? m->mb_bands : (a->return_comps ? m->desc.n_comps : 1);
    const bool rows_out = m->mb_bands > 0 || a->return_comps;
    const bool obs_ps = !a->ephemeris && a->n_obs == n && n > 1;
    const bool earth_ps = !a->ephemeris && a->n_earth == n && n > 1;
    for (Slot& s : m->slots) s.used = false;

    // single observer / Earth: upload once into slot-independent spots (tail of slot 0 input)
    // -> simpler: keep them at rows 3..5 / 6..8 of each slot, column 0.
    int64_t done = 0;
    int idx = 0;
    while (done < n) {
        const int64_t cn = std::min(chunk, n - done);
        Slot& s = m->slots[idx % kSlots];
        double* d_u = s.d_in;
        double* d_obs = s.d_in + 3 * m->ws_chunk;
        double* d_earth = s.d_in + 6 * m->ws_chunk;
        const size_t pitch = (size_t)m->ws_chunk * sizeof(double);
        if (ll) {  // 16 B per line of sight instead of 24: rows 0 / 1 of the direction slot
            CU_CHECK(cudaMemcpyAsync(d_u, ll->lon + done, (size_t)cn * sizeof(double), cudaMemcpyHostToDevice, s.stream));
            CU_CHECK(cudaMemcpyAsync(d_u + m->ws_chunk, ll->lat + done, (size_t)cn * sizeof(double),
                                     cudaMemcpyHostToDevice, s.stream));
        } else if (!hp)
            CU_CHECK(cudaMemcpy2DAsync(d_u, pitch, a->u + done, (size_t)a->u_stride * sizeof(double),
                                       (size_t)cn * sizeof(double), 3, cudaMemcpyHostToDevice, s.stream));
        const double* d_time = d_obs;
        if (a->ephemeris && !a->obstime)  // times already on the device (staged by zodi_ephemeris_stats)
            d_time = a->ephemeris->d_times + done;
        else if (a->ephemeris)  // obstime only (8 B per sample instead of 48 B of positions)
            CU_CHECK(cudaMemcpyAsync(d_obs, a->obstime + done, (size_t)cn * sizeof(double),
                                     cudaMemcpyHostToDevice, s.stream));
        else
        CU_CHECK(cudaMemcpy2DAsync(d_obs, pitch, a->obs + (obs_ps ? done : 0),
                                   (size_t)a->obs_stride * sizeof(double),
                                   (size_t)(obs_ps ? cn : 1) * sizeof(double), 3,
                                   cudaMemcpyHostToDevice, s.stream));
        if (m->m64.has_feature && !a->ephemeris)
            CU_CHECK(cudaMemcpy2DAsync(d_earth, pitch, a->earth + (earth_ps ? done : 0),
                                       (size_t)a->earth_stride * sizeof(double),
                                       (size_t)(earth_ps ? cn : 1) * sizeof(double), 3,
                                       cudaMemcpyHostToDevice, s.stream));
        LaunchArgs la;
        la.n = cn;
        la.shape_n = n;
        la.u = d_u; la.u_stride = m->ws_chunk;
        la.obs = d_obs; la.obs_stride = m->ws_chunk; la.obs_per_sample = obs_ps;
        la.earth = d_earth; la.earth_stride = m->ws_chunk; la.earth_per_sample = earth_ps;
        la.outside_mask = mask;
        la.return_comps = a->return_comps;
        la.out_f32 = a->out_dtype == ZODI_OUT_F32;
        la.out = s.d_out; la.out_stride = m->ws_chunk;
        la.n_peers = 0; la.peer_offset = 0; la.peer_stride = 0;
        set_healpix(la, hp, done);
        set_lonlat(la, ll, d_u, d_u + m->ws_chunk);
        if (a->ephemeris) set_ephemeris(la, a->ephemeris, d_time);
        if (!s.used) CU_CHECK(cudaEventRecord(s.k0, s.stream));
        CU_CHECK(launch_eval(m, la, a->precision, s.stream));
        CU_CHECK(cudaEventRecord(s.k1, s.stream));
        s.used = true;
        CU_CHECK(cudaMemcpy2DAsync((char*)a->out + (size_t)done * osz,
                                   (size_t)(rows_out ? a->out_stride : n) * osz, s.d_out,
                                   (size_t)m->ws_chunk * osz, (size_t)cn * osz, out_rows,
                                   cudaMemcpyDeviceToHost, s.stream));
        done += cn;
        ++idx;
    }
    double ms_total = 0.0;
    for (Slot& s : m->slots) {
        if (!s.used) continue;
        CU_CHECK(cudaStreamSynchronize(s.stream));
    }
    // Kernel-only time: only exact when a single chunk ran per slot; otherwise k0..k1 spans the
    // slot's whole activity, so report it for single-chunk calls and 0 otherwise.
    if (idx <= kSlots) {
        for (Slot& s : m->slots)
            if (s.used) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, s.k0, s.k1) == cudaSuccess) ms_total += ms;
            }
    }
    m->last_kernel_ms = ms_total;
    return ZODI_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int zodi_abi_version(void) { return ZODI_ABI_VERSION; }

const char* zodi_last_error(void) { return g_last_error.c_str(); }

int zodi_device_count(int* count) {
    if (!count) return fail(ZODI_ERR_INVALID, "count is NULL");
    *count = 0;
    CU_CHECK(cudaGetDeviceCount(count));
    return ZODI_OK;
}

int zodi_model_create(const zodi_model_desc* desc, int device, zodi_model_t* out) {
    if (!out) return fail(ZODI_ERR_INVALID, "out handle pointer is NULL");
    *out = nullptr;
    int rc = validate_desc(desc);
    if (rc) return rc;
    int count = 0;
    CU_CHECK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
        return fail(ZODI_ERR_CUDA, "device %d not available (%d CUDA devices)", device, count);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    zodi_model_s* m = new (std::nothrow) zodi_model_s();
    if (!m) return fail(ZODI_ERR_NOMEM, "out of host memory");
    m->device = device;
    rc = upload_model(m, desc);
    if (rc) { zodi_model_destroy(m); return rc; }
    *out = m;
    return ZODI_OK;
}

int zodi_model_update(zodi_model_t m, const zodi_model_desc* desc) {
    if (!m) return fail(ZODI_ERR_INVALID, "model handle is NULL");
    int rc = validate_desc(desc);
    if (rc) return rc;
    DeviceGuard guard(m->device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", m->device);
    CU_CHECK(cudaDeviceSynchronize());
    return upload_model(m, desc);
}

int zodi_model_destroy(zodi_model_t m) {
    if (!m) return ZODI_OK;
    {
        DeviceGuard guard(m->device);
        cudaDeviceSynchronize();
        free_workspace(m);
        cudaFree(m->d_table64); cudaFree(m->d_nodes64);
        cudaFree(m->d_table32); cudaFree(m->d_nodes32);
        cudaFree(m->d_scratch);
        cudaFree(m->d_tiles);
        cudaFree(m->d_mbtab64); cudaFree(m->d_mbtab32); cudaFree(m->d_mbrows32);
    }
    delete m;
    return ZODI_OK;
}

int zodi_flags_from_radius(zodi_model_t m, double r_max, uint8_t* flags) {
    if (!m || !flags) return fail(ZODI_ERR_INVALID, "NULL argument");
    flags_from_r(m, r_max, flags);
    return ZODI_OK;
}

int zodi_max_observer_radius(zodi_model_t m, const double* obs, int64_t n_obs, int64_t obs_stride,
                             int32_t memory, void* stream, double* r_max) {
    if (!m || !obs || !r_max || n_obs < 1 || obs_stride < n_obs)
        return fail(ZODI_ERR_INVALID, "bad argument");
    if (memory == ZODI_MEM_HOST) {
        *r_max = max_r_host(obs, n_obs, obs_stride);
        return ZODI_OK;
    }
    DeviceGuard guard(m->device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", m->device);
    return max_r_device(m, obs, n_obs, obs_stride, (cudaStream_t)stream, r_max);
}

static int evaluate_impl(zodi_model_t m, const zodi_eval_args* a, const zodi_healpix_args* hp,
                         const zodi_lonlat_args* ll = nullptr) {
    int rc = check_args(m, a, hp == nullptr && ll == nullptr);
    if (rc) return rc;
    if (ll && a->n > 0 && (!ll->lon || !ll->lat)) return fail(ZODI_ERR_INVALID, "lon/lat must be non-NULL");
    if (a->n == 0) return ZODI_OK;
    DeviceGuard guard(m->device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", m->device);

    uint8_t flags[2 * ZODI_MAX_COMPS];
    if (a->outside_flags) {
        std::memcpy(flags, a->outside_flags, 2 * (size_t)m->desc.n_comps);
    } else {
        double r_max = 0.0;
        if (a->memory == ZODI_MEM_HOST) {
            r_max = max_r_host(a->obs, a->n_obs, a->obs_stride);
        } else {
            rc = max_r_device(m, a->obs, a->n_obs, a->obs_stride, (cudaStream_t)a->stream, &r_max);
            if (rc) return rc;
        }
        flags_from_r(m, r_max, flags);
    }
    const uint32_t mask = flags_to_mask(flags, m->desc.n_comps);

    if (a->memory == ZODI_MEM_HOST) return evaluate_host(m, a, mask, hp, ll);

    LaunchArgs la;
    la.n = a->n;
    la.shape_n = a->n;
    la.u = a->u; la.u_stride = a->u_stride;
    la.obs = a->obs; la.obs_stride = a->obs_stride; la.obs_per_sample = (a->n_obs == a->n && a->n > 1);
    la.earth = a->earth; la.earth_stride = a->earth_stride;
    la.earth_per_sample = (a->n_earth == a->n && a->n > 1);
    la.outside_mask = mask;
    la.return_comps = a->return_comps;
    la.out_f32 = a->out_dtype == ZODI_OUT_F32;
    la.out = a->out; la.out_stride = a->out_stride;
    la.n_peers = a->n_peers; la.peer_offset = a->peer_offset; la.peer_stride = a->peer_stride;
    for (int p = 0; p < ZODI_MAX_PEERS; ++p) la.peer_out[p] = p < a->n_peers ? a->peer_out[p] : nullptr;
    set_healpix(la, hp, 0);
    if (ll) set_lonlat(la, ll, ll->lon, ll->lat);
    if (a->ephemeris) set_ephemeris(la, a->ephemeris, a->obstime);
    la.cyc_block = a->cyclic_block; la.cyc_parts = a->cyclic_parts; la.cyc_rank = a->cyclic_rank;
    la.cyc_shift = -1;
    if (la.cyc_block > 0 && (la.cyc_block & (la.cyc_block - 1)) == 0)
        for (int sh = 0; sh < 62; ++sh) if ((int64_t(1) << sh) == la.cyc_block) la.cyc_shift = sh;
    CU_CHECK(launch_eval(m, la, a->precision, (cudaStream_t)a->stream));
    return ZODI_OK;
}

int zodi_evaluate(zodi_model_t m, const zodi_eval_args* a) { return evaluate_impl(m, a, nullptr); }

// ---- on-device ephemeris ---------------------------------------------------------------------
static void spline_to_device_layout(int64_t n_knots, double t0, double dt, const double* knots,
                                    std::vector<double>& out) {
    std::vector<double> x(n_knots), c0, c1, c2, c3;
    for (int64_t k = 0; k < n_knots; ++k) x[k] = t0 + (double)k * dt;  // np.arange(t0, t1 + dt, dt)
    const int64_t nseg = n_knots - 1;
    out.assign((size_t)nseg * 12, 0.0);
    for (int axis = 0; axis < 3; ++axis) {
        cubic_spline_not_a_knot(x, knots + (size_t)axis * n_knots, c0, c1, c2, c3);
        for (int64_t i = 0; i < nseg; ++i) {
            double* o = &out[(size_t)i * 12 + axis * 4];
            o[0] = c0[i]; o[1] = c1[i]; o[2] = c2[i]; o[3] = c3[i];
        }
    }
}

int zodi_ephemeris_create(int device, const zodi_ephemeris_desc* d, zodi_ephemeris_t* out) {
    if (!out) return fail(ZODI_ERR_INVALID, "out handle pointer is NULL");
    *out = nullptr;
    if (!d || !d->earth_knots) return fail(ZODI_ERR_INVALID, "ephemeris descriptor / earth_knots is NULL");
    if (d->n_knots < 4) return fail(ZODI_ERR_INVALID, "n_knots=%lld < 4", (long long)d->n_knots);
    if (!(d->dt > 0.0)) return fail(ZODI_ERR_INVALID, "knot spacing dt must be positive");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    zodi_ephemeris_s* e = new (std::nothrow) zodi_ephemeris_s();
    if (!e) return fail(ZODI_ERR_NOMEM, "out of host memory");
    e->device = device; e->n_knots = d->n_knots; e->t0 = d->t0; e->dt = d->dt; e->obs_scale = d->obs_scale;
    spline_to_device_layout(d->n_knots, d->t0, d->dt, d->earth_knots, e->earth_c);
    const size_t bytes = e->earth_c.size() * sizeof(double);
    cudaError_t err = cudaMalloc((void**)&e->d_earth, bytes);
    if (err == cudaSuccess) err = cudaMemcpy(e->d_earth, e->earth_c.data(), bytes, cudaMemcpyHostToDevice);
    if (err == cudaSuccess && d->obs_knots) {
        std::vector<double> oc;
        spline_to_device_layout(d->n_knots, d->t0, d->dt, d->obs_knots, oc);
        err = cudaMalloc((void**)&e->d_obs, bytes);
        if (err == cudaSuccess) err = cudaMemcpy(e->d_obs, oc.data(), bytes, cudaMemcpyHostToDevice);
    }
    if (err == cudaSuccess) err = cudaMalloc((void**)&e->d_stats, 3 * sizeof(double));
    if (err != cudaSuccess) {
        zodi_ephemeris_destroy(e);
        return fail(ZODI_ERR_CUDA, "ephemeris upload failed: %s", cudaGetErrorString(err));
    }
    *out = e;
    return ZODI_OK;
}

int zodi_ephemeris_set_obs_scale(zodi_ephemeris_t e, double obs_scale) {
    if (!e) return fail(ZODI_ERR_INVALID, "ephemeris handle is NULL");
    e->obs_scale = obs_scale;
    return ZODI_OK;
}

int zodi_ephemeris_destroy(zodi_ephemeris_t e) {
    if (!e) return ZODI_OK;
    {
        DeviceGuard guard(e->device);
        cudaFree(e->d_earth); cudaFree(e->d_obs); cudaFree(e->d_stats); cudaFree(e->d_times);
    }
    delete e;
    return ZODI_OK;
}

int zodi_ephemeris_release_times(zodi_ephemeris_t e) {
    if (!e) return ZODI_OK;
    DeviceGuard guard(e->device);
    cudaFree(e->d_times);
    e->d_times = nullptr;
    e->n_times = 0;
    e->cap_times = 0;
    return ZODI_OK;
}

int zodi_ephemeris_coefficients(zodi_ephemeris_t e, double* c) {
    if (!e || !c) return fail(ZODI_ERR_INVALID, "NULL argument");
    const int64_t nseg = e->n_knots - 1;
    for (int k = 0; k < 4; ++k)
        for (int64_t i = 0; i < nseg; ++i)
            for (int axis = 0; axis < 3; ++axis) c[(k * nseg + i) * 3 + axis] = e->earth_c[(size_t)i * 12 + axis * 4 + k];
    return ZODI_OK;
}

// stage `t` on the device when it lives on the host; returns the device pointer to use
static int stage_times(const double* t, int64_t n, int32_t memory, cudaStream_t st, double** d_t, bool* owned) {
    *owned = false;
    *d_t = const_cast<double*>(t);
    if (memory == ZODI_MEM_DEVICE) return ZODI_OK;
    CU_CHECK(cudaMalloc((void**)d_t, (size_t)n * sizeof(double)));
    *owned = true;
    CU_CHECK(cudaMemcpyAsync(*d_t, t, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    return ZODI_OK;
}

int zodi_ephemeris_positions(zodi_ephemeris_t e, const double* t, int64_t n, int32_t memory, double* earth_out,
                             double* obs_out, void* stream) {
    if (!e || !t || n < 0) return fail(ZODI_ERR_INVALID, "bad argument");
    if (n == 0) return ZODI_OK;
    DeviceGuard guard(e->device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", e->device);
    cudaStream_t st = memory == ZODI_MEM_DEVICE ? (cudaStream_t)stream : nullptr;
    double* d_t; bool owned;
    int rc = stage_times(t, n, memory, st, &d_t, &owned);
    if (rc) return rc;
    double *d_e = earth_out, *d_o = obs_out;
    if (memory == ZODI_MEM_HOST) {
        if (earth_out) CU_CHECK(cudaMalloc((void**)&d_e, (size_t)3 * n * sizeof(double)));
        if (obs_out) CU_CHECK(cudaMalloc((void**)&d_o, (size_t)3 * n * sizeof(double)));
    }
    LaunchArgs la;
    std::memset(&la, 0, sizeof(la));
    la.n = n;
    set_ephemeris(la, e, d_t);
    zodi_ephemeris_positions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(la, d_e, d_o);
    g_launches.fetch_add(1);
    CU_CHECK(cudaGetLastError());
    if (memory == ZODI_MEM_HOST) {
        if (earth_out) CU_CHECK(cudaMemcpy(earth_out, d_e, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost));
        if (obs_out) CU_CHECK(cudaMemcpy(obs_out, d_o, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost));
        if (earth_out) cudaFree(d_e);
        if (obs_out) cudaFree(d_o);
    }
    if (owned) { cudaStreamSynchronize(st); cudaFree(d_t); }
    return ZODI_OK;
}

int zodi_ephemeris_stats(zodi_ephemeris_t e, const double* t, int64_t n, int32_t memory, void* stream,
                         double* stats) {
    if (!e || !t || !stats || n < 1) return fail(ZODI_ERR_INVALID, "bad argument");
    DeviceGuard guard(e->device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", e->device);
    cudaStream_t st = memory == ZODI_MEM_DEVICE ? (cudaStream_t)stream : nullptr;
    const double* d_t = t;
    if (memory == ZODI_MEM_HOST) {
        // keep the device copy: a following zodi_evaluate with obstime = NULL integrates from it
        // instead of uploading the times a second time
        e->n_times = 0;
        if (e->cap_times < n) {
            cudaFree(e->d_times);
            e->d_times = nullptr; e->cap_times = 0;
            CU_CHECK(cudaMalloc((void**)&e->d_times, (size_t)n * sizeof(double)));
            e->cap_times = n;
        }
        e->n_times = n;
        CU_CHECK(cudaMemcpyAsync(e->d_times, t, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
        d_t = e->d_times;
    }
    CU_CHECK(cudaMemsetAsync(e->d_stats, 0, 3 * sizeof(double), st));
    LaunchArgs la;
    std::memset(&la, 0, sizeof(la));
    la.n = n;
    set_ephemeris(la, e, d_t);
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
    zodi_ephemeris_stats_kernel<<<grid, 256, 0, st>>>(la, e->d_stats, reinterpret_cast<unsigned long long*>(e->d_stats + 1));
    g_launches.fetch_add(1);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaMemcpyAsync(stats, e->d_stats, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    if (!e->d_obs) stats[2] = e->obs_scale * e->obs_scale * stats[1];  // observer = scale * Earth
    return ZODI_OK;
}

static int check_healpix(int64_t nside, int64_t ipix_start, int64_t n, int nest) {
    if (nest && (nside & (nside - 1)) != 0)
        return fail(ZODI_ERR_INVALID, "NESTED ordering needs nside to be a power of two (got %lld)", (long long)nside);
    if (nside < 1 || nside > (1ll << 29)) return fail(ZODI_ERR_INVALID, "nside=%lld out of range", (long long)nside);
    const int64_t npix = 12 * nside * nside;
    if (ipix_start < 0 || n < 0 || ipix_start + n > npix)
        return fail(ZODI_ERR_INVALID, "pixel range [%lld, %lld) outside [0, %lld)", (long long)ipix_start,
                    (long long)(ipix_start + n), (long long)npix);
    return ZODI_OK;
}

int zodi_evaluate_healpix(zodi_model_t m, const zodi_healpix_args* hp) {
    if (!hp) return fail(ZODI_ERR_INVALID, "healpix args are NULL");
    int64_t span = hp->base.n;  // pixels [ipix_start, ipix_start + span) must exist
    if (hp->base.cyclic_block > 0 && hp->base.n > 0 && hp->base.cyclic_parts >= 1) {
        const int64_t j = hp->base.n - 1, lb = j / hp->base.cyclic_block;
        span = (lb * hp->base.cyclic_parts + hp->base.cyclic_rank) * hp->base.cyclic_block +
               (j - lb * hp->base.cyclic_block) + 1;
    }
    int rc = check_healpix(hp->nside, hp->ipix_start, span, hp->nest);
    if (rc) return rc;
    if (hp->base.n_obs != 1 && hp->base.n_obs != hp->base.n)
        return fail(ZODI_ERR_INVALID, "n_obs must be 1 or n");
    return evaluate_impl(m, &hp->base, hp);
}

int zodi_healpix_vectors(int device, int64_t nside, int32_t nest, int64_t ipix_start, int64_t n,
                         const double* rot, double* out, int64_t out_stride, int32_t memory, void* stream) {
    int rc = check_healpix(nside, ipix_start, n, nest);
    if (rc) return rc;
    if (!out || out_stride < n) return fail(ZODI_ERR_INVALID, "bad output buffer");
    if (n == 0) return ZODI_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    LaunchArgs la;
    std::memset(&la, 0, sizeof(la));
    la.n = n;
    zodi_healpix_args hp;
    std::memset(&hp, 0, sizeof(hp));
    hp.nside = nside; hp.ipix_start = ipix_start; hp.has_rot = rot != nullptr; hp.nest = nest;
    if (rot) std::memcpy(hp.rot, rot, sizeof(hp.rot));
    set_healpix(la, &hp, 0);
    cudaStream_t st = (cudaStream_t)stream;
    double* d_out = out;
    if (memory == ZODI_MEM_HOST) {
        st = nullptr;
        CU_CHECK(cudaMalloc((void**)&d_out, (size_t)3 * n * sizeof(double)));
    }
    const int64_t ld = memory == ZODI_MEM_HOST ? n : out_stride;
    zodi_healpix_vectors_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(la, d_out, ld);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && memory == ZODI_MEM_HOST)
        e = cudaMemcpy2D(out, (size_t)out_stride * sizeof(double), d_out, (size_t)n * sizeof(double),
                         (size_t)n * sizeof(double), 3, cudaMemcpyDeviceToHost);
    if (memory == ZODI_MEM_HOST) cudaFree(d_out);
    if (e != cudaSuccess) return fail(ZODI_ERR_CUDA, "healpix vectors failed: %s", cudaGetErrorString(e));
    return ZODI_OK;
}

// ---- directions from spherical sky coordinates ---------------------------------------------------
int zodi_evaluate_lonlat(zodi_model_t m, const zodi_lonlat_args* ll) {
    if (!ll) return fail(ZODI_ERR_INVALID, "lonlat args are NULL");
    return evaluate_impl(m, &ll->base, nullptr, ll);
}

int zodi_lonlat_vectors(int device, const double* lon, const double* lat, int64_t n, const double* rot,
                        double* out, int64_t out_stride, int32_t memory, void* stream) {
    if (n < 0 || !out || out_stride < n || (n > 0 && (!lon || !lat))) return fail(ZODI_ERR_INVALID, "bad argument");
    if (memory != ZODI_MEM_HOST && memory != ZODI_MEM_DEVICE) return fail(ZODI_ERR_INVALID, "unknown memory kind");
    if (n == 0) return ZODI_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    LaunchArgs la;
    std::memset(&la, 0, sizeof(la));
    la.n = n;
    set_healpix(la, nullptr, 0);
    zodi_lonlat_args ll;
    std::memset(&ll, 0, sizeof(ll));
    ll.has_rot = rot != nullptr;
    if (rot) std::memcpy(ll.rot, rot, sizeof(ll.rot));
    cudaStream_t st = (cudaStream_t)stream;
    double* d_buf = nullptr;  // host memory: [lon | lat | out(3, n)] staged in one allocation
    double* d_out = out;
    const double *d_lon = lon, *d_lat = lat;
    if (memory == ZODI_MEM_HOST) {
        st = nullptr;
        CU_CHECK(cudaMalloc((void**)&d_buf, (size_t)5 * n * sizeof(double)));
        cudaError_t e = cudaMemcpy(d_buf, lon, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_buf + n, lat, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(d_buf); return fail(ZODI_ERR_CUDA, "lonlat upload failed: %s", cudaGetErrorString(e)); }
        d_lon = d_buf; d_lat = d_buf + n; d_out = d_buf + 2 * n;
    }
    set_lonlat(la, &ll, d_lon, d_lat);
    const int64_t ld = memory == ZODI_MEM_HOST ? n : out_stride;
    zodi_healpix_vectors_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(la, d_out, ld);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && memory == ZODI_MEM_HOST)
        e = cudaMemcpy2D(out, (size_t)out_stride * sizeof(double), d_out, (size_t)n * sizeof(double),
                         (size_t)n * sizeof(double), 3, cudaMemcpyDeviceToHost);
    if (d_buf) cudaFree(d_buf);
    if (e != cudaSuccess) return fail(ZODI_ERR_CUDA, "lonlat vectors failed: %s", cudaGetErrorString(e));
    return ZODI_OK;
}

// ---- multi-band ------------------------------------------------------------------------------
struct zodi_multiband_s { zodi_model_s* model; };

static bool same_geometry(const zodi_model_desc& a, const zodi_model_desc& b) {
    if (a.kind != b.kind || a.n_comps != b.n_comps || a.n_nodes != b.n_nodes || a.n_temps != b.n_temps) return false;
    if (a.T_0 != b.T_0 || a.delta != b.delta) return false;
    for (int i = 0; i < a.n_temps; ++i) if (a.temps[i] != b.temps[i]) return false;
    for (int i = 0; i < a.n_nodes; ++i) if (a.nodes[i] != b.nodes[i] || a.weights[i] != b.weights[i]) return false;
    for (int c = 0; c < a.n_comps; ++c) {
        const zodi_component_desc &p = a.comps[c], &q = b.comps[c];
        if (p.type != q.type || p.cutoff_inner != q.cutoff_inner || p.cutoff_outer != q.cutoff_outer) return false;
        if (std::memcmp(p.x0, q.x0, sizeof(p.x0)) || std::memcmp(p.shape, q.shape, sizeof(p.shape))) return false;
        if (p.sin_Omega != q.sin_Omega || p.cos_Omega != q.cos_Omega || p.sin_i != q.sin_i || p.cos_i != q.cos_i) return false;
    }
    return true;
}

int zodi_multiband_create(const zodi_model_desc* descs, int32_t n_bands, int device, zodi_multiband_t* out) {
    if (!out) return fail(ZODI_ERR_INVALID, "out handle pointer is NULL");
    *out = nullptr;
    if (!descs || n_bands < 1 || n_bands > ZODI_MAX_BANDS)
        return fail(ZODI_ERR_INVALID, "n_bands=%d outside [1, %d]", n_bands, ZODI_MAX_BANDS);
    static_assert(ZODI_MAX_BANDS == kMaxBands, "band capacity");
    for (int b = 0; b < n_bands; ++b) {
        int rc = validate_desc(&descs[b]);
        if (rc) return rc;
        if (!same_geometry(descs[0], descs[b]))
            return fail(ZODI_ERR_INVALID, "band %d differs from band 0 in more than its spectral parameters", b);
    }
    zodi_model_t m = nullptr;
    int rc = zodi_model_create(&descs[0], device, &m);
    if (rc) return rc;
    if (m->kelsall_ok && descs[0].n_temps > kMultiBandMaxTemps) {
        zodi_model_destroy(m);
        return fail(ZODI_ERR_UNSUPPORTED, "multi-band evaluation supports at most %d table knots", kMultiBandMaxTemps);
    }
    if (!m->kelsall_ok) {
        zodi_model_destroy(m);
        return fail(ZODI_ERR_UNSUPPORTED, "multi-band evaluation needs a Kelsall-family model layout");
    }
    DeviceGuard guard(device);
    std::vector<Pair<double>> t64;
    std::vector<Pair<float>> t32;
    const int bad = build_multiband_model(descs, n_bands, m->mb64, m->mb32, t64, t32);
    if (bad >= 0) {
        zodi_model_destroy(m);
        return fail(ZODI_ERR_UNSUPPORTED, "band %d is not eligible for the fused kernel", bad);
    }
    cudaError_t e = cudaMalloc((void**)&m->d_mbtab64, t64.size() * sizeof(Pair<double>));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_mbtab64, t64.data(), t64.size() * sizeof(Pair<double>), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_mbtab32, t32.size() * sizeof(Pair<float>));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_mbtab32, t32.data(), t32.size() * sizeof(Pair<float>), cudaMemcpyHostToDevice);
    // knot-major copy for the packed kernel (MbRows<NB>::kRow pairs per row, bands past n_bands zero)
    const int nt = descs[0].n_temps, row = m->mb32.n_bands_padded + 2;
    std::vector<Pair<float>> rows((size_t)nt * row, Pair<float>{0.f, 0.f});
    for (int knot = 0; knot < nt; ++knot)
        for (int b = 0; b < n_bands; ++b) rows[(size_t)knot * row + b] = t32[(size_t)b * nt + knot];
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_mbrows32, rows.size() * sizeof(Pair<float>));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_mbrows32, rows.data(), rows.size() * sizeof(Pair<float>), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        zodi_model_destroy(m);
        return fail(ZODI_ERR_CUDA, "multi-band table upload failed: %s", cudaGetErrorString(e));
    }
    m->mb_bands = n_bands;
    zodi_multiband_s* h = new (std::nothrow) zodi_multiband_s{m};
    if (!h) { zodi_model_destroy(m); return fail(ZODI_ERR_NOMEM, "out of host memory"); }
    *out = h;
    return ZODI_OK;
}

int zodi_multiband_evaluate(zodi_multiband_t mb, const zodi_eval_args* a) {
    if (!mb) return fail(ZODI_ERR_INVALID, "multi-band handle is NULL");
    return evaluate_impl(mb->model, a, nullptr);
}

int zodi_multiband_evaluate_healpix(zodi_multiband_t mb, const zodi_healpix_args* hp) {
    if (!mb) return fail(ZODI_ERR_INVALID, "multi-band handle is NULL");
    return zodi_evaluate_healpix(mb->model, hp);
}

int zodi_multiband_evaluate_lonlat(zodi_multiband_t mb, const zodi_lonlat_args* ll) {
    if (!mb) return fail(ZODI_ERR_INVALID, "multiband handle is NULL");
    return zodi_evaluate_lonlat(mb->model, ll);
}

int zodi_multiband_destroy(zodi_multiband_t mb) {
    if (!mb) return ZODI_OK;
    zodi_model_destroy(mb->model);
    delete mb;
    return ZODI_OK;
}

int zodi_number_density(zodi_model_t m, const double* xyz, int64_t n, int64_t xyz_stride, const double* earth,
                        double* out, int64_t out_stride, int32_t memory, void* stream) {
    if (!m || !xyz || !earth || !out || n < 0 || xyz_stride < n || out_stride < n)
        return fail(ZODI_ERR_INVALID, "bad argument");
    if (n == 0) return ZODI_OK;
    DeviceGuard guard(m->device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", m->device);
    const int nc = m->desc.n_comps;
    const double* d_xyz = xyz;
    double* d_out = out;
    int64_t ld_in = xyz_stride, ld_out = out_stride;
    cudaStream_t st = (cudaStream_t)stream;
    if (memory == ZODI_MEM_HOST) {
        st = nullptr;
        double* tmp = nullptr;
        CU_CHECK(cudaMalloc((void**)&tmp, (size_t)(3 + nc) * n * sizeof(double)));
        d_xyz = tmp; d_out = tmp + 3 * n; ld_in = n; ld_out = n;
        cudaError_t e = cudaMemcpy2D(tmp, (size_t)n * sizeof(double), xyz, (size_t)xyz_stride * sizeof(double),
                                     (size_t)n * sizeof(double), 3, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(tmp); return fail(ZODI_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e)); }
    }
    // the density kernel reads the raw-to-device-form constants but needs amplitude-free semantics:
    // density<Real>() returns the full reference density (amplitudes included) for the generic model
    zodi_number_density_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->m64, d_xyz, n, ld_in, earth[0],
                                                                          earth[1], d_out, ld_out);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && memory == ZODI_MEM_HOST)
        e = cudaMemcpy2D(out, (size_t)out_stride * sizeof(double), d_out, (size_t)n * sizeof(double),
                         (size_t)n * sizeof(double), nc, cudaMemcpyDeviceToHost);
    if (memory == ZODI_MEM_HOST) cudaFree(const_cast<double*>(d_xyz));
    if (e != cudaSuccess) return fail(ZODI_ERR_CUDA, "number density failed: %s", cudaGetErrorString(e));
    return ZODI_OK;
}

const char* zodi_model_kernel_name(zodi_model_t m) {
    if (!m) return "";
    if (m->force_generic) return "zodi_los_generic_kernel";
    return m->kelsall_ok ? "zodi_los_kelsall_kernel" : (m->rrm_ok ? "zodi_los_rrm_kernel" : "zodi_los_generic_kernel");
}

int zodi_peer_buffer_alloc(int device, int64_t bytes, void** ptr, uint8_t handle[ZODI_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == ZODI_IPC_HANDLE_BYTES, "IPC handle size");
    if (!ptr || !handle || bytes <= 0) return fail(ZODI_ERR_INVALID, "bad argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    *ptr = nullptr;
    CU_CHECK(cudaMalloc(ptr, (size_t)bytes));
    CU_CHECK(cudaMemset(*ptr, 0, (size_t)bytes));  // flag arrays rely on it; maps are overwritten anyway
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) {
        cudaFree(*ptr);
        *ptr = nullptr;
        return fail(ZODI_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    std::memcpy(handle, &h, sizeof(h));
    return ZODI_OK;
}

int zodi_peer_buffer_open(int device, const uint8_t handle[ZODI_IPC_HANDLE_BYTES], void** ptr) {
    if (!ptr || !handle) return fail(ZODI_ERR_INVALID, "bad argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    CU_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return ZODI_OK;
}

int zodi_peer_buffer_close(int device, void* ptr) {
    if (!ptr) return ZODI_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    CU_CHECK(cudaIpcCloseMemHandle(ptr));
    return ZODI_OK;
}

int zodi_peer_buffer_free(int device, void* ptr) {
    if (!ptr) return ZODI_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    CU_CHECK(cudaFree(ptr));
    return ZODI_OK;
}

int zodi_peer_rendezvous(int device, void* const* peer_flags, int32_t n_peers, int32_t rank, uint32_t epoch,
                         void* stream) {
    if (!peer_flags || n_peers < 1 || n_peers > ZODI_MAX_PEERS || rank < 0 || rank >= n_peers)
        return fail(ZODI_ERR_INVALID, "bad rendezvous arguments");
    PeerFlagPtrs f;
    for (int p = 0; p < ZODI_MAX_PEERS; ++p) {
        f.p[p] = p < n_peers ? static_cast<uint32_t*>(peer_flags[p]) : nullptr;
        if (p < n_peers && !f.p[p]) return fail(ZODI_ERR_INVALID, "peer_flags[%d] is NULL", p);
    }
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    zodi_peer_rendezvous_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, n_peers, rank, epoch);
    g_launches.fetch_add(1);
    CU_CHECK(cudaGetLastError());
    return ZODI_OK;
}

const char* zodi_model_kernel_for(zodi_model_t m, int64_t n, int32_t precision) {
    if (!m) return "";
    if (m->mb_bands > 0)
        return (precision == ZODI_FP32 && multiband_takes_packed(m, n)) ? "zodi_los_multiband_x2_kernel"
                                                                         : "zodi_los_multiband_kernel";
    if (m->rrm_ok && !m->force_generic)
        return (precision == ZODI_FP32 && rrm_takes_packed(m, n)) ? "zodi_los_rrm_x2_kernel" : "zodi_los_rrm_kernel";
    if (!(m->kelsall_ok && !m->force_generic)) return "zodi_los_generic_kernel";
    if (precision == ZODI_FP32 && !m->no_x2) return "zodi_los_kelsall_x2_kernel";
    return "zodi_los_kelsall_kernel";
}

int64_t zodi_kernel_launch_count(void) { return g_launches.load(); }

double zodi_last_kernel_ms(zodi_model_t m) { return m ? m->last_kernel_ms : 0.0; }

int zodi_device_math(int device, int32_t op, int64_t n, const double* x, double aux, double* y) {
    if (!x || !y || n < 0) return fail(ZODI_ERR_INVALID, "bad argument");
    if (op < ZODI_MATH_LOG2_F64 || op > ZODI_MATH_LOG2_F32) return fail(ZODI_ERR_INVALID, "unknown math op %d", op);
    if (n == 0) return ZODI_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    double *d_x = nullptr, *d_y = nullptr;
    CU_CHECK(cudaMalloc((void**)&d_x, (size_t)n * sizeof(double)));
    if (cudaMalloc((void**)&d_y, (size_t)n * sizeof(double)) != cudaSuccess) {
        cudaFree(d_x);
        return fail(ZODI_ERR_NOMEM, "cannot allocate %lld doubles", (long long)n);
    }
    cudaMemcpy(d_x, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
    zodi_device_math_kernel<<<grid, 256>>>(op, n, d_x, aux, d_y);
    g_launches.fetch_add(1);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaMemcpy(y, d_y, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_x);
    cudaFree(d_y);
    if (err != cudaSuccess) return fail(ZODI_ERR_CUDA, "device math: %s", cudaGetErrorString(err));
    return ZODI_OK;
}

int zodi_peak_probe(int device, int32_t kind, double* per_second) {
    if (!per_second) return fail(ZODI_ERR_INVALID, "per_second is NULL");
    int count = 0;
    CU_CHECK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(ZODI_ERR_CUDA, "device %d not available", device);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ZODI_ERR_CUDA, "cannot select device %d", device);
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    cudaEvent_t e0, e1;
    CU_CHECK(cudaEventCreate(&e0));
    CU_CHECK(cudaEventCreate(&e1));
    void* buf = nullptr;
    void* buf2 = nullptr;
    double best = 0.0;
    int rc = ZODI_OK;
    const int64_t copy_bytes = 1ll << 30;
    if (kind == ZODI_PEAK_HBM_COPY) {
        if (cudaMalloc(&buf, copy_bytes) != cudaSuccess || cudaMalloc(&buf2, copy_bytes) != cudaSuccess)
            rc = fail(ZODI_ERR_NOMEM, "cannot allocate copy buffers");
        else
            cudaMemset(buf, 1, copy_bytes);
    } else if (cudaMalloc(&buf, (size_t)blocks * threads * 8) != cudaSuccess) {
        rc = fail(ZODI_ERR_NOMEM, "cannot allocate probe buffer");
    }
    for (int rep = 0; rc == ZODI_OK && rep < 5; ++rep) {
        const int iters = (kind == ZODI_PEAK_MUFU_EX2) ? 4096 : (kind == ZODI_PEAK_FP64_FMA ? 8192 : 16384);
        cudaEventRecord(e0);
        double work = 0.0;
        switch (kind) {
            case ZODI_PEAK_FP32_FMA:
                zodi_peak_fma_kernel<float><<<blocks, threads>>>((float*)buf, iters);
                work = 2.0 * 64.0 * iters * (double)blocks * threads;
                break;
            case ZODI_PEAK_FP64_FMA:
                zodi_peak_fma_kernel<double><<<blocks, threads>>>((double*)buf, iters);
                work = 2.0 * 64.0 * iters * (double)blocks * threads;
                break;
            case ZODI_PEAK_MUFU_EX2:
                zodi_peak_mufu_kernel<<<blocks, threads>>>((float*)buf, iters);
                work = 64.0 * iters * (double)blocks * threads;
                break;
            case ZODI_PEAK_HBM_COPY:
                zodi_peak_copy_kernel<<<blocks * 4, threads>>>((const float4*)buf, (float4*)buf2,
                                                              copy_bytes / 16);
                work = 2.0 * (double)copy_bytes;
                break;
            default:
                rc = fail(ZODI_ERR_INVALID, "unknown peak kind %d", kind);
        }
        g_launches.fetch_add(1);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            rc = fail(ZODI_ERR_CUDA, "peak probe kernel failed");
            break;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms > 0.f) best = std::max(best, work / (ms * 1e-3));
    }
    cudaFree(buf);
    cudaFree(buf2);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *per_second = best;
    return rc;
}

}  // extern "C"
