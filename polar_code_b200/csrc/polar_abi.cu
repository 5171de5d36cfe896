// polar_abi.cu -- C-ABI of libpolar_b200.so (see include/polar_b200.h) and kernel dispatch.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/polar_b200.h"
#include "polar_kernels.cuh"
#include "polar_sweep.cuh"
#include "polar_launch.h"

using namespace pb;

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
int pb200_set_error(int code, const char* msg) { g_err = msg; return code; }   // for the other ABI translation units
#define CUDA_TRY(x)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (x);                                                                         \
        if (_e != cudaSuccess) return fail(PB200_ECUDA, "%s failed: %s", #x, cudaGetErrorString(_e)); \
    } while (0)

extern "C" const char* pb200_last_error(void) { return g_err.c_str(); }
extern "C" int pb200_version(void) { return 100; }
extern "C" int pb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ---------------------------------------------------------------------------------------------------
// Code construction (host, float64): polar/polar.py:37-103
// ---------------------------------------------------------------------------------------------------
static int ilog2_exact(int N) {
    if (N <= 0 || (N & (N - 1))) return -1;
    int n = 0;
    while ((1 << n) < N) ++n;
    return n;
}

static double phi_inverse(double x) {  // polar.py:51-58
    if (x > 12.0) return 0.9861 * x - 2.3152;
    if (x > 3.5) return x * (0.009005 * x + 0.7694) - 0.9507;
    if (x > 1.0) return x * (0.062883 * x + 0.3678) - 0.1627;
    return x * (0.2202 * x + 0.06448);
}

extern "C" int pb200_construct_info_set(int N, int K, int method, double design_snr_db, int32_t* out) {
    const int n = ilog2_exact(N);
    if (n < 0) return fail(PB200_EINVAL, "N must be a power of two");
    if (!(0 < K && K <= N)) return fail(PB200_EINVAL, "K must satisfy 0 < K <= N");
    std::vector<double> metric(N, 0.0);
    if (method == 1) {  // polarization weights, polar.py:37-48
        for (int idx = 0; idx < N; ++idx) {
            double w = 0.0;
            for (int j = 0; j < n; ++j)
                if ((idx >> j) & 1) w += pow(2.0, j / 4.0);
            metric[idx] = w;
        }
    } else if (method == 0) {  // Gaussian approximation, polar.py:61-82
        const double sigma_sq = 1.0 / (2.0 * ((double)K / (double)N) * pow(10.0, design_snr_db / 10.0));
        metric[0] = 2.0 / sigma_sq;
        for (int level = 1; level <= n; ++level) {
            const int half = (1 << level) >> 1;
            for (int j = 0; j < half; ++j) {
                const double T = metric[j];
                metric[j] = phi_inverse(T);
                metric[half + j] = 2.0 * T;
            }
        }
        for (int i = 0; i < N; ++i) metric[i] = 0.5 - 0.5 * erf(sqrt(std::max(metric[i], 1e-12)) / 2.0);
    } else {
        return fail(PB200_EINVAL, "Unsupported construction method");
    }
    std::vector<int> order(N);
    for (int i = 0; i < N; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return metric[a] < metric[b]; });
    std::vector<int> best(order.begin(), order.begin() + K);
    std::sort(best.begin(), best.end());
    for (int i = 0; i < K; ++i) out[i] = best[i];
    return PB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// Engine
// ---------------------------------------------------------------------------------------------------
struct KernelCfg { int wpc, ctas_per_sm, smem, regs; };

struct pb200_engine {
    int device = 0, sms = 0;
    Code code{};
    std::vector<int> info_pos;
    unsigned long long poly = 0;   // CRC polynomial incl. leading 1 (0 = none)
    int K = 0;
    Tables tb{};
    int16_t* d_info_pos = nullptr;
    uint32_t* d_info_mask = nullptr;
    uint32_t* d_crc_tab = nullptr;
    int16_t* d_rm_src = nullptr;
    int8_t* d_rm_cnt = nullptr;
    int16_t* d_tx_src = nullptr;   // NR transmit gather: tx[e] = code[tx_src[e]] or pad (-1)
    std::map<int, uint32_t*> enc_tabs;   // sweep encoder tables, one per payload length (polar_abi_sweep.inl get_enc_tab)
    std::map<std::tuple<int, int, int>, KernelCfg> cfg_cache;
    // host-buffer pipeline
    cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
    float* d_stage_llr[3] = {nullptr, nullptr, nullptr};
    uint8_t* d_stage_bits[3] = {nullptr, nullptr, nullptr};
    uint8_t* d_stage_ok[3] = {nullptr, nullptr, nullptr};
    uint32_t* d_stage_flags[3] = {nullptr, nullptr, nullptr};
    uint16_t* d_stage_half[3] = {nullptr, nullptr, nullptr};     // binary16 ingest: rows land here and are widened into d_stage_llr
    int64_t stage_frames = 0, stage_half_frames = 0;
    int stage_len = 0;
    // NR + DL-SCL state
    int16_t* d_rm_dst = nullptr;   // [N] de-rate-matched position -> internal index
    unsigned char* d_q[2] = {nullptr, nullptr};
    size_t q_bytes = 0;
    unsigned int* d_q_counts = nullptr;
    int q_counts_n = 0;
    float* d_llr_store = nullptr;         // channel rows of frames in the DL-SCL retry queue (sweep mode)
    size_t llr_store_bytes = 0;
    float* d_abs_store = nullptr;         // |L0| rows of frames in the DL-SCL retry queue
    size_t abs_store_bytes = 0;
    int* d_bin_ring = nullptr;            // dl_bin_kernel: one ring of queue-entry indices per flip index
    size_t bin_ring_bytes = 0;
    unsigned int* d_bin_ctrl = nullptr;   // cursor, finished count, ring heads and tails
    int bin_ctrl_n = 0;
    int dl_binned = -1;                   // -1: not read yet (PB200_DL_BINNED, default 1)
    std::map<int, KernelCfg> dl_cfg;      // launch configuration of the binned retry kernel per MP
    // fraction of frames the previous DL-SCL piece queued for retries (read back asynchronously): picks the admission mode
    unsigned int* h_queued = nullptr;     // pinned: [0] queued frames of the last piece
    cudaEvent_t queued_ev = nullptr;
    long long queued_of = 0;              // frames of that piece (0: nothing in flight / known)
    double dl_fail_frac = -1.0;           // last known failure fraction (-1: unknown)
    long long dl_piece_max = 0;           // frames per DL-SCL piece, sized against free memory at the first DL-SCL sweep
    double* d_beta64 = nullptr;           // the caller's beta widened to fp64 for the current call [K,K]
    // per-warp global scratch of the decode kernels, one buffer per stream: kernels of one engine that are
    // enqueued on different streams may overlap on the device and must not share tree rows
    std::map<cudaStream_t, std::pair<unsigned char*, size_t>> scratch;
    // L2 residency control of the scratch (pin_scratch_in_l2): -1 = not probed, -2 = switched off by the user
    int l2_window_max = -1;
    std::map<cudaStream_t, std::pair<unsigned char*, size_t>> l2_window;
};
int sweep_build_tables(pb200_engine* e);

static int parse_poly(const char* s, unsigned long long* poly, int* deg) {
    if (!s || !*s) return fail(PB200_EINVAL, "CRC polynomial string must be non-empty");
    char* end = nullptr;
    const unsigned long long v = strtoull(s, &end, 16);
    if (end == s || *end != 0) return fail(PB200_EINVAL, "CRC polynomial must be a hex string");
    int len = 0;
    while (len < 64 && (v >> len)) ++len;
    if (len - 1 <= 0) return fail(PB200_EINVAL, "Polynomial degree must be positive");
    *poly = v;
    *deg = len - 1;
    return PB200_OK;
}

// x^e mod g(x) as a deg-bit integer
static unsigned long long xpow_mod(int e, unsigned long long poly, int deg) {
    const unsigned long long low = poly & ((1ull << deg) - 1ull);
    unsigned long long r = 1;  // x^0
    if (deg == 0) return 0;
    for (int i = 0; i < e; ++i) {
        const unsigned long long top = (r >> (deg - 1)) & 1ull;
        r = (r << 1) & ((1ull << deg) - 1ull);
        if (top) r ^= low;
    }
    return r;
}

extern "C" int pb200_create(pb200_engine** out, int device, int N, const int32_t* info_set, int K, const char* crc_poly) {
    if (!out) return fail(PB200_EINVAL, "out is NULL");
    *out = nullptr;
    const int n = ilog2_exact(N);
    if (n < 0) return fail(PB200_EINVAL, "N must be a power of two");
    if (N < 2) return fail(PB200_ENOSUP, "N must be at least 2");
    if (N > PB200_MAX_N) return fail(PB200_ENOSUP, "N > %d is not supported by this build", PB200_MAX_N);
    if (K <= 0 || K > N || !info_set) return fail(PB200_EINVAL, "info_set must hold 0 < K <= N indices");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(PB200_ECUDA, "no CUDA device: the polar_b200 engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(PB200_EINVAL, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    pb200_engine* e = new pb200_engine();
    e->device = device;
    cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, device);
    e->K = K;
    e->code.N = N; e->code.n = n; e->code.K = K; e->code.M = 1; e->code.crc_deg = 0;
    memset(e->code.info_mask, 0, sizeof e->code.info_mask);
    e->info_pos.assign(info_set, info_set + K);
    for (int j = 0; j < K; ++j) {
        const int p = info_set[j];
        if (p < 0 || p >= N) { delete e; return fail(PB200_EINVAL, "info_set indices out of range"); }
        if (j > 0 && info_set[j] <= info_set[j - 1]) { delete e; return fail(PB200_EINVAL, "info_set must be strictly increasing"); }
        e->code.info_mask[p >> 5] |= 1u << (p & 31);
    }
    int deg = 0;
    if (crc_poly) {
        int rc = parse_poly(crc_poly, &e->poly, &deg);
        if (rc) { delete e; return rc; }
        if (deg > 32) { delete e; return fail(PB200_ENOSUP, "CRC degree > 32 is not supported by the fused check"); }
        if (K <= deg) { delete e; return fail(PB200_EINVAL, "Message too short for the provided CRC polynomial"); }
    }
    e->code.crc_deg = deg;
    // tables
    std::vector<int16_t> pos16(K);
    for (int j = 0; j < K; ++j) pos16[j] = (int16_t)info_set[j];
    const int nn = N >= 4 ? N / 4 : 1;
    std::vector<uint32_t> crc_tab((size_t)nn * 16, 0);
    if (deg) {
        std::vector<uint32_t> S(N, 0);  // syndrome contribution of phase p
        for (int j = 0; j < K; ++j) S[info_set[j]] = (uint32_t)xpow_mod(K - 1 - j, e->poly, deg);
        for (int nib = 0; nib < nn; ++nib)
            for (int v = 0; v < 16; ++v) {
                uint32_t s = 0;
                for (int b = 0; b < 4; ++b)
                    if (((v >> b) & 1) && nib * 4 + b < N) s ^= S[nib * 4 + b];
                crc_tab[nib * 16 + v] = s;
            }
    }
    std::vector<int16_t> rm(N);
    for (int i = 0; i < N; ++i) rm[i] = (int16_t)i;
    std::vector<int8_t> rc1(N, 1);
    auto up = [&](void** d, const void* h, size_t bytes) -> cudaError_t {
        cudaError_t r = cudaMalloc(d, bytes);
        if (r != cudaSuccess) return r;
        return cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t ce;
    if ((ce = up((void**)&e->d_info_pos, pos16.data(), pos16.size() * 2)) != cudaSuccess ||
        (ce = up((void**)&e->d_crc_tab, crc_tab.data(), crc_tab.size() * 4)) != cudaSuccess ||
        (ce = up((void**)&e->d_info_mask, e->code.info_mask, sizeof e->code.info_mask)) != cudaSuccess ||
        (ce = up((void**)&e->d_rm_src, rm.data(), rm.size() * 2)) != cudaSuccess ||
        (ce = up((void**)&e->d_rm_cnt, rc1.data(), rc1.size())) != cudaSuccess) {
        pb200_destroy(e);
        return fail(PB200_ECUDA, "table upload failed: %s", cudaGetErrorString(ce));
    }
    e->tb.info_pos = e->d_info_pos;
    e->tb.info_mask = e->d_info_mask;
    e->tb.crc_tab = e->d_crc_tab;
    e->tb.rm_src = e->d_rm_src;
    e->tb.rm_cnt = e->d_rm_cnt;
    e->tb.E = 0;
    int rc = sweep_build_tables(e);
    if (rc) { pb200_destroy(e); return rc; }
    *out = e;
    return PB200_OK;
}

static void l2_release_window(int device);

extern "C" void pb200_destroy(pb200_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (auto& kv : e->l2_window)                  // hand this engine's share of the L2 set-aside back
        if (kv.second.first != nullptr && e->device >= 0 && e->device < 64) l2_release_window(e->device);
    cudaFree(e->d_info_pos); cudaFree(e->d_info_mask); cudaFree(e->d_crc_tab); cudaFree(e->d_rm_src); cudaFree(e->d_rm_cnt); cudaFree(e->d_tx_src);
    for (auto& kv : e->enc_tabs) cudaFree(kv.second);
    cudaFree(e->d_bin_ring); cudaFree(e->d_bin_ctrl);
    if (e->h_queued) cudaFreeHost(e->h_queued);
    if (e->queued_ev) cudaEventDestroy(e->queued_ev);
    cudaFree(e->d_llr_store); cudaFree(e->d_abs_store); cudaFree(e->d_rm_dst); cudaFree(e->d_beta64);
    for (auto& kv : e->scratch) cudaFree(kv.second.first); cudaFree(e->d_q[0]); cudaFree(e->d_q[1]); cudaFree(e->d_q_counts);
    for (int i = 0; i < 3; ++i) {
        if (e->hs[i]) cudaStreamDestroy(e->hs[i]);
        cudaFree(e->d_stage_llr[i]); cudaFree(e->d_stage_bits[i]); cudaFree(e->d_stage_ok[i]); cudaFree(e->d_stage_flags[i]); cudaFree(e->d_stage_half[i]);
    }
    delete e;
}

// nr/polar/interleaver.py:10-37 + rate_match.py:8-39 folded into two gather tables.
extern "C" int pb200_set_rate_matching(pb200_engine* e, int E) {
    if (!e) return fail(PB200_EINVAL, "engine is NULL");
    if (E < 0) return fail(PB200_EINVAL, "E must be >= 0");
    CUDA_TRY(cudaSetDevice(e->device));
    const int N = e->code.N;
    // 1. all host tables first (nothing of the engine is touched until every check and upload has succeeded)
    std::vector<int16_t> rm(N), rd(N), tx;
    std::vector<int8_t> rcnt(N, 1);
    if (E == 0) {
        for (int i = 0; i < N; ++i) { rm[i] = (int16_t)i; rd[i] = (int16_t)i; }
    } else {
        const int block = 32, nb = (N + block - 1) / block, total = nb * block;
        std::vector<int> order(total), inv(total);
        for (int i = 0; i < total; ++i) order[i] = (i % block) * nb + (i / block);   // interleaver.py:20
        for (int i = 0; i < total; ++i) inv[order[i]] = i;                           // argsort(order), :31-36
        // receive side: internal[i] = derated[inv[i]] if inv[i] < N else 0.0 (zero padding, :33-34)
        for (int i = 0; i < N; ++i) rm[i] = (int16_t)(inv[i] < N ? inv[i] : -1);
        // copies of de-rate-matched position p = rm[i] among the E transmitted symbols (rate_match.py:19-39)
        for (int i = 0; i < N; ++i) {
            const int pp = rm[i];
            const int copies = pp < 0 ? -1 : (pp < E ? (E - pp + N - 1) / N : 0);
            if (copies > 127) return fail(PB200_ENOSUP, "E > 127 N is not supported");
            rcnt[i] = (int8_t)copies;
        }
        for (int p = 0; p < N; ++p) rd[p] = (int16_t)(order[p] < N ? order[p] : -1);
        // transmit side: interleaved[k] = code[order[k]] or pad -1 (:17-21); rate_match: first E of the tiling (:8-16)
        tx.resize(E);
        for (int t = 0; t < E; ++t) {
            const int k = (E <= total) ? t : (t % total);
            tx[t] = (int16_t)(order[k] < N ? order[k] : -1);
        }
    }
    // 2. the engine must be idle: kernels of earlier calls (on any stream) may still read the tables
    CUDA_TRY(cudaDeviceSynchronize());
    int16_t* new_tx = nullptr;
    if (E > 0) {
        CUDA_TRY(cudaMalloc((void**)&new_tx, (size_t)E * 2));
        if (cudaMemcpy(new_tx, tx.data(), (size_t)E * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(new_tx);
            return fail(PB200_ECUDA, "upload of the transmit table failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    // 3. fixed-size tables are overwritten in place; on a failure the engine falls back to "no rate matching" with
    //    consistent identity tables instead of a half-updated state
    const bool ok = cudaMemcpy(e->d_rm_src, rm.data(), (size_t)N * 2, cudaMemcpyHostToDevice) == cudaSuccess &&
                    cudaMemcpy(e->d_rm_dst, rd.data(), (size_t)N * 2, cudaMemcpyHostToDevice) == cudaSuccess &&
                    cudaMemcpy(e->d_rm_cnt, rcnt.data(), (size_t)N, cudaMemcpyHostToDevice) == cudaSuccess;
    cudaFree(e->d_tx_src);
    e->d_tx_src = nullptr;
    if (!ok) {
        const char* why = cudaGetErrorString(cudaGetLastError());
        cudaFree(new_tx);
        std::vector<int16_t> id(N);
        for (int i = 0; i < N; ++i) id[i] = (int16_t)i;
        std::vector<int8_t> one(N, 1);
        cudaMemcpy(e->d_rm_src, id.data(), (size_t)N * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(e->d_rm_dst, id.data(), (size_t)N * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(e->d_rm_cnt, one.data(), (size_t)N, cudaMemcpyHostToDevice);
        e->tb.E = 0;
        return fail(PB200_ECUDA, "upload of the rate-matching tables failed (%s); rate matching is now off", why);
    }
    e->d_tx_src = new_tx;
    e->tb.E = E;
    return PB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// Kernel selection
// ---------------------------------------------------------------------------------------------------
static const void* pick_decode(int n, int MP, bool forced, bool metric, bool trace = false) {
    if (trace) return n == 7 ? pb_decode_kernel_7s_trace(MP) : n < 7 ? pb_decode_kernel_7_trace(MP) : pb_decode_kernel_9_trace(MP);
    if (n == 7) return pb_decode_kernel_7s(MP, forced, metric);      // the headline geometry N = 128: static code length
    return n < 7 ? pb_decode_kernel_7(MP, forced, metric) : pb_decode_kernel_9(MP, forced, metric);
}

static int round_mp(int M) { return M <= 1 ? 1 : M <= 2 ? 2 : M <= 4 ? 4 : 8; }

// per-warp global scratch; sized for the smallest split (HS = 5) so every kernel of the engine fits
static size_t warp_gbytes(int MP, int N, int K) {     // tree + channel rows, plus the trace rows behind them
    switch (MP) {
        case 1: return WarpMem<1, 5>::gbytes(N) + WarpMem<1, 5>::hbytes(K);
        case 2: return WarpMem<2, 5>::gbytes(N) + WarpMem<2, 5>::hbytes(K);
        case 4: return WarpMem<4, 5>::gbytes(N) + WarpMem<4, 5>::hbytes(K);
        default: return WarpMem<8, 5>::gbytes(N) + WarpMem<8, 5>::hbytes(K);
    }
}

// global scratch for `warps` resident warps of a launch on `st` (grows on demand; stays L2-resident across launches)
static int ensure_scratch(pb200_engine* e, cudaStream_t st, size_t warps, int MP, unsigned char** out) {
    const size_t need = warps * warp_gbytes(MP, e->code.N, e->code.K) + 256;
    auto& slot = e->scratch[st];
    if (slot.second < need) {
        if (slot.first) {
            CUDA_TRY(cudaStreamSynchronize(st));     // earlier launches on this stream may still use the old buffer
            cudaFree(slot.first);
        }
        slot.first = nullptr; slot.second = 0;
        CUDA_TRY(cudaMalloc((void**)&slot.first, need));
        slot.second = need;
    }
    *out = slot.first;
    return PB200_OK;
}

// Keep the dense tree/channel scratch of a launch resident in L2 (it is re-written and re-read every few microseconds)
// while everything else the stream touches -- the LLR rows in, the decisions out -- keeps the normal policy:
// an access-policy window over the scratch with the persisting property, backed by an L2 set-aside.  Best effort
// (older drivers / MIG slices without the feature just run without it).  Only the list kernels (MP >= 2) ask for it;
// a thread-per-frame launch (SC / M = 1) clears the window of ITS stream.
// The set-aside (cudaLimitPersistingL2CacheSize) is DEVICE-wide state, so it is managed in process-global, reference-
// counted state per device: it is taken when the first (engine, stream) window is set and given back -- persisting lines
// demoted -- when the last one is cleared (by a thread-per-frame launch on that stream or by pb200_destroy).  While any
// engine's list kernels hold a window nobody's launch path changes the limit.  Giving it back matters: an unused 79 MB
// set-aside leaves SC / M = 1 only 47 MB of L2 (measured: 0.69e9 instead of 1.6e9 frames/s).
// PB200_L2_PIN=0 switches the feature off; PB200_L2_SETASIDE_MB caps the set-aside (default: all the device allows).
struct L2DeviceState { bool probed = false; int window_max = 0, persist_max = 0, active_windows = 0; };
static L2DeviceState g_l2[64];
static std::mutex g_l2_mutex;

static void l2_release_window(int device) {            // one (engine, stream) window less; the last one frees the set-aside
    std::lock_guard<std::mutex> lock(g_l2_mutex);
    L2DeviceState& g = g_l2[device];
    if (g.active_windows > 0 && --g.active_windows == 0) {
        if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError();
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0) != cudaSuccess) cudaGetLastError();
    }
}

static void pin_scratch_in_l2(pb200_engine* e, cudaStream_t st, unsigned char* scratch, size_t bytes, bool enable) {
    if (e->l2_window_max == -1) {
        const char* env = getenv("PB200_L2_PIN");
        e->l2_window_max = (env && env[0] == '0') ? -2 : 0;      // -2: disabled by the user
    }
    if (e->l2_window_max == -2 || e->device < 0 || e->device >= 64) return;
    if (!enable) {
        // A thread-per-frame launch wants the whole L2: clear the windows of ALL streams of this engine (the host-buffer
        // path leaves one on each of its three internal streams), not only this stream's -- as long as one of them holds
        // the set-aside, SC / M = 1 run at 0.67e9 instead of 1.6e9 frames/s.  The next list launch on a stream sets its
        // window again.
        (void)st;
        for (auto& kv : e->l2_window) {
            if (kv.second.first == nullptr) continue;
            cudaStreamAttrValue v{};
            v.accessPolicyWindow.base_ptr = nullptr;
            v.accessPolicyWindow.num_bytes = 0;
            v.accessPolicyWindow.hitRatio = 0.f;
            v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
            v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            if (cudaStreamSetAttribute(kv.first, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
            kv.second = {nullptr, 0};
            l2_release_window(e->device);
        }
        return;
    }
    auto& cur = e->l2_window[st];
    L2DeviceState ds;
    {
        std::lock_guard<std::mutex> lock(g_l2_mutex);
        L2DeviceState& g = g_l2[e->device];
        if (!g.probed) {
            cudaDeviceGetAttribute(&g.window_max, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
            cudaDeviceGetAttribute(&g.persist_max, cudaDevAttrMaxPersistingL2CacheSize, e->device);
            if (const char* cap = getenv("PB200_L2_SETASIDE_MB")) {
                const long long mb = atoll(cap);
                if (mb >= 0 && mb * (1ll << 20) < g.persist_max) g.persist_max = (int)(mb << 20);
            }
            g.probed = true;
        }
        if (g.window_max > 0 && g.persist_max > 0 && cur.first == nullptr) {
            if (g.active_windows++ == 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)g.persist_max) != cudaSuccess)
                cudaGetLastError();
        }
        ds = g;
    }
    if (ds.window_max <= 0 || ds.persist_max <= 0) return;
    if (cur.first == scratch && cur.second == bytes) return;
    cudaStreamAttrValue v{};
    v.accessPolicyWindow.base_ptr = scratch;
    v.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)ds.window_max);
    v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)ds.persist_max / (double)std::max<size_t>(v.accessPolicyWindow.num_bytes, 1));
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
    cur = {scratch, bytes};
}

// shared bytes per warp; xk > 0: the |L0| rows of the DL-SCL retry kernel, tk > 0: the lineage bytes of a trace-recording
// kernel.  round = DL-SCL retry kernel (HS = 5), otherwise the kernel's default split
static size_t warp_bytes(int MP, int N, int xk, bool round = false, int tk = 0) {
    if (round) {
        switch (MP) {
            case 1: return WarpMem<1, 5>::bytes(N, xk, tk);
            case 2: return WarpMem<2, 5>::bytes(N, xk, tk);
            case 4: return WarpMem<4, 5>::bytes(N, xk, tk);
            default: return WarpMem<8, 5>::bytes(N, xk, tk);
        }
    }
    switch (MP) {
        case 1: return WarpMem<1>::bytes(N, xk, tk);
        case 2: return WarpMem<2>::bytes(N, xk, tk);
        case 4: return WarpMem<4>::bytes(N, xk, tk);
        default: return WarpMem<8>::bytes(N, xk, tk);
    }
}

// Choose warps per CTA so that resident warps per SM are maximal for this kernel's registers and smem.
static int choose_cfg(pb200_engine* e, const void* fn, int MP, int key_kind, size_t wb, KernelCfg* out) {
    auto key = std::make_tuple(MP, key_kind, (int)e->tb.E);
    auto it = e->cfg_cache.find(key);
    if (it != e->cfg_cache.end()) { *out = it->second; return PB200_OK; }
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, fn));
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
    KernelCfg best{0, 0, 0, fa.numRegs};
    int best_warps = 0;
    for (int wpc = 32; wpc >= 1; --wpc) {
        const size_t smem = wb * wpc;
        if (smem > (size_t)(227 * 1024) - fa.sharedSizeBytes) continue;
        if (wpc * 32 > fa.maxThreadsPerBlock) continue;
        int blocks = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, fn, wpc * 32, smem) != cudaSuccess) { cudaGetLastError(); continue; }
        if (blocks * wpc > best_warps) { best_warps = blocks * wpc; best = KernelCfg{wpc, blocks, (int)smem, fa.numRegs}; }
    }
    if (best_warps == 0) return fail(PB200_ECUDA, "no launch configuration fits (N=%d, MP=%d)", e->code.N, MP);
    e->cfg_cache[key] = best;
    *out = best;
    return PB200_OK;
}

static int launch_decode(pb200_engine* e, int M, bool metric, const DecodeArgs& a, cudaStream_t st) {
    const int MP = metric ? round_mp(M) : 1;
    const bool forced = a.force != nullptr;
    const bool trace = metric && a.info_llrs != nullptr;       // info_llrs of all paths: read off the leaf-LLR trace
    const void* fn = pick_decode(e->code.n, MP, forced, metric, trace);
    KernelCfg kc;
    int rc = choose_cfg(e, fn, MP, trace ? 8 : ((forced ? 1 : 0) | (metric ? 2 : 0)), warp_bytes(MP, e->code.N, 0, false, trace ? e->code.K : 0), &kc);
    if (rc) return rc;
    Code code = e->code;
    code.M = metric ? M : 1;
    const int fpw = 32 / MP;
    const int64_t groups = (a.B + fpw - 1) / fpw;
    const int64_t want = (groups + kc.wpc - 1) / kc.wpc;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)e->sms * kc.ctas_per_sm));
    DecodeArgs aa = a;
    rc = ensure_scratch(e, st, (size_t)grid * kc.wpc, MP, &aa.gscratch);
    if (rc) return rc;
    static const bool pin_m1 = [] { const char* v = getenv("PB200_L2_PIN_M1"); return v && v[0] == '1'; }();
    pin_scratch_in_l2(e, st, aa.gscratch, (size_t)grid * kc.wpc * warp_gbytes(MP, e->code.N, 0), MP >= 2 || pin_m1);
    void* args[3] = {(void*)&code, (void*)&e->tb, (void*)&aa};
    CUDA_TRY(cudaLaunchKernel(fn, dim3(grid), dim3(kc.wpc * 32), args, kc.smem, st));
    return PB200_OK;
}

static int check_decode_args(pb200_engine* e, const float* llr, int64_t B, int in_len, int M) {
    if (!e) return fail(PB200_EINVAL, "engine is NULL");
    if (M <= 0) return fail(PB200_EINVAL, "List size M must be positive");
    if (M > PB200_MAX_M) return fail(PB200_ENOSUP, "list size M > %d is not supported by this build", PB200_MAX_M);
    if (B < 0) return fail(PB200_EINVAL, "B must be >= 0");
    if (B > 0 && !llr) return fail(PB200_EINVAL, "llr is NULL");
    const int want = e->tb.E ? e->tb.E : e->code.N;
    if (in_len != want) return fail(PB200_EINVAL, "llr rows must have length %d (got %d)", want, in_len);
    return PB200_OK;
}

extern "C" int pb200_kernel_info(pb200_engine* e, int M, int* wpc, int* ctas, int* smem, int* regs) {
    if (!e) return fail(PB200_EINVAL, "engine is NULL");
    if (M <= 0 || M > PB200_MAX_M) return fail(PB200_EINVAL, "bad M");
    CUDA_TRY(cudaSetDevice(e->device));
    const int MP = round_mp(M);
    const void* fn = pick_decode(e->code.n, MP, false, true);
    KernelCfg kc;
    int rc = choose_cfg(e, fn, MP, 2, warp_bytes(MP, e->code.N, 0), &kc);
    if (rc) return rc;
    if (wpc) *wpc = kc.wpc;
    if (ctas) *ctas = kc.ctas_per_sm;
    if (smem) *smem = kc.smem;
    if (regs) *regs = kc.regs;
    return PB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// Encoder / CRC
// ---------------------------------------------------------------------------------------------------
extern "C" int pb200_encode_batch(pb200_engine* e, const uint8_t* msg, uint8_t* code, int64_t B, void* stream) {
    if (!e) return fail(PB200_EINVAL, "engine is NULL");
    if (B < 0 || (B > 0 && (!msg || !code))) return fail(PB200_EINVAL, "bad buffers");
    if (B == 0) return PB200_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    encode_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(e->code, e->tb, msg, code, B);
    CUDA_TRY(cudaGetLastError());
    return PB200_OK;
}

static int crc_common(const char* poly, const uint8_t* msg, int64_t B, int L, uint8_t* attach, uint8_t* ok, void* stream) {
    unsigned long long p;
    int deg;
    int rc = parse_poly(poly, &p, &deg);
    if (rc) return rc;
    if (deg > 63) return fail(PB200_ENOSUP, "CRC degree > 63 is not supported");
    if (L < 0 || B < 0) return fail(PB200_EINVAL, "bad sizes");
    if (ok && L <= deg) return fail(PB200_EINVAL, "Message too short for the provided CRC polynomial");
    if (B == 0) return PB200_OK;
    if (!msg) return fail(PB200_EINVAL, "msg is NULL");
    crc_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p, deg, msg, L, B, attach, ok);
    CUDA_TRY(cudaGetLastError());
    return PB200_OK;
}
extern "C" int pb200_crc_attach_batch(const char* poly, const uint8_t* msg, uint8_t* out, int64_t B, int L, void* stream) {
    if (!out && B > 0) return fail(PB200_EINVAL, "out is NULL");
    return crc_common(poly, msg, B, L, out, nullptr, stream);
}
extern "C" int pb200_crc_check_batch(const char* poly, const uint8_t* msg, uint8_t* ok, int64_t B, int L, void* stream) {
    if (!ok && B > 0) return fail(PB200_EINVAL, "ok is NULL");
    return crc_common(poly, msg, B, L, nullptr, ok, stream);
}

// ---------------------------------------------------------------------------------------------------
// Decoders
// ---------------------------------------------------------------------------------------------------
extern "C" int pb200_sc_decode_batch(pb200_engine* e, const float* llr, int64_t B, int in_len, uint8_t* bits, void* stream) {
    int rc = check_decode_args(e, llr, B, in_len, 1);
    if (rc) return rc;
    if (B == 0) return PB200_OK;
    if (!bits) return fail(PB200_EINVAL, "bits is NULL");
    CUDA_TRY(cudaSetDevice(e->device));
    DecodeArgs a{};
    a.llr = llr; a.B = B; a.in_len = in_len; a.best_bits = bits;
    return launch_decode(e, 1, false, a, (cudaStream_t)stream);
}

extern "C" int pb200_scl_decode_batch(pb200_engine* e, const float* llr, int64_t B, int in_len, const int8_t* force, int M,
                                      const pb200_scl_out* out, void* stream) {
    int rc = check_decode_args(e, llr, B, in_len, M);
    if (rc) return rc;
    if (!out) return fail(PB200_EINVAL, "out is NULL");
    if (B == 0) return PB200_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    DecodeArgs a{};
    a.llr = llr; a.B = B; a.in_len = in_len; a.force = force;
    a.cand = out->cand; a.metrics = out->metrics; a.info_llrs = out->info_llrs; a.n_cand = out->n_cand;
    a.best_idx = out->best_idx; a.best_bits = out->best_bits; a.best_words = out->best_words; a.crc_ok = out->crc_ok;
    a.flags = out->flags;
    return launch_decode(e, M, true, a, (cudaStream_t)stream);
}

// binary16 rows -> fp32 rows (exact).  The host-buffer path is bound by the host->device copy, so the widening pass over
// the staged chunk (0.75 KB of HBM traffic per frame) is free and the decode kernels keep a single input format.
static __global__ void widen_rows_kernel(const __half* __restrict__ in, float* __restrict__ out, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        const float2 v = __half22float2(*reinterpret_cast<const __half2*>(in + i));
        *reinterpret_cast<float2*>(out + i) = v;
    } else if (i < n) out[i] = __half2float(in[i]);
}

// Host buffers in, host buffers out: chunked, triple-buffered copy/compute overlap on internal streams.
// elem = bytes per LLR (4: fp32 rows, 2: binary16 rows).
static int decode_host_common(pb200_engine* e, const void* h_llr, int elem, int64_t B, int in_len, int M, uint8_t* h_bits,
                              uint8_t* h_ok, uint32_t* h_flags) {
    int rc = check_decode_args(e, reinterpret_cast<const float*>(h_llr), B, in_len, M);
    if (rc) return rc;
    if (B == 0) return PB200_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    const int K = e->code.K;
    // 2^16 frames (32 MiB of fp32 LLRs) per chunk: the pipeline's fill (first copy-in) and drain (last decode + copy-out)
    // are not overlapped, so small chunks keep them short; a chunk still fills the GPU (8 192 warps of work)
    const int64_t chunk = std::min<int64_t>(B, 1 << 16);
    if (e->stage_frames < chunk || e->stage_len != in_len) {
        // (the recorded geometry is cleared first: an allocation that fails half-way must not leave a size behind that
        //  makes the next call skip this block and run on freed / null buffers)
        e->stage_frames = 0; e->stage_half_frames = 0;
        for (int i = 0; i < 3; ++i) {
            if (e->hs[i]) CUDA_TRY(cudaStreamSynchronize(e->hs[i]));
            cudaFree(e->d_stage_llr[i]); cudaFree(e->d_stage_bits[i]); cudaFree(e->d_stage_ok[i]); cudaFree(e->d_stage_flags[i]);
            cudaFree(e->d_stage_half[i]);
            e->d_stage_llr[i] = nullptr; e->d_stage_bits[i] = nullptr; e->d_stage_ok[i] = nullptr; e->d_stage_flags[i] = nullptr;
            e->d_stage_half[i] = nullptr;
            if (!e->hs[i]) CUDA_TRY(cudaStreamCreateWithFlags(&e->hs[i], cudaStreamNonBlocking));
            CUDA_TRY(cudaMalloc((void**)&e->d_stage_llr[i], (size_t)chunk * in_len * 4));
            CUDA_TRY(cudaMalloc((void**)&e->d_stage_bits[i], (size_t)chunk * K));
            CUDA_TRY(cudaMalloc((void**)&e->d_stage_ok[i], (size_t)chunk));
            CUDA_TRY(cudaMalloc((void**)&e->d_stage_flags[i], (size_t)chunk * 4));
        }
        e->stage_frames = chunk;
        e->stage_len = in_len;
    }
    if (elem == 2 && e->stage_half_frames < e->stage_frames) {      // binary16 landing buffers: only when that ingest is used
        e->stage_half_frames = 0;
        for (int i = 0; i < 3; ++i) {
            cudaFree(e->d_stage_half[i]);
            e->d_stage_half[i] = nullptr;
            CUDA_TRY(cudaMalloc((void**)&e->d_stage_half[i], (size_t)e->stage_frames * in_len * 2));
        }
        e->stage_half_frames = e->stage_frames;
    }
    int64_t done = 0;
    int slot = 0;
    while (done < B) {
        const int64_t nb = std::min<int64_t>(chunk, B - done);
        cudaStream_t st = e->hs[slot];
        void* land = elem == 4 ? (void*)e->d_stage_llr[slot] : (void*)e->d_stage_half[slot];
        CUDA_TRY(cudaMemcpyAsync(land, reinterpret_cast<const unsigned char*>(h_llr) + (size_t)done * in_len * elem,
                                 (size_t)nb * in_len * elem, cudaMemcpyHostToDevice, st));
        if (elem == 2) {
            const size_t n = (size_t)nb * in_len;
            widen_rows_kernel<<<(unsigned)((n / 2 + 256) / 256), 256, 0, st>>>(reinterpret_cast<const __half*>(e->d_stage_half[slot]), e->d_stage_llr[slot], n);
            CUDA_TRY(cudaGetLastError());
        }
        DecodeArgs a{};
        a.llr = e->d_stage_llr[slot]; a.B = nb; a.in_len = in_len;
        a.best_bits = e->d_stage_bits[slot]; a.crc_ok = e->d_stage_ok[slot]; a.flags = e->d_stage_flags[slot];
        rc = launch_decode(e, M, true, a, st);
        if (rc) return rc;
        if (h_bits) CUDA_TRY(cudaMemcpyAsync(h_bits + done * K, e->d_stage_bits[slot], (size_t)nb * K, cudaMemcpyDeviceToHost, st));
        if (h_ok) CUDA_TRY(cudaMemcpyAsync(h_ok + done, e->d_stage_ok[slot], (size_t)nb, cudaMemcpyDeviceToHost, st));
        if (h_flags) CUDA_TRY(cudaMemcpyAsync(h_flags + done, e->d_stage_flags[slot], (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        done += nb;
        slot = (slot + 1) % 3;
    }
    for (int i = 0; i < 3; ++i) CUDA_TRY(cudaStreamSynchronize(e->hs[i]));
    return PB200_OK;
}

extern "C" int pb200_scl_decode_host(pb200_engine* e, const float* h_llr, int64_t B, int in_len, int M, uint8_t* h_bits,
                                     uint8_t* h_ok, uint32_t* h_flags) {
    return decode_host_common(e, h_llr, 4, B, in_len, M, h_bits, h_ok, h_flags);
}

extern "C" int pb200_scl_decode_host_f16(pb200_engine* e, const uint16_t* h_llr_f16, int64_t B, int in_len, int M, uint8_t* h_bits,
                                         uint8_t* h_ok, uint32_t* h_flags) {
    return decode_host_common(e, h_llr_f16, 2, B, in_len, M, h_bits, h_ok, h_flags);
}

#include "polar_abi_sweep.inl"
