// ldpc_abi.cu -- C-ABI of the toy NR-LDPC family (include/polar_b200.h, section "NR LDPC"); reference:
// dl_scl_polar/nr/ldpc/* and the nr_ldpc branch of eval/run_ber_sweep.py.  Kernels: ldpc_kernels.cuh.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/polar_b200.h"
#include "ldpc_kernels.cuh"

using namespace pb;

int pb200_set_error(int code, const char* msg);   // polar_abi.cu: sets the thread-local pb200_last_error text

static int lfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return pb200_set_error(code, buf);
}
#define LCUDA_TRY(x)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (x);                                                                           \
        if (_e != cudaSuccess) return lfail(PB200_ECUDA, "%s failed: %s", #x, cudaGetErrorString(_e));  \
    } while (0)

struct GenDev { LdpcGen g{}; uint32_t* d_G = nullptr; uint32_t* d_C = nullptr; };

struct pb200_ldpc {
    int device = 0, sms = 0;
    int m = 0, n = 0, nnz = 0;
    std::vector<uint8_t> H;            // dense [m][n] copy
    int* d_row_ptr = nullptr;
    int* d_col_idx = nullptr;
    std::map<int, GenDev> gens;        // per payload length k
    // layers of consecutive, mutually column-disjoint rows (group-per-frame kernels); G = 0: thread-per-frame only
    int nl = 0, G = 0, lgG = 0;
    bool weight4 = false;              // every row has exactly four ones (static-weight kernels)
    int* d_layer_ptr = nullptr;
    std::map<std::tuple<int, unsigned long long>, unsigned long long*> crc_tabs;   // (kp, poly) -> nibble table
    std::map<cudaStream_t, std::pair<unsigned char*, size_t>> scratch;   // global state for codes too large for smem
};

// nr/ldpc/basegraphs.py:19-42 (3x6 demo graph under both ids) + builder.py:10-30
extern "C" int pb200_ldpc_build_h(int bg, int Z, uint8_t* h_H, int* m_out, int* n_out) {
    static const int shifts[3][6] = {{0, 1, 2, 0, -1, -1}, {1, 0, 3, -1, 0, -1}, {2, 3, 0, -1, -1, 0}};
    if (bg != 1 && bg != 2) return lfail(PB200_EINVAL, "Unknown base graph: %d", bg);
    if (Z <= 0) return lfail(PB200_EINVAL, "Z must be positive");
    const int m = 3 * Z, n = 6 * Z;
    if (m_out) *m_out = m;
    if (n_out) *n_out = n;
    if (!h_H) return PB200_OK;         // size query
    memset(h_H, 0, (size_t)m * n);
    for (int br = 0; br < 3; ++br)
        for (int bc = 0; bc < 6; ++bc) {
            if (shifts[br][bc] < 0) continue;
            const int s = shifts[br][bc] % Z;
            for (int i = 0; i < Z; ++i) h_H[(size_t)(br * Z + i) * n + bc * Z + (i + s) % Z] = 1;
        }
    return PB200_OK;
}

// Greedy layering of the rows of H: a row joins the current layer while it shares no column with it and the layer has
// fewer than 32 rows.  lp = first row of every layer (+ m).  lanes = lanes per frame of the group kernels (layer width
// rounded up to a power of two, 4..32), or 0 when the layers are too narrow (< 4 rows) for that mapping to pay.
static void build_layers_host(const uint8_t* H, int m, int n, std::vector<int>& lp, int* lanes) {
    lp.assign(1, 0);
    std::vector<char> used(n, 0);
    int width = 0, maxw = 0;
    for (int r = 0; r < m; ++r) {
        bool clash = width >= 32;
        for (int c = 0; c < n && !clash; ++c) clash = H[(size_t)r * n + c] && used[c];
        if (clash) {
            lp.push_back(r);
            std::fill(used.begin(), used.end(), 0);
            width = 0;
        }
        for (int c = 0; c < n; ++c) if (H[(size_t)r * n + c]) used[c] = 1;
        ++width;
        maxw = std::max(maxw, width);
    }
    lp.push_back(m);
    int G = 0;
    if (maxw >= 4) { G = 4; while (G < maxw) G <<= 1; }
    *lanes = G;
}

extern "C" int pb200_ldpc_layers(const uint8_t* h_H, int m, int n, int32_t* h_layer_ptr, int* n_layers, int* group_lanes) {
    if (!h_H || m <= 0 || n <= 0) return lfail(PB200_EINVAL, "H must be a non-empty m x n matrix");
    std::vector<int> lp;
    int G = 0;
    build_layers_host(h_H, m, n, lp, &G);
    if (h_layer_ptr) for (size_t i = 0; i < lp.size(); ++i) h_layer_ptr[i] = lp[i];
    if (n_layers) *n_layers = (int)lp.size() - 1;
    if (group_lanes) *group_lanes = G;
    return PB200_OK;
}

extern "C" int pb200_ldpc_create(pb200_ldpc** out, int device, const uint8_t* h_H, int m, int n) {
    if (!out) return lfail(PB200_EINVAL, "out is NULL");
    *out = nullptr;
    if (!h_H || m <= 0 || n <= 0) return lfail(PB200_EINVAL, "H must be a non-empty m x n matrix");
    if (n > PB200_LDPC_MAX_N) return lfail(PB200_ENOSUP, "n > %d is not supported by this build", PB200_LDPC_MAX_N);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return lfail(PB200_ECUDA, "no CUDA device: the polar_b200 engine has no CPU fallback");
    if (device < 0 || device >= ndev) return lfail(PB200_EINVAL, "device %d out of range", device);
    LCUDA_TRY(cudaSetDevice(device));
    std::vector<int> rp(m + 1, 0), ci;
    for (int r = 0; r < m; ++r) {
        for (int c = 0; c < n; ++c) {
            const uint8_t v = h_H[(size_t)r * n + c];
            if (v > 1) return lfail(PB200_EINVAL, "H entries must be 0 or 1");
            if (v) ci.push_back(c);
        }
        rp[r + 1] = (int)ci.size();
    }
    pb200_ldpc* e = new pb200_ldpc();
    e->device = device;
    cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, device);
    e->m = m; e->n = n; e->nnz = (int)ci.size();
    e->H.assign(h_H, h_H + (size_t)m * n);
    if (ci.empty()) ci.push_back(0);
    cudaError_t ce;
    if ((ce = cudaMalloc((void**)&e->d_row_ptr, rp.size() * 4)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&e->d_col_idx, ci.size() * 4)) != cudaSuccess ||
        (ce = cudaMemcpy(e->d_row_ptr, rp.data(), rp.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (ce = cudaMemcpy(e->d_col_idx, ci.data(), ci.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
        pb200_ldpc_destroy(e);
        return lfail(PB200_ECUDA, "table upload failed: %s", cudaGetErrorString(ce));
    }
    {
        std::vector<int> lp;
        int G = 0;
        build_layers_host(h_H, m, n, lp, &G);
        e->nl = (int)lp.size() - 1;
        e->weight4 = true;
        for (int r = 0; r < m; ++r) e->weight4 = e->weight4 && (rp[r + 1] - rp[r] == 4);
        if (G > 0 && e->nnz > 0) {
            e->G = G;
            e->lgG = 0;
            while ((1 << e->lgG) < G) ++e->lgG;
            if ((ce = cudaMalloc((void**)&e->d_layer_ptr, lp.size() * 4)) != cudaSuccess ||
                (ce = cudaMemcpy(e->d_layer_ptr, lp.data(), lp.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
                pb200_ldpc_destroy(e);
                return lfail(PB200_ECUDA, "layer table upload failed: %s", cudaGetErrorString(ce));
            }
        }
    }
    *out = e;
    return PB200_OK;
}

extern "C" void pb200_ldpc_destroy(pb200_ldpc* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaFree(e->d_row_ptr); cudaFree(e->d_col_idx); cudaFree(e->d_layer_ptr);
    for (auto& kv : e->crc_tabs) cudaFree(kv.second);
    for (auto& kv : e->gens) { cudaFree(kv.second.d_G); cudaFree(kv.second.d_C); }
    for (auto& kv : e->scratch) cudaFree(kv.second.first);
    delete e;
}

// Generator of the parity part for payload length k.  encode.py:52-66 solves H_par p = H_sys s by Gauss-Jordan
// elimination with free variables left at 0 (:8-49); the row operations depend on H_par only, so p = G s with
// G = rows of (T H_sys) picked by the pivots (T = accumulated row operations), and the "no solution" test
// (:35-37) is  C s != 0  with C = the rows of T H_sys that belong to the all-zero rows of the reduced matrix.
// host part: G [n-k][kw] and the consistency rows Cc [nc][kw] (kw = max(1, ceil(k/32)))
static int build_generator_host(const uint8_t* H, int m, int n, int k, std::vector<uint32_t>& G, std::vector<uint32_t>& Cc, int* nc_out) {
    if (k < 0) return lfail(PB200_EINVAL, "payload length must be >= 0");
    if (n <= k) return lfail(PB200_EINVAL, "Parity-check matrix too small for payload length");
    const int np = n - k, kw = std::max(1, (k + 31) / 32), mw = (m + 31) / 32;
    std::vector<std::vector<uint8_t>> A(m, std::vector<uint8_t>(np));
    std::vector<std::vector<uint32_t>> T(m, std::vector<uint32_t>(mw, 0));
    for (int r = 0; r < m; ++r) {
        for (int c = 0; c < np; ++c) A[r][c] = H[(size_t)r * n + k + c] & 1;
        T[r][r >> 5] |= 1u << (r & 31);
    }
    std::vector<int> pivot_row(np, -1);
    int row = 0;
    for (int col = 0; col < np && row < m; ++col) {
        int pivot = -1;
        for (int r = row; r < m; ++r) if (A[r][col]) { pivot = r; break; }
        if (pivot < 0) continue;
        if (pivot != row) { std::swap(A[row], A[pivot]); std::swap(T[row], T[pivot]); }
        pivot_row[col] = row;
        for (int r = 0; r < m; ++r)
            if (r != row && A[r][col]) {
                for (int c = 0; c < np; ++c) A[r][c] ^= A[row][c];
                for (int w = 0; w < mw; ++w) T[r][w] ^= T[row][w];
            }
        ++row;
    }
    // rows of T * H_sys as bit rows over the k payload bits
    auto times_hsys = [&](const std::vector<uint32_t>& t, uint32_t* dst) {
        for (int w = 0; w < kw; ++w) dst[w] = 0;
        for (int i = 0; i < m; ++i)
            if ((t[i >> 5] >> (i & 31)) & 1u)
                for (int j = 0; j < k; ++j)
                    if (H[(size_t)i * n + j] & 1) dst[j >> 5] ^= 1u << (j & 31);
    };
    G.assign((size_t)np * kw, 0);
    Cc.clear();
    for (int col = 0; col < np; ++col)
        if (pivot_row[col] >= 0) times_hsys(T[pivot_row[col]], &G[(size_t)col * kw]);
    int nc = 0;
    for (int r = row; r < m; ++r) {
        bool zero = true;
        for (int c = 0; c < np; ++c) zero = zero && !A[r][c];
        if (!zero) continue;
        std::vector<uint32_t> tmp(kw);
        times_hsys(T[r], tmp.data());
        bool any = false;
        for (int w = 0; w < kw; ++w) any = any || tmp[w];
        if (any) { Cc.insert(Cc.end(), tmp.begin(), tmp.end()); ++nc; }
    }
    *nc_out = nc;
    return PB200_OK;
}

extern "C" int pb200_ldpc_parity_generator(const uint8_t* h_H, int m, int n, int k, uint32_t* h_G, uint32_t* h_C, int* n_check) {
    if (!h_H || m <= 0 || n <= 0) return lfail(PB200_EINVAL, "H must be a non-empty m x n matrix");
    std::vector<uint32_t> G, Cc;
    int nc = 0;
    int rc = build_generator_host(h_H, m, n, k, G, Cc, &nc);
    if (rc) return rc;
    if (h_G) memcpy(h_G, G.data(), G.size() * 4);
    if (h_C && nc) memcpy(h_C, Cc.data(), Cc.size() * 4);
    if (n_check) *n_check = nc;
    return PB200_OK;
}

static int get_generator(pb200_ldpc* e, int k, const GenDev** out) {
    auto it = e->gens.find(k);
    if (it != e->gens.end()) { *out = &it->second; return PB200_OK; }
    const int n = e->n;
    std::vector<uint32_t> G, Cc;
    int nc = 0;
    int rc = build_generator_host(e->H.data(), e->m, n, k, G, Cc, &nc);
    if (rc) return rc;
    const int np = n - k, kw = std::max(1, (k + 31) / 32);
    GenDev gd;
    LCUDA_TRY(cudaSetDevice(e->device));
    LCUDA_TRY(cudaMalloc((void**)&gd.d_G, G.size() * 4));
    LCUDA_TRY(cudaMemcpy(gd.d_G, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
    if (nc) {
        LCUDA_TRY(cudaMalloc((void**)&gd.d_C, Cc.size() * 4));
        LCUDA_TRY(cudaMemcpy(gd.d_C, Cc.data(), Cc.size() * 4, cudaMemcpyHostToDevice));
    }
    gd.g.k = k; gd.g.kw = kw; gd.g.np = np; gd.g.nc = nc; gd.g.G = gd.d_G; gd.g.Cc = gd.d_C;
    auto ins = e->gens.emplace(k, gd);
    *out = &ins.first->second;
    return PB200_OK;
}

static LdpcCode code_of(const pb200_ldpc* e, int k) {
    LdpcCode c;
    c.m = e->m; c.n = e->n; c.k = k; c.row_ptr = e->d_row_ptr; c.col_idx = e->d_col_idx;
    return c;
}

static unsigned grid_for(int64_t items, int threads, int cap_blocks) {
    const int64_t want = (items + threads - 1) / threads;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, cap_blocks));
}

extern "C" int pb200_ldpc_encode_batch(pb200_ldpc* e, const uint8_t* d_payload, int k, uint8_t* d_code, uint8_t* d_status,
                                       int64_t B, void* stream) {
    if (!e) return lfail(PB200_EINVAL, "engine is NULL");
    if (B < 0 || (B > 0 && (!d_payload || !d_code))) return lfail(PB200_EINVAL, "bad buffers");
    const GenDev* gd;
    int rc = get_generator(e, k, &gd);
    if (rc) return rc;
    if (B == 0) return PB200_OK;
    LCUDA_TRY(cudaSetDevice(e->device));
    const int threads = 128;
    ldpc_encode_kernel<<<grid_for(B, threads, e->sms * 8), threads, (size_t)gd->g.kw * threads * 4, (cudaStream_t)stream>>>(
        e->n, gd->g, d_payload, d_code, d_status, B);
    LCUDA_TRY(cudaGetLastError());
    return PB200_OK;
}

extern "C" int pb200_ldpc_rate_match_batch(const uint8_t* d_code, int N, int E, uint8_t* d_out, int64_t B, void* stream) {
    if (N <= 0 || E < 0 || B < 0) return lfail(PB200_EINVAL, "bad sizes");
    if (B == 0 || E == 0) return PB200_OK;
    if (!d_code || !d_out) return lfail(PB200_EINVAL, "bad buffers");
    ldpc_rate_match_kernel<<<grid_for(B * E, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(d_code, N, E, d_out, B);
    LCUDA_TRY(cudaGetLastError());
    return PB200_OK;
}

extern "C" int pb200_ldpc_derate_match_batch(const double* d_llr, int E, int N, double* d_out, int64_t B, void* stream) {
    if (N <= 0 || E < 0 || B < 0) return lfail(PB200_EINVAL, "bad sizes");
    if (B == 0) return PB200_OK;
    if ((E > 0 && !d_llr) || !d_out) return lfail(PB200_EINVAL, "bad buffers");
    ldpc_derate_kernel<<<grid_for(B * N, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(d_llr, E, N, d_out, B);
    LCUDA_TRY(cudaGetLastError());
    return PB200_OK;
}

// Launch shape of the thread-per-frame kernels: state in shared memory when at least one warp fits, else global.
struct LdpcLaunch { int threads, blocks; size_t smem; bool global; };

static int plan_launch(pb200_ldpc* e, const void* fn, size_t bytes_per_thread, int64_t frames, LdpcLaunch* L) {
    const size_t cap = 200 * 1024;
    int threads = (int)std::min<size_t>(256, (cap / bytes_per_thread) / 32 * 32);
    L->global = threads < 32;
    if (L->global) { threads = 128; L->smem = 0; }
    else L->smem = bytes_per_thread * threads;
    if (!L->global) LCUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    int per_sm = 0;
    LCUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, L->smem));
    if (per_sm < 1) return lfail(PB200_ECUDA, "LDPC kernel does not fit (n=%d)", e->n);
    if (L->global) per_sm = std::min(per_sm, 2);      // bounds the global scratch (state bytes x resident threads)
    L->threads = threads;
    L->blocks = (int)grid_for(frames, threads, e->sms * per_sm);
    return PB200_OK;
}

static int ensure_scratch(pb200_ldpc* e, cudaStream_t st, size_t need, unsigned char** out) {
    auto& slot = e->scratch[st];
    if (slot.second < need) {
        if (slot.first) { LCUDA_TRY(cudaStreamSynchronize(st)); cudaFree(slot.first); }
        slot.first = nullptr; slot.second = 0;
        LCUDA_TRY(cudaMalloc((void**)&slot.first, need));
        slot.second = need;
    }
    *out = slot.first;
    return PB200_OK;
}

// Launch shape of the group-per-frame kernels: warps per CTA such that the frame states fit in shared memory.
// ok = false: the state of even one warp does not fit -> use the thread-per-frame kernels.
static int plan_group_launch(pb200_ldpc* e, const void* fn, size_t frame_bytes, int64_t frames, LdpcLaunch* L, bool* ok) {
    *ok = false;
    if (e->G == 0) return PB200_OK;
    const int fpw = 32 / e->G;
    const size_t cap = 200 * 1024, per_warp = frame_bytes * fpw;
    int wpc = (int)std::min<size_t>(8, cap / per_warp);
    if (wpc < 1) return PB200_OK;
    L->threads = wpc * 32;
    L->smem = per_warp * wpc;
    L->global = false;
    LCUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    int per_sm = 0;
    LCUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, L->threads, L->smem));
    if (per_sm < 1) return PB200_OK;
    const int64_t batches = (frames + fpw - 1) / fpw;
    L->blocks = (int)grid_for(batches, wpc, e->sms * per_sm);
    *ok = true;
    return PB200_OK;
}

static LdpcLayers layers_of(const pb200_ldpc* e) {
    LdpcLayers L;
    L.nl = e->nl; L.G = e->G; L.lgG = e->lgG; L.layer_ptr = e->d_layer_ptr;
    return L;
}

extern "C" int pb200_ldpc_decode_batch(pb200_ldpc* e, const double* d_llr, int64_t B, int in_len, int max_iter, double alpha,
                                       int early_stop, uint8_t* d_hard, double* d_posterior, int32_t* d_iters, uint8_t* d_ok,
                                       void* stream) {
    if (!e) return lfail(PB200_EINVAL, "engine is NULL");
    if (B < 0 || in_len < 0) return lfail(PB200_EINVAL, "bad sizes");
    if (B == 0) return PB200_OK;
    if (!d_llr && in_len > 0) return lfail(PB200_EINVAL, "llr is NULL");
    LCUDA_TRY(cudaSetDevice(e->device));
    LdpcLaunch L;
    LdpcDecodeArgs a{};
    a.llr = d_llr; a.B = B; a.in_len = in_len; a.max_iter = max_iter; a.early_stop = early_stop; a.alpha = alpha;
    a.hard = d_hard; a.posterior = d_posterior; a.iters = d_iters; a.ok = d_ok;
    bool grouped = false;
    const void* gfn = e->weight4 ? (const void*)ldpc_decode_group_kernel<4> : (const void*)ldpc_decode_group_kernel<0>;
    int rc = plan_group_launch(e, gfn, ldpc_group_frame_bytes(e->n, e->m, 0, 0, 0), B, &L, &grouped);
    if (rc) return rc;
    if (grouped) {
        if (e->weight4) ldpc_decode_group_kernel<4><<<L.blocks, L.threads, L.smem, (cudaStream_t)stream>>>(code_of(e, e->n - e->m), layers_of(e), a);
        else ldpc_decode_group_kernel<0><<<L.blocks, L.threads, L.smem, (cudaStream_t)stream>>>(code_of(e, e->n - e->m), layers_of(e), a);
        LCUDA_TRY(cudaGetLastError());
        return PB200_OK;
    }
    rc = plan_launch(e, (const void*)ldpc_decode_kernel, (size_t)(e->n + e->m) * 8, B, &L);
    if (rc) return rc;
    if (L.global) {
        unsigned char* p;
        rc = ensure_scratch(e, (cudaStream_t)stream, (size_t)(e->n + e->m) * 8 * L.threads * L.blocks, &p);
        if (rc) return rc;
        a.gscratch = reinterpret_cast<double*>(p);
    }
    ldpc_decode_kernel<<<L.blocks, L.threads, L.smem, (cudaStream_t)stream>>>(code_of(e, e->n - e->m), a);
    LCUDA_TRY(cudaGetLastError());
    return PB200_OK;
}

// CRC nibble table for the grouped sweep: row (q, v) = remainder contribution of payload nibble q holding value v
// (crc.py:19-37 is linear with a zero initial register): bit j of the payload contributes x^(kp-1-j+deg) mod g.
static int get_crc_tab(pb200_ldpc* e, int kp, unsigned long long poly, int deg, LdpcCrcTab* out) {
    auto key = std::make_tuple(kp, poly);
    auto it = e->crc_tabs.find(key);
    const int nq = (kp + 3) / 4;
    if (it == e->crc_tabs.end()) {
        const unsigned long long mask = deg >= 64 ? ~0ull : ((1ull << deg) - 1ull), low = poly & mask;
        std::vector<unsigned long long> bitrem(kp);
        // x^deg mod g, then multiply by x step by step: bit kp-1 first (exponent deg), bit 0 last
        unsigned long long r = low;                       // x^deg mod g
        for (int j = kp - 1; j >= 0; --j) {
            bitrem[j] = r;
            const unsigned long long top = (r >> (deg - 1)) & 1ull;
            r = (r << 1) & mask;
            if (top) r ^= low;
        }
        std::vector<unsigned long long> tab((size_t)nq * 16, 0);
        for (int q = 0; q < nq; ++q)
            for (int v = 0; v < 16; ++v)
                for (int b = 0; b < 4; ++b)
                    if (((v >> b) & 1) && q * 4 + b < kp) tab[(size_t)q * 16 + v] ^= bitrem[q * 4 + b];
        unsigned long long* d = nullptr;
        LCUDA_TRY(cudaMalloc((void**)&d, tab.size() * 8));
        LCUDA_TRY(cudaMemcpy(d, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
        it = e->crc_tabs.emplace(key, d).first;
    }
    out->tab = it->second;
    out->nq = nq;
    return PB200_OK;
}

static int sweep_common(pb200_ldpc* e, const pb200_ldpc_sweep_cfg* c, int64_t* d_counters, uint16_t* d_fbe, uint16_t* d_fwork,
                        uint8_t* d_payload, double* d_llr, bool chan_only, void* stream) {
    if (!e || !c) return lfail(PB200_EINVAL, "engine / cfg is NULL");
    const int k = e->n - e->m;
    if (c->k_payload <= 0 || c->k_crc < 0) return lfail(PB200_EINVAL, "bad payload / CRC sizes");
    if (c->k_payload + c->k_crc != k) return lfail(PB200_EINVAL, "LDPC payload+CRC size mismatch with base graph");
    if (c->E <= 0) return lfail(PB200_EINVAL, "E must be positive");
    if (!(c->noise_var > 0.0)) return lfail(PB200_EINVAL, "noise_var must be positive");
    if (c->n_frames < 0) return lfail(PB200_EINVAL, "n_frames must be >= 0");
    if (c->max_iter > 65535) return lfail(PB200_ENOSUP, "max_iter > 65535 is not supported by the per-frame work counters");
    unsigned long long poly = 0;
    int deg = 0;
    if (c->k_crc > 0) {
        if (!c->crc_poly || !*c->crc_poly) return lfail(PB200_EINVAL, "CRC polynomial string must be non-empty");
        char* end = nullptr;
        poly = strtoull(c->crc_poly, &end, 16);
        if (end == c->crc_poly || *end != 0) return lfail(PB200_EINVAL, "CRC polynomial must be a hex string");
        while (deg < 64 && (poly >> deg)) ++deg;
        deg -= 1;
        if (deg <= 0) return lfail(PB200_EINVAL, "Polynomial degree must be positive");
        if (deg > 63) return lfail(PB200_ENOSUP, "CRC degree > 63 is not supported");
        if (deg < c->k_crc) return lfail(PB200_ENOSUP, "CRC degree below K_crc (message shorter than the systematic part) is not supported");
    }
    if (c->n_frames == 0) return PB200_OK;
    if (chan_only ? !d_llr : !d_counters) return lfail(PB200_EINVAL, "output buffer is NULL");
    const GenDev* gd;
    int rc = get_generator(e, k, &gd);
    if (rc) return rc;
    if (gd->g.nc) return lfail(PB200_EINVAL, "Linear system over GF(2) has no solution for some payloads of this H");
    LCUDA_TRY(cudaSetDevice(e->device));
    const int nw = (e->n + 31) / 32;
    LdpcLaunch L;
    const size_t bpt = (size_t)(e->n + e->m) * 8 + (size_t)(gd->g.kw + nw) * 4;
    bool grouped = false;
    const void* gfn = e->weight4 ? (const void*)ldpc_sweep_group_kernel<4> : (const void*)ldpc_sweep_group_kernel<0>;
    rc = plan_group_launch(e, gfn, ldpc_group_frame_bytes(e->n, e->m, c->E, gd->g.kw, nw), c->n_frames, &L, &grouped);
    if (rc) return rc;
    LdpcCrcTab ct{nullptr, 0};
    if (grouped && deg > 0) {
        rc = get_crc_tab(e, c->k_payload, poly, deg, &ct);
        if (rc) return rc;
    }
    if (!grouped) {
        rc = plan_launch(e, (const void*)ldpc_sweep_kernel, bpt, c->n_frames, &L);
        if (rc) return rc;
    }
    LdpcSweepArgs a{};
    a.frame_begin = c->frame_begin; a.n_frames = c->n_frames;
    a.k0 = (uint32_t)(c->seed & 0xffffffffu);
    a.k1 = (uint32_t)(c->seed >> 32) + 0x9E3779B9u * c->stream_id;
    a.sigma = (float)sqrt(c->noise_var);
    a.scale = (float)(2.0 / c->noise_var);
    a.kp = c->k_payload; a.E = c->E; a.poly = poly; a.deg = deg;
    a.max_iter = c->max_iter; a.early_stop = c->early_stop; a.alpha = c->alpha;
    a.counters = reinterpret_cast<unsigned long long*>(d_counters);
    a.frame_bit_errors = d_fbe; a.frame_work = d_fwork;
    a.payload_out = chan_only ? d_payload : nullptr;
    a.llr_out = chan_only ? d_llr : nullptr;
    if (grouped) {
        if (e->weight4) ldpc_sweep_group_kernel<4><<<L.blocks, L.threads, L.smem, (cudaStream_t)stream>>>(code_of(e, k), gd->g, layers_of(e), ct, a);
        else ldpc_sweep_group_kernel<0><<<L.blocks, L.threads, L.smem, (cudaStream_t)stream>>>(code_of(e, k), gd->g, layers_of(e), ct, a);
        LCUDA_TRY(cudaGetLastError());
        return PB200_OK;
    }
    if (L.global) {
        unsigned char* p;
        const size_t tot = (size_t)L.threads * L.blocks;
        rc = ensure_scratch(e, (cudaStream_t)stream, bpt * tot + 64, &p);
        if (rc) return rc;
        a.gscratch = reinterpret_cast<double*>(p);
        a.gwords = reinterpret_cast<uint32_t*>(p + (size_t)(e->n + e->m) * 8 * tot);
    }
    ldpc_sweep_kernel<<<L.blocks, L.threads, L.smem, (cudaStream_t)stream>>>(code_of(e, k), gd->g, a);
    LCUDA_TRY(cudaGetLastError());
    return PB200_OK;
}

extern "C" int pb200_ldpc_sweep(pb200_ldpc* e, const pb200_ldpc_sweep_cfg* cfg, int64_t* d_counters,
                                uint16_t* d_frame_bit_errors, uint16_t* d_frame_work, void* stream) {
    return sweep_common(e, cfg, d_counters, d_frame_bit_errors, d_frame_work, nullptr, nullptr, false, stream);
}

extern "C" int pb200_ldpc_channel_batch(pb200_ldpc* e, const pb200_ldpc_sweep_cfg* cfg, uint8_t* d_payload, double* d_llr,
                                        void* stream) {
    return sweep_common(e, cfg, nullptr, nullptr, nullptr, d_payload, d_llr, true, stream);
}
