// ldpc_kernels.cuh -- the toy NR-LDPC family on sm_100a (reference: dl_scl_polar/nr/ldpc/*).
//
// Work mapping: one THREAD per frame (the codes are tiny: n = 6Z, m = 3Z, row weight 4), float64 arithmetic with
// explicit round-to-nearest intrinsics (no FMA contraction), so every decision, iteration count and posterior is
// bit-identical to the reference's float64 NumPy decoder.
//
// Per-thread decoder state lives in shared memory, lane-interleaved: element i of thread t sits at st[i*T + t]
// (T = threads per CTA), so a warp's access to "element i" is one conflict-free 256 B row whatever the parity-check
// row asks for.  State = n posteriors + m check-node messages: the reference stores msg[r, idx] per EDGE
// (decode_nms.py:22), but every edge of a row is assigned the same `update` (decode_nms.py:32-33) and is only ever
// read at that row's own columns, so one scalar per row is an exact representation (m instead of m*n doubles).
// Codes too large for shared memory (n + m > ~900) use the same layout in a global scratch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "polar_sweep.cuh"   // philox4x32_10, normal4, kPurpose*

namespace pb {

struct LdpcCode {
    int m, n, k;                 // rows, columns, systematic bits (n - m for the sweep; encode takes any k < n)
    const int* row_ptr;          // [m+1] CSR of H (columns ascending inside a row, as np.where yields them)
    const int* col_idx;          // [nnz]
};

struct LdpcGen {                 // GF(2) generator of the parity part for one k (encode.py:52-66 is linear in the payload)
    int k, kw, np, nc;           // payload bits, words per row, parity columns (n-k), consistency rows
    const uint32_t* G;           // [np][kw]  parity[p] = <G[p], payload>
    const uint32_t* Cc;          // [nc][kw]  the system has no solution iff some <Cc[c], payload> = 1 (encode.py:35-37)
};

// Layered normalised min-sum (decode_nms.py:24-40) on the state column `st` (stride T): posteriors st[0..n),
// row messages st[n..n+m) (must be zero on entry).  Returns iters_used; parity_ok via `ok`.
__device__ __forceinline__ int nms_decode(const LdpcCode& c, double* st, int T, int max_iter, double alpha, bool early_stop,
                                          bool& ok) {
    const int n = c.n, m = c.m;
    int used = max_iter;
    bool clean = false;
    for (int it = 1; it <= max_iter; ++it) {
        for (int r = 0; r < m; ++r) {
            const int e0 = __ldg(c.row_ptr + r), e1 = __ldg(c.row_ptr + r + 1);
            if (e0 == e1) continue;                                       // decode_nms.py:27-28
            const double mr = st[(n + r) * T];
            double sign = 1.0, mag = __longlong_as_double(0x7ff0000000000000ll);
            for (int e = e0; e < e1; ++e) {
                const double ext = __dsub_rn(st[__ldg(c.col_idx + e) * T], mr);   // :29
                const double sg = ext > 0.0 ? 1.0 : (ext < 0.0 ? -1.0 : 0.0);    // np.sign
                sign = __dmul_rn(sign, sg);                                       // :30
                mag = fmin(mag, fabs(ext));                                       // :31
            }
            const double update = __dmul_rn(__dmul_rn(alpha, sign), mag);         // :32
            for (int e = e0; e < e1; ++e) {
                double* p = st + __ldg(c.col_idx + e) * T;
                *p = __dadd_rn(__dsub_rn(*p, mr), update);                        // :34
            }
            st[(n + r) * T] = update;                                             // :33
        }
        if (early_stop || it == max_iter) {                                       // :36-39
            bool bad = false;
            for (int r = 0; r < m; ++r) {
                uint32_t par = 0;
                for (int e = __ldg(c.row_ptr + r); e < __ldg(c.row_ptr + r + 1); ++e)
                    par ^= (uint32_t)(st[__ldg(c.col_idx + e) * T] < 0.0);
                bad |= par != 0;
            }
            clean = !bad;
            if (early_stop && clean) { used = it; break; }
        }
    }
    if (max_iter <= 0) {                                                          // :41 with the initial hard decisions
        bool bad = false;
        for (int r = 0; r < m; ++r) {
            uint32_t par = 0;
            for (int e = __ldg(c.row_ptr + r); e < __ldg(c.row_ptr + r + 1); ++e)
                par ^= (uint32_t)(st[__ldg(c.col_idx + e) * T] < 0.0);
            bad |= par != 0;
        }
        clean = !bad;
        used = max_iter;
    }
    ok = clean;
    return used;
}

// rate_match.py:18-38 derate_match_ldpc for position i of an E-long row (row-major, float64)
__device__ __forceinline__ double ldpc_derate_at(const double* row, int E, int N, int i) {
    if (E <= N) return i < E ? row[i] : 0.0;
    const int reps = E / N, rem = E - reps * N;
    double s = row[i];
    for (int r = 1; r < reps; ++r) s = __dadd_rn(s, row[(size_t)r * N + i]);
    double acc = __dadd_rn(0.0, s);
    int cnt = reps;
    if (i < rem) { acc = __dadd_rn(acc, row[(size_t)reps * N + i]); ++cnt; }
    return __ddiv_rn(acc, (double)cnt);
}

struct LdpcDecodeArgs {
    const double* llr;           // [B][in_len]
    long long B;
    int in_len;                  // n, or E (de-rate-matching fused into the load)
    int max_iter, early_stop;
    double alpha;
    uint8_t* hard;               // [B][n] or null
    double* posterior;           // [B][n] or null (final llr vector, for tests)
    int32_t* iters;              // [B] or null
    uint8_t* ok;                 // [B] or null
    double* gscratch;            // null = state in shared memory; else [(n+m)][gridDim.x*blockDim.x]
};

__global__ void ldpc_decode_kernel(LdpcCode c, LdpcDecodeArgs a) {
    extern __shared__ double ldpc_smem[];
    const int T = a.gscratch ? (int)(gridDim.x * blockDim.x) : (int)blockDim.x;
    double* st = a.gscratch ? a.gscratch + (size_t)blockIdx.x * blockDim.x + threadIdx.x : ldpc_smem + threadIdx.x;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < a.B; f += (long long)gridDim.x * blockDim.x) {
        const double* row = a.llr + (size_t)f * a.in_len;
        if (a.in_len == c.n) { for (int i = 0; i < c.n; ++i) st[i * T] = row[i]; }
        else { for (int i = 0; i < c.n; ++i) st[i * T] = ldpc_derate_at(row, a.in_len, c.n, i); }
        for (int r = 0; r < c.m; ++r) st[(c.n + r) * T] = 0.0;
        bool ok;
        const int used = nms_decode(c, st, T, a.max_iter, a.alpha, a.early_stop != 0, ok);
        if (a.hard) for (int i = 0; i < c.n; ++i) a.hard[(size_t)f * c.n + i] = (uint8_t)(st[i * T] < 0.0);
        if (a.posterior) for (int i = 0; i < c.n; ++i) a.posterior[(size_t)f * c.n + i] = st[i * T];
        if (a.iters) a.iters[f] = used;
        if (a.ok) a.ok[f] = (uint8_t)ok;
    }
}

__global__ void ldpc_derate_kernel(const double* llr, int E, int N, double* out, long long B) {
    const long long total = B * N;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long f = e / N;
        const int i = (int)(e - f * N);
        out[e] = ldpc_derate_at(llr + (size_t)f * E, E, N, i);
    }
}

// rate_match.py:8-15 rate_match_ldpc: out[B][E] = first E entries of the tiled codeword
__global__ void ldpc_rate_match_kernel(const uint8_t* code, int N, int E, uint8_t* out, long long B) {
    const long long total = B * E;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long f = e / E;
        const int t = (int)(e - f * E);
        out[e] = code[(size_t)f * N + (t % N)];
    }
}

// encode.py:52-66 encode_ldpc over rows: payload[B][k] -> code[B][n]; status[f] = 1 when the system has no solution
__global__ void ldpc_encode_kernel(int n, LdpcGen g, const uint8_t* payload, uint8_t* code, uint8_t* status, long long B) {
    extern __shared__ uint32_t ldpc_wsmem[];
    const int T = blockDim.x;
    uint32_t* w = ldpc_wsmem + threadIdx.x;               // payload words of this thread: w[j*T]
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < B; f += (long long)gridDim.x * blockDim.x) {
        const uint8_t* p = payload + (size_t)f * g.k;
        for (int j = 0; j < g.kw; ++j) {
            uint32_t v = 0;
            for (int b = 0; b < 32 && j * 32 + b < g.k; ++b) v |= (uint32_t)(p[j * 32 + b] & 1u) << b;
            w[j * T] = v;
        }
        uint8_t* o = code + (size_t)f * n;
        for (int j = 0; j < g.k; ++j) o[j] = p[j] & 1u;
        for (int q = 0; q < g.np; ++q) {
            uint32_t acc = 0;
            for (int j = 0; j < g.kw; ++j) acc ^= __ldg(g.G + (size_t)q * g.kw + j) & w[j * T];
            o[g.k + q] = (uint8_t)(__popc(acc) & 1);
        }
        if (status) {
            uint32_t bad = 0;
            for (int q = 0; q < g.nc; ++q) {
                uint32_t acc = 0;
                for (int j = 0; j < g.kw; ++j) acc ^= __ldg(g.Cc + (size_t)q * g.kw + j) & w[j * T];
                bad |= __popc(acc) & 1;
            }
            status[f] = (uint8_t)bad;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused Monte-Carlo sweep for --scheme nr_ldpc (eval/run_ber_sweep.py:127-166 with the encoder / decoder of
// :258-271): payload -> [CRC] -> encode -> rate match -> BPSK + AWGN -> LLR -> de-rate-match -> NMS decode ->
// counters, all inside one thread per frame.  Same Philox4x32-10 convention as the polar sweep (counter =
// global frame index, key = (seed, stream)), so results do not depend on chunking or on the number of ranks.
// ---------------------------------------------------------------------------------------------------
struct LdpcSweepArgs {
    long long frame_begin, n_frames;
    uint32_t k0, k1;             // Philox key
    float sigma, scale;          // noise sigma, 2/sigma^2 (fp32 channel as in the polar sweep; widened to f64 LLRs)
    int kp, E;                   // payload bits, transmitted bits
    unsigned long long poly;     // CRC polynomial incl. leading 1 (deg = 0: no CRC)
    int deg;
    int max_iter, early_stop;
    double alpha;
    unsigned long long* counters;   // [16] : [0] frames [1] frame errors [2] bit errors [7] iterations (work)
    uint16_t* frame_bit_errors;     // [n_frames] or null (exact: k <= n <= 4096)
    uint16_t* frame_work;           // [n_frames] or null
    uint8_t* payload_out;           // channel-only mode: [n_frames][kp] or null
    double* llr_out;                // channel-only mode: [n_frames][E]; non-null = do not decode
    double* gscratch;
    uint32_t* gwords;               // global word scratch when gscratch is used
};

__global__ void ldpc_sweep_kernel(LdpcCode c, LdpcGen g, LdpcSweepArgs a) {
    extern __shared__ double ldpc_smem[];
    const int n = c.n, m = c.m, k = g.k;
    const int nw = (n + 31) / 32, kw = g.kw;
    const int T = a.gscratch ? (int)(gridDim.x * blockDim.x) : (int)blockDim.x;
    const int tid = a.gscratch ? (int)(blockIdx.x * blockDim.x + threadIdx.x) : (int)threadIdx.x;
    double* st = (a.gscratch ? a.gscratch : ldpc_smem) + tid;
    uint32_t* wb = (a.gscratch ? a.gwords : reinterpret_cast<uint32_t*>(ldpc_smem + (size_t)(n + m) * blockDim.x)) + tid;
    uint32_t* msg = wb;                       // message words  msg[j*T], j < kw
    uint32_t* cw = wb + (size_t)kw * T;       // codeword words cw[j*T],  j < nw
    const uint2 key = make_uint2(a.k0, a.k1);
    unsigned long long n_frames = 0, n_fe = 0, n_be = 0, n_work = 0;
    for (long long fi = (long long)blockIdx.x * blockDim.x + threadIdx.x; fi < a.n_frames; fi += (long long)gridDim.x * blockDim.x) {
        const long long fr = a.frame_begin + fi;
        // ---- payload (run_ber_sweep.py:128) and optional CRC (:134-136; crc.py:19-37), first k bits kept (:264) ----
        for (int j = 0; j < kw; ++j) msg[j * T] = 0;
        const int pwn = (a.kp + 31) / 32;
        for (int w = 0; w < pwn; w += 4) {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)(w >> 2), kPurposePayload), key);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (w + q < pwn) {
                    uint32_t v = rr[q];
                    const int rem = a.kp - (w + q) * 32;
                    if (rem < 32) v &= (1u << rem) - 1u;
                    if (w + q < kw) {
                        const int keep = k - (w + q) * 32;      // bits of this word that are inside the first k
                        msg[(w + q) * T] = keep >= 32 ? v : (keep > 0 ? (v & ((1u << keep) - 1u)) : 0u);
                    }
                    if (a.payload_out) {
                        for (int b = 0; b < 32 && (w + q) * 32 + b < a.kp; ++b)
                            a.payload_out[(size_t)fi * a.kp + (w + q) * 32 + b] = (uint8_t)((v >> b) & 1u);
                    }
                }
            }
        }
        if (a.deg > 0) {
            // the CRC runs over the whole payload (which may be longer than k only in degenerate set-ups we reject on the host)
            const unsigned long long low = a.poly & ((1ull << a.deg) - 1ull);
            unsigned long long reg = 0;
            for (int j = 0; j < a.kp; ++j) {
                const unsigned long long b = (msg[(j >> 5) * T] >> (j & 31)) & 1u;
                const unsigned long long top = ((reg >> (a.deg - 1)) & 1ull) ^ b;
                reg = (reg << 1) & ((1ull << a.deg) - 1ull);
                if (top) reg ^= low;
            }
            for (int t = 0; t < a.deg && a.kp + t < k; ++t) {
                const int j = a.kp + t;
                const uint32_t b = (uint32_t)((reg >> (a.deg - 1 - t)) & 1ull);
                msg[(j >> 5) * T] |= b << (j & 31);
            }
        }
        // ---- systematic encode (encode.py:52-66): codeword = [message | G message] -----------------------------
        for (int j = 0; j < nw; ++j) cw[j * T] = j < kw ? msg[j * T] : 0u;
        for (int q = 0; q < g.np; ++q) {
            uint32_t acc = 0;
            for (int j = 0; j < kw; ++j) acc ^= __ldg(g.G + (size_t)q * kw + j) & msg[j * T];
            const int pos = k + q;
            cw[(pos >> 5) * T] |= (uint32_t)(__popc(acc) & 1) << (pos & 31);
        }
        // ---- rate match, BPSK + AWGN, LLR (run_ber_sweep.py:139-142), de-rate-match accumulated in transmit order ----
        const bool chan_only = a.llr_out != nullptr;
        int pos = 0, rep = 0;                 // t = rep*n + pos
        for (int tb = 0; tb * 4 < a.E; ++tb) {
            float z[4];
            normal4(philox4x32_10(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)tb, kPurposeNoise), key), z);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = tb * 4 + q;
                if (t < a.E) {
                    const uint32_t bit = (cw[(pos >> 5) * T] >> (pos & 31)) & 1u;
                    const double llr = (double)(fmaf(a.sigma, z[q], 1.0f - 2.0f * (float)bit) * a.scale);
                    if (chan_only) a.llr_out[(size_t)fi * a.E + t] = llr;
                    else st[pos * T] = rep == 0 ? llr : __dadd_rn(st[pos * T], llr);
                    if (++pos == n) { pos = 0; ++rep; }
                }
            }
        }
        if (chan_only) continue;
        if (a.E <= n) { for (int i = a.E; i < n; ++i) st[i * T] = 0.0; }                 // rate_match.py:21-24
        else {
            const int reps = a.E / n, rem = a.E - reps * n;
            for (int i = 0; i < n; ++i)                                                  // rate_match.py:32-38
                st[i * T] = __ddiv_rn(__dadd_rn(0.0, st[i * T]) , (double)(reps + (i < rem ? 1 : 0)));
        }
        for (int r = 0; r < m; ++r) st[(n + r) * T] = 0.0;
        bool ok;
        const int used = nms_decode(c, st, T, a.max_iter, a.alpha, a.early_stop != 0, ok);
        // ---- payload errors (run_ber_sweep.py:77-82,153-157) -----------------------------------------------
        uint32_t be = 0;
        for (int j = 0; j < a.kp; ++j) be += (uint32_t)(st[j * T] < 0.0) ^ ((msg[(j >> 5) * T] >> (j & 31)) & 1u);
        n_frames += 1; n_fe += be > 0; n_be += be; n_work += (unsigned)used;
        if (a.frame_bit_errors) a.frame_bit_errors[fi] = (uint16_t)be;
        if (a.frame_work) a.frame_work[fi] = (uint16_t)(used > 65535 ? 65535 : used);
    }
    if (a.llr_out) return;
    // warp-level reduction, one atomic per warp and counter
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_frames += __shfl_xor_sync(0xffffffffu, n_frames, o);
        n_fe += __shfl_xor_sync(0xffffffffu, n_fe, o);
        n_be += __shfl_xor_sync(0xffffffffu, n_be, o);
        n_work += __shfl_xor_sync(0xffffffffu, n_work, o);
    }
    if ((threadIdx.x & 31) == 0 && n_frames) {
        atomicAdd(a.counters + 0, n_frames);
        atomicAdd(a.counters + 1, n_fe);
        atomicAdd(a.counters + 2, n_be);
        atomicAdd(a.counters + 7, n_work);
    }
}


// ===================================================================================================
// Group-per-frame kernels for lifted (layered) codes.
//
// Rows of H are split on the host into LAYERS of consecutive, mutually column-disjoint rows (for a lifted base graph a
// layer is a block row of Z circulant rows, capped at 32).  The rows of a layer read and write disjoint posteriors, so
// running them in parallel -- one lane per row -- gives exactly the values of the reference's sequential row loop
// (decode_nms.py:25-34).  A frame is owned by a group of G lanes (G = layer width rounded up to a power of two,
// 4..32), a warp holds 32/G frames, and the frame state (n posteriors + m row messages, float64) sits in shared
// memory: for Z = 32 that is 2.3 KB per FRAME instead of per THREAD, i.e. 32 resident warps instead of 2.
// ===================================================================================================
struct LdpcLayers {
    int nl, G, lgG;              // layers, lanes per frame (power of two), log2 G
    const int* layer_ptr;        // [nl+1] first row of each layer
};

// per-frame shared-memory footprint of the grouped kernels (bytes; E = 0 for the LLR-in decoder)
__host__ __device__ inline size_t ldpc_group_frame_bytes(int n, int m, int E, int kw, int nw) {
    size_t b = (size_t)(n + m) * 8 + (size_t)(kw + nw) * 4 + (size_t)E * 4;
    return (b + 15) & ~(size_t)15;
}

// Layered NMS on the frame state `st` (contiguous: posteriors [0,n), row messages [n,n+m)) by the group of this lane.
// `active` is group-uniform; all lanes of the warp must call this together.  Returns iters_used, parity via `ok`.
// W = 4: every row of H has exactly four ones (all lifts of the demo base graph): the four column indices of a row
// are one 16-byte load and the edge loops disappear; W = 0: general CSR rows.
// The sign product of decode_nms.py:30 is kept as (parity of negatives, any zero): alpha * sign is then +-alpha or 0,
// the same value the reference multiplies by the magnitude (the sign of a zero update cannot change any decision).
template <int W>
__device__ __forceinline__ int nms_decode_group(const LdpcCode& c, const LdpcLayers& L, double* st, int li, uint32_t gmask_shift,
                                                uint32_t gmask, bool active, int max_iter, double alpha, bool early_stop, bool& ok) {
    const int n = c.n;
    int used = max_iter;
    bool running = active, clean = false;
    auto syndrome_bad = [&]() {
        bool bad = false;
        for (int l = 0; l < L.nl; ++l) {
            const int r = __ldg(L.layer_ptr + l) + li;
            if (r < __ldg(L.layer_ptr + l + 1)) {
                if constexpr (W == 4) {
                    const int4 q = __ldg(reinterpret_cast<const int4*>(c.col_idx) + r);
                    bad |= ((st[q.x] < 0.0) != (st[q.y] < 0.0)) != ((st[q.z] < 0.0) != (st[q.w] < 0.0));
                } else {
                    uint32_t par = 0;
                    const int e1 = __ldg(c.row_ptr + r + 1);
                    for (int e = __ldg(c.row_ptr + r); e < e1; ++e) par ^= (uint32_t)(st[__ldg(c.col_idx + e)] < 0.0);
                    bad |= par != 0;
                }
            }
        }
        return ((__ballot_sync(0xffffffffu, bad) >> gmask_shift) & gmask) != 0;
    };
    for (int it = 1; it <= max_iter; ++it) {
        if (!__any_sync(0xffffffffu, running)) break;
        for (int l = 0; l < L.nl; ++l) {
            const int r = __ldg(L.layer_ptr + l) + li;
            if (running && r < __ldg(L.layer_ptr + l + 1)) {
                if constexpr (W == 4) {
                    const int4 q = __ldg(reinterpret_cast<const int4*>(c.col_idx) + r);
                    const double mr = st[n + r];
                    const double x0 = __dsub_rn(st[q.x], mr), x1 = __dsub_rn(st[q.y], mr);
                    const double x2 = __dsub_rn(st[q.z], mr), x3 = __dsub_rn(st[q.w], mr);
                    const bool neg = ((x0 < 0.0) != (x1 < 0.0)) != ((x2 < 0.0) != (x3 < 0.0));
                    const bool zero = (x0 == 0.0) || (x1 == 0.0) || (x2 == 0.0) || (x3 == 0.0);
                    const double mag = fmin(fmin(fabs(x0), fabs(x1)), fmin(fabs(x2), fabs(x3)));
                    const double as = zero ? 0.0 : (neg ? -alpha : alpha);
                    const double update = __dmul_rn(as, mag);
                    st[q.x] = __dadd_rn(x0, update); st[q.y] = __dadd_rn(x1, update);
                    st[q.z] = __dadd_rn(x2, update); st[q.w] = __dadd_rn(x3, update);
                    st[n + r] = update;
                } else {
                    const int e0 = __ldg(c.row_ptr + r), e1 = __ldg(c.row_ptr + r + 1);
                    if (e0 != e1) {
                        const double mr = st[n + r];
                        bool neg = false, zero = false;
                        double mag = __longlong_as_double(0x7ff0000000000000ll);
                        for (int e = e0; e < e1; ++e) {
                            const double ext = __dsub_rn(st[__ldg(c.col_idx + e)], mr);
                            neg ^= ext < 0.0;
                            zero |= ext == 0.0;
                            mag = fmin(mag, fabs(ext));
                        }
                        const double update = __dmul_rn(zero ? 0.0 : (neg ? -alpha : alpha), mag);
                        for (int e = e0; e < e1; ++e) {
                            double* p = st + __ldg(c.col_idx + e);
                            *p = __dadd_rn(__dsub_rn(*p, mr), update);
                        }
                        st[n + r] = update;
                    }
                }
            }
            __syncwarp();
        }
        if (early_stop || it == max_iter) {
            const bool bad = syndrome_bad();
            if (running) {
                clean = !bad;
                if (early_stop && clean) { used = it; running = false; }
            }
        }
    }
    if (max_iter <= 0) { const bool bad = syndrome_bad(); clean = !bad; }
    ok = clean;
    return used;
}

template <int W>
__global__ void ldpc_decode_group_kernel(LdpcCode c, LdpcLayers L, LdpcDecodeArgs a) {
    extern __shared__ double ldpc_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int G = L.G, fpw = 32 >> L.lgG, li = lane & (G - 1), fme = lane >> L.lgG;
    const uint32_t gshift = (uint32_t)(lane & ~(G - 1)), gmask = G >= 32 ? 0xffffffffu : ((1u << G) - 1u);
    const size_t fb = ldpc_group_frame_bytes(c.n, c.m, 0, 0, 0);
    double* st = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ldpc_smem) + ((size_t)warp * fpw + fme) * fb);
    const long long nbatch = (a.B + fpw - 1) / fpw;
    for (long long b = (long long)blockIdx.x * wpc + warp; b < nbatch; b += (long long)gridDim.x * wpc) {
        const long long f = b * fpw + fme;
        const bool valid = f < a.B;
        if (valid) {
            const double* row = a.llr + (size_t)f * a.in_len;
            if (a.in_len == c.n) { for (int i = li; i < c.n; i += G) st[i] = row[i]; }
            else { for (int i = li; i < c.n; i += G) st[i] = ldpc_derate_at(row, a.in_len, c.n, i); }
            for (int r = li; r < c.m; r += G) st[c.n + r] = 0.0;
        }
        __syncwarp();
        bool ok;
        const int used = nms_decode_group<W>(c, L, st, li, gshift, gmask, valid, a.max_iter, a.alpha, a.early_stop != 0, ok);
        if (valid) {
            if (a.hard) for (int i = li; i < c.n; i += G) a.hard[(size_t)f * c.n + i] = (uint8_t)(st[i] < 0.0);
            if (a.posterior) for (int i = li; i < c.n; i += G) a.posterior[(size_t)f * c.n + i] = st[i];
            if (li == 0) {
                if (a.iters) a.iters[f] = used;
                if (a.ok) a.ok[f] = (uint8_t)ok;
            }
        }
        __syncwarp();
    }
}

struct LdpcCrcTab { const unsigned long long* tab; int nq; };   // [nq][16] CRC contribution of payload nibble q with value v

template <int W>
__global__ void ldpc_sweep_group_kernel(LdpcCode c, LdpcGen g, LdpcLayers L, LdpcCrcTab ct, LdpcSweepArgs a) {
    extern __shared__ double ldpc_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int G = L.G, fpw = 32 >> L.lgG, li = lane & (G - 1), fme = lane >> L.lgG;
    const uint32_t gshift = (uint32_t)(lane & ~(G - 1)), gmask = G >= 32 ? 0xffffffffu : ((1u << G) - 1u);
    const int n = c.n, m = c.m, k = g.k, kw = g.kw, nw = (n + 31) / 32;
    const bool chan_only = a.llr_out != nullptr;
    const size_t fb = ldpc_group_frame_bytes(n, m, a.E, kw, nw);
    unsigned char* base = reinterpret_cast<unsigned char*>(ldpc_smem) + ((size_t)warp * fpw + fme) * fb;
    double* st = reinterpret_cast<double*>(base);
    uint32_t* msg = reinterpret_cast<uint32_t*>(base + (size_t)(n + m) * 8);
    uint32_t* cw = msg + kw;
    float* raw = reinterpret_cast<float*>(cw + nw);
    const uint2 key = make_uint2(a.k0, a.k1);
    unsigned long long n_frames = 0, n_fe = 0, n_be = 0, n_work = 0;
    const long long nbatch = (a.n_frames + fpw - 1) / fpw;
    for (long long b = (long long)blockIdx.x * wpc + warp; b < nbatch; b += (long long)gridDim.x * wpc) {
        const long long fi = b * fpw + fme;
        const bool valid = fi < a.n_frames;
        const long long fr = a.frame_begin + fi;
        // ---- payload words (Philox), zero the rest --------------------------------------------------------
        if (valid) {
            for (int j = li; j < kw + nw; j += G) msg[j] = 0;
        }
        __syncwarp();
        const int pwn = (a.kp + 31) / 32;
        if (valid) {
            for (int w4 = li * 4; w4 < pwn; w4 += G * 4) {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)(w4 >> 2), kPurposePayload), key);
                const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int w = w4 + q;
                    if (w < pwn) {
                        uint32_t v = rr[q];
                        const int rem = a.kp - w * 32;
                        if (rem < 32) v &= (1u << rem) - 1u;
                        msg[w] = v;                       // kp <= k: all payload words lie inside the message
                        if (a.payload_out)
                            for (int bb = 0; bb < 32 && w * 32 + bb < a.kp; ++bb)
                                a.payload_out[(size_t)fi * a.kp + w * 32 + bb] = (uint8_t)((v >> bb) & 1u);
                    }
                }
            }
        }
        __syncwarp();
        // ---- CRC (crc.py:19-37) as a XOR of nibble-table rows, reduced over the group ------------------------
        if (a.deg > 0) {
            unsigned long long reg = 0;
            if (valid)
                for (int q = li; q < ct.nq; q += G)
                    reg ^= __ldg(ct.tab + (size_t)q * 16 + ((msg[q >> 3] >> (4 * (q & 7))) & 15u));
            for (int o = 1; o < G; o <<= 1) reg ^= __shfl_xor_sync(0xffffffffu, reg, o);
            if (valid && li == 0) {
                for (int t = 0; t < a.deg && a.kp + t < k; ++t) {
                    const int j = a.kp + t;
                    msg[j >> 5] |= (uint32_t)((reg >> (a.deg - 1 - t)) & 1ull) << (j & 31);
                }
            }
            __syncwarp();
        }
        // ---- systematic encode: codeword = [message | G message] ------------------------------------------
        if (valid) {
            for (int j = li; j < kw; j += G) {
                uint32_t v = msg[j];
                const int keep = k - j * 32;
                if (keep < 32) v &= (1u << keep) - 1u;
                atomicOr(&cw[j], v);
            }
            for (int q = li; q < g.np; q += G) {
                uint32_t acc = 0;
                for (int j = 0; j < kw; ++j) acc ^= __ldg(g.G + (size_t)q * kw + j) & msg[j];
                const int pos = k + q;
                if (__popc(acc) & 1) atomicOr(&cw[pos >> 5], 1u << (pos & 31));
            }
        }
        __syncwarp();
        // ---- rate match, BPSK + AWGN, LLR: raw[t] (fp32, as in the thread-per-frame kernel) ---------------------
        if (valid) {
            for (int tb = li; tb * 4 < a.E; tb += G) {
                float z[4];
                normal4(philox4x32_10(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)tb, kPurposeNoise), key), z);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = tb * 4 + q;
                    if (t < a.E) {
                        const int pos = t % n;
                        const uint32_t bit = (cw[pos >> 5] >> (pos & 31)) & 1u;
                        const float llr = fmaf(a.sigma, z[q], 1.0f - 2.0f * (float)bit) * a.scale;
                        if (chan_only) a.llr_out[(size_t)fi * a.E + t] = (double)llr;
                        else raw[t] = llr;
                    }
                }
            }
        }
        __syncwarp();
        if (chan_only) continue;
        // ---- de-rate-match in the reference's summation order (rate_match.py:18-38), zero the row messages ---------
        if (valid) {
            const int reps = a.E / n, rem = a.E - reps * n;
            for (int i = li; i < n; i += G) {
                double v;
                if (a.E <= n) v = i < a.E ? (double)raw[i] : 0.0;
                else {
                    double s = (double)raw[i];
                    for (int r = 1; r < reps; ++r) s = __dadd_rn(s, (double)raw[r * n + i]);
                    double acc = __dadd_rn(0.0, s);
                    int cnt = reps;
                    if (i < rem) { acc = __dadd_rn(acc, (double)raw[reps * n + i]); ++cnt; }
                    v = __ddiv_rn(acc, (double)cnt);
                }
                st[i] = v;
            }
            for (int r = li; r < m; r += G) st[n + r] = 0.0;
        }
        __syncwarp();
        bool ok;
        const int used = nms_decode_group<W>(c, L, st, li, gshift, gmask, valid, a.max_iter, a.alpha, a.early_stop != 0, ok);
        // ---- payload errors, counters ------------------------------------------------------------------------
        uint32_t be = 0;
        if (valid)
            for (int j = li; j < a.kp; j += G) be += (uint32_t)(st[j] < 0.0) ^ ((msg[j >> 5] >> (j & 31)) & 1u);
        for (int o = 1; o < G; o <<= 1) be += __shfl_xor_sync(0xffffffffu, be, o);
        if (valid && li == 0) {
            n_frames += 1; n_fe += be > 0; n_be += be; n_work += (unsigned)used;
            if (a.frame_bit_errors) a.frame_bit_errors[fi] = (uint16_t)be;
            if (a.frame_work) a.frame_work[fi] = (uint16_t)(used > 65535 ? 65535 : used);
        }
        __syncwarp();
    }
    if (chan_only) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_frames += __shfl_xor_sync(0xffffffffu, n_frames, o);
        n_fe += __shfl_xor_sync(0xffffffffu, n_fe, o);
        n_be += __shfl_xor_sync(0xffffffffu, n_be, o);
        n_work += __shfl_xor_sync(0xffffffffu, n_work, o);
    }
    if (lane == 0 && n_frames) {
        atomicAdd(a.counters + 0, n_frames);
        atomicAdd(a.counters + 1, n_fe);
        atomicAdd(a.counters + 2, n_be);
        atomicAdd(a.counters + 7, n_work);
    }
}

}  // namespace pb
