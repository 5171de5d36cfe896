// polar_kernels.cuh -- __global__ kernels: batched decode (API), encoder, CRC.
#pragma once
#include "polar_decode.cuh"

namespace pb {

// Device-side tables of one engine (all read-only).
struct Tables {
    const int16_t* info_pos;     // [K] phase index of info bit j
    const uint32_t* info_mask;   // [16] device copy of Code::info_mask for dynamic indexing
    const uint32_t* crc_tab;     // [N/4][16] syndrome contribution of nibble value v at nibble position p
    const int16_t* rm_src;       // [N] NR: position in the de-rate-matched vector feeding internal LLR i, -1 = 0.0
    const int8_t* rm_cnt;        // [N] NR: number of transmitted copies combined into internal LLR i; 0 = never sent (-1.0), -1 = pad (0.0)
    int E;                       // NR transmitted length (0 = off)
};

struct DecodeArgs {
    const float* llr;            // [B, in_len]
    int64_t B;
    int in_len;
    const int8_t* force;         // [B,K] or null
    uint8_t* cand;               // [B,M,K]
    double* metrics;             // [B,M]
    float* info_llrs;            // [B,M,K]
    int32_t* n_cand;
    int32_t* best_idx;
    uint8_t* best_bits;          // [B,K]
    uint32_t* best_words;        // [B,XWn]
    uint8_t* crc_ok;
    uint32_t* flags;
    unsigned char* gscratch;     // per-warp global scratch (WarpMem::gbytes each)
};

// NR rate matching: stage the de-rate-matched + de-interleaved row of each of the warp's FPW frames in wm.chan
// (frame-interleaved like every staged channel row).  rate_match.py:19-39: mean of the repeats, -1.0 where nothing was
// sent; interleaver.py:26-37: gather through rm_src.
template <int MP, typename WM>
__device__ __forceinline__ void load_channel(const Code& code, const Tables& tb, const WM& wm, const float* llr,
                                             int in_len, int64_t frame0, int64_t B, int lane) {
    constexpr int FPW = 32 / MP;
    const int N = code.N;
    for (int e = lane; e < FPW * N; e += 32) {
        const int f = e & (FPW - 1), i = e / FPW;          // e = i * FPW + f: the stores are coalesced
        const int64_t frame = frame0 + f;
        float v = 0.f;
        if (frame < B) {
            const int p = tb.rm_src[i];
            if (p >= 0) {
                float acc = 0.f;
                int cnt = 0;
                for (int q = p; q < tb.E; q += N) { acc += llr[frame * (int64_t)in_len + q]; ++cnt; }
                v = cnt ? acc / (float)cnt : -1.0f;
            }
        }
        wm.chan[e] = v;
    }
    __syncwarp();
}

// CRC syndrome of u-hat through the nibble table; 0 <=> check_crc passes (crc.py:40-56).
template <int XW>
__device__ __forceinline__ uint32_t crc_syndrome(const Code& code, const Tables& tb, const uint32_t (&u)[XW]) {
    uint32_t syn = 0;
    const int nn = code.N >= 4 ? code.N / 4 : 1;
#pragma unroll
    for (int w = 0; w < XW; ++w) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int nib = w * 8 + k;
            if (nib < nn) syn ^= __ldg(&tb.crc_tab[nib * 16 + ((u[w] >> (4 * k)) & 15u)]);
        }
    }
    return syn;
}

// Write information bits [j0, j1) of the word stashed at stash[w*32] (w = phase/32) as bytes dst[j].
__device__ __forceinline__ void write_info_bits(const Code& code, const Tables& tb, const float* stash, uint8_t* dst,
                                                int j0 = 0, int j1 = -1) {
    const int K = (j1 < 0) ? code.K : j1;
    const bool al = ((reinterpret_cast<uintptr_t>(dst + j0) & 3) == 0);
    int j = j0;
    for (; j + 4 <= K && al; j += 4) {
        uint32_t w = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int pos = __ldg(&tb.info_pos[j + t]);
            w |= ((__float_as_uint(stash[(pos >> 5) * 32]) >> (pos & 31)) & 1u) << (8 * t);
        }
        *reinterpret_cast<uint32_t*>(dst + j) = w;
    }
    for (; j < K; ++j) {
        const int pos = __ldg(&tb.info_pos[j]);
        dst[j] = (uint8_t)((__float_as_uint(stash[(pos >> 5) * 32]) >> (pos & 31)) & 1u);
    }
}

// Build the per-frame force masks over phases from force[frame, 0..K) (scl.py:126-144).
template <int XW>
__device__ __forceinline__ void load_force(const Code& code, const int8_t* force, int64_t frame, bool valid,
                                           uint32_t (&fmask)[XW], uint32_t (&fval)[XW], uint32_t& flags) {
    int j = 0;
#pragma unroll
    for (int w = 0; w < XW; ++w) {
        fmask[w] = 0; fval[w] = 0;
        if (w * 32 < code.N) {
            for (int b = 0; b < 32 && w * 32 + b < code.N; ++b) {
                if ((code.info_mask[w] >> b) & 1u) {
                    if (valid && force != nullptr) {
                        const int v = force[frame * (int64_t)code.K + j];
                        if (v == 0 || v == 1) { fmask[w] |= 1u << b; fval[w] |= (uint32_t)v << b; }
                        else if (v != -1) flags |= 4u;
                    }
                    ++j;
                }
            }
        }
    }
}

// info-mask words of the plain list decode kernels: 0 = measured best (chain of selects on laundered words for every list
// size since v19; until v18 M = 4 preferred the indexed local copy of the parameters, form 1, which costs 88 bytes of local
// memory per thread), 1 / 2 force one form
#ifndef PB_DECODE_UMASK
#define PB_DECODE_UMASK 0
#endif
#ifndef PB_LIST_THREADS
#define PB_LIST_THREADS 1024
#endif
// N <= 128: compiled for 1024 threads per CTA, i.e. 64 registers -> 32 resident warps per SM (no spills for the plain
// kernels, a few dozen bytes for the forced ones); N = 256 / 512 keep 16 partial-sum words per path and get 128 registers.
// TRACE (instantiated for FORCED kernels only; a.force may then be null): the list decode records the leaf-LLR trace and
// info_llrs[B,M,K] (scl.py:159,167,203-209) of all M paths is read off it -- K shared-memory reads and K/MP loads per
// path instead of one SC replay per path.
template <int MP, int LOGMAX, bool FORCED, bool METRIC, int NS = 0, bool TRACE = false, int HS = DefaultHS<MP>::value>
__global__ void __launch_bounds__(LOGMAX <= 7 ? (MP > 1 ? PB_LIST_THREADS : 1024) : 512) decode_kernel(const Code code_, const Tables tb, const DecodeArgs a) {
    const Code code = with_static_n<NS>(code_);
    using Dec = ListDecoder<MP, LOGMAX, FORCED, METRIC, HS>;
    using WM = WarpMem<MP, HS>;
    using PathT = typename Dec::PathT;
    constexpr int FPW = 32 / MP, XW = PathT::XW;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    WM wm;
    wm.carve(smem + (size_t)warp * WM::bytes(code.N, 0, TRACE ? code.K : 0), WM::warp_scratch(a.gscratch, code.N), code.N, 0,
             TRACE ? WM::warp_trace(a.gscratch, code.N, code.K) : nullptr, TRACE ? code.K : 0);
    const int K = code.K, M = code.M;
    const int xwn = code.N >= 32 ? code.N / 32 : 1;
    const int64_t ngroups = (a.B + FPW - 1) / FPW;
    for (int64_t g = (int64_t)blockIdx.x * wpc + warp; g < ngroups; g += (int64_t)gridDim.x * wpc) {
        const int64_t frame0 = g * FPW;
        const int64_t frame = frame0 + lane / MP;
        const bool valid = frame < a.B;
        if (tb.E == 0) {
            // whole rows by 16-byte cp.async when the tile fits (N = 128, MP >= 4: one DRAM round trip per frame group),
            // else 32 columns at a time
#ifndef PB_DECODE_ASYNC_ROWS
#define PB_DECODE_ASYNC_ROWS 1
#endif
            bool staged = false;
            if (PB_DECODE_ASYNC_ROWS) staged = stage_channel_rows_async<MP>(wm, code.N, lane, valid ? a.llr + frame * (int64_t)a.in_len : nullptr);
            if (!staged) {
                if (frame0 + FPW <= a.B && (code.N & 31) == 0) stage_channel_block<MP>(wm, code.N, lane, a.llr + frame0 * (int64_t)code.N);
                else stage_channel_rows<MP>(wm, code.N, lane, [&](int f) -> const float* {
                    return frame0 + f < a.B ? a.llr + (frame0 + f) * (int64_t)a.in_len : nullptr; });
            }
        } else load_channel<MP, WM>(code, tb, wm, a.llr, a.in_len, frame0, a.B, lane);
        const float* chanf = wm.chan + lane / MP;
        uint32_t flags = 0;
        uint32_t fmask[XW], fval[XW];
        if constexpr (FORCED) load_force<XW>(code, a.force, frame, valid, fmask, fval, flags);
        PathT p;
        Dec::init(p, lane, valid);
        Dec::template run<TRACE, false, (MP > 1 && !FORCED && !TRACE && METRIC) ? (PB_DECODE_UMASK ? PB_DECODE_UMASK : 2) : 0>(code, tb.info_mask, wm, p, lane, chanf, fmask, fval, flags);

        // u-hat = x-hat * F^{(x)n}
        uint32_t u[XW];
#pragma unroll
        for (int k = 0; k < XW; ++k) u[k] = p.alive ? p.xh[k] : 0u;
        transform_words<XW>(u, code.n);
        const bool pass = p.alive && (code.crc_deg == 0 || crc_syndrome<XW>(code, tb, u) == 0);

        // candidate order is metric order; first CRC pass wins, else index 0 (scl.py:190-197)
        uint32_t v = (pass && code.crc_deg > 0) ? p.r : 0xffu;
#pragma unroll
        for (int o = 1; o < MP; o <<= 1) v = min(v, __shfl_xor_sync(kFull, v, o));
        const uint32_t best_r = (v == 0xffu) ? 0u : v;
        const uint32_t gmask = (__ballot_sync(kFull, p.alive) >> (lane & ~(MP - 1))) & Dec::GM;
        uint32_t fl = flags;
#pragma unroll
        for (int o = 1; o < MP; o <<= 1) fl |= __shfl_xor_sync(kFull, fl, o);

        if constexpr (TRACE) {
            if (a.info_llrs != nullptr) {
                // one lineage walk per path of the group: candidate p.r of the frame = the path that ended in slot s
#pragma unroll
                for (int s = 0; s < MP; ++s) {
                    const int end_lane = (lane & ~(MP - 1)) + s;
                    const bool al = __shfl_sync(kFull, (int)p.alive, end_lane) != 0;
                    const uint32_t rs = __shfl_sync(kFull, p.r, end_lane);
                    float* dst = a.info_llrs + ((frame * M + rs) * (int64_t)K);
                    Dec::trace_walk(code, wm, lane, end_lane, [&](int j, float L) { if (al && valid) dst[j] = L; });
                }
                __syncwarp();
            }
        }
        // stash u-hat words in the (now dead) tree area of the own slot for dynamic bit addressing
        float* stash = wm.scr + lane;
#pragma unroll
        for (int k = 0; k < XW; ++k) if (k < xwn) stash[k * 32] = __uint_as_float(u[k]);
        if (a.best_bits) {
            // the MP lanes of a group share the unpacking of the winning word: lane s writes bytes [s*q, (s+1)*q)
            const uint32_t bm = (__ballot_sync(kFull, p.alive && p.r == best_r) >> (lane & ~(MP - 1))) & Dec::GM;
            const int bl = (lane & ~(MP - 1)) + (bm ? __ffs(bm) - 1 : 0);
            __syncwarp();
            const int q = ((K + MP - 1) / MP + 3) & ~3, s = lane & (MP - 1);
            if (valid && s * q < K) write_info_bits(code, tb, wm.scr + bl, a.best_bits + frame * (int64_t)K, s * q, min(K, (s + 1) * q));
        }
        if (p.alive) {
            if (a.cand) write_info_bits(code, tb, stash, a.cand + (frame * M + p.r) * (int64_t)K);
            if (a.metrics) a.metrics[frame * M + p.r] = p.m;
            if (p.r == best_r) {
                if (a.best_words) for (int k = 0; k < xwn; ++k) a.best_words[frame * xwn + k] = __float_as_uint(stash[k * 32]);
                if (a.best_idx) a.best_idx[frame] = (int32_t)best_r;
                if (a.crc_ok) a.crc_ok[frame] = (uint8_t)(code.crc_deg > 0 ? pass : 0);
                if (a.n_cand) a.n_cand[frame] = __popc(gmask);
                if (a.flags) a.flags[frame] = fl;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Encoder (polar.py:106-119): one thread per frame, msg[B,K] u8 -> code[B,N] u8.
// ---------------------------------------------------------------------------
static __global__ void encode_kernel(const Code code, const Tables tb, const uint8_t* __restrict__ msg, uint8_t* __restrict__ out, int64_t B) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= B) return;
    uint32_t u[kMaxWords];
#pragma unroll
    for (int k = 0; k < kMaxWords; ++k) u[k] = 0;
    for (int j = 0; j < code.K; ++j) {
        const int pos = tb.info_pos[j];
        const uint32_t b = msg[f * code.K + j] & 1u;
#pragma unroll
        for (int k = 0; k < kMaxWords; ++k) if (k == (pos >> 5)) u[k] |= b << (pos & 31);
    }
    transform_words<kMaxWords>(u, code.n);
    for (int i = 0; i < code.N; ++i) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < kMaxWords; ++k) if (k == (i >> 5)) w = u[k];
        out[f * code.N + i] = (uint8_t)((w >> (i & 31)) & 1u);
    }
}

// ---------------------------------------------------------------------------
// CRC (crc.py:19-56): bit-serial long division, MSB first, zero init, one thread per frame.
// poly has degree `deg` (<= 63) with the leading 1 at bit `deg`.
// ---------------------------------------------------------------------------
static __global__ void crc_kernel(unsigned long long poly, int deg, const uint8_t* __restrict__ msg, int L, int64_t B,
                           uint8_t* __restrict__ out_attach, uint8_t* __restrict__ out_ok) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= B) return;
    const unsigned long long low = poly & ((deg >= 64) ? ~0ull : ((1ull << deg) - 1ull));
    unsigned long long reg = 0;  // remainder register, deg bits
    if (out_attach) {
        // remainder of msg(x) * x^deg
        for (int i = 0; i < L; ++i) {
            const unsigned long long b = msg[f * L + i] & 1u;
            const unsigned long long top = ((reg >> (deg - 1)) & 1ull) ^ b;
            reg = (reg << 1) & ((1ull << deg) - 1ull);
            if (top) reg ^= low;
            out_attach[f * (L + deg) + i] = (uint8_t)b;
        }
        for (int t = 0; t < deg; ++t) out_attach[f * (L + deg) + L + t] = (uint8_t)((reg >> (deg - 1 - t)) & 1ull);
    } else {
        // remainder of msg(x) itself: feed the bits through the same register; the last deg bits enter unmultiplied
        for (int i = 0; i < L; ++i) {
            const unsigned long long b = msg[f * L + i] & 1u;
            const unsigned long long top = (reg >> (deg - 1)) & 1ull;
            reg = ((reg << 1) | b) & ((1ull << deg) - 1ull);
            if (top) reg ^= low;
        }
        out_ok[f] = (uint8_t)(reg == 0);
    }
}

}  // namespace pb
