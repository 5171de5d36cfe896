// polar_core.cuh -- device-side SC / SCL decoder core for sm_100a.
//
// Work mapping ("thread per path"): a frame is decoded by a group of MP
// adjacent lanes (MP = list capacity rounded up to a power of two, 1..8), so a
// warp holds FPW = 32/MP frames and never synchronises with other warps.  All
// frames follow the same static SC schedule (the frozen pattern is a property
// of the code, not of the frame), so the warp runs divergence-free.
//
// What replaces the reference's per-fork deep copies
// (dl_scl_polar/polar/scl.py:52-62, 77 % of its run time):
//   * LLR tree, lane-interleaved: element (height h, index i) of slot `lane`
//     lives at base[(2^h-2+i)*32 + lane], so every access of a warp is one
//     conflict-free / fully coalesced 128 B row whatever slot each path points
//     to.  Heights < HS sit in shared memory, heights >= HS in an L2-resident
//     global scratch (WarpMem), the height-1 pair of a phase pair in registers.
//   * Lazy copy: a path keeps one 4-bit slot pointer per tree height (word P).
//     Because every live path recomputes heights <= c in the same phase, a
//     path always writes its OWN slot and a fork copies one register.
//   * Partial sums (scl.py:84-99) are bit-packed: for each height h the 2^h
//     bits of the finished left child wait in a register field; the upward
//     XOR propagation is a handful of shifts per decided bit.
//   * u-hat is never stored: after the last phase the propagated word is the
//     re-encoded codeword x^, and u^ = x^ * F^{(x)n} (F^{(x)n} is an involution).
//
// Reference semantics kept bit-for-bit (scl.py:108-209): both children of a
// free bit are scored with the exact softplus metric (scl.py:102-105) and the
// list is truncated to the M best with the reference's stable order (parent
// rank, bit).  The stable order is obtained by ranking unique 64-bit keys =
// (IEEE bits of the fp64 metric & ~15) | (2*rank+bit).  Metrics accumulate in
// fp64; LLR arithmetic is fp32 (f exact, g one rounding).  Frames in which two
// competing metrics come within ~1e-6 relative are flagged (PB_FLAG_NEAR_TIE)
// -- these are the only frames allowed to differ from the float64 reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PB_FLAG_NEAR_TIE 1u      // two competing path metrics within ~1e-6 relative at some prune
#define PB_FLAG_RANK_TIE 2u      // DL-SCL: best and runner-up flip scores within ~1e-6 relative

namespace pb {

constexpr int kMaxLog = 9;            // N <= 512
constexpr int kMaxWords = 16;         // N/32
constexpr uint32_t kFull = 0xffffffffu;

// Code description handed to every kernel by value (lives in the constant bank).
// Kernels instantiated with a static code length (template parameter NS = log2 N, 0 = any) replace N and n by
// compile-time constants: the schedule arithmetic of the phase loop (which tree height a phase recomputes, whether it
// reads the channel rows, the width of the partial-sum words) folds away for the headline geometry N = 128.
struct Code {
    int N, n, K, M;                   // code length, log2, info bits, list size (M <= MP)
    int crc_deg;                      // 0 = no CRC
    uint32_t info_mask[kMaxWords];    // bit phi set <=> phase phi is an information bit (index it with STATIC indices only)
};

template <int NS> __device__ __forceinline__ Code with_static_n(const Code& c) {
    Code r = c;
    if constexpr (NS > 0) { r.N = 1 << NS; r.n = NS; }
    return r;
}

#ifndef PB_F_XORSIGN
#define PB_F_XORSIGN 1
#endif
__device__ __forceinline__ float f_op(float a, float b) {
    // polar.py:122-123 : sign(a) sign(b) min(|a|,|b|)   (exact in fp32)
    // One instruction: min.xorsign.abs (FMNMX.XORSIGN |a|, |b|) takes the smaller magnitude and the XOR of the two sign bits --
    // bit for bit what fminf(|a|,|b|) with the sign OR-ed in gives (three instructions), including -0 for a zero operand.
#if PB_F_XORSIGN
    float r;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
#else
    float mn = fminf(fabsf(a), fabsf(b));
    uint32_t s = (__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u;
    return __uint_as_float(__float_as_uint(mn) | s);
#endif
}
__device__ __forceinline__ float g_op(float a, float b, uint32_t bit) {
    // polar.py:126-127 : b + (1-2c) a
    return b + __uint_as_float(__float_as_uint(a) ^ (bit << 31));
}
// The same with the partial-sum bit taken straight from its packed word: bit `pos` of `word` is moved onto the sign
// position (one shift), masked and XOR-ed into a in one LOP3 -- three instructions per g instead of five.
__device__ __forceinline__ float g_op_packed(float a, float b, uint32_t word, int pos) {
    return b + __uint_as_float(__float_as_uint(a) ^ ((word << (31 - pos)) & 0x80000000u));
}

// log(1+exp(-|L|)) -- the part both softplus branches share (scl.py:102-105; numpy logaddexp = max + log1p(exp(-|d|))).
// t = exp(-|L|) in (0,1]; log1p(t) = t*q(t) with q a degree-9 near-minimax polynomial of log1p(t)/t on [0,1]
// (Chebyshev-node fit, max relative error 5e-9; 1.6e-7 after fp32 Horner) -- as accurate as log1pf(expf()) in
// fp32 (2.9e-7) at a third of the instructions; the t*q form keeps full relative accuracy as t -> 0.
__device__ __forceinline__ float softplus_tail(float L) {
    // t = exp(-|L|) as one FMUL + MUFU.EX2 (ex2.approx.ftz: results below 2^-126, i.e. |L| > 87.3, flush to zero -- the
    // reference's float64 tail is < 1.2e-38 there; two metrics that differ only by such tails compare as a tie here and
    // the frame is flagged PB_FLAG_NEAR_TIE like any other near-tie)
    float t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(L) * -1.4426950408889634f));
    float q = -3.176057010e-03f;
    q = fmaf(q, t, 1.954252722e-02f);
    q = fmaf(q, t, -5.637361275e-02f);
    q = fmaf(q, t, 1.054362379e-01f);
    q = fmaf(q, t, -1.526966707e-01f);
    q = fmaf(q, t, 1.966327426e-01f);
    q = fmaf(q, t, -2.495161626e-01f);
    q = fmaf(q, t, 3.332971050e-01f);
    q = fmaf(q, t, -4.999989265e-01f);
    q = fmaf(q, t, 9.999999947e-01f);
    return q * t;
}


// ---------------------------------------------------------------------------
// Per-warp memory view.
//   shared : tree heights 2..HS-1 (lane-interleaved), rank-exchange area, DL-SCL |L0| rows
//   global : tree heights HS..n-1 (same lane-interleaved layout: one 128 B line per warp access) and the
//            staged channel rows when the LLRs are generated / de-rate-matched on the fly.  The scratch of all
//            resident warps (~16 KB each) stays L2-resident; moving the two tall heights out of shared memory is
//            what lifts occupancy from 11 to 32 warps per SM for N = 128.
// ---------------------------------------------------------------------------
// Heights >= HS live in the global scratch.  HS = 5 (heights 5.. global, 30 rows = 3.8 KB of shared memory per warp)
// lets the register count (64), not shared memory, set the occupancy: 32 warps per SM.  Because the tall levels are
// evaluated depth-first (Tree::produce) the global rows are written once and read once per use.
#ifndef PB_DEFAULT_HS
#define PB_DEFAULT_HS 5
#endif
// Channel LLRs of a warp's FPW frames are ALWAYS staged in the warp's global scratch, frame-interleaved:
// element i of frame f at chan[i * FPW + f].  The MP lanes of a group read the same word and the FPW groups read FPW
// consecutive words, so a warp-wide channel load is one 32-byte sector (one L1 wavefront) instead of one line per
// frame -- reading the caller's rows in place cost a third of all L1 wavefronts of the list kernels, which is what
// bounds them (l1tex data-pipe wavefronts 85 % of peak, profiles/r01_v9_*).
__host__ __device__ inline int chan_stride(int N) { return N + 4; }       // floats reserved per frame (sizing only)

template <int MP> struct DefaultHS { static constexpr int value = PB_DEFAULT_HS; };

__host__ __device__ inline int tree_rows_shared(int N, int hs) {
    const int all = N >= 6 ? N - 2 : 4;                 // rows of heights 1..n-1 (at least 4 rows of scratch)
    const int cap = (1 << hs) - 2;
    return all < cap ? all : cap;
}
__host__ __device__ inline int tree_rows_global(int N, int hs) {
    const int all = N >= 6 ? N - 2 : 4;
    const int cap = (1 << hs) - 2;
    return all > cap ? all - cap : 0;
}

// exchange area of a warp: candidate keys [2][32][2] u64 (one buffer for even, one for odd phases) + the rank table of
// the prune: one word per rank and group, (2*MP + 1) * (32/MP) <= 80 words (see ListDecoder::run)
constexpr int kXchgBytes = 2 * 32 * 16 + 80 * 4;

template <int MP, int HS = DefaultHS<MP>::value>
struct WarpMem {
    float* ts;            // shared tree base:  element (h,i), h <  HS, at ts[((2^h-2)+i)*32 + lane]
    float* tg;            // global tree base (pre-offset): element (h,i), h >= HS, at tg[((2^h-2)+i)*32 + lane]
    float* chan;          // [N][FPW] staged channel LLRs, frame-interleaved (global scratch)
    float* scr;           // >= 3*max(N/32,1) lane-interleaved rows of scratch for the encoder / bit stash
    unsigned long long* xchg;  // [32][2] candidate keys for the rank exchange / small per-frame scratch (shared)
    float* absl;          // [FPW][absl_stride(xk)] DL-SCL only: |L0| of the reference path (flip.py:102) (shared)
    uint32_t* lin;        // [tk][4] trace mode only: ballot words (bit planes 0..2) of the slot each surviving path came
                          // from at info phase j (shared; written by lane 0)
    float* hist;          // [K][32] trace mode only: leaf LLR every slot saw at info phase j (global scratch)
    static constexpr int FPW = 32 / MP;
    // Shared layout of a warp, all offsets compile-time constants (the hot loop re-derives these pointers from the warp
    // base whenever registers are short -- with N-dependent sizes that re-derivation was ~8 % of all instructions):
    //   [0, kTreeBytes) tree rows of heights < HS (always the full 2^HS - 2 rows)   [kTreeBytes, +kXchgBytes) xchg
    //   then the lineage words (trace kernels), then the |L0| rows unless they alias the tree area
    static constexpr int kTreeRows = (1 << HS) - 2;
    static constexpr size_t kTreeBytes = (size_t)kTreeRows * 32 * 4;
    __host__ __device__ static size_t tree_bytes(int) { return kTreeBytes; }
    __host__ __device__ static size_t lin_bytes(int tk) { return (tk && MP > 1) ? (size_t)tk * 16 : 0; }
    // The |L0| rows of the DL-SCL retry kernel are only live BETWEEN two list decodes (trace walk -> beta scoring), when
    // the tree rows are dead: they alias the shared tree area behind the stash rows whenever they fit there.
    // Row stride of the |L0| rows: the FP64 tensor-core scoring (MP >= 4) reads element i0 + lane % 4 of frame lane / MP,
    // conflict-free when consecutive frames are 4 banks apart (stride = 4 mod 32); the scalar scoring of MP < 4 reads the
    // same element of all FPW frames (odd stride).
    __host__ __device__ static int absl_stride(int xk) { return MP >= 4 ? (xk <= 4 ? 4 : ((xk - 4 + 31) / 32) * 32 + 4) : xk + 1; }
    __host__ __device__ static size_t absl_bytes(int xk) { return xk ? (((size_t)FPW * absl_stride(xk) * 4 + 15) & ~(size_t)15) : 0; }
    // (the head of the tree area doubles as the bit stash of the output stage and as the coalescing tile of
    //  stage_channel_rows, both of which may run while the |L0| rows are live)
    static constexpr size_t kTileBytes = (size_t)FPW * (32 + MP) * 4;
    __host__ __device__ static size_t stash_bytes(int N) {
        const size_t st = (size_t)3 * (N >= 32 ? N / 32 : 1) * 32 * 4;
        return st > kTileBytes ? st : kTileBytes;
    }
    __host__ __device__ static bool absl_aliases_tree(int N, int xk) {
        return xk && kTreeBytes >= stash_bytes(N) + absl_bytes(xk);
    }
    // shared bytes per warp: xk > 0 adds the |L0| rows, tk > 0 the lineage words of the trace (info_llrs without a replay)
    __host__ __device__ static size_t bytes(int N, int xk = 0, int tk = 0) {
        size_t x = absl_aliases_tree(N, xk) ? 0 : absl_bytes(xk);
        return kTreeBytes + kXchgBytes + lin_bytes(tk) + x;
    }
    __host__ __device__ static size_t gbytes(int N) {                     // global scratch bytes per warp (tree + channel rows)
        size_t t = (size_t)tree_rows_global(N, HS) * 32 * 4;
        size_t ch = (((size_t)FPW * chan_stride(N) * 4) + 127) & ~(size_t)127;
        return t + ch;
    }
    __host__ __device__ static size_t hbytes(int K) { return (size_t)K * 32 * 4; }   // trace rows per warp
    // The scratch buffer of a launch holds the tree/channel areas of all resident warps first (dense, so that the
    // plain kernels' working set stays L2-resident) and the trace rows of all warps behind them.
    static __device__ unsigned char* warp_scratch(unsigned char* scratch, int N) {
        return scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * gbytes(N);
    }
    static __device__ unsigned char* warp_trace(unsigned char* scratch, int N, int K) {
        const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
        return scratch + warps * gbytes(N) + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * hbytes(K);
    }
    __device__ void carve(unsigned char* sbase, unsigned char* gbase, int N, int xk = 0, unsigned char* hbase = nullptr, int tk = 0) {
        ts = reinterpret_cast<float*>(sbase);
        xchg = reinterpret_cast<unsigned long long*>(sbase + kTreeBytes);
        lin = reinterpret_cast<uint32_t*>(sbase + kTreeBytes + kXchgBytes);
        absl = absl_aliases_tree(N, xk) ? reinterpret_cast<float*>(sbase + stash_bytes(N))
                                        : reinterpret_cast<float*>(sbase + kTreeBytes + kXchgBytes + lin_bytes(tk));
        float* g = reinterpret_cast<float*>(gbase);
        tg = g - ((1 << HS) - 2) * 32;
        chan = g + (size_t)tree_rows_global(N, HS) * 32;
        hist = reinterpret_cast<float*>(hbase);
        const int need = 3 * (N >= 32 ? N / 32 : 1);
        scr = (kTreeRows >= need) ? ts : g;
    }
};

// Stage the channel rows of the warp's FPW frames into wm.chan (frame-interleaved, see chan_stride above).
// row_of(f) -> pointer to the N floats of frame f of this warp, or nullptr (row of zeros); must be warp-uniform per f.
// 32 columns at a time through a shared-memory tile [FPW][32 + MP] (the tree rows are dead before phase 0): rows are
// read coalesced (one 128 B line per frame and chunk), the tile is read column-wise without bank conflicts
// (bank = f * MP + column offset covers all 32 banks) and written out as full 128 B lines: 4 L1 wavefronts per 32
// staged floats.
template <int MP, typename WM, typename RowOf>
__device__ __forceinline__ void stage_channel_rows(const WM& wm, int N, int lane, RowOf&& row_of) {
    constexpr int FPW = 32 / MP, TS = 32 + MP;
    static_assert((size_t)FPW * TS * 4 <= WM::kTreeBytes + kXchgBytes, "staging tile must fit the tree + exchange area");
    float* tile = wm.ts;
    for (int c0 = 0; c0 < N; c0 += 32) {
        const int C = N - c0 < 32 ? N - c0 : 32;
#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            const float* r = row_of(f);
            tile[f * TS + lane] = (r != nullptr && lane < C) ? r[c0 + lane] : 0.f;
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < FPW; ++it) {
            const int t = it * 32 + lane, f = t & (FPW - 1), di = t / FPW;
            if (di < C) wm.chan[(c0 + di) * FPW + f] = tile[f * TS + di];
        }
        __syncwarp();
    }
}
// The same for whole rows at once, when the tile [FPW][N + MP] fits the tree + exchange area (N = 128: MP >= 4) and every
// row is 16-byte aligned: each row arrives with ONE 16-byte cp.async per lane (global -> shared, no registers, all FPW
// rows in flight together -- one DRAM round trip instead of one per 32 columns), then the tile is written out
// frame-interleaved as full 128 B lines.  Returns false (nothing done) when the conditions do not hold.
template <int MP, typename WM>
__device__ __forceinline__ bool stage_channel_rows_async(const WM& wm, int N, int lane, const float* row /* this lane's frame */) {
    constexpr int FPW = 32 / MP;
    const int TS = N + MP;
    if ((size_t)FPW * TS * 4 > WM::kTreeBytes + kXchgBytes || (N & 127) != 0) return false;
    const bool ok = row == nullptr || (reinterpret_cast<uintptr_t>(row) & 15) == 0;
    if (!__all_sync(kFull, ok)) return false;
    float* tile = wm.ts;
    for (int c0 = 0; c0 < N; c0 += 128) {
#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            const float* r = reinterpret_cast<const float*>(__shfl_sync(kFull, (unsigned long long)reinterpret_cast<uintptr_t>(row), f * MP));
            float* d = tile + f * TS + c0 + 4 * lane;
            if (r != nullptr) {
                const unsigned int sa = (unsigned int)__cvta_generic_to_shared(d);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(r + c0 + 4 * lane) : "memory");
            } else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    for (int t = lane; t < FPW * N; t += 32) {
        const int f = t & (FPW - 1), di = t / FPW;
        wm.chan[t] = tile[f * TS + di];                       // chan[di * FPW + f]; bank of the read = f * MP + di (mod 32): conflict-free
    }
    __syncwarp();
    return true;
}
// The common case: the FPW rows are CONTIGUOUS in memory (frames g*FPW .. g*FPW+FPW-1 of a [B, N] buffer), all valid and
// N is a multiple of 32 -- no per-row pointers, no guards.
template <int MP, typename WM>
__device__ __forceinline__ void stage_channel_block(const WM& wm, int N, int lane, const float* rows) {
    constexpr int FPW = 32 / MP, TS = 32 + MP;
    float* tile = wm.ts;
    const float* src = rows + lane;
    float* dst = wm.chan + lane;
    const float* trd = tile + (lane & (FPW - 1)) * TS + lane / FPW;       // tile[f * TS + di], di = it * MP + lane / FPW
    for (int c0 = 0; c0 < N; c0 += 32) {
#pragma unroll
        for (int f = 0; f < FPW; ++f) tile[f * TS + lane] = src[(size_t)f * N + c0];
        __syncwarp();
#pragma unroll
        for (int it = 0; it < FPW; ++it) dst[c0 * FPW + it * 32] = trd[it * MP];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Bit-packed partial sums.  Height h holds 2^h bits:
//   h = 0..4 -> word 0, bit offset 2^h - 1 ; h = 5 -> word 1 ; h = 6 -> words 2,3 ;
//   h = 7 -> words 4..7 ; h = 8 -> words 8..15.
// ---------------------------------------------------------------------------
template <int LOGMAX> struct BitsCfg { static constexpr int BW = (LOGMAX <= 7) ? 4 : 16; static constexpr int XW = (1 << LOGMAX) / 32 > 0 ? (1 << LOGMAX) / 32 : 1; };

template <int H, int BW>
__device__ __forceinline__ uint32_t left_bit(const uint32_t (&bw)[BW], int i) {
    // bit i of the left-child buffer at height H (H < 5: static field of word 0)
    if constexpr (H < 5) return (bw[0] >> (((1 << H) - 1) + i)) & 1u;
    else return 0;  // H >= 5 handled word-wise by the callers
}
// g with bit i of the left-child buffer at height H < 5 (static position in word 0)
template <int H, int BW>
__device__ __forceinline__ float g_left(float a, float b, const uint32_t (&bw)[BW], int i) {
    return g_op_packed(a, b, bw[0], ((1 << H) - 1) + i);
}

// Upward propagation after deciding `bit` at a phase with T trailing ones
// (scl.py:84-99).  cur (2^T bits) is returned in cw[]; the caller stores it as
// the left buffer of height T, or takes it as x^ when T == n.
template <int T, int BW, int CW>
__device__ __forceinline__ void ascend(const uint32_t (&bw)[BW], uint32_t bit, uint32_t (&cw)[CW]) {
    uint32_t cur = bit;
#pragma unroll
    for (int h = 0; h < (T < 5 ? T : 5); ++h) {
        uint32_t L = (bw[0] >> ((1 << h) - 1)) & ((h == 5) ? 0xffffffffu : ((1u << (1 << h)) - 1u));
        cur = (L ^ cur) | (cur << (1 << h));
    }
    cw[0] = cur;
    if constexpr (T >= 6) {           // height 5 buffer = word 1
        cw[0] = bw[1] ^ cur; cw[1] = cur;
    }
    if constexpr (T >= 7) {           // height 6 buffer = words 2,3
        uint32_t c0 = cw[0], c1 = cw[1];
        cw[0] = bw[2] ^ c0; cw[1] = bw[3] ^ c1; cw[2] = c0; cw[3] = c1;
    }
    if constexpr (T >= 8 && BW >= 8) {  // height 7 buffer = words 4..7
        uint32_t c[4] = {cw[0], cw[1], cw[2], cw[3]};
#pragma unroll
        for (int k = 0; k < 4; ++k) { cw[k] = bw[4 + k] ^ c[k]; cw[4 + k] = c[k]; }
    }
    if constexpr (T >= 9 && BW >= 16) {  // height 8 buffer = words 8..15
        uint32_t c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = cw[k];
#pragma unroll
        for (int k = 0; k < 8; ++k) { cw[k] = bw[8 + k] ^ c[k]; cw[8 + k] = c[k]; }
    }
}

template <int T, int BW, int CW>
__device__ __forceinline__ void store_height(uint32_t (&bw)[BW], const uint32_t (&cw)[CW]) {
    if constexpr (T < 5) {
        const uint32_t mask = ((1u << (1 << T)) - 1u) << ((1 << T) - 1);
        bw[0] = (bw[0] & ~mask) | (cw[0] << ((1 << T) - 1));
    } else if constexpr (T == 5) { bw[1] = cw[0]; }
    else if constexpr (T == 6) { bw[2] = cw[0]; bw[3] = cw[1]; }
    else if constexpr (T == 7 && BW >= 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) bw[4 + k] = cw[k];
    } else if constexpr (T == 8 && BW >= 16) {
#pragma unroll
        for (int k = 0; k < 8; ++k) bw[8 + k] = cw[k];
    }
}

// GF(2) polar transform of an N-bit word held in XW registers (polar.py:17-29).
template <int XW>
__device__ __forceinline__ void transform_words(uint32_t (&x)[XW], int n) {
    // stages inside a word
    const uint32_t m[5] = {0x55555555u, 0x33333333u, 0x0f0f0f0fu, 0x00ff00ffu, 0x0000ffffu};
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        if (s < n) {
#pragma unroll
            for (int w = 0; w < XW; ++w) x[w] ^= (x[w] >> (1 << s)) & m[s];
        }
    }
    // stages across words: stage s>=5 pairs word w with w + 2^(s-5)
#pragma unroll
    for (int s = 5; s < 5 + 4; ++s) {
        const int d = 1 << (s - 5);
        if (s < n && d < XW) {
#pragma unroll
            for (int w = 0; w < XW; ++w)
                if ((w & d) == 0 && w + d < XW) x[w] ^= x[w + d];
        }
    }
}

// ---------------------------------------------------------------------------
// Path state (registers of one lane).
// ---------------------------------------------------------------------------
template <int LOGMAX>
struct Path {
    static constexpr int BW = BitsCfg<LOGMAX>::BW;
    static constexpr int XW = BitsCfg<LOGMAX>::XW;
    uint32_t P;            // 4-bit slot pointer per height h (field h-1), heights 1..n-1
    uint32_t bw[BW];       // partial-sum left buffers
    uint32_t xh[XW];       // x^ after the last phase
    double m;              // path metric (scl.py:19), fp64 accumulation
    uint32_t r;            // rank in the reference's sorted list
    bool alive;
};

// ---------------------------------------------------------------------------
// LLR tree evaluation.  Every routine stops at the height-1 pair (a, b): the two leaves of a phase pair are
// f(a,b) and g(a,b,u_even) and never touch memory.
// ---------------------------------------------------------------------------
template <int MP, int LOGMAX, int HS>
struct Tree {
    using WM = WarpMem<MP, HS>;
    using PathT = Path<LOGMAX>;
    static constexpr int BW = PathT::BW;

    template <int H> static __device__ __forceinline__ float* base(const WM& wm) { return (H >= HS) ? wm.tg : wm.ts; }

    // 2^H register values (H >= 1) -> height-1 pair; heights H-1..2 are stored in the own slot
    template <int H>
    static __device__ __forceinline__ void reg_chain(float (&v)[1 << H], const WM& wm, int lane, float& a, float& b) {
        if constexpr (H == 1) { a = v[0]; b = v[1]; }
        else {
            float w[(1 << H) / 2];
            float* own = base<H - 1>(wm) + lane;
#pragma unroll
            for (int i = 0; i < (1 << (H - 1)); ++i) {
                w[i] = f_op(v[i], v[i + (1 << (H - 1))]);
                if constexpr (H - 1 >= 2) own[(((1 << (H - 1)) - 2) + i) * 32] = w[i];
            }
            reg_chain<H - 1>(w, wm, lane, a, b);
        }
    }

    // bit e of the left-child buffer of height H (H in 4..6), for e = g + G*k with runtime g < G = 2^H/8 and
    // static k: the word index depends on k only, so every register index stays static.
    template <int H, int K>
    static __device__ __forceinline__ float g_strided(float a, float b, const uint32_t (&bw)[BW], int g) {
        constexpr int G = (1 << H) / 8;
        if constexpr (H == 4) return g_op_packed(a, b, bw[0], 15 + g + G * K);
        else if constexpr (H == 5) return g_op_packed(a, b, bw[1], g + G * K);
        else return g_op_packed(a, b, bw[2 + (K >= 4 ? 1 : 0)], g + G * (K & 3));     // H == 6: e < 32 iff K < 4
    }

    // Produce height H (>= 1) from height H+1 held at src[i*STRIDE] (a tree slot, STRIDE 32 with the lane folded into
    // src, or the frame-interleaved channel rows, STRIDE FPW with the frame folded into src)
    // with OP 0 = f, 1 = g using the left bits of height H; store it in the own slot and
    // continue with f down to the height-1 pair.
    //   H <= 3 : everything in registers.
    //   H 4..6 : depth-first in groups of 8 strided elements {g + G*k}: they reduce in registers to 4 values of
    //            height H-1, 2 of H-2 and 1 of H-3, so each level is stored once and never re-loaded (the tall rows
    //            live in the L2-resident scratch; re-reading them was the dominant stall).
    //   H >= 7 : (N >= 256) plain level loop, then recurse.
    //   COOP   : (phase 0 of a list decode, H = 5 or 6: every frame has ONE live path, in slot 0) the MP lanes of a group
    //            share the reduction of the channel row instead of repeating it on dead slots: the lane of slot s takes the
    //            groups g = s, s + MP, ... and writes heights H..H-3 into SLOT 0 (src must then be the frame's channel row,
    //            the same for all lanes of the group); from height H-3 down every lane continues on its own.  Same f
    //            operations on the same inputs, a quarter (M = 4) of the instructions and of the channel-row wavefronts.
    template <int H, int OP, int STRIDE, bool COOP = false>
    static __device__ __forceinline__ void produce(const float* src, const uint32_t (&bw)[BW], const WM& wm, int lane,
                                                   float& a, float& b) {
        constexpr int S = 1 << H;
        static_assert(!COOP || (H >= 5 && H <= 6 && MP > 1 && OP == 0), "cooperative form: phase 0 of a list decode only");
        const int col = COOP ? (lane & ~(MP - 1)) : lane;          // column (slot) the rows are written to
        float* own = base<H>(wm) + col;
        if constexpr (H <= 3) {
            float v[S];
#pragma unroll
            for (int i = 0; i < S; ++i) {
                const float x = src[i * STRIDE], y = src[(i + S) * STRIDE];
                v[i] = OP ? g_left<H, BW>(x, y, bw, i) : f_op(x, y);
                if constexpr (H >= 2) own[((S - 2) + i) * 32] = v[i];
            }
            reg_chain<H>(v, wm, lane, a, b);
        } else if constexpr (H <= 6) {
            constexpr int G = S / 8;
            float* o0 = own + (S - 2) * 32;
            float* o1 = base<H - 1>(wm) + col + (S / 2 - 2) * 32;
            float* o2 = base<H - 2>(wm) + col + (S / 4 - 2) * 32;
            float* o3 = base<H - 3>(wm) + col + (S / 8 - 2) * 32;
            float r0 = 0.f, r1 = 0.f;
            // one group = the 8 strided elements {g + G*k}: reduce them to heights H-1, H-2, H-3 in registers
            auto group = [&](int g, const float (&x)[8], const float (&y)[8]) {
                float v[8];
                v[0] = OP ? g_strided<H, 0>(x[0], y[0], bw, g) : f_op(x[0], y[0]);
                v[1] = OP ? g_strided<H, 1>(x[1], y[1], bw, g) : f_op(x[1], y[1]);
                v[2] = OP ? g_strided<H, 2>(x[2], y[2], bw, g) : f_op(x[2], y[2]);
                v[3] = OP ? g_strided<H, 3>(x[3], y[3], bw, g) : f_op(x[3], y[3]);
                v[4] = OP ? g_strided<H, 4>(x[4], y[4], bw, g) : f_op(x[4], y[4]);
                v[5] = OP ? g_strided<H, 5>(x[5], y[5], bw, g) : f_op(x[5], y[5]);
                v[6] = OP ? g_strided<H, 6>(x[6], y[6], bw, g) : f_op(x[6], y[6]);
                v[7] = OP ? g_strided<H, 7>(x[7], y[7], bw, g) : f_op(x[7], y[7]);
#pragma unroll
                for (int k = 0; k < 8; ++k) o0[(g + G * k) * 32] = v[k];
                float w[4], z[2];
#pragma unroll
                for (int k = 0; k < 4; ++k) { w[k] = f_op(v[k], v[k + 4]); o1[(g + G * k) * 32] = w[k]; }
#pragma unroll
                for (int k = 0; k < 2; ++k) { z[k] = f_op(w[k], w[k + 2]); o2[(g + G * k) * 32] = z[k]; }
                const float r = f_op(z[0], z[1]);
                if constexpr (H - 3 >= 2) o3[g * 32] = r;
                else { if (g == 0) r0 = r; else r1 = r; }
            };
            {
#pragma unroll 1
                for (int g = COOP ? (lane & (MP - 1)) : 0; g < G; g += COOP ? MP : 1) {
                    float x[8], y[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { x[k] = src[(g + G * k) * STRIDE]; y[k] = src[(g + G * k + S) * STRIDE]; }
                    group(g, x, y);
                }
            }
            if constexpr (COOP) __syncwarp();                       // slot 0's height H-3 row is complete
            if constexpr (H - 3 >= 2) chain_from<H - 3>(bw, wm, lane, col, a, b);
            else { a = r0; b = r1; }
        } else {
            float* dst = own + (S - 2) * 32;
            constexpr int W0 = (H == 7) ? 4 : 8;  // first word of height H
#pragma unroll
            for (int w = 0; w < S / 32; ++w) {
                const uint32_t bits = OP ? bw[(W0 + w) < BW ? (W0 + w) : 0] : 0u;
#pragma unroll 8
                for (int j = 0; j < 32; ++j) {
                    const int i = w * 32 + j;
                    const float x = src[i * STRIDE], y = src[(i + S) * STRIDE];
                    dst[i * 32] = OP ? g_op_packed(x, y, bits, j) : f_op(x, y);
                }
            }
            produce<H - 1, 0, 32>(dst, bw, wm, lane, a, b);
        }
    }

    // f-chain from height H (>= 2) of the slot in column `col` down to the pair (heights below H go to the own slot)
    template <int H>
    static __device__ __forceinline__ void chain_from(const uint32_t (&bw)[BW], const WM& wm, int lane, int col, float& a, float& b) {
        produce<H - 1, 0, 32>(base<H>(wm) + col + (((1 << H) - 2) * 32), bw, wm, lane, a, b);
    }
};

}  // namespace pb
