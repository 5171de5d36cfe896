// polar_decode.cuh -- the SC / SCL phase loop on top of polar_core.cuh.
// Reference: dl_scl_polar/polar/scl.py:108-209 (decode_scl), polar/polar.py:130-168 (sc_decode).
#pragma once
#include <type_traits>

#include "polar_core.cuh"

namespace pb {

template <int MP, int LOGMAX, bool FORCED, bool METRIC, int HS = DefaultHS<MP>::value>
struct ListDecoder {
    using PathT = Path<LOGMAX>;
    using TreeT = Tree<MP, LOGMAX, HS>;
    using WM = WarpMem<MP, HS>;
    static constexpr int BW = PathT::BW;
    static constexpr int XW = PathT::XW;
    static constexpr int FPW = 32 / MP;
    static constexpr uint32_t GM = (MP >= 32) ? 0xffffffffu : ((1u << MP) - 1u);
    // Phase 0 of a list decode has one live path per frame (slot 0): instead of repeating the reduction of the channel row
    // on the MP - 1 dead slots (10 % of all f/g work at M = 4, 12 % at M = 8), the lanes of a group split it (Tree::produce,
    // COOP) and all paths start with their tall heights pointing at slot 0 -- the state the first fork would produce anyway.
#ifndef PB_COOP0
#define PB_COOP0 1
#endif
    static constexpr bool kCoop0 = PB_COOP0 && MP > 1;

    // height-1 pair (a, b) of this lane's path for the EVEN phase `phi` (lazy form of scl.py:64-82).
    // C1 = true: the caller knows phi = 2 (mod 4), i.e. only height 1 is recomputed (g of the height-2 row).
    template <bool C1 = false>
    static __device__ __forceinline__ void pair_llr(const Code& code, const WM& wm, PathT& p, int phi, int lane,
                                                    const float* chanf, float& a, float& b) {
        const int n = code.n;
        const int slot = lane & (MP - 1), gbase = lane & ~(MP - 1);
        if constexpr (C1) {
            if (n == 2) TreeT::template produce<1, 1, FPW>(chanf, p.bw, wm, lane, a, b);    // N = 4: straight from the channel row
            else {
                const uint32_t q = (p.P >> 4) & 0xfu;                                         // slot holding height 2
                TreeT::template produce<1, 1, 32>(wm.ts + 2 * 32 + gbase + q, p.bw, wm, lane, a, b);
            }
            p.P = (p.P & ~0xfu) | (uint32_t)slot;
            return;
        }
        if (n == 1) { a = chanf[0]; b = chanf[FPW]; return; }    // N = 2: the channel row is the pair
        // first height produced: c = n-1 at phi = 0, else the number of trailing zeros of phi (>= 1)
        // One case per height, everything inside it static: the slot field of P, the source row, the pointer update
        // (heights 2..c live in the own slot afterwards).  Height n-1 comes straight from the staged channel rows
        // (stride FPW): f at phi = 0, g at phi = N/2.
        // (an if-chain on the bits of phi, most frequent height first -- height c is visited N / 2^(c+1) times per
        //  frame -- is cheaper than the two-level jump table the compiler builds for a switch on c)
#define PB_CASE(CC)                                                                                                   \
    if ((phi & (1 << CC)) || (phi == 0 && CC == n - 1)) {                                                             \
        if constexpr (CC < LOGMAX) {                                                                                  \
            if (CC == n - 1) {                                                                                        \
                if (phi == 0) {                                                                                       \
                    if constexpr (kCoop0 && CC >= 5 && CC <= 6) {                                                     \
                        /* one live path per frame: the group shares the row reduction, heights CC-3.. in slot 0 */   \
                        TreeT::template produce<CC, 0, FPW, true>(chanf, p.bw, wm, lane, a, b);                       \
                        p.P = ((uint32_t)slot * 0x11111111u) & ((1u << (4 * (CC >= 4 ? CC - 4 : 0))) - 1u);           \
                        return;                                                                                       \
                    } else TreeT::template produce<CC, 0, FPW>(chanf, p.bw, wm, lane, a, b);                          \
                }                                                                                                     \
                else TreeT::template produce<CC, 1, FPW>(chanf, p.bw, wm, lane, a, b);                                \
            } else if constexpr (CC + 1 < LOGMAX) {                                                                   \
                const uint32_t q = (p.P >> (4 * CC)) & 0xfu;         /* slot holding height CC+1 (field CC) */        \
                const float* src = ((CC + 1 >= HS) ? wm.tg : wm.ts) + (((2 << CC) - 2) * 32) + gbase + q;             \
                TreeT::template produce<CC, 1, 32>(src, p.bw, wm, lane, a, b);                                        \
            }                                                                                                         \
            constexpr uint32_t mask = (CC >= 8) ? 0xffffffffu : ((1u << (4 * CC)) - 1u);                              \
            p.P = (p.P & ~mask) | ((slot * 0x11111111u) & mask);                                                      \
        }                                                                                                             \
        return;                                                                                                       \
    }
        PB_CASE(1) PB_CASE(2) PB_CASE(3) PB_CASE(4) PB_CASE(5) PB_CASE(6) PB_CASE(7) PB_CASE(8)
#undef PB_CASE
    }

    // scl.py:84-99 in packed form, for an ODD phase (at least one trailing one).
    // T1 = true: the caller knows phi = 1 (mod 4), i.e. exactly one trailing one.
    template <bool T1 = false>
    static __device__ __forceinline__ void set_bit_odd(const Code& code, PathT& p, int phi, uint32_t bit) {
        if constexpr (T1) {
            uint32_t cw[1];
            ascend<1, BW, 1>(p.bw, bit, cw);
            if (code.n == 1) p.xh[0] = cw[0];
            else store_height<1, BW, 1>(p.bw, cw);
            return;
        }
        // t = trailing ones of phi (>= 1): an if-chain on the bits of phi, shortest run first
#define PB_CASE(TT)                                                                     \
    if (!(phi & (1 << TT))) {                                                           \
        if constexpr (TT <= LOGMAX) {                                                   \
            constexpr int CW = ((1 << TT) / 32) > 0 ? ((1 << TT) / 32) : 1;             \
            uint32_t cw[CW];                                                            \
            ascend<TT, BW, CW>(p.bw, bit, cw);                                          \
            if (TT == code.n) {                                                         \
                _Pragma("unroll") for (int k = 0; k < CW; ++k) p.xh[k] = cw[k];        \
            } else if constexpr (TT < LOGMAX) store_height<TT, BW, CW>(p.bw, cw);       \
        }                                                                               \
        return;                                                                         \
    }
        PB_CASE(1) PB_CASE(2) PB_CASE(3) PB_CASE(4) PB_CASE(5) PB_CASE(6) PB_CASE(7) PB_CASE(8) PB_CASE(9)
#undef PB_CASE
    }

    // r += (x < y) for doubles: one DSETP + one predicated IADD
    static __device__ __forceinline__ void inc_if_lt(uint32_t& r, double x, double y) {
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(r) : "d"(x), "d"(y));
    }

    // acc += C if x < y (doubles): one DSETP + one predicated IADD with an immediate
    template <uint32_t C>
    static __device__ __forceinline__ void add_if_lt(uint32_t& acc, double x, double y) {
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}" : "+r"(acc) : "d"(x), "d"(y), "n"(C));
    }

    static __device__ __forceinline__ void init(PathT& p, int lane, bool frame_valid) {
        p.P = 0;
#pragma unroll
        for (int k = 0; k < BW; ++k) p.bw[k] = 0;
#pragma unroll
        for (int k = 0; k < XW; ++k) p.xh[k] = 0;
        p.r = 0;
        p.alive = frame_valid && ((lane & (MP - 1)) == 0);   // scl.py:135 one initial path
        // List kernels: a lane without a live path carries the DEAD metric 2^1000.  Adding softplus terms leaves it
        // unchanged, its candidates sort behind every real one, and "alive" is just (metric < 2^992): no alive
        // flags, no selects in the key construction.
        p.m = (MP > 1 && !p.alive) ? kDeadMetric : 0.0;
    }
    static constexpr double kDeadMetric = 1.0715086071862673e+301;      // 2^1000 = 0x7e70000000000000
    static constexpr uint32_t kDeadHigh = 0x7e000000u;                   // high word of 2^992: alive <=> high word below
    static __device__ __forceinline__ bool metric_alive(double m) { return (uint32_t)__double2hiint(m) < kDeadHigh; }

    // Partial-sum state at the even phase phi_start, given the decided bits of all phases below it (prefix; bits at
    // frozen phases are 0): for every set bit h of phi_start the left buffer of height h is the polar transform of the
    // aligned 2^h-bit block that ends where the block of phi_start begins (scl.py:84-99 run to completion on that block).
    // After stages 0..h-1 of the word-parallel transform every aligned 2^h block holds its own transform.
    static __device__ __forceinline__ uint32_t word_at(const uint32_t (&x)[XW], int idx) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < XW; ++k) if (k == idx) v = x[k];
        return v;
    }
    // one height of jump_partial_sums (H is a compile-time constant: the packed fields of bw are static)
    template <int H>
    static __device__ __forceinline__ void jump_level(const Code& code, PathT& p, uint32_t (&x)[XW], int phi_start) {
        if constexpr (H < LOGMAX) {
            if (H >= 1 && H < code.n && ((phi_start >> H) & 1)) {
                const int o = phi_start & ~((2 << H) - 1);             // first phase of the finished left block of height H
                constexpr int CW = ((1 << H) / 32) > 0 ? ((1 << H) / 32) : 1;
                if constexpr (CW <= BW) {
                    uint32_t cw[CW];
                    if constexpr (H < 5) cw[0] = (word_at(x, o >> 5) >> (o & 31)) & ((1u << (1 << H)) - 1u);
                    else {
#pragma unroll
                        for (int k = 0; k < CW; ++k) cw[k] = word_at(x, (o >> 5) + k);
                    }
                    store_height<H, BW, CW>(p.bw, cw);
                }
            }
            // transform stage H (polar.py:17-29)
            if constexpr (H < 5) {
                constexpr uint32_t m = H == 0 ? 0x55555555u : H == 1 ? 0x33333333u : H == 2 ? 0x0f0f0f0fu : H == 3 ? 0x00ff00ffu : 0x0000ffffu;
#pragma unroll
                for (int w = 0; w < XW; ++w) x[w] ^= (x[w] >> (1 << H)) & m;
            } else {
                constexpr int d = 1 << (H - 5);
#pragma unroll
                for (int w = 0; w < XW; ++w) if ((w & d) == 0 && w + d < XW) x[w] ^= x[w + d];
            }
            jump_level<H + 1>(code, p, x, phi_start);
        }
    }
    static __device__ __forceinline__ void jump_partial_sums(const Code& code, PathT& p, const uint32_t (&prefix)[XW], int phi_start) {
        uint32_t x[XW];
#pragma unroll
        for (int w = 0; w < XW; ++w) {
            const int lo = w * 32;
            x[w] = prefix[w] & (phi_start >= lo + 32 ? 0xffffffffu : (phi_start <= lo ? 0u : ((1u << (phi_start - lo)) - 1u)));
        }
        jump_level<0>(code, p, x, phi_start);
    }

    // Decode the FPW frames of this warp; `chanf` = wm.chan + (frame of this lane within the warp): the staged,
    // frame-interleaved channel row (element i at chanf[i * FPW], see stage_channel_rows).
    // fmask/fval (FORCED): per-frame masks over phases -- bit phi of fmask set <=> u_phi is forced to bit phi of
    // fval (scl.py:138-144,155-161).
    // TRACE: also record, per information phase j, the leaf LLR every slot saw (wm.hist[j][lane]) and the slot each
    // surviving path came from (wm.lin[j][lane]); trace_walk() then yields a path's info_llrs (scl.py:159,167)
    // without the SC replay.
    // JUMP (DL-SCL retries whose forced prefix is known, see dl_bin_kernel): the decode STARTS at the even phase
    // phi_start (warp-uniform; every phase below it is forced for every frame of the warp).  The state the sequential
    // schedule would have reached there is rebuilt directly: the partial sums are polar transforms of aligned blocks of
    // the prefix bits (fval), and the LLR rows of the tree node that contains phi_start follow from one f / g level per
    // height -- the same fp32 operations on the same inputs, so bit-identical to decoding the prefix -- by running the
    // tree part of the "pseudo-phases" phi_start with its low bits cleared.  The path metric restarts at 0 at the
    // frame's OWN start phase my_start (even, >= phi_start; below it the frame's single path is forced): the forced
    // prefix adds the same constant to every path of the frame, and with the reset at my_start
    // the result does not depend on which other frames share the warp.  jstart = number of information phases below
    // phi_start (first trace row written).
    // UMASK: fetch the info-mask words from the kernel parameters (see the phase code) -- measured on B200: +3 % (M = 4) /
    // +9 % (M = 8) for the plain list decode kernel, but -6 % for the thread-per-frame kernels and -13 % for the sweep /
    // retry kernels (the by-value Code then lives in local memory next to their larger state), so only decode_kernel's
    // plain list instantiations ask for it.
    template <bool TRACE = false, bool JUMP = false, int UMASK = 0>
    static __device__ __forceinline__ void run(const Code& code, const uint32_t* __restrict__ imask, const WM& wm, PathT& p,
                                               int lane, const float* chanf, const uint32_t (&fmask)[XW],
                                               const uint32_t (&fval)[XW], uint32_t& flags, int phi_start = 0, int jstart = 0,
                                               int my_start = 0) {
        const int N = code.N;
        const uint32_t M = (uint32_t)code.M;
        const int slot = lane & (MP - 1), gbase = lane & ~(MP - 1);
        // rank table of the prune: entry (group g, rank r) at word r * GN + g (GN groups per warp), so that the reads of the
        // entries slot / slot + 1 by all 32 lanes are conflict-free; a lane beyond the list size M reads the last rank,
        // which is then always a dead candidate
        constexpr int GN = 32 / MP;
        uint32_t* tab = reinterpret_cast<uint32_t*>(wm.xchg + 128) + lane / MP;
        const int ridx = ((uint32_t)slot < M ? slot : 2 * MP - 1) * GN;
        uint32_t tie = 0;
        uint32_t cur_info = 0, cur_fm = 0, cur_fv = 0;   // word phi/32 of the info / force masks
        float a = 0.f, b = 0.f;                          // height-1 pair of the current phase pair
        int jinfo = JUMP ? jstart : 0;                   // index of the current information phase (TRACE)
        if constexpr (JUMP) {
            if (phi_start > 0) jump_partial_sums(code, p, fval, phi_start);
            cur_info = __ldg(imask + (phi_start >> 5)) >> (phi_start & 31);
#pragma unroll
            for (int k = 0; k < XW; ++k) if (k == (phi_start >> 5)) { cur_fm = fmask[k] >> (phi_start & 31); cur_fv = fval[k] >> (phi_start & 31); }
        }
        // One phase.  R = 0..3: R = phi mod 4 is a compile-time constant, so the even/odd split, the height-1
        // recomputation of phases 2 (mod 4) and the one-level partial-sum update of phases 1 (mod 4) need no dispatch.
        // R = 10 / 11: even / odd phase of a pair whose kind (`half`) is a run-time, warp-uniform flag.
        // R < 0: everything is read from phi -- the
        // forced / trace-recording kernels keep the compact loop, their code is large enough as it is (measured:
        // the four-fold body costs them more in instruction fetch than the dispatch it removes).
        auto phase = [&](const int phi, auto rc, const int half, const bool warm = false) {
            constexpr int R = decltype(rc)::value;
            // info / force masks as shift registers: bit 0 is the current phase (reloaded every 32 phases, shifted by one
            // at the end of every phase)
            if (R <= 0 || R == 10) {
                if ((phi & 31) == 0 && !(JUMP && warm)) {
                    if constexpr (UMASK == 1) {
                        // word phi/32 of the info mask from the kernel parameters instead of a global load
                        uint32_t mw = 0;
#pragma unroll
                        for (int k = 0; k < XW; ++k) if (k == (phi >> 5)) mw = code.info_mask[k];
                        cur_info = mw;
                    } else if constexpr (UMASK == 2) {
                        // the same with each word laundered through a register, so that the chain stays a chain of selects
                        uint32_t mw = 0;
#pragma unroll
                        for (int k = 0; k < XW; ++k) {
                            uint32_t wk;
                            asm("mov.u32 %0, %1;" : "=r"(wk) : "r"(code.info_mask[k]));
                            if (k == (phi >> 5)) mw = wk;
                        }
                        cur_info = mw;
                    } else cur_info = __ldg(imask + (phi >> 5));
                    if constexpr (FORCED) {
#pragma unroll
                        for (int k = 0; k < XW; ++k) if (k == (phi >> 5)) { cur_fm = fmask[k]; cur_fv = fval[k]; }
                    }
                }
            }
            const bool odd = R < 0 ? (phi & 1) : (R & 1);
            // Lanes without a live path run the same code on their own (unused) slot: no divergence, no merges.
            float L;
            if (!odd) {
                if constexpr (R == 10) {                         // even phase of a pair: half = 1 <=> phi = 2 (mod 4)
                    if (half) pair_llr<true>(code, wm, p, phi, lane, chanf, a, b);
                    else pair_llr<false>(code, wm, p, phi, lane, chanf, a, b);
                } else pair_llr<R == 2>(code, wm, p, phi, lane, chanf, a, b);
                if constexpr (JUMP) {
                    if (warm) return;                            // pseudo-phase of a jump start: LLR rows only
                    if (phi == my_start) p.m = (MP == 1 || metric_alive(p.m)) ? 0.0 : p.m;
                }
                L = f_op(a, b);
            }
            else L = g_op_packed(a, b, p.bw[0], 0);              // u_{phi-1} sits in the height-0 field
            const bool is_info = cur_info & 1u;
            const bool is_forced = FORCED && is_info && (cur_fm & 1u);
            const uint32_t forced_val = cur_fv & 1u;
            cur_info >>= 1;
            if constexpr (FORCED) { cur_fm >>= 1; cur_fv >>= 1; }
            uint32_t bit = 0;
            if constexpr (!METRIC) {
                // plain SC (polar.py:147-153): frozen -> 0 else L < 0
                bit = (is_info && L < 0.f) ? 1u : 0u;
                if (is_forced) bit = forced_val;
            } else {
                const float tail = softplus_tail(L);
                const double dtail = (double)tail;
                double m0 = p.m + ((double)fmaxf(-L, 0.f) + dtail);                // bit 0: logaddexp(0,-L)
                if constexpr (MP == 1) {
                    bool a0 = p.alive, a1 = p.alive && is_info;
                    if (is_forced) {
                        a0 = a0 && (forced_val == 0);
                        a1 = a1 && (forced_val == 1);
                    }
                    if constexpr (TRACE) { if (is_info) { wm.hist[jinfo * 32 + lane] = L; ++jinfo; } }
                    const double m1 = p.m + ((double)fmaxf(L, 0.f) + dtail);       // bit 1: logaddexp(0, L)
                    bool pick1 = a1 && (!a0 || m1 < m0);
                    if (a0 && a1) {
                        const uint32_t h0 = (uint32_t)(__double_as_longlong(m0) >> 32), h1 = (uint32_t)(__double_as_longlong(m1) >> 32);
                        if ((uint32_t)(h0 - h1 + 2u) <= 4u) tie = 1;
                    }
                    p.m = pick1 ? m1 : m0;
                    bit = pick1 ? 1u : 0u;
                } else {
                    if (!is_info) {
                        // frozen phase (scl.py:149-153): one child per path.  The reference re-sorts here too (:173),
                        // but a rank is only ever used as the tie-break between EXACTLY equal metrics, so the
                        // re-ranking is deferred to the next prune / the final ordering below.
                        p.m = m0;                                 // (a dead lane's metric is never read)
                        bit = 0;
                    } else {
                        // Keys: IEEE bits of the (non-negative) fp64 metric with the stable-sort tie-break 2*rank+bit in
                        // the 4 lowest mantissa bits (the path in slot s IS the path of rank s, see below).  They order
                        // identically as integers and as doubles, so the rank compares run on the otherwise idle FP64
                        // pipe (one DSETP each).  Dead lanes carry the dead metric, so every candidate has a unique
                        // rank (dead ones last, in slot order).
                        double m1 = p.m + ((double)fmaxf(L, 0.f) + dtail);         // bit 1: logaddexp(0, L)
                        if constexpr (FORCED) {                                    // scl.py:155-161: a forced bit kills the other child
                            if (is_forced) { if (forced_val) m0 = kDeadMetric; else m1 = kDeadMetric; }
                        }
                        const uint32_t h0 = (uint32_t)__double2hiint(m0), h1 = (uint32_t)__double2hiint(m1);
                        const uint32_t l0 = ((uint32_t)__double2loint(m0) & ~15u) | (2u * slot);
                        const uint32_t l1 = ((uint32_t)__double2loint(m1) & ~15u) | (2u * slot + 1u);
                        // the keys of even and odd phases live in two buffers: a lane may already publish the keys of
                        // the next phase while another one still fetches its new metric from this phase's keys
                        uint4* keys = reinterpret_cast<uint4*>(wm.xchg + (odd ? 64 : 0));
                        keys[lane] = make_uint4(l0, h0, l1, h1);
                        __syncwarp();
                        const double d0 = __hiloint2double((int)h0, (int)l0), d1 = __hiloint2double((int)h1, (int)l1);
                        // rank of a candidate = number of smaller keys among the group's 2 MP: its sibling (in registers)
                        // and the pairs of the MP - 1 other lanes (lane slot ^ j, one LDS.128 each)
                        uint32_t rank0 = 0, rank1 = 0;
#ifndef PB_RANK_SELF
#define PB_RANK_SELF 1
#endif
#ifndef PB_RANK_ANTISYM_MP
#define PB_RANK_ANTISYM_MP 4          // smallest group size that uses the antisymmetric exchange (16: never)
#endif
                        if constexpr (MP >= PB_RANK_ANTISYM_MP && MP >= 4) {
                            // Keys are unique, so every comparison between two lanes' candidates is needed only once:
                            // a lane compares its pair with the pairs of the lanes at slot + 1 .. slot + MP/2 - 1 (mod MP)
                            // and counts BOTH sides -- (ox < d0) raises my rank0, and the partner's rank of ox is
                            // 2 - (ox < d0) - (ox < d1) -- in one packed word per partner (bytes: my rank0, my rank1,
                            // partner's ox count, partner's oy count); the partner's half comes back with one shuffle.
                            // The lane opposite (slot ^ MP/2) is compared by both sides for their own ranks only.
                            // MP = 8: 4 LDS.128 + 3 SHFL instead of 7 LDS.128, 18 compares instead of 30.
                            constexpr int H = MP / 2;
                            uint32_t own = 0, lows = 0, backs = 0;
                            add_if_lt<0x1u>(own, d1, d0); add_if_lt<0x100u>(own, d0, d1);
                            uint32_t part[H - 1];
#pragma unroll
                            for (int d = 1; d < H; ++d) {
                                const uint4 o = keys[gbase | ((slot + d) & (MP - 1))];
                                const double ox = __hiloint2double((int)o.y, (int)o.x), oy = __hiloint2double((int)o.w, (int)o.z);
                                uint32_t acc = 0;
                                add_if_lt<0x00010001u>(acc, ox, d0); add_if_lt<0x00010100u>(acc, ox, d1);
                                add_if_lt<0x01000001u>(acc, oy, d0); add_if_lt<0x01000100u>(acc, oy, d1);
                                part[d - 1] = acc;
                                lows += acc;
                            }
                            {
                                const uint4 o = keys[lane ^ H];
                                const double ox = __hiloint2double((int)o.y, (int)o.x), oy = __hiloint2double((int)o.w, (int)o.z);
                                add_if_lt<0x1u>(own, ox, d0); add_if_lt<0x100u>(own, ox, d1);
                                add_if_lt<0x1u>(own, oy, d0); add_if_lt<0x100u>(own, oy, d1);
                            }
#pragma unroll
                            for (int d = 1; d < H; ++d) backs += __shfl_sync(kFull, part[d - 1], gbase | ((slot - d) & (MP - 1)));
                            // (byte fields never carry: ranks < 16, each returned count <= 2 and at most MP/2 - 1 of them)
                            own += (lows & 0xffffu) + (uint32_t)(0x0202u * (H - 1)) - (backs >> 16);
                            rank0 = own & 0xffu; rank1 = (own >> 8) & 0xffu;
                        } else {
#if PB_RANK_SELF
                        inc_if_lt(rank0, d1, d0); inc_if_lt(rank1, d0, d1);
#pragma unroll
                        for (int j = 1; j < MP; ++j) {
                            const uint4 o = keys[lane ^ j];
#else
#pragma unroll
                        for (int j = 0; j < MP; ++j) {
                            const uint4 o = keys[gbase + j];
#endif
                            const double ox = __hiloint2double((int)o.y, (int)o.x), oy = __hiloint2double((int)o.w, (int)o.z);
                            inc_if_lt(rank0, ox, d0); inc_if_lt(rank0, oy, d0);
                            inc_if_lt(rank1, ox, d1); inc_if_lt(rank1, oy, d1);
                        }
                        }   // (all-partners exchange)
                        // scl.py:173-174: the sorted list, truncated to M.  Every candidate publishes itself under its
                        // rank -- one word: high key word (for the near-tie test) and candidate id 2*slot+bit -- and
                        // lane s then BECOMES the candidate of rank s: it reads entry s and pulls that candidate's
                        // state from its parent lane.  Survivors therefore always sit in rank order (slot = rank): no
                        // rank register, no survive/clone bookkeeping, and the neighbour s+1 of the near-tie test is a
                        // fixed address.
                        tab[rank0 * GN] = h0 * 16u + 2u * slot;
                        tab[rank1 * GN] = h1 * 16u + (2u * slot + 1u);
                        __syncwarp();
                        const uint32_t w = tab[ridx], wn = tab[ridx + GN];
                        const int e = (int)(w & 15u);                 // candidate id: 2 * (slot of the parent) + bit
                        const int src = gbase + (e >> 1);
                        if constexpr (TRACE) {
                            wm.hist[jinfo * 32 + lane] = L;
                            {   // lineage of all 32 lanes as bit planes: three ballots, one 16-byte store
                                const uint32_t b0 = __ballot_sync(kFull, e & 2), b1 = __ballot_sync(kFull, e & 4);
                                const uint32_t b2 = MP > 4 ? __ballot_sync(kFull, e & 8) : 0u;
                                if (lane == 0) *reinterpret_cast<uint4*>(wm.lin + jinfo * 4) = make_uint4(b0, b1, b2, 0u);
                            }
                            ++jinfo;
                        }
                        // metric of the candidate = its key without the tie-break bits
                        const uint2 kk = reinterpret_cast<const uint2*>(keys)[2 * gbase + e];
                        p.m = __hiloint2double((int)kk.y, (int)(kk.x & ~15u));
                        // near-tie: a kept (alive) candidate and its successor within ~1e-6 relative (high key words
                        // <= 2 apart; the table holds them shifted by 4, i.e. modulo 2^28 -- far beyond any metric ratio)
                        if (kk.y < kDeadHigh && (uint32_t)((wn | 15u) - w) <= 47u) tie = 1;
                        // Every lane reads the path state from `src` (possibly itself).  For N <= 128 all four
                        // partial-sum words travel unconditionally (a branch around a shuffle costs more than the
                        // shuffle); the wide buffers of N = 256 / 512 travel only while they are live (a left buffer
                        // of height h is live while bit h of phi is set).
                        p.P = __shfl_sync(kFull, p.P, src);
#pragma unroll
                        for (int k = 0; k < (BW < 4 ? BW : 4); ++k) p.bw[k] = __shfl_sync(kFull, p.bw[k], src);
                        if constexpr (BW >= 8) {
                            if (phi & 128) {
#pragma unroll
                                for (int k = 4; k < 8; ++k) p.bw[k] = __shfl_sync(kFull, p.bw[k], src);
                            }
                        }
                        if constexpr (BW >= 16) {
                            if (phi & 256) {
#pragma unroll
                                for (int k = 8; k < 16; ++k) p.bw[k] = __shfl_sync(kFull, p.bw[k], src);
                            }
                        }
                        if (!odd) { a = __shfl_sync(kFull, a, src); b = __shfl_sync(kFull, b, src); }
                        bit = (uint32_t)(e & 1);
                    }
                }
            }
            if (!odd) p.bw[0] = (p.bw[0] & ~1u) | bit;           // height-0 left buffer
            else if constexpr (R == 11) {                        // odd phase of a pair: half = 0 <=> phi = 1 (mod 4)
                if (half) set_bit_odd<false>(code, p, phi, bit);
                else set_bit_odd<true>(code, p, phi, bit);
            } else set_bit_odd<R == 1>(code, p, phi, bit);
        };
        // Loop forms.  4 = fully static blocks of four phases (even/odd split, the height-1 recompute of phases 2 mod 4 and
        // the one-level partial-sum update of phases 1 mod 4 need no dispatch; four copies of the phase body),
        // 2 = phase pairs with one warp-uniform pair-kind branch (two copies), 1 = compact loop (one copy).
        // Measured on B200 (round 2, static-N kernels, 1 Mi+ frames): plain list kernels 4 > 2 by 3-4 % (M = 2, 4, 8 and the
        // fused sweep; in round 1 the larger body of that build stalled on instruction fetch and pairs won); forced /
        // trace-recording kernels 2 > 1 by 5-7 % on the DL-SCL legs.
#ifndef PB_LIST_FORM
#define PB_LIST_FORM 4
#endif
#ifndef PB_TRACE_FORM
#define PB_TRACE_FORM 2
#endif
        constexpr int FORM = JUMP ? 0 : (!FORCED && !TRACE) ? (MP == 1 ? 4 : PB_LIST_FORM) : PB_TRACE_FORM;
        if constexpr (JUMP) {
            // phases from phi_start on, preceded by the pseudo-phases {phi_start with its low bits cleared}: first the
            // top bit alone (0 or N/2: height n-1 from the channel row by f or g), then one more set bit at a time.
            // PB_JUMP_FORM 2: phase pairs (two copies of the phase body), 1: compact loop (one copy).
#ifndef PB_JUMP_FORM
#define PB_JUMP_FORM 2
#endif
            int phi0 = phi_start & (N >> 1);
#pragma unroll 1
            while (phi0 < N) {
                const bool warm = phi0 < phi_start;
                if constexpr (PB_JUMP_FORM == 2) {
                    const int half = (phi0 >> 1) & 1;
                    phase(phi0, std::integral_constant<int, 10>{}, half, warm);
                    if (warm) { phi0 |= 1 << (31 - __clz(phi_start ^ phi0)); continue; }
                    phase(phi0 + 1, std::integral_constant<int, 11>{}, half);
                    phi0 += 2;
                } else {
                    phase(phi0, std::integral_constant<int, -1>{}, 0, warm);
                    if (warm) phi0 |= 1 << (31 - __clz(phi_start ^ phi0));
                    else ++phi0;
                }
            }
        } else if constexpr (FORM == 4) {
#pragma unroll 1
            for (int phi0 = 0; phi0 < N; phi0 += 4) {
                phase(phi0, std::integral_constant<int, 0>{}, 0);
                phase(phi0 + 1, std::integral_constant<int, 1>{}, 0);
                if (phi0 + 2 < N) {                              // (N = 2 has a single phase pair)
                    phase(phi0 + 2, std::integral_constant<int, 2>{}, 1);
                    phase(phi0 + 3, std::integral_constant<int, 3>{}, 1);
                }
            }
        } else if constexpr (FORM == 2) {
#pragma unroll 1
            for (int phi0 = 0; phi0 < N; phi0 += 2) {
                const int half = (phi0 >> 1) & 1;
                phase(phi0, std::integral_constant<int, 10>{}, half);
                phase(phi0 + 1, std::integral_constant<int, 11>{}, half);
            }
        } else {
            for (int phi = 0; phi < N; ++phi) phase(phi, std::integral_constant<int, -1>{}, 0);
        }
        // final list order = metric order (scl.py:173-174,183-188), ties by the last computed rank
        if constexpr (MP > 1 && METRIC) {
            p.alive = metric_alive(p.m);
            const unsigned long long kf = p.alive ? (((unsigned long long)__double_as_longlong(p.m) & ~15ull) | (uint32_t)slot) : ~0ull;
            wm.xchg[lane] = kf;
            __syncwarp();
            uint32_t rank = 0, cnt = 0;
            const uint32_t hf = (uint32_t)(kf >> 32);
#pragma unroll
            for (int j = 0; j < MP; ++j) {
                const unsigned long long o = wm.xchg[gbase + j];
                rank += (o < kf);
                cnt += ((uint32_t)((uint32_t)(o >> 32) - hf + 2u) <= 4u);
            }
            if (p.alive) { p.r = rank; if (cnt >= 2) tie = 1; }
            __syncwarp();
        }
        if (tie) flags |= PB_FLAG_NEAR_TIE;
    }

    // info_llrs of the path that ended in lane `end_lane` (group-uniform), from the trace of a run<true>():
    // sink(j, L) is called on the group lane with slot == j mod MP.  Blocks of 32 information phases: first the
    // lineage chain (shared-memory reads only), then this lane's 32/MP trace loads back to back, then the sinks --
    // the global loads of a block are all in flight together instead of one L2 round trip per phase.
    template <typename Sink>
    static __device__ __forceinline__ void trace_walk(const Code& code, const WM& wm, int lane, int end_lane, Sink&& sink, int jlow = 0) {
        const int slot = lane & (MP - 1), gbase = lane & ~(MP - 1);
        constexpr int PER = 32 / MP;                        // phases of a block this lane is responsible for
        int s = end_lane & (MP - 1);
        for (int jtop = code.K - 1; jtop >= jlow; jtop -= 32) {      // (jlow > 0: the trace of a jump-started decode begins there)
            unsigned long long mine = 0;                    // slots of this lane's phases, 4 bits each, earliest phase lowest
            // (unrolled by 4, not 32: the walk runs once per traced decode, and its 32-fold straight-line form was 6 KB of
            //  the trace-recording kernels' instruction footprint -- DL-SCL at 4 dB +1.8 %, 5 dB +0.7 %)
#ifndef PB_WALK_UNROLL
#define PB_WALK_UNROLL 4
#endif
            constexpr int kWalkUnroll = PB_WALK_UNROLL;
#pragma unroll kWalkUnroll
            for (int t = 0; t < 32; ++t) {
                const int j = jtop - t;
                if (j >= jlow) {
                    int w = 0;
                    if constexpr (MP > 1) {
                        const uint4 q = *reinterpret_cast<const uint4*>(wm.lin + j * 4);      // broadcast read
                        const int sh = gbase + s;
                        w = (int)(((q.x >> sh) & 1u) | (((q.y >> sh) & 1u) << 1) | (((q.z >> sh) & 1u) << 2));
                    }
                    if ((j & (MP - 1)) == slot) mine = (mine << 4) | (unsigned long long)w;
                    s = w;
                }
            }
            // this lane's phases of the block, ascending: j0, j0 + MP, ...
            const int jlo = jtop - 31 > jlow ? jtop - 31 : jlow;
            const int j0 = jlo + ((slot - jlo) & (MP - 1));
            float v[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int j = j0 + i * MP;
                v[i] = (j <= jtop) ? wm.hist[j * 32 + gbase + (int)((mine >> (4 * i)) & 15ull)] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int j = j0 + i * MP;
                if (j <= jtop) sink(j, v[i]);
            }
        }
    }
};

}  // namespace pb
