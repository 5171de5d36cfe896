// polar_decode.cuh -- the SC / SCL phase loop on top of polar_core.cuh.
// Reference: dl_scl_polar/polar/scl.py:108-209 (decode_scl), polar/polar.py:130-168 (sc_decode).
#pragma once
#include "polar_core.cuh"

namespace pb {

template <int MP, int LOGMAX, bool FORCED, bool METRIC>
struct ListDecoder {
    using PathT = Path<LOGMAX>;
    using TreeT = Tree<MP, LOGMAX>;
    static constexpr int BW = PathT::BW;
    static constexpr int XW = PathT::XW;
    static constexpr int FPW = 32 / MP;
    static constexpr uint32_t GM = (MP >= 32) ? 0xffffffffu : ((1u << MP) - 1u);

    // leaf LLR of `phi` for this lane's path (lazy form of scl.py:64-82)
    static __device__ __forceinline__ float leaf_llr(const Code& code, const WarpMem<MP>& wm, PathT& p, int phi,
                                                     int lane) {
        const int n = code.n;
        const int slot = lane & (MP - 1), gbase = lane & ~(MP - 1);
        float* own = wm.tree + lane;
        const float* chanf = wm.chan + (lane / MP) * (code.N + 1);
        float leaf = 0.f;
        int c;
        if (phi == 0) {
            c = n - 1;
            switch (n) {
#define PB_CASE(NN) case NN: if constexpr (NN <= LOGMAX) leaf = TreeT::template produce<NN - 1, 0>(chanf, 1, p.bw, own); break;
                PB_CASE(1) PB_CASE(2) PB_CASE(3) PB_CASE(4) PB_CASE(5) PB_CASE(6) PB_CASE(7) PB_CASE(8) PB_CASE(9)
#undef PB_CASE
                default: break;
            }
        } else {
            c = __ffs(phi) - 1;
            const float* src;
            int stride;
            if (c + 1 == n) { src = chanf; stride = 1; }
            else {
                const uint32_t q = (p.P >> (4 * c)) & 0xfu;            // slot holding height c+1 (field c)
                src = wm.tree + (((2 << c) - 2) * 32) + gbase + q;
                stride = 32;
            }
            switch (c) {
#define PB_CASE(CC) case CC: if constexpr (CC < LOGMAX) leaf = TreeT::template produce<CC, 1>(src, stride, p.bw, own); break;
                PB_CASE(0) PB_CASE(1) PB_CASE(2) PB_CASE(3) PB_CASE(4) PB_CASE(5) PB_CASE(6) PB_CASE(7) PB_CASE(8)
#undef PB_CASE
                default: break;
            }
        }
        // heights 1..c now live in the own slot
        const uint32_t mask = (c >= 8) ? 0xffffffffu : ((1u << (4 * c)) - 1u);
        p.P = (p.P & ~mask) | ((slot * 0x11111111u) & mask);
        return leaf;
    }

    // scl.py:84-99 in packed form
    static __device__ __forceinline__ void set_bit(const Code& code, PathT& p, int phi, uint32_t bit) {
        const int t = __ffs(~phi) - 1;  // trailing ones of phi
        switch (t) {
#define PB_CASE(TT)                                                                     \
    case TT:                                                                            \
        if constexpr (TT <= LOGMAX) {                                                   \
            constexpr int CW = ((1 << TT) / 32) > 0 ? ((1 << TT) / 32) : 1;             \
            uint32_t cw[CW];                                                            \
            ascend<TT, BW, CW>(p.bw, bit, cw);                                          \
            if (TT == code.n) {                                                         \
                _Pragma("unroll") for (int k = 0; k < CW; ++k) p.xh[k] = cw[k];        \
            } else if constexpr (TT < LOGMAX) store_height<TT, BW, CW>(p.bw, cw);       \
        }                                                                               \
        break;
            PB_CASE(0) PB_CASE(1) PB_CASE(2) PB_CASE(3) PB_CASE(4) PB_CASE(5) PB_CASE(6) PB_CASE(7) PB_CASE(8) PB_CASE(9)
#undef PB_CASE
            default: break;
        }
    }

    static __device__ __forceinline__ void init(PathT& p, int lane, bool frame_valid) {
        p.P = 0;
#pragma unroll
        for (int k = 0; k < BW; ++k) p.bw[k] = 0;
#pragma unroll
        for (int k = 0; k < XW; ++k) p.xh[k] = 0;
        p.m = 0.0;
        p.r = 0;
        p.alive = frame_valid && ((lane & (MP - 1)) == 0);   // scl.py:135 one initial path
    }

    // Decode the FPW frames whose channel LLRs are in wm.chan.  fmask/fval (FORCED): per-frame masks over
    // phases -- bit phi of fmask set <=> u_phi is forced to bit phi of fval (scl.py:138-144,155-161).
    static __device__ __forceinline__ void run(const Code& code, const WarpMem<MP>& wm, PathT& p, int lane,
                                               const uint32_t (&fmask)[XW], const uint32_t (&fval)[XW], uint32_t& flags) {
        const int N = code.N;
        const uint32_t M = (uint32_t)code.M;
        const int slot = lane & (MP - 1), gbase = lane & ~(MP - 1);
        uint32_t tie = 0;
        uint32_t cur_info = 0, cur_fm = 0, cur_fv = 0;   // word phi/32 of the info / force masks
        for (int phi = 0; phi < N; ++phi) {
            if ((phi & 31) == 0) {
                cur_info = code.info_mask[phi >> 5];
                if constexpr (FORCED) {
#pragma unroll
                    for (int k = 0; k < XW; ++k) if (k == (phi >> 5)) { cur_fm = fmask[k]; cur_fv = fval[k]; }
                }
            }
            float L = 0.f;
            if (p.alive) L = leaf_llr(code, wm, p, phi, lane);
            const bool is_info = (cur_info >> (phi & 31)) & 1u;
            const bool is_forced = FORCED && is_info && ((cur_fm >> (phi & 31)) & 1u);
            const uint32_t forced_val = (cur_fv >> (phi & 31)) & 1u;
            uint32_t bit = 0;
            if constexpr (!METRIC) {
                // plain SC (polar.py:147-153): frozen -> 0 else L < 0
                bit = (is_info && L < 0.f) ? 1u : 0u;
                if (is_forced) bit = forced_val;
            } else {
                const float tail = softplus_tail(L);
                const double m0 = p.m + ((double)fmaxf(-L, 0.f) + (double)tail);   // bit 0: logaddexp(0,-L)
                const double m1 = p.m + ((double)fmaxf(L, 0.f) + (double)tail);    // bit 1: logaddexp(0, L)
                bool a0 = p.alive, a1 = p.alive && is_info;
                if (is_forced) {
                    a0 = a0 && (forced_val == 0);
                    a1 = a1 && (forced_val == 1);
                }
                if constexpr (MP == 1) {
                    bool pick1 = a1 && (!a0 || m1 < m0);
                    if (a0 && a1) {
                        const uint32_t h0 = (uint32_t)(__double_as_longlong(m0) >> 32), h1 = (uint32_t)(__double_as_longlong(m1) >> 32);
                        if ((uint32_t)(h0 - h1 + 2u) <= 4u) tie = 1;
                    }
                    p.m = pick1 ? m1 : m0;
                    bit = pick1 ? 1u : 0u;
                } else {
                    const unsigned long long dead = ~0ull;
                    if (!is_info) {
                        // frozen phase (scl.py:149-153): one child per path.  The reference re-sorts here too (:173),
                        // but a rank is only ever used as the tie-break between EXACTLY equal metrics, so the
                        // re-ranking is deferred to the next prune / the final ordering (rank_final below).
                        if (a0) p.m = m0;
                        bit = 0;
                    } else {
                        const unsigned long long k0 = a0 ? (((unsigned long long)__double_as_longlong(m0) & ~15ull) | (2u * p.r)) : dead;
                        const unsigned long long k1 = a1 ? (((unsigned long long)__double_as_longlong(m1) & ~15ull) | (2u * p.r + 1u)) : dead;
                        reinterpret_cast<ulonglong2*>(wm.xchg)[lane] = make_ulonglong2(k0, k1);
                        __syncwarp();
                        uint32_t rank0 = 0, rank1 = 0, cnt0 = 0, cnt1 = 0;
                        const uint32_t h0 = (uint32_t)(k0 >> 32), h1 = (uint32_t)(k1 >> 32);
#pragma unroll
                        for (int j = 0; j < MP; ++j) {
                            const ulonglong2 o = reinterpret_cast<const ulonglong2*>(wm.xchg)[gbase + j];
                            const uint32_t ox = (uint32_t)(o.x >> 32), oy = (uint32_t)(o.y >> 32);
                            rank0 += (o.x < k0) + (o.y < k0);
                            rank1 += (o.x < k1) + (o.y < k1);
                            cnt0 += ((uint32_t)(ox - h0 + 2u) <= 4u) + ((uint32_t)(oy - h0 + 2u) <= 4u);
                            cnt1 += ((uint32_t)(ox - h1 + 2u) <= 4u) + ((uint32_t)(oy - h1 + 2u) <= 4u);
                        }
                        const bool s0 = a0 && rank0 < M, s1 = a1 && rank1 < M;   // scl.py:174 keep the M best
                        if ((s0 && cnt0 >= 2) || (s1 && cnt1 >= 2)) tie = 1;
                        const bool dbl = s0 && s1, fre = !s0 && !s1;
                        const uint32_t dm = (__ballot_sync(kFull, dbl) >> gbase) & GM;
                        const uint32_t fm = (__ballot_sync(kFull, fre) >> gbase) & GM;
                        int src = lane;
                        bool take = false;
                        if (fre) {
                            const int k = __popc(fm & ((1u << slot) - 1u));
                            uint32_t d = dm;
#pragma unroll
                            for (int i = 0; i < MP / 2; ++i) if (i < k) d &= d - 1;
                            if (d) { src = gbase + __ffs(d) - 1; take = true; }
                        }
                        // the second child of a doubly-surviving path moves into a freed slot
                        const uint32_t P2 = __shfl_sync(kFull, p.P, src);
                        uint32_t b2[BW];
#pragma unroll
                        for (int k = 0; k < BW; ++k) b2[k] = __shfl_sync(kFull, p.bw[k], src);
                        const double m2 = __shfl_sync(kFull, m1, src);
                        const uint32_t r2 = __shfl_sync(kFull, rank1, src);
                        if (take) {
                            p.P = P2;
#pragma unroll
                            for (int k = 0; k < BW; ++k) p.bw[k] = b2[k];
                            p.m = m2; p.r = r2; bit = 1; p.alive = true;
                        } else if (s0) { p.m = m0; p.r = rank0; bit = 0; }
                        else if (s1) { p.m = m1; p.r = rank1; bit = 1; }
                        else p.alive = false;
                    }
                }
            }
            if (p.alive) set_bit(code, p, phi, bit);
            if constexpr (MP > 1) __syncwarp();
        }
        // final list order = metric order (scl.py:173-174,183-188), ties by the last computed rank
        if constexpr (MP > 1) {
            const unsigned long long kf = p.alive ? (((unsigned long long)__double_as_longlong(p.m) & ~15ull) | p.r) : ~0ull;
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

    // SC pass along the known bits `u` (own slot only), reporting the leaf LLR of every information phase:
    // the values a path saw during list decoding (scl.py:159,167 info_llrs), recomputed instead of copied.
    template <typename Sink>
    static __device__ __forceinline__ void replay(const Code& code, const WarpMem<MP>& wm, int lane, bool active,
                                                  const uint32_t (&u)[XW], Sink&& sink) {
        PathT q;
        init(q, lane, true);
        q.alive = active;
        q.P = (lane & (MP - 1)) * 0x11111111u;
        int j = 0;
        uint32_t cur_info = 0, cur_u = 0;
        for (int phi = 0; phi < code.N; ++phi) {
            if ((phi & 31) == 0) {
                cur_info = code.info_mask[phi >> 5];
#pragma unroll
                for (int k = 0; k < XW; ++k) if (k == (phi >> 5)) cur_u = u[k];
            }
            float L = 0.f;
            if (active) L = leaf_llr(code, wm, q, phi, lane);
            if ((cur_info >> (phi & 31)) & 1u) { if (active) sink(j, L); ++j; }
            if (active) set_bit(code, q, phi, (cur_u >> (phi & 31)) & 1u);
        }
    }
};

}  // namespace pb
