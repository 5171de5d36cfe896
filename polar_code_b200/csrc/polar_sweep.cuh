// polar_sweep.cuh -- channel generation (Philox4x32-10 + Box-Muller), the fused Monte-Carlo sweep kernel and
// DL-SCL flip retries as one persistent, work-stealing retry kernel.
//
// Reference paths: eval/run_fer_sweep.py:60-121 (channel, counters), eval/run_ber_sweep.py:112-181,
// dlscl/flip.py:65-141 (retry controller), nr/polar/scl_nr.py:23-57 (rate-matched chain).
//
// DL-SCL on the GPU: the baseline pass decodes every frame and records the leaf-LLR trace; frames whose best path
// fails the CRC are appended to a queue (frame id, best u-hat, transmitted word; their LLR row goes to the LLR store,
// the |L0| vector of their best path -- read off the trace -- to the |L0| store).  ONE persistent retry kernel then
// drains the queue: every lane group owns a frame through all of its retries -- score q = |L0| @ beta in fp64
// (flip.py:104-108), force the prefix + flipped bit (flip.py:30-34), list-decode with the trace on, finish or walk
// the trace for the next |L0| (flip.py:102,133) and go on -- and pulls the next entry when its frame is done, so all
// lanes stay busy whatever the SNR and nothing is regenerated, replayed or re-queued between retries.
#pragma once
#include "polar_kernels.cuh"

namespace pb {

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  counter = (frame_lo, frame_hi, block, purpose), key = (k0, k1).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

enum { kPurposePayload = 0, kPurposeNoise = 1, kPurposeUncoded = 2 };

// four N(0,1) samples from one Philox block (two Box-Muller pairs).  SFU forms: ln u = ln2 * lg2(u) (MUFU.LG2),
// sqrt(t) = t * rsqrt(t) (MUFU.RSQ), sin / cos of an angle in [-pi, pi) (MUFU.SIN / MUFU.COS, absolute error ~5e-7):
// ~20 instructions per block instead of ~100 for logf / sqrtf / sincospif, at an error of 1e-6 sigma per sample.
__device__ __forceinline__ void normal4(uint4 r, float (&z)[4]) {
    const float u1 = fmaf((float)r.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (x+0.5)/2^32 in (0,1]
    const float u3 = fmaf((float)r.z, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float t1 = -1.3862943611198906f * __log2f(u1), t2 = -1.3862943611198906f * __log2f(u3);   // -2 ln u >= 0
    const float r1 = t1 > 0.f ? t1 * rsqrtf(t1) : 0.f, r2 = t2 > 0.f ? t2 * rsqrtf(t2) : 0.f;
    // angle 2*pi*(u2 - 1/2), u2 = y/2^32: uniform on [-pi, pi)
    const float a1 = fmaf((float)r.y, 1.4629180792671596e-09f, -3.14159265358979f);
    const float a2 = fmaf((float)r.w, 1.4629180792671596e-09f, -3.14159265358979f);
    z[0] = r1 * __cosf(a1); z[1] = r1 * __sinf(a1);
    z[2] = r2 * __cosf(a2); z[3] = r2 * __sinf(a2);
}

struct ChanCfg {
    uint32_t k0, k1;           // Philox key (seed, stream)
    float sigma, scale;        // noise sigma, 2/sigma^2
    float sigma_u, scale_u;    // uncoded branch
    int kp;                    // payload bits (K - kp CRC bits attached)
    int include_uncoded;
    unsigned long long poly;   // CRC polynomial incl. leading 1
    int deg;
    const uint32_t* enc_tab;   // [ceil(kp/4)][16][XW] u-word contribution of payload nibble q with value v (CRC bits included)
    const int16_t* tx_src;     // [E] NR transmit gather (code index or -1 = pad symbol +3)
    const int16_t* rm_dst;     // [N] NR: internal index fed by de-rate-matched position p (-1 none)
    int has_pads;              // NR: some internal positions are fed by interleaver pads (N not a multiple of 32)
};

struct DlEntryHdr { long long frame; uint32_t flags; uint32_t n_tried; };

struct SweepArgs {
    // source: llr != null -> LLR-in mode (rows indexed by frame), else Philox channel
    const float* llr;
    int in_len;
    long long frame_begin, n_frames;
    ChanCfg cc;
    int retries, run_scl, fe_mode;
    uint32_t be_mask[kMaxWords];   // phases whose bits are compared for bit errors
    const float* beta;             // caller's beta [K,K] f32 (null: |L0| ranking)
    const double* beta64;          // the same matrix widened to fp64 once per call (no conversion inside the scoring loop)
    unsigned long long* counters;
    uint16_t* frame_bit_errors;   // exact per-frame counts (K <= 512, retries <= 65535)
    uint16_t* frame_work;
    // DL API outputs (indexed by frame - frame_begin)
    uint8_t* best_bits;
    uint32_t* best_words;
    uint8_t* success;
    int32_t* n_attempts;
    int32_t* tried;
    int R;
    uint32_t* flags;
    // retry queues: entry = hdr + u[XW] + tried[XW]
    unsigned char* q_in;
    unsigned char* q_out;
    unsigned int* q_in_count;
    unsigned int* q_out_count;
    unsigned int q_capacity;
    float* llr_store;              // [q_capacity][N] channel rows of the frames that entered the retry queue (Philox mode)
    float* abs_store;              // [q_capacity][K] |L0| of the baseline best path of every queued frame (flip.py:102)
    unsigned char* gscratch;       // per-warp global scratch (WarpMem::gbytes each)
    // dl_bin_kernel: frames waiting for their next retry, binned by the information index they will flip
    int* bin_ring;                 // [K][bin_cap] queue-entry indices (-1 = not written yet), one ring per flip index
    unsigned int* bin_ctrl;        // 128-byte lines: [0] admission cursor, [32] finished frames, [64..72) statistics, [96 + 32 b + {0,1,2}] head / tail / fill of ring b
    unsigned int bin_cap;          // ring capacity (power of two)
    unsigned int inflight_target;  // frames admitted but not finished that the kernel tries to keep in the bins
    int replay_admission;          // 1: the baseline pass left no |L0| rows, admission replays the baseline decode with the trace on
};

// retry-queue entry: best u-hat of the latest attempt, tried set (by info index), the transmitted word (for the
// counters) and the row of the LLR store that holds this frame's channel LLRs (written once by the baseline pass)
template <int XW> struct DlEntry { DlEntryHdr h; uint32_t u[XW]; uint32_t tried[XW]; uint32_t u_sent[XW]; uint32_t store; uint32_t pad; };

// counters indices (include/polar_b200.h)
// Counter column of a frame group in shared memory (c-th counter of group g of the warp at base[c*FPW + g]; only the
// group leader touches it): keeps the 12 running totals out of the register file of the persistent kernels; flush()
// reduces the leaders' columns over the warp and adds them to the global block.
struct AccRef {
    uint32_t* p;
    int stride;
    __device__ __forceinline__ uint32_t& operator[](int c) const { return p[c * stride]; }
};
__host__ __device__ constexpr int acc_bytes(int MP) { return 12 * (32 / MP) * 4; }

enum { cFrames = 0, cSclFe, cSclBe, cDlFe, cDlBe, cUncFe, cUncBe, cDlWork, cNearTie, cSclUndet, cDlUndet, cRankTie, cNum };

// ---------------------------------------------------------------------------------------------------
// Channel generation for the FPW frames of a warp.  `my_frame` = global frame index of this lane's group (<0 none).
// Leaves the channel LLRs in wm.chan, returns the transmitted u (all lanes of the group) and the frame's uncoded
// bit-error count.  Scratch: the (not yet used) tree area.
// ---------------------------------------------------------------------------------------------------
#ifndef PB_RETRY_THREADS
#define PB_RETRY_THREADS 1024
#endif
#ifndef PB_SWEEP_THREADS
#define PB_SWEEP_THREADS 1024
#endif
template <int MP, int XW, typename WM>
__device__ __forceinline__ void gen_channel(const Code& code, const Tables& tb, const ChanCfg& cc, const WM& wm,
                                            long long my_frame, int lane, uint32_t (&u_sent)[XW], uint32_t& unc_err,
                                            bool want_chan, float* raw_out = nullptr, long long raw_base = 0) {
    constexpr int FPW = 32 / MP;
    const int N = code.N;
    const int xwn = N >= 32 ? N / 32 : 1;
    const int slot = lane & (MP - 1), fme = lane / MP;
    uint32_t* scr = reinterpret_cast<uint32_t*>(wm.scr);       // per-lane column: scr[w*32 + lane]
    const uint2 key = make_uint2(cc.k0, cc.k1);
    const int pwn = (cc.kp + 31) / 32;
    // ---- payload -> CRC -> u, on every lane of the group (no staging): the payload words come from Philox and
    // u = XOR over the payload nibbles of one row of the encoder table (payload bits at their information positions
    // plus their linear contribution to the CRC bits, crc.py:19-37 + polar.py:116-117)
    const int R2 = 2 * xwn;                                    // scratch rows: payload [0,xwn), codeword [2xwn,3xwn)
    uint32_t pw[XW];
#pragma unroll
    for (int w = 0; w < XW; ++w) { pw[w] = 0; u_sent[w] = 0; }
    if (my_frame >= 0) {
#pragma unroll
        for (int w4 = 0; w4 < XW; w4 += 4) {
            if (w4 < pwn) {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)my_frame, (uint32_t)(my_frame >> 32), (uint32_t)(w4 >> 2), kPurposePayload), key);
                const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (w4 + c < XW && w4 + c < pwn) {
                        uint32_t v = rr[c];
                        const int rem = cc.kp - (w4 + c) * 32;
                        if (rem < 32) v &= (1u << rem) - 1u;
                        pw[w4 + c] = v;
                    }
                }
            }
        }
        const int nq = (cc.kp + 3) >> 2;
#pragma unroll
        for (int w = 0; w < XW; ++w) {
            if (w < pwn) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int q = w * 8 + k;
                    if (q < nq) {
                        const uint32_t* row = cc.enc_tab + ((size_t)q * 16 + ((pw[w] >> (4 * k)) & 15u)) * XW;
                        if constexpr (XW % 4 == 0) {
#pragma unroll
                            for (int x4 = 0; x4 < XW; x4 += 4) {
                                const uint4 t = __ldg(reinterpret_cast<const uint4*>(row + x4));
                                u_sent[x4] ^= t.x; u_sent[x4 + 1] ^= t.y; u_sent[x4 + 2] ^= t.z; u_sent[x4 + 3] ^= t.w;
                            }
                        } else {
#pragma unroll
                            for (int x = 0; x < XW; ++x) u_sent[x] ^= __ldg(row + x);
                        }
                    }
                }
            }
        }
    }
    uint32_t x[XW];
#pragma unroll
    for (int w = 0; w < XW; ++w) x[w] = u_sent[w];
    transform_words<XW>(x, code.n);                           // codeword (polar.py:118)
    // leaders stage the codeword (and, for the uncoded reference, the payload) words so any lane can read any frame's bits
    if (slot == 0) {
#pragma unroll
        for (int w = 0; w < XW; ++w) if (w < xwn) { scr[(R2 + w) * 32 + lane] = x[w]; if (cc.include_uncoded) scr[w * 32 + lane] = pw[w]; }
    }
    __syncwarp();
    unc_err = 0;
    // frame ids of all groups, staged in smem so the item loops below can address any frame
    long long* fids = reinterpret_cast<long long*>(wm.xchg + 16);     // xchg has 64 u64; use the upper half
    if (slot == 0) fids[fme] = my_frame;
    __syncwarp();
    // ---- uncoded BPSK reference (run_fer_sweep.py:111-121) --------------------------------------------
    if (cc.include_uncoded) {
        uint32_t* cnt = reinterpret_cast<uint32_t*>(wm.xchg);
        if (lane < FPW) cnt[lane] = 0;
        __syncwarp();
        const int nblk = (cc.kp + 3) / 4;
        for (int item = lane; item < FPW * nblk; item += 32) {
            const int f = item / nblk, jb = item - f * nblk;
            const long long fr = fids[f];
            if (fr < 0) continue;
            float z[4];
            normal4(philox4x32_10(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)jb, kPurposeUncoded), key), z);
            uint32_t err = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = jb * 4 + c;
                if (j < cc.kp) {
                    const uint32_t b = (scr[(j >> 5) * 32 + f * MP] >> (j & 31)) & 1u;
                    const float y = fmaf(cc.sigma_u, z[c], 1.0f - 2.0f * (float)b);
                    err += (uint32_t)((y * cc.scale_u < 0.f) ? 1u : 0u) ^ b;
                }
            }
            if (err) atomicAdd(&cnt[f], err);
        }
        __syncwarp();
        unc_err = cnt[fme];
        __syncwarp();
    }
    if (!want_chan) return;
    // ---- coded channel: BPSK + AWGN -> LLR (run_fer_sweep.py:83-87), NR chain fused (scl_nr.py:31-35,47-48) --
    const bool nr = tb.E != 0;
    const int Eeff = nr ? tb.E : N;
    const int nblk = N >= 4 ? N / 4 : 1;
    const int rounds = (Eeff + N - 1) / N;
    for (int item = lane; item < FPW * nblk; item += 32) {
        const int f = item & (FPW - 1), jb = item / FPW;             // frame fastest: the interleaved stores are full sectors
        const long long fr = fids[f];
        if (fr < 0) continue;
        if (!nr && N >= 4) {
            // plain mother code: the four symbols of this block are four neighbouring codeword bits
            float z[4];
            normal4(philox4x32_10(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)jb, kPurposeNoise), key), z);
            const uint32_t nib = (scr[(R2 + (jb >> 3)) * 32 + f * MP] >> ((jb & 7) * 4)) & 15u;
            float4 v;
            v.x = fmaf(cc.sigma, z[0], 1.0f - 2.0f * (float)(nib & 1u)) * cc.scale;
            v.y = fmaf(cc.sigma, z[1], 1.0f - 2.0f * (float)((nib >> 1) & 1u)) * cc.scale;
            v.z = fmaf(cc.sigma, z[2], 1.0f - 2.0f * (float)((nib >> 2) & 1u)) * cc.scale;
            v.w = fmaf(cc.sigma, z[3], 1.0f - 2.0f * (float)((nib >> 3) & 1u)) * cc.scale;
            if (raw_out) *reinterpret_cast<float4*>(raw_out + (fr - raw_base) * (long long)N + jb * 4) = v;
            else {
                float* d = wm.chan + (jb * 4) * FPW + f;                 // frame-interleaved staging (stage_channel_rows)
                d[0] = v.x; d[FPW] = v.y; d[2 * FPW] = v.z; d[3 * FPW] = v.w;
            }
            continue;
        }
        // general path (NR rate matching, N = 2): this lane owns the positions p = 4 jb + c of EVERY repetition round,
        // so the copies of a position are summed in registers, round by round (= the order of load_channel), and
        // the mean is stored once (rate_match.py:19-39 + interleaver.py:26-37)
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < rounds; ++k) {
            float z[4];
            normal4(philox4x32_10(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)(k * nblk + jb), kPurposeNoise), key), z);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int p = jb * 4 + c;          // position inside this round
                const int t = k * N + p;           // transmitted position
                if (p < N && t < Eeff) {
                    const int src = nr ? (int)__ldg(&cc.tx_src[t]) : t;
                    float s = 3.0f;                // BPSK of the interleaver pad value -1 (interleaver.py:17)
                    if (src >= 0) s = 1.0f - 2.0f * (float)((scr[(R2 + (src >> 5)) * 32 + f * MP] >> (src & 31)) & 1u);
                    const float llr = fmaf(cc.sigma, z[c], s) * cc.scale;
                    if (raw_out) raw_out[(fr - raw_base) * (long long)Eeff + t] = llr;
                    acc[c] = k == 0 ? llr : acc[c] + llr;
                }
            }
        }
        if (!raw_out) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int p = jb * 4 + c;
                if (p < N) {
                    const int dst = nr ? (int)__ldg(&cc.rm_dst[p]) : p;
                    if (dst >= 0) {
                        const int cnt = nr ? (int)__ldg(&tb.rm_cnt[dst]) : 1;
                        wm.chan[dst * FPW + f] = cnt == 0 ? -1.0f : (cnt == 1 ? acc[c] : acc[c] / (float)cnt);
                    }
                }
            }
        }
    }
    __syncwarp();
    if (nr && !raw_out && cc.has_pads) {
        // internal positions fed by an interleaver pad carry 0.0 (interleaver.py:33-34)
        for (int e = lane; e < FPW * N; e += 32) {
            const int f = e & (FPW - 1), i = e / FPW;
            if (fids[f] >= 0 && __ldg(&tb.rm_cnt[i]) < 0) wm.chan[e] = 0.f;
        }
        __syncwarp();
    }
}


__device__ __forceinline__ const float* shfl_ptr(const float* p, int src) {
    return reinterpret_cast<const float*>(__shfl_sync(kFull, (unsigned long long)reinterpret_cast<uintptr_t>(p), src));
}

// Stage the channel LLRs of arbitrary frames (LLR-in mode): `row` = this lane's frame's row in the caller's buffer
// (nullptr = none).  Plain rows go through the coalescing tile; with NR rate matching the de-rate-matched row is
// gathered (rate_match.py:19-39, interleaver.py:26-37) straight into the frame-interleaved layout.
// WHOLE_ROWS: the whole tree + exchange area of the warp is free (no |L0| rows live), so plain rows may be fetched in one go.
template <int MP, typename WM, bool WHOLE_ROWS = false>
__device__ __forceinline__ void load_channel_ids(const Code& code, const Tables& tb, const WM& wm, const float* row, int lane) {
    constexpr int FPW = 32 / MP;
    const int N = code.N;
    if (tb.E == 0) {
        bool done = false;
        if constexpr (WHOLE_ROWS) done = stage_channel_rows_async<MP>(wm, N, lane, row);
        if (!done) stage_channel_rows<MP>(wm, N, lane, [&](int f) { return shfl_ptr(row, f * MP); });
        return;
    }
    for (int e = lane; e < FPW * N; e += 32) {
        const int f = e & (FPW - 1), i = e / FPW;
        const float* r = shfl_ptr(row, f * MP);              // (every lane runs the same number of iterations)
        float v = 0.f;
        if (r != nullptr) {
            const int p = tb.rm_src[i];
            if (p >= 0) {
                float acc = 0.f;
                int cnt = 0;
                for (int q = p; q < tb.E; q += N) { acc += r[q]; ++cnt; }
                v = cnt ? acc / (float)cnt : -1.0f;
            }
        }
        wm.chan[e] = v;
    }
    __syncwarp();
}

template <int MP, int LOGMAX, int HS>
struct Sweep {
    using DecU = ListDecoder<MP, LOGMAX, false, true, HS>;
    using DecF = ListDecoder<MP, LOGMAX, true, true, HS>;
    using WM = WarpMem<MP, HS>;
    using PathT = Path<LOGMAX>;
    static constexpr int XW = PathT::XW;
    static constexpr int FPW = 32 / MP;
    static constexpr uint32_t GM = DecU::GM;
    using Entry = DlEntry<XW>;

    struct Best { uint32_t u[XW]; bool pass; uint32_t flags; int lane; };   // identical on all lanes of a group

    // best candidate of a finished list decode (scl.py:183-197), broadcast to the group
    static __device__ __forceinline__ void pick_best(const Code& code, const Tables& tb, const PathT& p, int lane, uint32_t flags, Best& b) {
        const int gbase = lane & ~(MP - 1);
        uint32_t u[XW];
#pragma unroll
        for (int k = 0; k < XW; ++k) u[k] = p.alive ? p.xh[k] : 0u;
        transform_words<XW>(u, code.n);
        const bool pass = p.alive && (code.crc_deg == 0 || crc_syndrome<XW>(code, tb, u) == 0);
        uint32_t v = (pass && code.crc_deg > 0) ? p.r : 0xffu;
#pragma unroll
        for (int o = 1; o < MP; o <<= 1) v = min(v, __shfl_xor_sync(kFull, v, o));
        const uint32_t best_r = (v == 0xffu) ? 0u : v;
        const uint32_t bm = (__ballot_sync(kFull, p.alive && p.r == best_r) >> gbase) & GM;
        const int bl = gbase + (bm ? __ffs(bm) - 1 : 0);
#pragma unroll
        for (int k = 0; k < XW; ++k) b.u[k] = __shfl_sync(kFull, u[k], bl);
        b.pass = __shfl_sync(kFull, (int)pass, bl) != 0;
        b.lane = bl;
#pragma unroll
        for (int o = 1; o < MP; o <<= 1) flags |= __shfl_xor_sync(kFull, flags, o);
        b.flags = flags;
    }

    static __device__ __forceinline__ uint32_t bit_errors(const Code& code, const SweepArgs& a, const Best& b, const uint32_t (&u_sent)[XW], bool& wrong) {
        uint32_t e = 0, d = 0;
#pragma unroll
        for (int w = 0; w < XW; ++w) {
            const uint32_t x = (b.u[w] ^ u_sent[w]) & code.info_mask[w];
            d |= x;
            e += __popc(x & a.be_mask[w]);
        }
        wrong = d != 0;
        return e;
    }

    // write the per-frame results of a finished DL-SCL frame (flip.py:137-141) / count it (run_fer_sweep.py:100-109)
    static __device__ __forceinline__ void finish_dl(const Code& code, const Tables& tb, const SweepArgs& a, const WM& wm, int lane,
                                                     long long frame, const Best& b, uint32_t n_tried, const uint32_t (&u_sent)[XW], const AccRef& acc) {
        const long long idx = frame - a.frame_begin;
        if (a.llr == nullptr) {
            bool wrong;
            const uint32_t be = bit_errors(code, a, b, u_sent, wrong);
            acc[cDlBe] += be;
            acc[cDlFe] += (a.fe_mode == 0) ? (b.pass ? 0u : 1u) : (be ? 1u : 0u);
            acc[cDlUndet] += (b.pass && wrong) ? 1u : 0u;
            acc[cDlWork] += n_tried;
            if (a.frame_bit_errors) a.frame_bit_errors[idx] = (uint16_t)be;
            if (a.frame_work) a.frame_work[idx] = (uint16_t)min(n_tried, 65535u);
        }
        acc[cRankTie] += (b.flags & PB_FLAG_RANK_TIE) ? 1u : 0u;
        const int xwn = code.N >= 32 ? code.N / 32 : 1;
        if (a.best_words) for (int k = 0; k < xwn; ++k) { uint32_t w = 0;
#pragma unroll
            for (int q = 0; q < XW; ++q) if (q == k) w = b.u[q];
            a.best_words[idx * xwn + k] = w; }
        if (a.best_bits) {
            float* stash = wm.scr + lane;
#pragma unroll
            for (int k = 0; k < XW; ++k) if (k < xwn) stash[k * 32] = __uint_as_float(b.u[k]);
            write_info_bits(code, tb, stash, a.best_bits + idx * (long long)code.K);
        }
        if (a.success) a.success[idx] = (uint8_t)b.pass;
        if (a.n_attempts) a.n_attempts[idx] = (int32_t)(1 + n_tried);
        if (a.flags) a.flags[idx] = b.flags;
    }

    // Warp-aggregated append of the leaders with `need` to the output queue.  `store` < 0: first time this frame is
    // queued -> its staged channel row (chanf, element i at chanf[i * FPW]) is copied to llr_store[slot] and slot
    // becomes its store index.
    // Returns the queue slot of this lane's entry (-1: none).
    static __device__ __forceinline__ long long enqueue(const Code& code, const SweepArgs& a, int lane, bool need, long long frame, const Best& b,
                                                        const uint32_t (&tried)[XW], uint32_t n_tried, const uint32_t (&u_sent)[XW],
                                                        long long store, const float* chanf) {
        const uint32_t m = __ballot_sync(kFull, need);
        if (m == 0) return -1;
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(a.q_out_count, (unsigned int)__popc(m));
        base = __shfl_sync(kFull, base, 0);
        long long slot = -1;
        if (need) {
            const unsigned int s = base + __popc(m & ((1u << lane) - 1u));
            if (s < a.q_capacity) {
                slot = s;
                Entry* e = reinterpret_cast<Entry*>(a.q_out) + s;
                e->h.frame = frame; e->h.flags = b.flags; e->h.n_tried = n_tried;
#pragma unroll
                for (int k = 0; k < XW; ++k) { e->u[k] = b.u[k]; e->tried[k] = tried[k]; e->u_sent[k] = u_sent[k]; }
                e->store = (uint32_t)(store >= 0 ? store : s);
                e->pad = 0;
            }
        }
        if (a.llr_store != nullptr) {
            // group lanes copy the row of a newly queued frame
            const long long gslot = __shfl_sync(kFull, (store < 0) ? slot : -1ll, lane & ~(MP - 1));
            if (gslot >= 0) {
                float* dst = a.llr_store + gslot * (long long)code.N;
                for (int i = lane & (MP - 1); i < code.N; i += MP) dst[i] = chanf[i * FPW];
            }
        }
        return slot;
    }

    // rank_indices (flip.py:104-108): first untried index of argsort(|L0| @ beta) = argmin over untried.  `ab` = |L0| of
    // the reference path of this group's frame (shared memory), `tried` = the frame's tried set (by info index).
    // Returns the index to flip (0 for an invalid group); sets PB_FLAG_RANK_TIE in eflags when the best and the
    // runner-up score are within 2e-6 relative.
    static __device__ __forceinline__ int next_flip_index(const Code& code, const SweepArgs& a, const float* ab, const uint32_t (&tried)[XW],
                                                          int lane, bool valid, uint32_t& eflags) {
        const int K = code.K;
        const int slot = lane & (MP - 1);
        double m1 = 1e300, m2 = 1e300;
        int a1 = 0x7fffffff;
        auto offer = [&](double q, int j) {              // candidate score q of info index j (skipped when tried before)
            uint32_t tw = 0;
#pragma unroll
            for (int w = 0; w < XW; ++w) if (w == (j >> 5)) tw = tried[w];
            if (j < K && !((tw >> (j & 31)) & 1u)) {
                if (q < m1 || (q == m1 && j < a1)) { m2 = m1; m1 = q; a1 = j; }
                else if (q < m2) m2 = q;
            }
        };
        if (a.beta64 != nullptr && MP >= 4) {
            // q = |L0| @ beta of the warp's frames on the FP64 tensor cores: D[8 x 8] += A[8 x 4] B[4 x 8]
            // (mma.m8n8k4.f64).  Row r = lane / 4 of A and D belongs to the frame of lane 4r (MP = 4: one row per frame;
            // MP = 8: two identical rows per frame), so every lane loads |L0| of ITS OWN frame and receives scores of
            // its own frame: A[r][k] = |L0|[i0 + lane % 4], B[k][n] = beta[i0 + lane % 4][j0 + lane / 4], and the lane
            // ends up with columns j0 + 2 (lane % 4) + {0, 1}.  128 DMMAs replace 4 096 load + convert + DFMA triples
            // per retry (flip.py:104-108 is K^2 multiply-adds; the summation order inside a DMMA differs from numpy's
            // BLAS order as any other order would -- scores closer than 2e-6 relative are flagged PB_FLAG_RANK_TIE).
            const int kq = lane & 3, nq = lane >> 2;
            for (int j0 = 0; j0 < K; j0 += 8) {
                double c0 = 0.0, c1 = 0.0;
                const bool bj = j0 + nq < K;
                const double* bp = a.beta64 + j0 + nq;
                for (int i0 = 0; i0 < K; i0 += 4) {
                    const int i = i0 + kq;
                    const double av = i < K ? (double)ab[i] : 0.0;
                    const double bv = (i < K && bj) ? __ldg(bp + (size_t)i * K) : 0.0;
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
                }
                offer(c0, j0 + 2 * kq);
                offer(c1, j0 + 2 * kq + 1);
            }
        } else {
#pragma unroll
            for (int w = 0; w < XW; ++w) {
                if (w * 32 < K) {
                    double q[32 / MP];
#pragma unroll
                    for (int k = 0; k < 32 / MP; ++k) q[k] = 0.0;
                    if (a.beta64) {
                        for (int i = 0; i < K; ++i) {
                            const double x = (double)ab[i];
#pragma unroll
                            for (int k = 0; k < 32 / MP; ++k) {
                                const int j = w * 32 + slot + MP * k;
                                if (j < K) q[k] += x * __ldg(&a.beta64[(size_t)i * K + j]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32 / MP; ++k) {
                            const int j = w * 32 + slot + MP * k;
                            if (j < K) q[k] = (double)ab[j];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 32 / MP; ++k) offer(q[k], w * 32 + slot + MP * k);
                }
            }
        }
        // (tensor-core path, MP = 8: lanes l and l ^ 4 hold the SAME columns of the same frame -- reducing over them too
        //  would make every best score its own runner-up)
        const int red = (a.beta64 != nullptr && MP >= 4) ? 4 : MP;
#pragma unroll
        for (int o = 1; o < MP; o <<= 1) {
            if (o >= red) break;
            const double om1 = __shfl_xor_sync(kFull, m1, o), om2 = __shfl_xor_sync(kFull, m2, o);
            const int oa1 = __shfl_xor_sync(kFull, a1, o);
            if (om1 < m1 || (om1 == m1 && oa1 < a1)) { m2 = fmin(m1, om2); m1 = om1; a1 = oa1; }
            else m2 = fmin(m2, om1);
        }
        if (valid && m2 < 1e299 && (m2 - m1) <= 2e-6 * fmax(fabs(m1), fabs(m2))) eflags |= PB_FLAG_RANK_TIE;
        return (valid && a1 < K) ? a1 : 0;
    }

    static __device__ __forceinline__ void flush(const SweepArgs& a, int lane, const AccRef& acc) {
        if (a.counters == nullptr) return;
#pragma unroll
        for (int c = 0; c < cNum; ++c) {
            const uint32_t s = __reduce_add_sync(kFull, (lane & (MP - 1)) == 0 ? acc[c] : 0u);
            if (lane == 0 && s) atomicAdd(&a.counters[c], (unsigned long long)s);
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// Baseline pass: channel -> SCL(M) -> counters; failing frames go to the retry queue.
// ---------------------------------------------------------------------------------------------------
// TRACE (chosen when DL-SCL retries follow): the list decode records the leaf-LLR trace, and every frame that enters
// the retry queue gets the |L0| vector of its best path written to abs_store (no SC replay anywhere in DL-SCL).
template <int MP, int LOGMAX, bool TRACE = false, int NS = 0, int HS = DefaultHS<MP>::value>
__global__ void __launch_bounds__(PB_SWEEP_THREADS) sweep_kernel(const Code code_, const Tables tb, const SweepArgs a) {
    const Code code = with_static_n<NS>(code_);
    using S = Sweep<MP, LOGMAX, HS>;
    using WM = WarpMem<MP, HS>;
    using PathT = typename S::PathT;
    constexpr int FPW = 32 / MP, XW = S::XW;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    WM wm;
    const size_t kWarpBytes = WM::bytes(code.N, 0, TRACE ? code.K : 0);
    wm.carve(smem + (size_t)warp * kWarpBytes, WM::warp_scratch(a.gscratch, code.N), code.N, 0,
             TRACE ? WM::warp_trace(a.gscratch, code.N, code.K) : nullptr, TRACE ? code.K : 0);
    const bool leader = (lane & (MP - 1)) == 0;
    static_assert(cNum <= 12, "counter column");
    const AccRef acc{reinterpret_cast<uint32_t*>(smem + (size_t)wpc * kWarpBytes + (size_t)warp * acc_bytes(MP)) + lane / MP, FPW};
    if ((lane & (MP - 1)) == 0) {
#pragma unroll
        for (int c = 0; c < cNum; ++c) acc[c] = 0;
    }
    __syncwarp();
    const long long ngroups = (a.n_frames + FPW - 1) / FPW;
    for (long long g = (long long)blockIdx.x * wpc + warp; g < ngroups; g += (long long)gridDim.x * wpc) {
        const long long idx = g * FPW + lane / MP;
        const bool valid = idx < a.n_frames;
        const long long my_frame = valid ? a.frame_begin + idx : -1;
        uint32_t u_sent[XW];
        uint32_t unc = 0;
        if (a.llr) {
#pragma unroll
            for (int k = 0; k < XW; ++k) u_sent[k] = 0;
            load_channel_ids<MP, WM, true>(code, tb, wm, valid ? a.llr + idx * (long long)a.in_len : nullptr, lane);
        } else {
            gen_channel<MP, XW, WM>(code, tb, a.cc, wm, my_frame, lane, u_sent, unc, true);
        }
        uint32_t flags = 0;
        uint32_t fm[XW], fv[XW];
        PathT p;
        S::DecU::init(p, lane, valid);
        const float* chanf = wm.chan + lane / MP;
#ifndef PB_SWEEP_UMASK
#define PB_SWEEP_UMASK 0
#endif
        S::DecU::template run<TRACE, false, (TRACE || MP == 1) ? 0 : PB_SWEEP_UMASK>(code, tb.info_mask, wm, p, lane, chanf, fm, fv, flags);
        typename S::Best b;
        S::pick_best(code, tb, p, lane, flags, b);
        bool need = false;
        uint32_t tried[XW];
#pragma unroll
        for (int k = 0; k < XW; ++k) tried[k] = 0;
        if (leader && valid) {
            acc[cFrames] += 1;
            acc[cNearTie] += (b.flags & PB_FLAG_NEAR_TIE) ? 1u : 0u;
            if (a.llr == nullptr) {
                bool wrong;
                const uint32_t be = S::bit_errors(code, a, b, u_sent, wrong);
                if (a.run_scl) {
                    acc[cSclBe] += be;
                    acc[cSclFe] += (a.fe_mode == 0) ? (b.pass ? 0u : 1u) : (be ? 1u : 0u);
                    acc[cSclUndet] += (b.pass && wrong) ? 1u : 0u;
                }
                if (a.cc.include_uncoded) { acc[cUncBe] += unc; acc[cUncFe] += unc ? 1u : 0u; }
                if (a.retries < 0) {
                    if (a.frame_bit_errors) a.frame_bit_errors[idx] = (uint16_t)be;
                    if (a.frame_work) a.frame_work[idx] = 0;
                }
            }
            if (a.retries >= 0) {
                const bool pass = code.crc_deg == 0 ? true : b.pass;       // flip.py:82-88 _passes
                need = !(pass || a.retries == 0);                          // flip.py:90
            }
        }
        if (a.retries >= 0) {
            if (leader && valid && !need) S::finish_dl(code, tb, a, wm, lane, my_frame, b, 0, u_sent, acc);
            long long qslot = S::enqueue(code, a, lane, need, my_frame, b, tried, 0, u_sent, -1, chanf);
            if constexpr (TRACE) {
                // |L0| of the best path of a queued frame (flip.py:102), straight from the trace
                qslot = __shfl_sync(kFull, qslot, lane & ~(MP - 1));
                if (qslot >= 0) {
                    float* dst = a.abs_store + qslot * (long long)code.K;
                    S::DecU::trace_walk(code, wm, lane, b.lane, [&](int j, float L) { dst[j] = fabsf(L); });
                }
            }
        }
        __syncwarp();
    }
    S::flush(a, lane, acc);
}

// ---------------------------------------------------------------------------------------------------
// DL-SCL retries (flip.py:110-135): ONE persistent launch over the queue of frames whose baseline decode failed.
// Every lane group owns one frame at a time and keeps it through all of its retries (reference bits, tried set,
// flags live in registers); a group whose frame is finished pulls the next queue entry (warp-aggregated atomic),
// so all lanes stay busy until the queue is drained and there are no per-round launches or queue round trips.
// Per retry: rank q = |L0| @ beta in fp64 (flip.py:104-108), force prefix + flipped bit (flip.py:30-34), list-decode
// (flip.py:53), then finish or go on (flip.py:127-135).  |L0| of the reference path (flip.py:102,133) comes from the
// trace the previous list decode of that frame left behind (baseline: abs_store; retries: trace_walk) -- the leaf
// LLRs are never recomputed by an SC replay.
// ---------------------------------------------------------------------------------------------------
template <int MP, int LOGMAX, int NS = 0, int HS = 5>
__global__ void __launch_bounds__(MP >= 4 ? PB_RETRY_THREADS : 512) dl_retry_kernel(const Code code_, const Tables tb, const SweepArgs a) {
    const Code code = with_static_n<NS>(code_);
    using S = Sweep<MP, LOGMAX, HS>;
    using WM = WarpMem<MP, HS>;
    using PathT = typename S::PathT;
    using Entry = typename S::Entry;
    constexpr int FPW = 32 / MP, XW = S::XW;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    WM wm;
    const size_t kWarpBytes = WM::bytes(code.N, code.K, code.K);
    wm.carve(smem + (size_t)warp * kWarpBytes, WM::warp_scratch(a.gscratch, code.N), code.N, code.K, WM::warp_trace(a.gscratch, code.N, code.K), code.K);
    const int slot = lane & (MP - 1), fme = lane / MP, gbase = lane & ~(MP - 1);
    float* ab = wm.absl + fme * WM::absl_stride(code.K);      // |L0| of the reference path of this group's frame
    const bool leader = slot == 0;
    const int K = code.K;
    static_assert(cNum <= 12, "counter column");
    const AccRef acc{reinterpret_cast<uint32_t*>(smem + (size_t)wpc * kWarpBytes + (size_t)warp * acc_bytes(MP)) + lane / MP, FPW};
    if ((lane & (MP - 1)) == 0) {
#pragma unroll
        for (int c = 0; c < cNum; ++c) acc[c] = 0;
    }
    __syncwarp();
    const unsigned int n_in = min(*a.q_in_count, a.q_capacity);

    // state of the frame this group is working on (identical on all lanes of the group)
    bool active = false, exhausted = false;
    long long my_frame = -1, store = 0;
    uint32_t eflags = 0, n_tried = 0;
    uint32_t u_ref[XW], tried[XW], u_sent[XW];
#pragma unroll
    for (int k = 0; k < XW; ++k) { u_ref[k] = 0; tried[k] = 0; u_sent[k] = 0; }

    for (;;) {
        // ---- idle groups pull the next queue entry --------------------------------------------------------
        const uint32_t want = __ballot_sync(kFull, leader && !active && !exhausted);
        if (want) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(a.q_out_count, (unsigned int)__popc(want));     // q_out_count = "next entry" cursor
            base = __shfl_sync(kFull, base, 0);
            long long idx = -1;
            if (leader && !active && !exhausted) idx = (long long)base + __popc(want & ((1u << lane) - 1u));
            idx = __shfl_sync(kFull, idx, gbase);
            if (idx >= 0) {
                if (idx < (long long)n_in) {
                    const Entry* e = reinterpret_cast<const Entry*>(a.q_in) + idx;
                    my_frame = e->h.frame; eflags = e->h.flags; n_tried = e->h.n_tried; store = e->store;
#pragma unroll
                    for (int k = 0; k < XW; ++k) { u_ref[k] = e->u[k]; tried[k] = e->tried[k]; u_sent[k] = e->u_sent[k]; }
                    active = true;
                    const float* src = a.abs_store + store * (long long)K;     // written by the baseline pass
                    for (int j = slot; j < K; j += MP) ab[j] = src[j];
                } else exhausted = true;
            }
            __syncwarp();
        }
        if (!__any_sync(kFull, active)) break;
        const bool valid = active;
        // Channel rows: the LLR store (sweep mode) or the caller's buffer (API mode; with NR rate matching the
        // de-rate-matched row).  They are re-staged only when some group of the warp took a new frame -- the staged rows
        // of the other groups are unchanged (the decode never writes wm.chan).
        if (want) {
            const float* row = nullptr;
            if (valid) row = (a.llr == nullptr) ? a.llr_store + store * (long long)code.N
                                                : a.llr + (my_frame - a.frame_begin) * (long long)a.in_len;
            load_channel_ids<MP, WM>(code, tb, wm, row, lane);
        }
        const float* chanf = wm.chan + fme;
        const int jf = S::next_flip_index(code, a, ab, tried, lane, valid, eflags);
        const int pf = (int)__ldg(&tb.info_pos[jf]);
        // _force_vector (flip.py:30-34): prefix of the reference bits, then the flipped bit, rest free
        uint32_t fm[XW], fv[XW];
#pragma unroll
        for (int w = 0; w < XW; ++w) {
            const int lo = w * 32;
            const uint32_t below = (pf >= lo + 32) ? 0xffffffffu : (pf <= lo ? 0u : ((1u << (pf - lo)) - 1u));
            const uint32_t bit = (pf >= lo && pf < lo + 32) ? (1u << (pf - lo)) : 0u;
            fm[w] = code.info_mask[w] & (below | bit);
            fv[w] = (u_ref[w] & below) | (~u_ref[w] & bit);
            if (valid && jf >= lo && jf < lo + 32) tried[w] |= 1u << (jf - lo);      // tried set is indexed by info index
        }
        if (valid) n_tried += 1;
        if (leader && valid && a.tried) a.tried[(my_frame - a.frame_begin) * (long long)a.R + (n_tried - 1)] = jf;
        uint32_t flags = 0;
        PathT p;
        S::DecF::init(p, lane, valid);
        S::DecF::template run<true>(code, tb.info_mask, wm, p, lane, chanf, fm, fv, flags); // retry_with_flip (flip.py:37-62)
        typename S::Best b;
        S::pick_best(code, tb, p, lane, flags | eflags, b);
        if (valid) {
            if (leader) acc[cNearTie] += ((b.flags & PB_FLAG_NEAR_TIE) && !(eflags & PB_FLAG_NEAR_TIE)) ? 1u : 0u;
            const bool pass = code.crc_deg == 0 ? true : b.pass;
            const bool more = !(pass || (int)n_tried >= a.retries || (int)n_tried >= K);   // flip.py:111,134
            if (!more) {
                if (leader) S::finish_dl(code, tb, a, wm, lane, my_frame, b, n_tried, u_sent, acc);
                active = false;
            } else {
                // the next attempt starts from THIS attempt's best path (flip.py:127-133)
                eflags = b.flags;
#pragma unroll
                for (int k = 0; k < XW; ++k) u_ref[k] = b.u[k];
                S::DecF::trace_walk(code, wm, lane, b.lane, [&](int j, float L) { ab[j] = fabsf(L); });   // flip.py:133
            }
        }
        __syncwarp();
    }
    S::flush(a, lane, acc);
}

// ---------------------------------------------------------------------------------------------------
// DL-SCL retries, binned by flip position (the default retry kernel; dl_retry_kernel above is kept as the
// reference implementation, PB200_DL_BINNED=0).
//
// A retry re-decodes the frame with the prefix of the latest best path forced and one bit flipped
// (flip.py:30-62, scl.py:138-161).  Below the flipped phase exactly one path is alive and every decision is known,
// so those phases only rebuild state the previous attempt already had: they are SKIPPED.  The decode jump-starts
// at the flipped phase (ListDecoder::run<TRACE, JUMP>: partial sums from the prefix bits, one f / g level per tree
// height from the channel row, metric restarting at 0), the leaf LLRs of the prefix phases are taken from the
// frame's |L0| row (they were traced by the attempt that produced the prefix) and only rows >= the jump are
// re-traced.  With the shipped beta the flipped position is nearly input-independent (attempt 1 flips info index 17
// in 94 % of the frames, attempt 2 index 18, ...; mean flipped phase 66 of 128), so 40 % of the information phases and
// half of the f / g work of a retry disappear -- provided the FPW frames of a warp start at the same phase, since the
// schedule is warp-uniform.  Hence the bins: a frame waiting for its next attempt sits in the ring of the information
// index it will flip; a warp pops FPW frames of ONE ring (or, when no ring holds a full batch, the highest rings,
// starting at the lowest flipped phase of the batch), decodes them, scores the next flip (|L0| @ beta) and pushes
// every unfinished frame into its next ring.  Frames therefore migrate between warps: their state lives in the queue
// entry + the |L0| store (global, read with ld.cg after an acquire on the ring slot).
//
// Rings are multi-producer / multi-consumer and retry-free (thousands of warps pop the same ring at once):
// push = atomicAdd on the tail, write the slot (it was -1), then atomicAdd on the ring's fill count;
// pop = atomicSub of the wanted number from the fill count (anything beyond what was there is given back at once),
// atomicAdd on the head for what was got, then wait for each claimed slot to be written and reset it to -1.
// A frame enters a ring at most once (the tried set), at most inflight_target + resident frames are in the rings at
// any time and bin_cap exceeds that, so a ring never wraps onto an unconsumed slot.
// ---------------------------------------------------------------------------------------------------
// polling loads of the scheduler words: relaxed, GPU scope (served by L2, never hoisted or cached in L1)
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) { unsigned int v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_volatile_s32(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
// Release store at GPU scope: everything this thread wrote (and what it observed through __syncwarp) is visible in L2
// before the value is.  Unlike __threadfence() -- MEMBAR.SC + CCTL.IVALL, which throws away the whole L1 of the SM, beta
// and the code tables with it, three times per batch -- it invalidates nothing; the consumers read with ld.cg (L2).
__device__ __forceinline__ void st_release_s32(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

template <int MP, int LOGMAX, int NS = 0, int HS = 5>
__global__ void __launch_bounds__(MP >= 4 ? PB_RETRY_THREADS : 512) dl_bin_kernel(const Code code_, const Tables tb, const SweepArgs a) {
    const Code code = with_static_n<NS>(code_);
    using S = Sweep<MP, LOGMAX, HS>;
    using WM = WarpMem<MP, HS>;
    using PathT = typename S::PathT;
    using Entry = typename S::Entry;
    constexpr int FPW = 32 / MP, XW = S::XW;
    constexpr int EW = (int)(sizeof(Entry) / 8);                 // entry size in 8-byte words
    static_assert(sizeof(Entry) % 8 == 0, "queue entries are moved as 8-byte words");
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    WM wm;
    const size_t kWarpBytes = WM::bytes(code.N, code.K, code.K);
    wm.carve(smem + (size_t)warp * kWarpBytes, WM::warp_scratch(a.gscratch, code.N), code.N, code.K, WM::warp_trace(a.gscratch, code.N, code.K), code.K);
    const int slot = lane & (MP - 1), fme = lane / MP, gbase = lane & ~(MP - 1);
    float* ab = wm.absl + fme * WM::absl_stride(code.K);
    const bool leader = slot == 0;
    const int K = code.K;
    const AccRef acc{reinterpret_cast<uint32_t*>(smem + (size_t)wpc * kWarpBytes + (size_t)warp * acc_bytes(MP)) + lane / MP, FPW};
    if (leader) {
#pragma unroll
        for (int c = 0; c < cNum; ++c) acc[c] = 0;
    }
    __syncwarp();
    const unsigned int n_in = min(*a.q_in_count, a.q_capacity);
    // control words, one 128-byte line each so that the rings do not share lines: cursor, finished count, statistics,
    // then per ring {head, tail, fill}
    unsigned int* cursor = a.bin_ctrl;
    unsigned int* done = a.bin_ctrl + 32;
    unsigned int* stats = a.bin_ctrl + 64;    // [0] waits, [1] lost claims, [2] batches, [3] frame decodes, [4] sum of phi_start, [5] mixed plans
    auto ring_head = [&](int b) { return a.bin_ctrl + 96 + (size_t)b * 32; };
    auto ring_tail = [&](int b) { return a.bin_ctrl + 96 + (size_t)b * 32 + 1; };
    auto ring_fill = [&](int b) { return reinterpret_cast<int*>(a.bin_ctrl + 96 + (size_t)b * 32 + 2); };
    const unsigned int cap_mask = a.bin_cap - 1;
    // scratch of the scheduling warp in its (dead between decodes) tree rows: the ring fill counts
    unsigned int* cnt = reinterpret_cast<unsigned int*>(wm.ts);              // [K]
    Entry* const entries = reinterpret_cast<Entry*>(a.q_in);

    // score the next flip of this group's frame from ab[] and queue it (group-uniform arguments).  The leaders of a warp
    // that push into the same ring share one atomicAdd on its tail and one on its fill count.
    auto push_next = [&](bool go, long long eidx, const uint32_t (&tried)[XW], uint32_t eflags) {
        uint32_t fl = eflags;
        const int jn = S::next_flip_index(code, a, ab, tried, lane, go, fl);
        const bool me = go && leader;
        if (me && fl != eflags) entries[eidx].h.flags = fl;                  // PB_FLAG_RANK_TIE of this ranking
        __syncwarp();                                                         // the group's |L0| stores happen before the leader's release
        const uint32_t peers = __match_any_sync(kFull, me ? jn : -1 - lane);  // leaders pushing into the same ring
        const int first = __ffs(peers) - 1;
        unsigned int base = 0;
        if (me && lane == first) base = atomicAdd(ring_tail(jn), (unsigned int)__popc(peers));
        base = __shfl_sync(kFull, base, first);
        if (me) {
            const unsigned int pos = base + __popc(peers & ((1u << lane) - 1u));
            int* sl = a.bin_ring + (size_t)jn * a.bin_cap + (pos & cap_mask);
            while (ld_volatile_s32(sl) != -1) __nanosleep(100);               // (never taken: see the capacity argument above)
            st_release_s32(sl, (int)eidx);                                    // entry + |L0| row before the ring slot
        }
        if (me && lane == first) atomicAdd(ring_fill(jn), __popc(peers));         // (a claim that overtakes the slot write waits on the slot)
    };

    // Scheduler: ONE warp per CTA and step talks to the rings and claims work for the whole CTA (FPW frames per warp), the
    // other warps wait at a barrier -- 148-296 pollers instead of 4 736, two atomics per CTA step on the ring's own cache
    // line, and the warps of a CTA decode the same ring in lock step (same start phase, same instruction stream).  The
    // scheduler of a step is the warp that finishes the previous step FIRST, so its claim (a few L2 round trips) runs
    // while the slower warps are still pushing.
    // Fast path: keep popping the ring popped last, so rings are drained one after the other and the frames advance level
    // by level.  Slow path, when that ring runs dry: read the cursor, the finished count and every ring's fill count,
    // then leave / admit new frames / pick the fullest ring (or, when no ring holds FPW frames, the highest rings) / wait.
    __shared__ int s_cmd;                     // 0 decode, 1 admit, 2 leave
    __shared__ unsigned int s_arrive;         // warps that finished the previous step
    __shared__ int s_sticky;                  // ring popped last
    __shared__ unsigned int s_base;           // admission: first queue entry of this CTA step
    __shared__ int s_plan[1024];              // ring of each frame slot of the CTA (-1 none)
    __shared__ unsigned int s_pos[1024];      // claimed ring position
    __shared__ unsigned char s_ok[1024];      // position valid
    const int CTA_F = FPW * wpc;
    unsigned int st_batches = 0, st_decodes = 0, st_phi = 0, st_mixed = 0;
    unsigned long long st_sched_ns = 0, t_kernel0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_kernel0));
    unsigned int backoff = 500u;
    if (threadIdx.x == 0) { s_arrive = 0; s_sticky = -1; }
    __syncthreads();
    for (;;) {
        bool push_go = false;                 // what the common tail of the iteration pushes (one call site)
        long long push_idx = -1;
        uint32_t push_flags = 0;
        uint32_t tried[XW];
#pragma unroll
        for (int k = 0; k < XW; ++k) tried[k] = 0;
        unsigned int arrived = 0;
        if (lane == 0) arrived = atomicAdd(&s_arrive, 1u);
        const bool i_schedule = __shfl_sync(kFull, arrived, 0) == 0;
        if (i_schedule) {
            int cmd = 0;
            unsigned int base = 0;
            int sticky = s_sticky;
            unsigned long long t_sched0 = 0;
            if (lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_sched0));
            // claim up to CTA_F frames of ring b; fills the CTA plan; returns the number got
            auto claim_ring = [&](int b) -> int {
                int got = 0;
                unsigned int h = 0;
                if (lane == 0) {
                    const int had = atomicSub(ring_fill(b), CTA_F);
                    got = had >= CTA_F ? CTA_F : (had > 0 ? had : 0);
                    if (got < CTA_F) atomicAdd(ring_fill(b), CTA_F - got);
                    if (got > 0) h = atomicAdd(ring_head(b), (unsigned int)got);
                }
                got = __shfl_sync(kFull, got, 0); h = __shfl_sync(kFull, h, 0);
                for (int i = lane; i < CTA_F; i += 32) { s_plan[i] = b; s_ok[i] = i < got; s_pos[i] = h + i; }
                return got;
            };
            for (;;) {
                if (sticky >= 0) {
                    const int got = claim_ring(sticky);
                    if (got < CTA_F) sticky = -1;                             // ran dry: look around next time
                    if (got > 0) { cmd = 0; break; }
                }
                // one round trip: cursor, finished count and the fill counts of all rings (transiently negative while a
                // pop gives back what it could not get)
                unsigned int cur = 0, fin = 0;
                if (lane == 0) { cur = ld_volatile_u32(cursor); fin = ld_volatile_u32(done); }
                unsigned int bestc = 0, total = 0;
                int bestb = 0;
                {
                    constexpr int NB = (1 << LOGMAX) / 32;                    // K <= N: NB loads per lane, all in flight together
                    int f[NB];
#pragma unroll
                    for (int i = 0; i < NB; ++i) f[i] = (i * 32 + lane < K) ? ld_volatile_s32(ring_fill(i * 32 + lane)) : 0;
#pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        const int b = i * 32 + lane;
                        const unsigned int c = f[i] > 0 ? (unsigned int)f[i] : 0u;
                        if (b < K) cnt[b] = c;
                        total += c;
                        if (c > bestc) { bestc = c; bestb = b; }
                    }
                }
                cur = __shfl_sync(kFull, cur, 0); fin = __shfl_sync(kFull, fin, 0);
                if (fin >= n_in) { cmd = 2; break; }
                const unsigned int admitted = cur < n_in ? cur : n_in;
                if (cur < n_in && admitted - fin < a.inflight_target) {
                    if (lane == 0) base = atomicAdd(cursor, (unsigned int)CTA_F);
                    base = __shfl_sync(kFull, base, 0);
                    cmd = 1;
                    break;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned int oc = __shfl_xor_sync(kFull, bestc, o);
                    const int ob = __shfl_xor_sync(kFull, bestb, o);
                    total += __shfl_xor_sync(kFull, total, o);
                    if (oc > bestc || (oc == bestc && ob > bestb)) { bestc = oc; bestb = ob; }
                }
                int got_any = 0;
                if (total > 0) {
                    __syncwarp();
                    if (bestc >= (unsigned int)FPW) {
                        const int got = claim_ring(bestb);
                        if (got == CTA_F) sticky = bestb;
                        got_any = got;
                    } else {
                        // no ring holds a warp's worth: the highest rings first (similar start phases end up in one warp)
                        if (lane == 0) {
                            int g = 0;
                            for (int b = K - 1; b >= 0 && g < CTA_F; --b)
                                for (unsigned int c = cnt[b]; c > 0 && g < CTA_F; --c) s_plan[g++] = b;
                            for (; g < CTA_F; ++g) s_plan[g] = -1;
                            g = 0;
                            while (g < CTA_F) {                               // one claim per run of equal rings
                                const int b = s_plan[g];
                                if (b < 0) { s_ok[g] = 0; ++g; continue; }
                                int n = 1;
                                while (g + n < CTA_F && s_plan[g + n] == b) ++n;
                                const int had = atomicSub(ring_fill(b), n);
                                const int got = had >= n ? n : (had > 0 ? had : 0);
                                if (got < n) atomicAdd(ring_fill(b), n - got);
                                unsigned int h = 0;
                                if (got > 0) h = atomicAdd(ring_head(b), (unsigned int)got);
                                for (int i = 0; i < n; ++i) { s_ok[g + i] = i < got; s_pos[g + i] = h + i; }
                                got_any += got;
                                g += n;
                            }
                        }
                        got_any = __shfl_sync(kFull, got_any, 0);
                    }
                }
                if (got_any > 0) { cmd = 0; break; }
                if (lane == 0) atomicAdd(&stats[total == 0 ? 0 : 1], 1u);      // frames are in flight elsewhere / lost the race
                __nanosleep(backoff);
                backoff = backoff < 8000u ? backoff * 2u : backoff;
            }
            backoff = 500u;
            if (lane == 0) { s_cmd = cmd; s_base = base; s_sticky = sticky; }
            if (lane == 0) {                                                  // statistics: time this CTA's scheduler spent
                unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                st_sched_ns += t1 - t_sched0;
            }
        }
        __syncthreads();
        const int cmd = s_cmd;
        if (cmd == 2) break;
        const int myi = warp * FPW + fme;                                     // this group's frame slot in the CTA plan
        const int myb = s_plan[myi];
        bool valid = cmd == 0 && s_ok[myi] != 0;
        const unsigned int mypos = s_pos[myi];
        const unsigned int adm_base = s_base;
        if (i_schedule && lane == 0) s_arrive = 0;                            // (every warp of the CTA has arrived: it is past the barrier)
        __syncthreads();                                                      // the plan may be rewritten from here on
        // Admission (cmd 1): the baseline pass queued these frames WITHOUT a trace (tracing every frame costs a quarter of
        // the baseline pass, and most frames never need it).  Their first |L0| row comes from replaying the baseline
        // decode here, trace on, nothing forced -- "attempt 0", the same decode bit for bit -- after which they are
        // ranked and queued like any other attempt.
        // (replay_admission = 0: the baseline pass traced every frame and left the |L0| rows in abs_store; admission then
        //  only ranks them)
        const bool fresh = cmd == 1 && a.replay_admission != 0;
        long long eidx = -1;
        if (cmd == 1 && !fresh) {
            push_idx = (long long)adm_base + myi;
            push_go = push_idx < (long long)n_in;
            if (push_go) {
                push_flags = __ldcg(&entries[push_idx].h.flags);
                const float* src = a.abs_store + push_idx * (long long)K;
                for (int j = slot; j < K; j += MP) ab[j] = __ldcg(src + j);
            }
        } else if (fresh) {
            eidx = (long long)adm_base + myi;
            valid = eidx < (long long)n_in;
        } else if (valid && leader) {
            int* sl = a.bin_ring + (size_t)myb * a.bin_cap + (mypos & cap_mask);
            int v;
            while ((v = ld_volatile_s32(sl)) < 0) __nanosleep(50);
            *reinterpret_cast<volatile int*>(sl) = -1;
            eidx = v;                                                         // the entry / |L0| loads below depend on it and read L2 (ld.cg)
        }
        if (__any_sync(kFull, valid)) {
            if (!fresh) eidx = __shfl_sync(kFull, eidx, gbase);
            __syncwarp();

            // ---- the frames of this batch ----------------------------------------------------------------------
            long long my_frame = -1, store = 0;
            uint32_t eflags = 0, n_tried = 0;
            uint32_t u_ref[XW], u_sent[XW];
#pragma unroll
            for (int k = 0; k < XW; ++k) { u_ref[k] = 0; u_sent[k] = 0; }
            if (valid) {
                unsigned long long w[EW];
                const unsigned long long* ep = reinterpret_cast<const unsigned long long*>(entries + eidx);
#pragma unroll
                for (int k = 0; k < EW; ++k) w[k] = __ldcg(ep + k);
                Entry e;
                memcpy(&e, w, sizeof(Entry));
                my_frame = e.h.frame; eflags = e.h.flags; n_tried = e.h.n_tried; store = e.store;
#pragma unroll
                for (int k = 0; k < XW; ++k) { u_ref[k] = e.u[k]; tried[k] = e.tried[k]; u_sent[k] = e.u_sent[k]; }
            }
            {
                const float* row = nullptr;
                // (sweep mode: the LLR-store row of a frame is its queue slot, so the row loads do not wait for the entry)
                if (valid) row = (a.llr == nullptr) ? a.llr_store + eidx * (long long)code.N
                                                    : a.llr + (my_frame - a.frame_begin) * (long long)a.in_len;
                load_channel_ids<MP, WM, true>(code, tb, wm, row, lane);
            }
            const float* chanf = wm.chan + fme;
            const int jf = (valid && !fresh) ? myb : 0;                           // the ring IS the index to flip
            const int pf = fresh ? 0 : (int)__ldg(&tb.info_pos[jf]);
            // start of the warp's decode: the lowest flipped phase of the batch, rounded down to a phase pair
            int pmin = valid ? pf : code.N;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pmin = min(pmin, __shfl_xor_sync(kFull, pmin, o));
            const int phi_start = pmin & ~1;
            {   // statistics of this warp (flushed once, at the end)
                const uint32_t vm = __ballot_sync(kFull, valid && leader);
                const uint32_t same = __ballot_sync(kFull, !valid || pf == pmin);
                st_batches += 1; st_decodes += __popc(vm); st_phi += (unsigned int)phi_start; st_mixed += same != kFull ? 1u : 0u;
            }
            int jstart = 0;
#pragma unroll
            for (int w = 0; w < XW; ++w) {
                const int lo = w * 32;
                const uint32_t below = (phi_start >= lo + 32) ? 0xffffffffu : (phi_start <= lo ? 0u : ((1u << (phi_start - lo)) - 1u));
                jstart += __popc(code.info_mask[w] & below);
            }
            // _force_vector (flip.py:30-34): prefix of the reference bits, then the flipped bit, rest free
            uint32_t fm[XW], fv[XW];
#pragma unroll
            for (int w = 0; w < XW; ++w) {
                const int lo = w * 32;
                const uint32_t below = (pf >= lo + 32) ? 0xffffffffu : (pf <= lo ? 0u : ((1u << (pf - lo)) - 1u));
                const uint32_t bit = (pf >= lo && pf < lo + 32) ? (1u << (pf - lo)) : 0u;
                fm[w] = fresh ? 0u : (code.info_mask[w] & (below | bit));
                fv[w] = fresh ? 0u : ((u_ref[w] & below) | (~u_ref[w] & bit));
                if (valid && !fresh && jf >= lo && jf < lo + 32) tried[w] |= 1u << (jf - lo);      // tried set is indexed by info index
            }
            if (valid && !fresh) n_tried += 1;
            if (leader && valid && !fresh && a.tried) a.tried[(my_frame - a.frame_begin) * (long long)a.R + (n_tried - 1)] = jf;
            uint32_t flags = 0;
            PathT p;
            S::DecF::init(p, lane, valid);
            S::DecF::template run<true, true>(code, tb.info_mask, wm, p, lane, chanf, fm, fv, flags, phi_start, jstart, pf & ~1);   // retry_with_flip (flip.py:37-62)
            typename S::Best b;
            S::pick_best(code, tb, p, lane, flags | eflags, b);
            bool more = false;
            if (valid) {
                if (leader) acc[cNearTie] += ((b.flags & PB_FLAG_NEAR_TIE) && !(eflags & PB_FLAG_NEAR_TIE)) ? 1u : 0u;
                const bool pass = code.crc_deg == 0 ? true : b.pass;
                more = !(pass || (int)n_tried >= a.retries || (int)n_tried >= K);          // flip.py:111,134
                if (!more && leader) S::finish_dl(code, tb, a, wm, lane, my_frame, b, n_tried, u_sent, acc);
            }
            {
                const uint32_t fmask_done = __ballot_sync(kFull, valid && !more && leader);
                if (lane == 0 && fmask_done) atomicAdd(done, (unsigned int)__popc(fmask_done));
            }
            __syncwarp();
            if (more) {
                // the next attempt starts from THIS attempt's best path (flip.py:127-133): new entry, new |L0| rows >= jstart
                if (leader) {
                    Entry e;
                    e.h.frame = my_frame; e.h.flags = b.flags; e.h.n_tried = n_tried;
#pragma unroll
                    for (int k = 0; k < XW; ++k) { e.u[k] = b.u[k]; e.tried[k] = tried[k]; e.u_sent[k] = u_sent[k]; }
                    e.store = (uint32_t)store; e.pad = 0;
                    unsigned long long w[EW];
                    memcpy(w, &e, sizeof(Entry));
                    unsigned long long* ep = reinterpret_cast<unsigned long long*>(entries + eidx);
#pragma unroll
                    for (int k = 0; k < EW; ++k) ep[k] = w[k];
                }
                float* dst = a.abs_store + eidx * (long long)K;
                S::DecF::trace_walk(code, wm, lane, b.lane, [&](int j, float L) { const float v = fabsf(L); ab[j] = v; dst[j] = v; }, jstart);   // flip.py:133
                for (int j = slot; j < jstart; j += MP) ab[j] = __ldcg(dst + j);          // prefix rows: traced by earlier attempts
            }
            push_go = more; push_idx = eidx; push_flags = b.flags;
        }   // (decode path)
        __syncwarp();
        push_next(push_go, push_idx, tried, push_flags);
        __syncwarp();
    }
    if (lane == 0) {
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        atomicMax(&stats[6], (unsigned int)(st_sched_ns >> 10));              // longest time a warp spent scheduling (~us)
        atomicMax(&stats[7], (unsigned int)((t1 - t_kernel0) >> 10));         // longest warp lifetime (~us)
    }
    if (lane == 0 && st_batches) {
        atomicAdd(&stats[2], st_batches); atomicAdd(&stats[3], st_decodes); atomicAdd(&stats[4], st_phi); atomicAdd(&stats[5], st_mixed);
    }
    S::flush(a, lane, acc);
}

// Channel only: msg[B,K] u8 and raw LLRs [B, E or N] (same Philox stream as sweep_kernel).
template <int LOGMAX>
__global__ void channel_kernel(const Code code, const Tables tb, const SweepArgs a, uint8_t* msg, float* llr) {
    constexpr int MP = 4, FPW = 8;
    constexpr int XW = BitsCfg<LOGMAX>::XW;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    using WM = WarpMem<MP, 5>;
    WM wm;
    wm.carve(smem + (size_t)warp * WM::bytes(code.N), WM::warp_scratch(a.gscratch, code.N), code.N);
    const long long ngroups = (a.n_frames + FPW - 1) / FPW;
    for (long long g = (long long)blockIdx.x * wpc + warp; g < ngroups; g += (long long)gridDim.x * wpc) {
        const long long idx = g * FPW + lane / MP;
        const bool valid = idx < a.n_frames;
        const long long my_frame = valid ? a.frame_begin + idx : -1;
        uint32_t u_sent[XW];
        uint32_t unc;
        gen_channel<MP, XW, WM>(code, tb, a.cc, wm, my_frame, lane, u_sent, unc, true, llr, a.frame_begin);
        if (msg && valid && (lane & (MP - 1)) == 0) {
            float* stash = wm.scr + lane;
            const int xwn = code.N >= 32 ? code.N / 32 : 1;
#pragma unroll
            for (int k = 0; k < XW; ++k) if (k < xwn) stash[k * 32] = __uint_as_float(u_sent[k]);
            write_info_bits(code, tb, stash, msg + idx * (long long)code.K);
        }
        __syncwarp();
    }
}

// choose_flip_index (flip.py:13-27): one block per row, argmin(abs_l0 @ beta) / argmin(abs_l0), first minimum.
static __global__ void flip_index_kernel(const float* __restrict__ abs_l0, const float* __restrict__ beta, int32_t* __restrict__ out, int K) {
    extern __shared__ double qs[];
    const long long row = blockIdx.x;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        double q = 0.0;
        if (beta) for (int i = 0; i < K; ++i) q += (double)abs_l0[row * K + i] * (double)beta[(size_t)i * K + j];
        else q = (double)abs_l0[row * K + j];
        qs[j] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 0;
        for (int j = 1; j < K; ++j) if (qs[j] < qs[best]) best = j;
        out[row] = best;
    }
}

// encode_rate_matched (scl_nr.py:23-35): payload[B,Kp] -> CRC -> encode -> interleave -> rate-match; one thread per frame.
static __global__ void nr_encode_kernel(const Code code, const Tables tb, const ChanCfg cc, const uint8_t* __restrict__ payload, int8_t* __restrict__ tx,
                                 long long B, int E) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= B) return;
    uint32_t u[kMaxWords];
#pragma unroll
    for (int k = 0; k < kMaxWords; ++k) u[k] = 0;
    const unsigned long long low = cc.poly & ((1ull << cc.deg) - 1ull);
    unsigned long long reg = 0;
    auto put = [&](int j, uint32_t b) {
        const int pos = tb.info_pos[j];
#pragma unroll
        for (int k = 0; k < kMaxWords; ++k) if (k == (pos >> 5)) u[k] |= b << (pos & 31);
    };
    for (int j = 0; j < cc.kp; ++j) {
        const uint32_t b = payload[f * cc.kp + j] & 1u;
        put(j, b);
        const unsigned long long top = ((reg >> (cc.deg - 1)) & 1ull) ^ b;
        reg = (reg << 1) & ((1ull << cc.deg) - 1ull);
        if (top) reg ^= low;
    }
    if (cc.deg > 0) for (int t = 0; t < cc.deg && cc.kp + t < code.K; ++t) put(cc.kp + t, (uint32_t)((reg >> (cc.deg - 1 - t)) & 1ull));
    transform_words<kMaxWords>(u, code.n);
    for (int t = 0; t < E; ++t) {
        const int src = cc.tx_src[t];
        int8_t v = -1;
        if (src >= 0) {
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < kMaxWords; ++k) if (k == (src >> 5)) w = u[k];
            v = (int8_t)((w >> (src & 31)) & 1u);
        }
        tx[f * E + t] = v;
    }
}

}  // namespace pb
