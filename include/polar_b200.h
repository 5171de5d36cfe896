/*
 * polar_b200.h -- C-ABI of the B200-native polar decoding engine (libpolar_b200.so).
 *
 * The reference (heimrih/polar_code, package dl_scl_polar) has no FFI layer: its boundary is a set
 * of per-frame Python functions over 1-D NumPy arrays.  Each entry point below is
 * the BATCHED form of one of those functions and names the reference interface it replaces
 * (file:line relative to /root/reference/).  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add; polar_code_b200/dl_scl_polar/ is that stub written out.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - `d_` pointers are DEVICE pointers, `h_` pointers are HOST pointers; all buffers are caller-owned
 *   - every call returns 0 on success or a negative PB200_E* code; pb200_last_error() gives the text
 *   - calls only enqueue work on `stream` unless stated otherwise (the *_host calls synchronise)
 *   - a handle is not thread-safe; use one handle per GPU / rank.  Decode calls of one handle may be enqueued on
 *     different streams (each stream gets its own scratch); pb200_sweep / pb200_dlscl_decode_batch share the
 *     handle's retry queues and must be stream-ordered with respect to each other
 *   - there is NO CPU fallback: without a CUDA device pb200_create fails with PB200_ECUDA
 */
#ifndef POLAR_B200_H
#define POLAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_OK 0
#define PB200_EINVAL (-1)   /* bad argument: the Python mirror raises ValueError              */
#define PB200_ECUDA (-2)    /* CUDA runtime failure / no device                                */
#define PB200_ERANGE (-3)   /* index out of range: the Python mirror raises IndexError         */
#define PB200_ENOSUP (-4)   /* valid for the reference but outside this build's limits         */

#define PB200_MAX_N 512     /* code length limit (power of two)                                */
#define PB200_MAX_M 8       /* list size limit                                                 */

/* per-frame flag bits written to `flags` outputs */
#define PB200_FLAG_NEAR_TIE 1u  /* two competing path metrics within ~1e-6 relative at a prune   */
#define PB200_FLAG_RANK_TIE 2u  /* DL-SCL: two flip scores within ~1e-6 relative when ranking    */
#define PB200_FLAG_BAD_FORCE 4u /* force_info_bits entry outside {-1,0,1} (scl.py:141-144)       */

typedef struct pb200_engine pb200_engine;

const char *pb200_last_error(void);
int pb200_version(void);
int pb200_device_count(void);

/* ---- code construction: polar/polar.py:85-103 construct_info_set (host, float64) -------------
 * method 0 = "gaussian", 1 = "polarization"; writes K sorted int32 indices. */
int pb200_construct_info_set(int N, int K, int method, double design_snr_db, int32_t *h_info_set);

/* ---- engine: one per (device, N, info_set, crc polynomial) -----------------------------------
 * crc_poly: hex string incl. the leading 1 (polar/crc.py:10-16), or NULL for crc=None
 * (scl.py:190-197 then always returns candidate 0). */
int pb200_create(pb200_engine **out, int device, int N, const int32_t *h_info_set, int K, const char *crc_poly);
void pb200_destroy(pb200_engine *e);

/* NR rate matching fused into the LLR load: nr/polar/rate_match.py:19-39 derate_match_polar followed
 * by nr/polar/interleaver.py:26-37 subblock_deinterleave.  After this call the decode entry points
 * take E-long inputs (in_len = E).  E = 0 switches it off again. */
int pb200_set_rate_matching(pb200_engine *e, int E);

/* ---- encoder / CRC ---------------------------------------------------------------------------
 * polar/polar.py:106-119 encode (and run_ber_sweep.py:65-70, scl_nr.py:17-20): msg[B,K] -> code[B,N] */
int pb200_encode_batch(pb200_engine *e, const uint8_t *d_msg, uint8_t *d_code, int64_t B, void *stream);
/* polar/crc.py:19-37 attach_crc: msg[B,L] -> out[B,L+deg];  polar/crc.py:40-56 check_crc: msg[B,L] -> ok[B] */
int pb200_crc_attach_batch(const char *poly, const uint8_t *d_msg, uint8_t *d_out, int64_t B, int L, void *stream);
int pb200_crc_check_batch(const char *poly, const uint8_t *d_msg, uint8_t *d_ok, int64_t B, int L, void *stream);
/* nr/polar/scl_nr.py:23-35 encode_rate_matched: payload[B,Kp] -> tx[B,E] as int8 (pad value -1 kept) */
int pb200_nr_encode_batch(pb200_engine *e, const uint8_t *d_payload, int8_t *d_tx, int64_t B, int E, void *stream);

/* ---- decoders --------------------------------------------------------------------------------
 * Outputs of one decode_scl call per frame (polar/scl.py:203-209), any pointer may be NULL:
 *   cand[B,M,K] u8        candidates in metric order       metrics[B,M] f64 (rows >= n_cand untouched)
 *   info_llrs[B,M,K] f32  leaf LLR seen at each info phase n_cand[B], best_idx[B] i32
 *   best_bits[B,K] u8     = cand[best_idx]                 best_words[B,max(N/32,1)] u32: u-hat packed
 *   crc_ok[B] u8          check_crc(best_bits)             flags[B] u32 (PB200_FLAG_*)
 */
typedef struct {
    uint8_t *cand;
    double *metrics;
    float *info_llrs;
    int32_t *n_cand;
    int32_t *best_idx;
    uint8_t *best_bits;
    uint32_t *best_words;
    uint8_t *crc_ok;
    uint32_t *flags;
} pb200_scl_out;

/* polar/polar.py:130-168 sc_decode: llr[B,in_len] f32 -> bits[B,K] u8 */
int pb200_sc_decode_batch(pb200_engine *e, const float *d_llr, int64_t B, int in_len, uint8_t *d_bits, void *stream);

/* polar/scl.py:108-209 decode_scl; d_force: NULL or int8[B,K] (-1 free, 0/1 forced) */
int pb200_scl_decode_batch(pb200_engine *e, const float *d_llr, int64_t B, int in_len, const int8_t *d_force, int M,
                           const pb200_scl_out *out, void *stream);

/* dlscl/flip.py:65-141 decode_with_retries; d_beta: NULL (|L0| ranking) or f32[K,K].
 * Returns the LAST attempt (flip.py:137): best_bits[B,K], success[B] u8, n_attempts[B] i32 (incl.
 * baseline), tried[B,max(retries,1)] i32 (-1 padded), flags[B]; any output may be NULL. */
typedef struct {
    uint8_t *best_bits;
    uint32_t *best_words;
    uint8_t *success;
    int32_t *n_attempts;
    int32_t *tried;
    uint32_t *flags;
} pb200_dl_out;
int pb200_dlscl_decode_batch(pb200_engine *e, const float *d_llr, int64_t B, int in_len, int M, int retries,
                             const float *d_beta, const pb200_dl_out *out, void *stream);

/* dlscl/flip.py:13-27 choose_flip_index over rows: abs_l0[B,K] f32 (+ beta[K,K]) -> idx[B] i32 */
int pb200_choose_flip_index_batch(const float *d_abs_l0, const float *d_beta, int32_t *d_idx, int64_t B, int K,
                                  void *stream);

/* Host-buffer convenience used for end-to-end timing: h_llr[B,in_len] (pinned for full speed) is copied
 * in chunks on internal streams, decoded (decode_scl, list size M) and best_bits[B,K], crc_ok[B],
 * flags[B] are copied back.  Synchronous. */
int pb200_scl_decode_host(pb200_engine *e, const float *h_llr, int64_t B, int in_len, int M, uint8_t *h_best_bits,
                          uint8_t *h_crc_ok, uint32_t *h_flags);
/* The same with the LLR rows given as IEEE binary16 (OPTIONAL ingest format: half the host->device bytes, which is what
 * bounds the host-buffer path).  Rows are widened to fp32 on load -- exactly -- so the result equals
 * pb200_scl_decode_host on the same values as fp32; quantising float LLRs to binary16 is the CALLER's decision and is
 * outside the parity contract of scl.py:108-209. */
int pb200_scl_decode_host_f16(pb200_engine *e, const uint16_t *h_llr_f16, int64_t B, int in_len, int M,
                              uint8_t *h_best_bits, uint8_t *h_crc_ok, uint32_t *h_flags);

/* ---- Monte-Carlo sweeps (channel + decode + counters fused on the GPU) -------------------------
 * eval/run_fer_sweep.py:60-121 and eval/run_ber_sweep.py:112-181.  Frames are numbered globally;
 * frame f of stream `stream_id` always draws the same Philox4x32-10 numbers, so results do not depend
 * on how [frame_begin, frame_begin+n_frames) is split over launches or ranks.
 *
 * counters (int64, ADDED to, so a rank can accumulate and then all-reduce the block):
 *   [0] frames            [1] scl_frame_errors   [2] scl_bit_errors   [3] dl_frame_errors
 *   [4] dl_bit_errors     [5] uncoded_frame_err  [6] uncoded_bit_err  [7] dl_attempts_minus_1 (sum)
 *   [8] near_tie_frames   [9] scl_undetected (CRC pass but wrong word)   [10] dl_undetected
 *   [11] rank_tie_frames  [12..15] reserved
 */
#define PB200_NCOUNTERS 16
typedef struct {
    int M;                 /* list size                                                          */
    int retries;           /* DL-SCL retries; < 0 = do not run DL-SCL                            */
    int run_scl;           /* also count plain SCL (run_fer_sweep runs both, :36-37)             */
    int k_payload;         /* payload bits; K - k_payload CRC bits are attached (crc.py:19-37)   */
    int E;                 /* transmitted bits; 0 or N = no rate matching; else NR chain         */
    int frame_error_mode;  /* 0: CRC failure of the returned word (run_fer_sweep.py:91-94)
                              1: payload mismatch (run_ber_sweep.py:156-157)                    */
    int bit_error_span;    /* number of leading info bits compared (K for FER sweep, k_payload BER) */
    int include_uncoded;   /* run_fer_sweep.py:111-121                                           */
    double noise_var;      /* sigma^2 of the coded channel                                       */
    double noise_var_uncoded;
    uint64_t seed;
    uint32_t stream_id;    /* e.g. SNR-point index                                               */
    int64_t frame_begin;
    int64_t n_frames;
} pb200_sweep_cfg;
/* d_beta: NULL or f32[K,K]; d_counters: int64[PB200_NCOUNTERS] on the device;
 * d_frame_bit_errors: NULL or u16[n_frames] per-frame bit errors of the LAST decoder run (for the
 * adaptive stop of run_ber_sweep.py:127; exact: K <= 512 < 65536), d_frame_work: NULL or u16[n_frames] attempts-1. */
int pb200_sweep(pb200_engine *e, const pb200_sweep_cfg *cfg, const float *d_beta, int64_t *d_counters,
                uint16_t *d_frame_bit_errors, uint16_t *d_frame_work, void *stream);

/* Generate the channel only (payload -> CRC -> encode -> [NR] -> BPSK + AWGN -> LLR), same Philox
 * stream as pb200_sweep: msg[B,K] u8 (NULL ok), llr[B,E or N] f32. */
int pb200_channel_batch(pb200_engine *e, const pb200_sweep_cfg *cfg, uint8_t *d_msg, float *d_llr, void *stream);

/* Tuning aid: scheduler statistics of the engine's last binned DL-SCL retry launch (out8: waits on empty rings, lost
 * claims, batches, frame decodes, sum of the batches' start phases, batches that mixed rings, 0, 0).  Synchronises. */
int pb200_debug_bin_stats(pb200_engine *e, unsigned int *out8);

/* Introspection for benchmarks: resident warps per SM, frames per warp, dynamic smem per CTA of the
 * decode kernel that (M, forced) selects. */
int pb200_kernel_info(pb200_engine *e, int M, int *warps_per_cta, int *ctas_per_sm, int *smem_bytes, int *regs);


/* ==== NR LDPC (toy family of the comparison CLI) ===========================
 * Reference: dl_scl_polar/nr/ldpc/{basegraphs,builder,encode,rate_match,decode_nms}.py and the nr_ldpc branch
 * of eval/run_ber_sweep.py (:134-136,146-149,163-164,258-271).  One thread per frame, FLOAT64 arithmetic with
 * the reference's operation order: hard decisions, iteration counts and posteriors are bit-identical to the
 * reference's on the same float64 LLRs. */
#define PB200_LDPC_MAX_N 4096

typedef struct pb200_ldpc pb200_ldpc;

/* basegraphs.py:39-42 load_base_graph + builder.py:20-30 build_h_matrix: dense H u8[3Z][6Z] on the HOST
 * (h_H may be NULL to query the shape). */
int pb200_ldpc_build_h(int bg, int Z, uint8_t *h_H, int *m, int *n);

/* handle for one dense 0/1 parity-check matrix h_H[m][n] (host, row-major) */
int pb200_ldpc_create(pb200_ldpc **out, int device, const uint8_t *h_H, int m, int n);
void pb200_ldpc_destroy(pb200_ldpc *e);

/* Host side of the encoder and of the launch planning (no GPU needed; exported so that the host logic is testable
 * on its own).  parity_generator: the elimination of encode.py:8-49 depends on H only, so parity = G * payload with
 * h_G[(n-k)][kw] bit rows (kw = max(1, ceil(k/32))); h_C[n_check][kw] are the rows whose non-zero product reproduces
 * the reference's "Linear system over GF(2) has no solution" (h_G / h_C may be NULL; h_C needs room for m rows).
 * layers: first row of every layer of consecutive, mutually column-disjoint rows (h_layer_ptr[n_layers+1], room for
 * m+1), and the lanes per frame of the group-per-frame kernels (0 = thread-per-frame mapping). */
int pb200_ldpc_parity_generator(const uint8_t *h_H, int m, int n, int k, uint32_t *h_G, uint32_t *h_C, int *n_check);
int pb200_ldpc_layers(const uint8_t *h_H, int m, int n, int32_t *h_layer_ptr, int *n_layers, int *group_lanes);

/* encode.py:52-66 encode_ldpc: payload[B,k] u8 -> code[B,n] u8 (systematic, parity by GF(2) elimination with free
 * variables 0).  d_status: NULL or u8[B], 1 where the reference raises "Linear system over GF(2) has no solution". */
int pb200_ldpc_encode_batch(pb200_ldpc *e, const uint8_t *d_payload, int k, uint8_t *d_code, uint8_t *d_status, int64_t B,
                            void *stream);
/* rate_match.py:8-15 rate_match_ldpc: code[B,N] -> out[B,E];  :18-38 derate_match_ldpc: llr[B,E] f64 -> out[B,N] f64 */
int pb200_ldpc_rate_match_batch(const uint8_t *d_code, int N, int E, uint8_t *d_out, int64_t B, void *stream);
int pb200_ldpc_derate_match_batch(const double *d_llr, int E, int N, double *d_out, int64_t B, void *stream);

/* decode_nms.py:8-40 decode_ldpc_nms: llr[B,in_len] f64 (in_len = n, or E with the de-rate-matching fused into the
 * load) -> hard[B,n] u8, posterior[B,n] f64, iters_used[B] i32, parity_ok[B] u8; any output may be NULL. */
int pb200_ldpc_decode_batch(pb200_ldpc *e, const double *d_llr, int64_t B, int in_len, int max_iter, double alpha,
                            int early_stop, uint8_t *d_hard, double *d_posterior, int32_t *d_iters, uint8_t *d_ok,
                            void *stream);

/* Fused Monte-Carlo sweep of run_ber_sweep.py:112-181 for --scheme nr_ldpc; Philox convention of pb200_sweep.
 * counters (int64[PB200_NCOUNTERS], ADDED to): [0] frames [1] frame errors [2] payload bit errors [7] iterations. */
typedef struct {
    int k_payload;         /* payload bits                                                        */
    int k_crc;             /* CRC bits appended (0 = none); k_payload + k_crc must equal n - m    */
    int E;                 /* transmitted bits (rate_match_ldpc)                                  */
    int max_iter;
    int early_stop;
    double alpha;
    const char *crc_poly;  /* hex string incl. the leading 1; used when k_crc > 0                 */
    double noise_var;
    uint64_t seed;
    uint32_t stream_id;
    int64_t frame_begin;
    int64_t n_frames;
} pb200_ldpc_sweep_cfg;
int pb200_ldpc_sweep(pb200_ldpc *e, const pb200_ldpc_sweep_cfg *cfg, int64_t *d_counters, uint16_t *d_frame_bit_errors,
                     uint16_t *d_frame_work, void *stream);
/* channel only, same Philox stream: payload[B,k_payload] u8 (NULL ok), llr[B,E] f64 */
int pb200_ldpc_channel_batch(pb200_ldpc *e, const pb200_ldpc_sweep_cfg *cfg, uint8_t *d_payload, double *d_llr,
                             void *stream);

#ifdef __cplusplus
}
#endif
#endif /* POLAR_B200_H */
