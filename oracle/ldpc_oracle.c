/*
 * ldpc_oracle.c -- CPU ORACLE for the toy NR-LDPC family.  TEST INFRASTRUCTURE, not the
 * product: only tests/, bench.py's CPU legs and __graft_entry__.smoke() may load it, and only as the checker.
 *
 * Plain C / float64 restatement of the reference package dl_scl_polar/nr/ldpc (file:line cited per function,
 * relative to /root/reference/).  Pinned against the reference itself: oracle/gen_golden.py imports the reference
 * and writes tests/golden/ldpc.npz (H matrices, encodings, de-rate-matched vectors, decode_ldpc_nms outputs);
 * tests/test_oracle_golden.py compares this file with those vectors bit for bit (float64 included).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* nr/ldpc/basegraphs.py:19-30 -- the 3x6 demo base graph (both bg=1 and bg=2 map to it, :33-36) */
static const int kDemoShifts[3][6] = {
    {0, 1, 2, 0, -1, -1},
    {1, 0, 3, -1, 0, -1},
    {2, 3, 0, -1, -1, 0},
};

/* nr/ldpc/basegraphs.py:39-42 load_base_graph + builder.py:10-30 build_h_matrix.
 * H is written row-major as int8 [3Z][6Z]; returns 0, or -1 for an unknown base graph / bad Z. */
int po_ldpc_build_h(int bg, int Z, int8_t *H, int *m_out, int *n_out) {
    if (bg != 1 && bg != 2) return -1;
    if (Z <= 0) return -1;
    const int m = 3 * Z, n = 6 * Z;
    memset(H, 0, (size_t)m * n);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 6; ++c) {
            int shift = kDemoShifts[r][c];
            if (shift < 0) continue;              /* builder.py:12-13 zero block */
            shift %= Z;                           /* builder.py:14 */
            for (int i = 0; i < Z; ++i)           /* builder.py:15-16 */
                H[(size_t)(r * Z + i) * n + c * Z + (i + shift) % Z] = 1;
        }
    if (m_out) *m_out = m;
    if (n_out) *n_out = n;
    return 0;
}

/* nr/ldpc/encode.py:8-49 _gauss_solve_gf2: A [m][n], b [m] -> x [n]; returns 0 or -2 (no solution). */
static int gauss_solve_gf2(const uint8_t *A_in, const uint8_t *b_in, int m, int n, uint8_t *x) {
    uint8_t *A = (uint8_t *)malloc((size_t)m * n);
    uint8_t *b = (uint8_t *)malloc((size_t)m);
    int *pivot_row = (int *)malloc(sizeof(int) * (size_t)n);
    for (size_t i = 0; i < (size_t)m * n; ++i) A[i] = A_in[i] & 1;
    for (int i = 0; i < m; ++i) b[i] = b_in[i] & 1;
    for (int c = 0; c < n; ++c) pivot_row[c] = -1;
    memset(x, 0, (size_t)n);
    int row = 0;
    for (int col = 0; col < n && row < m; ++col) {            /* :16-33 (break when row == m) */
        int pivot = -1;
        for (int r = row; r < m; ++r)
            if (A[(size_t)r * n + col]) { pivot = r; break; }
        if (pivot < 0) continue;
        if (pivot != row) {                                   /* :24-26 */
            for (int c = 0; c < n; ++c) {
                uint8_t t = A[(size_t)row * n + c];
                A[(size_t)row * n + c] = A[(size_t)pivot * n + c];
                A[(size_t)pivot * n + c] = t;
            }
            uint8_t t = b[row]; b[row] = b[pivot]; b[pivot] = t;
        }
        pivot_row[col] = row;
        for (int r = 0; r < m; ++r)                           /* :28-31 */
            if (r != row && A[(size_t)r * n + col]) {
                for (int c = 0; c < n; ++c) A[(size_t)r * n + c] ^= A[(size_t)row * n + c];
                b[r] ^= b[row];
            }
        ++row;
    }
    int rc = 0;
    for (int r = row; r < m && rc == 0; ++r) {                /* :35-37 */
        int any = 0;
        for (int c = 0; c < n; ++c) any |= A[(size_t)r * n + c];
        if (!any && b[r]) rc = -2;
    }
    if (rc == 0) {
        for (int col = n - 1; col >= 0; --col) {              /* :39-45 */
            const int r = pivot_row[col];
            if (r < 0) continue;
            uint8_t val = b[r];
            for (int c = col + 1; c < n; ++c)
                if (A[(size_t)r * n + c]) val ^= x[c];
            x[col] = val;
        }
    }
    free(A); free(b); free(pivot_row);
    return rc;
}

/* nr/ldpc/encode.py:52-66 encode_ldpc: payload [k] -> codeword [n].  -1: n <= k, -2: no solution. */
int po_ldpc_encode(const int8_t *payload, int k, const int8_t *H, int m, int n, int8_t *codeword) {
    if (n <= k) return -1;                                    /* :57-58 */
    const int np_ = n - k;
    uint8_t *Hpar = (uint8_t *)malloc((size_t)m * np_);
    uint8_t *syn = (uint8_t *)malloc((size_t)m);
    uint8_t *par = (uint8_t *)malloc((size_t)np_);
    for (int r = 0; r < m; ++r) {
        unsigned s = 0;
        for (int c = 0; c < k; ++c) s += (unsigned)((H[(size_t)r * n + c] & 1) & (payload[c] & 1));   /* :63 */
        syn[r] = (uint8_t)(s & 1u);
        for (int c = 0; c < np_; ++c) Hpar[(size_t)r * np_ + c] = (uint8_t)(H[(size_t)r * n + k + c] & 1);
    }
    const int rc = gauss_solve_gf2(Hpar, syn, m, np_, par);   /* :64 */
    if (rc == 0) {
        for (int c = 0; c < k; ++c) codeword[c] = (int8_t)(payload[c] & 1);
        for (int c = 0; c < np_; ++c) codeword[k + c] = (int8_t)par[c];
    }
    free(Hpar); free(syn); free(par);
    return rc;
}

/* nr/ldpc/rate_match.py:18-38 derate_match_ldpc: llr [E] -> out [N] */
void po_ldpc_derate(const double *llr, int E, int N, double *out) {
    if (E <= N) {                                             /* :21-24 zero padding */
        for (int i = 0; i < N; ++i) out[i] = i < E ? llr[i] : 0.0;
        return;
    }
    const int reps = E / N, rem = E % N;
    for (int i = 0; i < N; ++i) {
        double s = llr[i];                                    /* :30-31 shaped.sum(axis=0): rows added in order */
        for (int r = 1; r < reps; ++r) s = s + llr[(size_t)r * N + i];
        double acc = 0.0 + s;                                 /* :32 accum += ... */
        int count = reps;
        if (i < rem) { acc = acc + llr[(size_t)reps * N + i]; ++count; }   /* :34-36 */
        out[i] = acc / (double)count;                         /* :38 */
    }
}

/* numpy sign(): -1, 0, +1 */
static double np_sign(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }

/* nr/ldpc/decode_nms.py:8-40 decode_ldpc_nms (layered normalised min-sum, float64). */
int po_ldpc_decode_nms(const double *llr_in, const int8_t *H, int m, int n, int max_iter, double alpha, int early_stop,
                       int8_t *hard, int *iters_used, int *parity_ok) {
    double *llr = (double *)malloc(sizeof(double) * (size_t)n);
    double *msg = (double *)calloc((size_t)m * n, sizeof(double));
    int *idx = (int *)malloc(sizeof(int) * (size_t)n);
    memcpy(llr, llr_in, sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) hard[i] = llr[i] < 0.0;       /* :21 */
    int done = 0, used = max_iter;
    for (int it = 1; it <= max_iter && !done; ++it) {
        for (int r = 0; r < m; ++r) {                         /* :25-34 */
            int w = 0;
            for (int c = 0; c < n; ++c)
                if (H[(size_t)r * n + c] == 1) idx[w++] = c;
            if (w == 0) continue;
            double sign = 1.0, mag = INFINITY;
            for (int j = 0; j < w; ++j) {
                const double ext = llr[idx[j]] - msg[(size_t)r * n + idx[j]];   /* :29 */
                sign = sign * np_sign(ext);                   /* :30 */
                const double a = fabs(ext);
                if (a < mag) mag = a;                         /* :31 */
            }
            const double update = alpha * sign * mag;         /* :32 (left to right) */
            for (int j = 0; j < w; ++j) {
                const double ext = llr[idx[j]] - msg[(size_t)r * n + idx[j]];
                msg[(size_t)r * n + idx[j]] = update;         /* :33 */
                llr[idx[j]] = ext + update;                   /* :34 */
            }
        }
        for (int i = 0; i < n; ++i) hard[i] = llr[i] < 0.0;   /* :36 */
        int any = 0;
        for (int r = 0; r < m; ++r) {                         /* :37 */
            unsigned s = 0;
            for (int c = 0; c < n; ++c) s += (unsigned)(H[(size_t)r * n + c] * hard[c]);
            any |= (int)(s & 1u);
        }
        if (early_stop && !any) { used = it; done = 1; }      /* :38-39 */
    }
    int any = 0;
    for (int r = 0; r < m; ++r) {                             /* :41 */
        unsigned s = 0;
        for (int c = 0; c < n; ++c) s += (unsigned)(H[(size_t)r * n + c] * hard[c]);
        any |= (int)(s & 1u);
    }
    *iters_used = used;
    *parity_ok = !any;
    free(llr); free(msg); free(idx);
    return 0;
}

/* batched form of the decoder: llr [B][n] -> hard [B][n], iters [B], ok [B] */
int po_ldpc_decode_batch(const double *llr, int B, const int8_t *H, int m, int n, int max_iter, double alpha,
                         int early_stop, int8_t *hard, int32_t *iters, int32_t *ok) {
    for (int f = 0; f < B; ++f) {
        int it = 0, pk = 0;
        po_ldpc_decode_nms(llr + (size_t)f * n, H, m, n, max_iter, alpha, early_stop, hard + (size_t)f * n, &it, &pk);
        iters[f] = it;
        ok[f] = pk;
    }
    return 0;
}
