/*
 * polar_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C, float64 restatement of the heimrih/polar_code hot path
 * (package dl_scl_polar).  Every function cites the reference file:line it
 * follows (paths relative to /root/reference/).  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load
 * this library, and only as the checker.  The shipped path is the CUDA
 * library in polar_code_b200/csrc and never links or calls this file.
 *
 * Parity status: PINNED.  oracle/gen_golden.py imports the reference itself
 * (in the build container) and tests/test_oracle_golden.py checks this file
 * against those vectors and against the published results/fer_M{1,4,8}.csv.
 *
 * Arithmetic is IEEE double with the same operation order as the NumPy code:
 *   f  = sign(a)*sign(b)*min(|a|,|b|)          (polar/polar.py:122-123)
 *   g  = b + (1-2c)*a                          (polar/polar.py:126-127)
 *   PM += logaddexp(0, -+L)                    (polar/scl.py:102-105)
 * logaddexp follows numpy's npy_logaddexp (x==y -> x+ln2; else the larger
 * argument + log1p(exp(-|d|))), evaluated with libm exp/log1p as numpy does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define PO_MAXN 1024
#define PO_MAXLOG 10
#define PO_OK 0
#define PO_EINVAL (-1)

/* ------------------------------------------------------------------ */
/* helpers                                                            */
/* ------------------------------------------------------------------ */

static int ilog2_exact(int N) {
    int n = 0;
    if (N <= 0 || (N & (N - 1)) != 0) return -1; /* polar/polar.py:32-34 */
    while ((1 << n) < N) n++;
    return n;
}

/* numpy npy_logaddexp(x, y) in double; used as logaddexp(0.0, v). */
static double np_logaddexp(double x, double y) {
    const double LOGE2 = 0.693147180559945309417232121458176568;
    if (x == y) return x + LOGE2;
    {
        const double tmp = x - y;
        if (tmp > 0) return x + log1p(exp(-tmp));
        else if (tmp <= 0) return y + log1p(exp(tmp));
        return tmp; /* NaN */
    }
}

/* polar/scl.py:102-105 */
static double update_metric(double metric, double llr, int bit) {
    if (bit) return metric + np_logaddexp(0.0, llr);
    return metric + np_logaddexp(0.0, -llr);
}

/* np.sign for doubles: -1, 0, +1 (NaN ignored). */
static double np_sign(double v) { return (v > 0) - (v < 0); }

/* polar/polar.py:122-123 */
static double f_fn(double a, double b) {
    const double fa = fabs(a), fb = fabs(b);
    return np_sign(a) * np_sign(b) * (fa < fb ? fa : fb);
}

/* polar/polar.py:126-127 */
static double g_fn(double a, double b, int c) { return b + (double)(1 - 2 * c) * a; }

/* ------------------------------------------------------------------ */
/* polar/polar.py : construction, transform                           */
/* ------------------------------------------------------------------ */

/* polar/polar.py:17-29 -- x = u F^{(x)n}, natural order, in place. */
void po_polar_transform(int8_t *x, int N) {
    int n = ilog2_exact(N);
    for (int stage = 0; stage < n; stage++) {
        int step = 1 << stage, block = step << 1;
        for (int start = 0; start < N; start += block)
            for (int i = 0; i < step; i++) x[start + i] ^= x[start + step + i];
    }
}

/* polar/polar.py:51-58 */
static double phi_inv(double x) {
    if (x > 12.0) return 0.9861 * x - 2.3152;
    if (x > 3.5) return x * (0.009005 * x + 0.7694) - 0.9507;
    if (x > 1.0) return x * (0.062883 * x + 0.3678) - 0.1627;
    return x * (0.2202 * x + 0.06448);
}

/* stable argsort of doubles (np.argsort(kind="stable"), polar.py:95,98) */
static void stable_argsort(const double *v, int n, int *order) {
    for (int i = 0; i < n; i++) order[i] = i;
    for (int i = 1; i < n; i++) { /* insertion sort is stable */
        int oi = order[i];
        int j = i - 1;
        while (j >= 0 && v[order[j]] > v[oi]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = oi;
    }
}

/* polar/polar.py:85-103.  method 0 = "gaussian" (:61-82), 1 = "polarization" (:37-48). */
int po_construct_info_set(int N, int K, int method, double design_snr_db, int32_t *out) {
    int n = ilog2_exact(N);
    if (n < 0 || N > PO_MAXN) return PO_EINVAL;
    if (!(0 < K && K <= N)) return PO_EINVAL;
    double *metric = (double *)malloc(sizeof(double) * N);
    int *order = (int *)malloc(sizeof(int) * N);
    if (method == 1) {
        for (int idx = 0; idx < N; idx++) { /* :37-48 */
            double w = 0.0;
            int bits = idx;
            for (int j = 0; j < n; j++) {
                if (bits & 1) w += pow(2.0, j / 4.0);
                bits >>= 1;
            }
            metric[idx] = w;
        }
    } else if (method == 0) {
        double rate = (double)K / (double)N; /* :62 */
        double snr = pow(10.0, design_snr_db / 10.0);
        double sigma_sq = 1.0 / (2.0 * rate * snr);
        double *m = metric;
        for (int i = 0; i < N; i++) m[i] = 0.0;
        m[0] = 2.0 / sigma_sq;
        for (int level = 1; level <= n; level++) { /* :69-75 */
            int B = 1 << level, half = B >> 1;
            for (int j = 0; j < half; j++) {
                double T = m[j];
                m[j] = phi_inv(T);
                m[half + j] = 2.0 * T;
            }
        }
        for (int i = 0; i < N; i++) { /* :79-81 */
            double val = m[i] > 1e-12 ? m[i] : 1e-12;
            m[i] = 0.5 - 0.5 * erf(sqrt(val) / 2.0);
        }
    } else {
        free(metric); free(order);
        return PO_EINVAL;
    }
    stable_argsort(metric, N, order);
    /* np.sort(order[:K]) :102 */
    uint8_t *mark = (uint8_t *)calloc(N, 1);
    for (int i = 0; i < K; i++) mark[order[i]] = 1;
    int k = 0;
    for (int i = 0; i < N; i++) if (mark[i]) out[k++] = i;
    free(mark); free(metric); free(order);
    return PO_OK;
}

/* polar/polar.py:106-119 generalised to an explicit (N, info_set):
 * also eval/run_ber_sweep.py:65-70 and nr/polar/scl_nr.py:17-20. */
int po_encode(const int8_t *msg, int K, const int32_t *info_set, int N, int8_t *x) {
    if (ilog2_exact(N) < 0) return PO_EINVAL;
    memset(x, 0, N);
    for (int j = 0; j < K; j++) x[info_set[j]] = msg[j] & 1;
    po_polar_transform(x, N);
    return PO_OK;
}

/* ------------------------------------------------------------------ */
/* polar/crc.py                                                       */
/* ------------------------------------------------------------------ */

/* polar/crc.py:10-16 -- hex string (may carry 0x) -> MSB-first bit list. */
static int poly_to_bits(const char *poly, int8_t *bits /* >=65 */) {
    if (!poly || !*poly) return PO_EINVAL;
    char *end = NULL;
    unsigned long long value = strtoull(poly, &end, 16);
    if (end == poly) return PO_EINVAL;
    int len = 0;
    while (len < 64 && (value >> len)) len++;
    for (int i = 0; i < len; i++) bits[i] = (int8_t)((value >> (len - 1 - i)) & 1);
    return len;
}

/* polar/crc.py:19-37 ; out has len+degree entries; returns degree or <0. */
int po_crc_attach(const int8_t *msg, int len, const char *poly, int8_t *out) {
    int8_t pb[65];
    int plen = poly_to_bits(poly, pb);
    if (plen < 0) return PO_EINVAL;
    int degree = plen - 1;
    if (degree <= 0) return PO_EINVAL; /* :27-28 */
    int8_t *buf = (int8_t *)calloc((size_t)len + degree, 1);
    for (int i = 0; i < len; i++) buf[i] = msg[i] & 1;
    for (int i = 0; i < len; i++) { /* :32-35 */
        if (buf[i] == 0) continue;
        for (int j = 0; j <= degree; j++) buf[i + j] ^= pb[j];
    }
    for (int i = 0; i < len; i++) out[i] = msg[i] & 1;
    for (int i = 0; i < degree; i++) out[len + i] = buf[len + i];
    free(buf);
    return degree;
}

/* polar/crc.py:40-56 ; returns 1 pass, 0 fail, <0 error (too short / bad poly). */
int po_crc_check(const int8_t *msg_crc, int len, const char *poly) {
    int8_t pb[65];
    int plen = poly_to_bits(poly, pb);
    if (plen < 0) return PO_EINVAL;
    int degree = plen - 1;
    if (len <= degree) return PO_EINVAL; /* :48-49 */
    int8_t *buf = (int8_t *)malloc((size_t)len);
    for (int i = 0; i < len; i++) buf[i] = msg_crc[i] & 1;
    for (int i = 0; i < len - degree; i++) { /* :52-55 */
        if (buf[i] == 0) continue;
        for (int j = 0; j <= degree; j++) buf[i + j] ^= pb[j];
    }
    int any = 0;
    for (int i = len - degree; i < len; i++) any |= buf[i];
    free(buf);
    return any ? 0 : 1;
}

/* ------------------------------------------------------------------ */
/* polar/polar.py:130-168 : recursive SC                              */
/* ------------------------------------------------------------------ */

typedef struct { const uint8_t *frozen; int8_t *u_hat; } sc_ctx;

/* (gcc cannot see that half >= 1 fills tmp[0..half) before the recursive call reads it) */
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wmaybe-uninitialized"
static void sc_segment(sc_ctx *c, const double *seg, int size, int start, int8_t *bits_out) {
    if (size == 1) { /* :147-154 */
        int8_t bit = c->frozen[start] ? 0 : (int8_t)(seg[0] < 0);
        c->u_hat[start] = bit;
        bits_out[0] = bit;
        return;
    }
    int half = size / 2;
    double tmp[PO_MAXN / 2];
    int8_t lb[PO_MAXN / 2], rb[PO_MAXN / 2];
    for (int i = 0; i < half; i++) tmp[i] = f_fn(seg[i], seg[half + i]); /* :157 */
    sc_segment(c, tmp, half, start, lb);
    for (int i = 0; i < half; i++) tmp[i] = g_fn(seg[i], seg[half + i], lb[i]); /* :160 */
    sc_segment(c, tmp, half, start + half, rb);
    for (int i = 0; i < half; i++) { bits_out[i] = lb[i] ^ rb[i]; bits_out[half + i] = rb[i]; } /* :163 */
}
#pragma GCC diagnostic pop

int po_sc_decode(const double *llr, int N, const int32_t *info_set, int K, int8_t *out_bits) {
    if (ilog2_exact(N) < 0 || N > PO_MAXN) return PO_EINVAL;
    uint8_t frozen[PO_MAXN];
    int8_t u_hat[PO_MAXN], bits[PO_MAXN];
    memset(frozen, 1, N);
    for (int j = 0; j < K; j++) {
        if (info_set[j] < 0 || info_set[j] >= N) return PO_EINVAL; /* :139-140 */
        frozen[info_set[j]] = 0;
    }
    memset(u_hat, 0, N);
    sc_ctx c = { frozen, u_hat };
    sc_segment(&c, llr, N, 0, bits);
    for (int j = 0; j < K; j++) out_bits[j] = u_hat[info_set[j]];
    return PO_OK;
}

/* ------------------------------------------------------------------ */
/* polar/scl.py : _PathState + decode_scl                             */
/* ------------------------------------------------------------------ */

/* polar/scl.py:16-50.  llr[level] holds the block of the node currently
 * being visited at that level (the reference keeps every node and
 * recomputes root->leaf each phase, scl.py:64-78; the values are the same
 * because every recomputation has identical inputs).  bits[level][.] keeps
 * the whole level like bit_layers. */
typedef struct {
    double metric;
    int n_info;
    int N, n;
    double *llr;       /* level l block at llr + llr_off(l): sizes N, N/2, ..., 1 (2N-1 doubles) */
    double *info_llrs; /* [N] */
    int8_t *bits;      /* [(n+1)][N] like bit_layers */
    int8_t *u;         /* [N] */
} path_t;

static size_t path_bytes(int N, int n) {
    return sizeof(path_t) + sizeof(double) * (size_t)(2 * N) + sizeof(double) * (size_t)N +
           (size_t)(n + 1) * N + (size_t)N;
}
static void path_fix(path_t *p) {
    char *base = (char *)(p + 1);
    p->llr = (double *)base;
    p->info_llrs = p->llr + 2 * p->N;
    p->bits = (int8_t *)(p->info_llrs + p->N);
    p->u = p->bits + (size_t)(p->n + 1) * p->N;
}
static path_t *path_new(int N, int n) { /* scl.py:25-50 */
    path_t *p = (path_t *)calloc(1, path_bytes(N, n));
    p->N = N; p->n = n;
    path_fix(p);
    return p;
}
static path_t *path_clone(const path_t *src) { /* scl.py:52-62 */
    size_t nb = path_bytes(src->N, src->n);
    path_t *p = (path_t *)malloc(nb);
    memcpy(p, src, nb);
    path_fix(p);
    return p;
}
/* offset of level l's block inside llr: N + N/2 + ... */
static int llr_off(int N, int level) { return 2 * N - (2 * N >> level); }

/* scl.py:64-82 -- leaf LLR for `phase`; only levels below the lowest
 * changed ancestor need recomputing (lazy form of _ensure_alpha). */
static double path_llr_for_phase(path_t *p, int n, int phase) {
    int first = 1;
    if (phase != 0) {
        int tz = 0;
        while (!((phase >> tz) & 1)) tz++;
        first = n - tz;
    }
    for (int level = first; level <= n; level++) {
        int node = phase >> (n - level);
        int half = 1 << (n - level);
        const double *parent = p->llr + llr_off(p->N, level - 1);
        double *target = p->llr + llr_off(p->N, level);
        if ((node & 1) == 0) { /* :74-75 */
            for (int i = 0; i < half; i++) target[i] = f_fn(parent[i], parent[half + i]);
        } else { /* :76-78 */
            const int8_t *lb = &p->bits[(size_t)level * p->N + (node - 1) * half];
            for (int i = 0; i < half; i++) target[i] = g_fn(parent[i], parent[half + i], lb[i]);
        }
    }
    return p->llr[llr_off(p->N, n)];
}

/* scl.py:84-99 */
static void path_set_bit(path_t *p, int n, int phase, int bit) {
    bit &= 1;
    p->u[phase] = (int8_t)bit;
    p->bits[(size_t)n * p->N + phase] = (int8_t)bit;
    int level = n, node = phase;
    while (level > 0 && (node & 1)) {
        int size = 1 << (n - level);
        int parent = node >> 1;
        const int8_t *left = &p->bits[(size_t)level * p->N + (node - 1) * size];
        const int8_t *right = &p->bits[(size_t)level * p->N + node * size];
        int8_t *pv = &p->bits[(size_t)(level - 1) * p->N + parent * 2 * size];
        for (int i = 0; i < size; i++) { pv[i] = left[i] ^ right[i]; pv[size + i] = right[i]; }
        node = parent;
        level--;
    }
}

typedef struct {
    int n_cand;
    int best_idx;
    double min_rel_gap; /* smallest relative gap between competing metrics at any sort */
    double min_rel_gap_prune; /* same, only at information-bit sorts and the final order */
} scl_info;

/*
 * polar/scl.py:108-209.
 *   force: NULL or int8[K] with -1 free / 0,1 forced (:126-130,138-144)
 *   crc:   NULL or hex polynomial (:190-197)
 * outputs (caller-allocated): cand[M*K], metrics[M], info_llrs[M*K] (may be NULL).
 * returns PO_OK, PO_EINVAL, or -2 for a force entry outside {-1,0,1} (ValueError :144).
 */
int po_scl_decode(const double *llr, int N, const int32_t *info_set, int K, int M, const char *crc,
                  const int8_t *force, int8_t *cand, double *metrics, double *info_llrs,
                  scl_info *info) {
    int n = ilog2_exact(N);
    if (M <= 0 || n < 0 || N > PO_MAXN) return PO_EINVAL; /* :118-119 */
    uint8_t is_info[PO_MAXN];
    memset(is_info, 0, N);
    for (int j = 0; j < K; j++) is_info[info_set[j]] = 1;

    int cap = 2 * M;
    path_t **paths = (path_t **)malloc(sizeof(path_t *) * cap);
    path_t **next = (path_t **)malloc(sizeof(path_t *) * cap);
    int np = 1, rc = PO_OK;
    paths[0] = path_new(N, n); /* :135 */
    for (int i = 0; i < N; i++) paths[0]->llr[i] = llr[i];
    double min_gap = INFINITY, min_gap_prune = INFINITY;
    int info_index = 0;

    for (int phase = 0; phase < N && rc == PO_OK; phase++) { /* :136 */
        int frozen = !is_info[phase];
        int forced = -1;
        if (!frozen && force) { /* :138-144 */
            int val = force[info_index];
            if (val == 0 || val == 1) forced = val;
            else if (val != -1) { rc = -2; break; }
        }
        int nn = 0;
        for (int pi = 0; pi < np; pi++) { /* :147 */
            path_t *p = paths[pi];
            double L = path_llr_for_phase(p, n, phase);
            if (frozen) { /* :149-153 */
                p->metric = update_metric(p->metric, L, 0);
                path_set_bit(p, n, phase, 0);
                next[nn++] = p;
            } else if (forced >= 0) { /* :156-161 */
                p->metric = update_metric(p->metric, L, forced);
                path_set_bit(p, n, phase, forced);
                p->info_llrs[p->n_info++] = L;
                next[nn++] = p;
            } else { /* :163-168 : clone for bit 0 then bit 1 */
                for (int bit = 0; bit < 2; bit++) {
                    path_t *b = path_clone(p);
                    b->metric = update_metric(b->metric, L, bit);
                    path_set_bit(b, n, phase, bit);
                    b->info_llrs[b->n_info++] = L;
                    next[nn++] = b;
                }
                free(p);
            }
        }
        if (!frozen) info_index++; /* :169-170 */
        /* stable sort by metric (list.sort, :173) */
        for (int i = 1; i < nn; i++) {
            path_t *pi_ = next[i];
            int j = i - 1;
            while (j >= 0 && pi_->metric < next[j]->metric) { next[j + 1] = next[j]; j--; }
            next[j + 1] = pi_;
        }
        /* competing-metric gap: adjacent pairs whose order decides a kept slot */
        for (int i = 0; i + 1 < nn && i < M; i++) {
            double a = next[i]->metric, b = next[i + 1]->metric;
            double den = fabs(b) > fabs(a) ? fabs(b) : fabs(a);
            double gap = den > 0 ? (b - a) / den : 0.0;
            if (gap < min_gap) min_gap = gap;
            if ((!frozen || phase == N - 1) && gap < min_gap_prune) min_gap_prune = gap;
        }
        np = nn < M ? nn : M; /* :174 */
        for (int i = 0; i < nn; i++) {
            if (i < np) paths[i] = next[i];
            else free(next[i]);
        }
    }

    if (rc == PO_OK) {
        int best = -1;
        for (int i = 0; i < np; i++) { /* :183-188 */
            for (int j = 0; j < K; j++) cand[i * K + j] = paths[i]->u[info_set[j]];
            metrics[i] = paths[i]->metric;
            if (info_llrs)
                for (int j = 0; j < K; j++) info_llrs[i * K + j] = paths[i]->info_llrs[j];
        }
        if (crc) { /* :190-194 */
            for (int i = 0; i < np; i++) {
                int ok = po_crc_check(&cand[i * K], K, crc);
                if (ok < 0) { rc = PO_EINVAL; break; }
                if (ok) { best = i; break; }
            }
        }
        if (best < 0 && np > 0) best = 0; /* :196-197 */
        info->n_cand = np;
        info->best_idx = best;
        info->min_rel_gap = min_gap;
        info->min_rel_gap_prune = min_gap_prune;
    }
    for (int i = 0; i < np; i++) free(paths[i]);
    free(paths); free(next);
    return rc;
}

/* ------------------------------------------------------------------ */
/* dlscl/flip.py                                                      */
/* ------------------------------------------------------------------ */

/* dlscl/flip.py:13-27 : argmin(abs_l0 @ beta) or argmin(abs_l0). beta is float32 [K,K]. */
int po_choose_flip_index(const double *abs_l0, int K, const float *beta) {
    int best = 0;
    double bestv = 0;
    for (int j = 0; j < K; j++) {
        double q;
        if (beta) {
            q = 0.0;
            for (int i = 0; i < K; i++) q += abs_l0[i] * (double)beta[i * K + j];
        } else q = abs_l0[j];
        if (j == 0 || q < bestv) { bestv = q; best = j; }
    }
    return best;
}

typedef struct {
    int success;       /* flip.py:140 */
    int n_attempts;    /* len(attempts) incl. baseline */
    int n_tried;       /* len(tried_indices) */
    int n_cand;        /* candidates in the returned (last) attempt */
    int best_idx;
    double min_rel_gap;      /* over all SCL sorts of all attempts */
    double min_rank_gap;     /* smallest relative gap between the chosen q and the runner-up */
} dl_info;

/*
 * dlscl/flip.py:65-141 decode_with_retries.
 * Returns the LAST attempt's candidates/metrics/info_llrs (flip.py:137),
 * tried[retries] (flip.py:110-116).
 */
int po_decode_with_retries(const double *llr, int N, const int32_t *info_set, int K, int M,
                           int retries, const char *crc, const float *beta, int8_t *cand,
                           double *metrics, double *info_llrs, int32_t *tried, dl_info *out) {
    scl_info si;
    double *ill = (double *)malloc(sizeof(double) * (size_t)M * K);
    int8_t *ref_bits = (int8_t *)malloc(K);
    int8_t *forced = (int8_t *)malloc(K);
    double *abs_l0 = (double *)malloc(sizeof(double) * K);
    double *q = (double *)malloc(sizeof(double) * K);
    int rc = po_scl_decode(llr, N, info_set, K, M, crc, NULL, cand, metrics, ill, &si); /* :77 */
    out->n_attempts = 1;
    out->n_tried = 0;
    out->min_rel_gap = si.min_rel_gap;
    out->min_rank_gap = INFINITY;
    int pass = 0;
    if (rc == PO_OK) {
        /* _passes :82-88 */
        pass = crc ? (po_crc_check(&cand[si.best_idx * K], K, crc) == 1) : 1;
        if (!(pass || retries <= 0)) { /* :90 */
            memcpy(ref_bits, &cand[si.best_idx * K], K); /* :97-98 */
            for (int j = 0; j < K; j++) abs_l0[j] = fabs(ill[si.best_idx * K + j]); /* :102 */
            while (out->n_tried < retries && out->n_tried < K) { /* :111 */
                /* rank_indices :104-108 ; first untried index of argsort(q) == argmin over untried */
                for (int j = 0; j < K; j++) {
                    if (beta) {
                        double s = 0.0;
                        for (int i = 0; i < K; i++) s += abs_l0[i] * (double)beta[i * K + j];
                        q[j] = s;
                    } else q[j] = abs_l0[j];
                }
                int idx = -1;
                double second = INFINITY;
                for (int j = 0; j < K; j++) {
                    int seen = 0;
                    for (int t = 0; t < out->n_tried; t++) if (tried[t] == j) seen = 1;
                    if (seen) continue;
                    if (idx < 0 || q[j] < q[idx]) { if (idx >= 0) second = q[idx]; idx = j; }
                    else if (q[j] < second) second = q[j];
                }
                if (idx < 0) break; /* :114-115 */
                if (second < INFINITY) {
                    double den = fabs(second) > fabs(q[idx]) ? fabs(second) : fabs(q[idx]);
                    double gap = den > 0 ? (second - q[idx]) / den : 0.0;
                    if (gap < out->min_rank_gap) out->min_rank_gap = gap;
                }
                tried[out->n_tried++] = idx; /* :116 */
                /* _force_vector :30-34 */
                for (int j = 0; j < K; j++) forced[j] = -1;
                for (int j = 0; j < idx; j++) forced[j] = ref_bits[j];
                forced[idx] = (int8_t)(1 - ref_bits[idx]);
                rc = po_scl_decode(llr, N, info_set, K, M, crc, forced, cand, metrics, ill, &si); /* :53 */
                if (rc != PO_OK) break;
                out->n_attempts++;
                if (si.min_rel_gap < out->min_rel_gap) out->min_rel_gap = si.min_rel_gap;
                memcpy(ref_bits, &cand[si.best_idx * K], K); /* :127-132 */
                for (int j = 0; j < K; j++) abs_l0[j] = fabs(ill[si.best_idx * K + j]); /* :133 */
                pass = crc ? (po_crc_check(&cand[si.best_idx * K], K, crc) == 1) : 1;
                if (pass) break; /* :134-135 */
            }
        }
        out->success = pass;
        out->n_cand = si.n_cand;
        out->best_idx = si.best_idx;
        if (info_llrs) memcpy(info_llrs, ill, sizeof(double) * (size_t)M * K);
    }
    free(ill); free(ref_bits); free(forced); free(abs_l0); free(q);
    return rc;
}

/* ------------------------------------------------------------------ */
/* nr/polar                                                           */
/* ------------------------------------------------------------------ */

/* nr/polar/interleaver.py:10-23 : order[i] = (i%32)*nb + i/32 over nb*32 slots.
 * Writes the gather order (size nb*32) and returns that size. */
int po_interleave_order(int len, int32_t *order) {
    const int block = 32;
    int nb = (len + block - 1) / block;
    int total = nb * block;
    for (int i = 0; i < total; i++) order[i] = (i % block) * nb + (i / block);
    return total;
}

/* nr/polar/interleaver.py:10-23 on doubles; pad value -1 ; out has nb*32 entries. */
int po_subblock_interleave(const double *in, int len, double *out) {
    int32_t order[PO_MAXN + 32];
    int total = po_interleave_order(len, order);
    for (int i = 0; i < total; i++) {
        int src = order[i];
        out[i] = src < len ? in[src] : -1.0;
    }
    return total;
}

/* nr/polar/interleaver.py:26-37 : out[:original_len] of padded[argsort(order)]. */
int po_subblock_deinterleave(const double *in, int in_len, int original_len, double *out) {
    int32_t order[PO_MAXN + 32], inv[PO_MAXN + 32];
    int total = po_interleave_order(original_len, order);
    for (int i = 0; i < total; i++) inv[order[i]] = i; /* argsort of a permutation */
    for (int i = 0; i < original_len; i++) {
        int src = inv[i];
        out[i] = src < in_len ? in[src] : 0.0; /* padded = zeros; padded[:size] = bits */
    }
    return original_len;
}

/* nr/polar/rate_match.py:19-39 */
void po_derate_match(const double *in, int E, int N, double *out) {
    if (E <= N) { /* :22-25 */
        for (int i = 0; i < N; i++) out[i] = i < E ? in[i] : -1.0;
        return;
    }
    int reps = E / N, rem = E % N;
    for (int i = 0; i < N; i++) {
        double acc = 0.0;
        int cnt = 0;
        for (int r = 0; r < reps; r++) acc += in[r * N + i]; /* shaped.sum(axis=0) :31-33 */
        cnt += reps;
        if (i < rem) { acc += in[reps * N + i]; cnt += 1; } /* :34-37 */
        if (cnt == 0) cnt = 1;
        out[i] = acc / (double)cnt; /* :39 */
    }
}

/* nr/polar/scl_nr.py:38-57 : de-rate-match + de-interleave + SCL; returns best bits in `bits`. */
int po_decode_rate_matched_scl(const double *llr_E, int E, const char *crc, int N,
                               const int32_t *info_set, int K, int M, int8_t *bits, int *crc_pass,
                               double *min_rel_gap) {
    double a[PO_MAXN], b[PO_MAXN];
    po_derate_match(llr_E, E, N, a);
    po_subblock_deinterleave(a, N, N, b);
    int8_t *cand = (int8_t *)malloc((size_t)M * K);
    double *metrics = (double *)malloc(sizeof(double) * M);
    scl_info si;
    int rc = po_scl_decode(b, N, info_set, K, M, crc, NULL, cand, metrics, NULL, &si);
    if (rc == PO_OK) {
        memcpy(bits, &cand[si.best_idx * K], K);
        *crc_pass = po_crc_check(bits, K, crc) == 1;
        if (min_rel_gap) *min_rel_gap = si.min_rel_gap;
    }
    free(cand); free(metrics);
    return rc;
}

/* ------------------------------------------------------------------ */
/* batched drivers (pthreads over frames) -- used by tests and by      */
/* bench.py's cpu_baseline leg to time the CPU path on all host cores. */
/* ------------------------------------------------------------------ */

int po_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef void (*frame_fn)(void *ctx, int b);
typedef struct { frame_fn fn; void *ctx; int B; int next; int chunk; pthread_mutex_t mu; } pf_t;

static void *pf_worker(void *arg) {
    pf_t *pf = (pf_t *)arg;
    for (;;) {
        pthread_mutex_lock(&pf->mu);
        int lo = pf->next;
        pf->next += pf->chunk;
        pthread_mutex_unlock(&pf->mu);
        if (lo >= pf->B) break;
        int hi = lo + pf->chunk < pf->B ? lo + pf->chunk : pf->B;
        for (int b = lo; b < hi; b++) pf->fn(pf->ctx, b);
    }
    return NULL;
}

static void parallel_frames(frame_fn fn, void *ctx, int B, int nthreads) {
    if (nthreads <= 0) nthreads = po_num_threads();
    if (nthreads > 256) nthreads = 256;
    if (nthreads <= 1 || B < 2) { for (int b = 0; b < B; b++) fn(ctx, b); return; }
    pf_t pf = { fn, ctx, B, 0, 8, PTHREAD_MUTEX_INITIALIZER };
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads; t++)
        if (pthread_create(&th[started], NULL, pf_worker, &pf) == 0) started++;
    if (started == 0) pf_worker(&pf);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

typedef struct {
    const double *llr; int N; const int32_t *info_set; int K; int M; int retries; int E;
    const char *crc; const int8_t *force; const float *beta;
    int8_t *cand; double *metrics; double *info_llrs; int32_t *n_cand; int32_t *best_idx;
    double *min_gap; double *min_rank_gap; int8_t *bits; int32_t *success; int32_t *n_attempts;
    int32_t *tried; int32_t *n_tried; int32_t *crc_pass; int rc;
} bctx;

static void sc_one(void *v, int b) {
    bctx *c = (bctx *)v;
    int r = po_sc_decode(c->llr + (size_t)b * c->N, c->N, c->info_set, c->K, c->bits + (size_t)b * c->K);
    if (r != PO_OK) c->rc = r;
}

int po_sc_decode_batch(const double *llr, int B, int N, const int32_t *info_set, int K,
                       int8_t *out_bits, int nthreads) {
    bctx c; memset(&c, 0, sizeof c);
    c.llr = llr; c.N = N; c.info_set = info_set; c.K = K; c.bits = out_bits;
    parallel_frames(sc_one, &c, B, nthreads);
    return c.rc;
}

static void scl_one(void *v, int b) {
    bctx *c = (bctx *)v;
    scl_info si;
    size_t MK = (size_t)c->M * c->K;
    int r = po_scl_decode(c->llr + (size_t)b * c->N, c->N, c->info_set, c->K, c->M, c->crc,
                          c->force ? c->force + (size_t)b * c->K : NULL, c->cand + b * MK,
                          c->metrics + (size_t)b * c->M,
                          c->info_llrs ? c->info_llrs + b * MK : NULL, &si);
    if (r != PO_OK) { c->rc = r; return; }
    c->n_cand[b] = si.n_cand;
    c->best_idx[b] = si.best_idx;
    c->min_gap[b] = si.min_rel_gap_prune;
}

/* cand[B,M,K] metrics[B,M] info_llrs[B,M,K]|NULL n_cand[B] best_idx[B] min_gap[B]; force[B,K]|NULL */
int po_scl_decode_batch(const double *llr, int B, int N, const int32_t *info_set, int K, int M,
                        const char *crc, const int8_t *force, int8_t *cand, double *metrics,
                        double *info_llrs, int32_t *n_cand, int32_t *best_idx, double *min_gap,
                        int nthreads) {
    bctx c; memset(&c, 0, sizeof c);
    c.llr = llr; c.N = N; c.info_set = info_set; c.K = K; c.M = M; c.crc = crc; c.force = force;
    c.cand = cand; c.metrics = metrics; c.info_llrs = info_llrs; c.n_cand = n_cand;
    c.best_idx = best_idx; c.min_gap = min_gap;
    parallel_frames(scl_one, &c, B, nthreads);
    return c.rc;
}

static void dl_one(void *v, int b) {
    bctx *c = (bctx *)v;
    int R = c->retries > 0 ? c->retries : 1;
    int8_t *cand = (int8_t *)malloc((size_t)c->M * c->K);
    double *metrics = (double *)malloc(sizeof(double) * c->M);
    int32_t *tr = (int32_t *)malloc(sizeof(int32_t) * R);
    dl_info di;
    int r = po_decode_with_retries(c->llr + (size_t)b * c->N, c->N, c->info_set, c->K, c->M,
                                   c->retries, c->crc, c->beta, cand, metrics, NULL, tr, &di);
    if (r != PO_OK) c->rc = r;
    else {
        memcpy(c->bits + (size_t)b * c->K, &cand[di.best_idx * c->K], c->K);
        c->success[b] = di.success;
        c->n_attempts[b] = di.n_attempts;
        c->n_tried[b] = di.n_tried;
        for (int t = 0; t < c->retries; t++) c->tried[(size_t)b * R + t] = t < di.n_tried ? tr[t] : -1;
        c->min_gap[b] = di.min_rel_gap;
        c->min_rank_gap[b] = di.min_rank_gap;
    }
    free(cand); free(metrics); free(tr);
}

/* best_bits[B,K] success[B] n_attempts[B] tried[B,max(retries,1)] n_tried[B] min_gap[B] min_rank_gap[B] */
int po_dlscl_decode_batch(const double *llr, int B, int N, const int32_t *info_set, int K, int M,
                          int retries, const char *crc, const float *beta, int8_t *best_bits,
                          int32_t *success, int32_t *n_attempts, int32_t *tried, int32_t *n_tried,
                          double *min_gap, double *min_rank_gap, int nthreads) {
    bctx c; memset(&c, 0, sizeof c);
    c.llr = llr; c.N = N; c.info_set = info_set; c.K = K; c.M = M; c.retries = retries; c.crc = crc;
    c.beta = beta; c.bits = best_bits; c.success = success; c.n_attempts = n_attempts;
    c.tried = tried; c.n_tried = n_tried; c.min_gap = min_gap; c.min_rank_gap = min_rank_gap;
    parallel_frames(dl_one, &c, B, nthreads);
    return c.rc;
}

static void nr_one(void *v, int b) {
    bctx *c = (bctx *)v;
    int pass = 0;
    double mg = 0;
    int r = po_decode_rate_matched_scl(c->llr + (size_t)b * c->E, c->E, c->crc, c->N, c->info_set,
                                       c->K, c->M, c->bits + (size_t)b * c->K, &pass, &mg);
    if (r != PO_OK) c->rc = r;
    c->crc_pass[b] = pass;
    c->min_gap[b] = mg;
}

/* NR batch: llr_E[B,E] -> bits[B,K], crc_pass[B], min_gap[B] */
int po_nr_decode_batch(const double *llr_E, int B, int E, const char *crc, int N,
                       const int32_t *info_set, int K, int M, int8_t *bits, int32_t *crc_pass,
                       double *min_gap, int nthreads) {
    bctx c; memset(&c, 0, sizeof c);
    c.llr = llr_E; c.N = N; c.E = E; c.info_set = info_set; c.K = K; c.M = M; c.crc = crc;
    c.bits = bits; c.crc_pass = crc_pass; c.min_gap = min_gap;
    parallel_frames(nr_one, &c, B, nthreads);
    return c.rc;
}
