// polar_abi_sweep.inl -- host side of the Monte-Carlo sweep, DL-SCL retry kernel, channel generator, NR encoder.
// Included at the end of polar_abi.cu.

int sweep_build_tables(pb200_engine* e) {
    const int N = e->code.N;
    std::vector<int16_t> id(N);
    for (int i = 0; i < N; ++i) id[i] = (int16_t)i;
    if (cudaMalloc((void**)&e->d_rm_dst, (size_t)N * 2) != cudaSuccess) return fail(PB200_ECUDA, "cudaMalloc rm_dst failed");
    if (cudaMemcpy(e->d_rm_dst, id.data(), (size_t)N * 2, cudaMemcpyHostToDevice) != cudaSuccess) return fail(PB200_ECUDA, "upload rm_dst failed");
    return PB200_OK;
}

static size_t entry_bytes(const pb200_engine* e) { return e->code.n <= 7 ? sizeof(DlEntry<4>) : sizeof(DlEntry<16>); }

static int ensure_queues(pb200_engine* e, long long frames, int retries, bool want_store) {
    const size_t need = (size_t)frames * entry_bytes(e);
    const size_t need_store = want_store ? (size_t)frames * e->code.N * sizeof(float) : 0;
    if (e->llr_store_bytes < need_store) {
        cudaFree(e->d_llr_store);
        e->d_llr_store = nullptr; e->llr_store_bytes = 0;
        CUDA_TRY(cudaMalloc((void**)&e->d_llr_store, need_store));
        e->llr_store_bytes = need_store;
    }
    const size_t need_abs = retries > 0 ? (size_t)frames * e->code.K * sizeof(float) : 0;
    if (e->abs_store_bytes < need_abs) {
        cudaFree(e->d_abs_store);
        e->d_abs_store = nullptr; e->abs_store_bytes = 0;
        CUDA_TRY(cudaMalloc((void**)&e->d_abs_store, need_abs));
        e->abs_store_bytes = need_abs;
    }
    if (e->q_bytes < need) {
        cudaFree(e->d_q[0]);
        e->d_q[0] = nullptr; e->q_bytes = 0;
        CUDA_TRY(cudaMalloc((void**)&e->d_q[0], need));
        e->q_bytes = need;
    }
    if (e->q_counts_n < retries + 2) {
        cudaFree(e->d_q_counts);
        e->d_q_counts = nullptr; e->q_counts_n = 0;
        CUDA_TRY(cudaMalloc((void**)&e->d_q_counts, sizeof(unsigned int) * (retries + 2)));
        e->q_counts_n = retries + 2;
    }
    return PB200_OK;
}

// Encoder table of the sweep for payload length kp: row (q, v) = the u word (XW u32) produced by payload nibble q
// holding value v -- its bits at their information positions (polar.py:116-117) plus, when CRC bits follow the
// payload, their contribution to each CRC bit (crc.py:19-37 is linear with a zero initial register).
static int get_enc_tab(pb200_engine* e, int kp, const uint32_t** out) {
    auto it = e->enc_tabs.find(kp);
    if (it != e->enc_tabs.end()) { *out = it->second; return PB200_OK; }
    const int K = e->code.K, deg = e->code.crc_deg, XWk = e->code.n <= 7 ? 4 : 16;
    const int nq = (kp + 3) / 4;
    std::vector<uint32_t> tab((size_t)nq * 16 * XWk, 0);
    std::vector<std::vector<uint32_t>> bitrow(kp, std::vector<uint32_t>(XWk, 0));   // u contribution of payload bit j
    for (int j = 0; j < kp; ++j) {
        auto put = [&](int msg_index) { const int pos = e->info_pos[msg_index]; bitrow[j][pos >> 5] ^= 1u << (pos & 31); };
        put(j);
        if (kp < K && deg > 0) {
            const unsigned long long rem = xpow_mod(kp - 1 - j + deg, e->poly, deg);   // remainder of x^(kp-1-j) * x^deg
            for (int t = 0; t < deg && kp + t < K; ++t)
                if ((rem >> (deg - 1 - t)) & 1ull) put(kp + t);
        }
    }
    for (int q = 0; q < nq; ++q)
        for (int v = 0; v < 16; ++v)
            for (int b = 0; b < 4; ++b)
                if (((v >> b) & 1) && q * 4 + b < kp)
                    for (int w = 0; w < XWk; ++w) tab[((size_t)q * 16 + v) * XWk + w] ^= bitrow[q * 4 + b][w];
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, tab.size() * 4));
    CUDA_TRY(cudaMemcpy(d, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    e->enc_tabs[kp] = d;
    *out = d;
    return PB200_OK;
}

static int fill_chan(pb200_engine* e, const pb200_sweep_cfg* c, ChanCfg* cc) {
    cc->k0 = (uint32_t)(c->seed & 0xffffffffu);
    cc->k1 = (uint32_t)(c->seed >> 32) + 0x9E3779B9u * c->stream_id;
    cc->sigma = (float)sqrt(c->noise_var);
    cc->scale = (float)(2.0 / c->noise_var);
    const double nvu = c->noise_var_uncoded > 0 ? c->noise_var_uncoded : 1.0;
    cc->sigma_u = (float)sqrt(nvu);
    cc->scale_u = (float)(2.0 / nvu);
    cc->kp = c->k_payload;
    cc->include_uncoded = c->include_uncoded;
    cc->poly = e->poly;
    cc->deg = e->code.crc_deg;
    cc->tx_src = e->d_tx_src;
    cc->rm_dst = e->d_rm_dst;
    cc->has_pads = (e->tb.E != 0 && (e->code.N % 32) != 0) ? 1 : 0;
    return get_enc_tab(e, c->k_payload, &cc->enc_tab);
}

static int check_sweep_cfg(pb200_engine* e, const pb200_sweep_cfg* c) {
    if (!e || !c) return fail(PB200_EINVAL, "engine/cfg is NULL");
    if (c->M <= 0) return fail(PB200_EINVAL, "List size M must be positive");
    if (c->M > PB200_MAX_M) return fail(PB200_ENOSUP, "list size M > %d is not supported by this build", PB200_MAX_M);
    if (!(c->noise_var > 0)) return fail(PB200_EINVAL, "noise_var must be positive");
    if (c->k_payload <= 0 || c->k_payload > e->code.K) return fail(PB200_EINVAL, "k_payload must be in 1..K");
    if (c->k_payload < e->code.K) {
        if (e->code.crc_deg == 0) return fail(PB200_EINVAL, "k_payload < K needs a CRC polynomial");
        if (e->code.K - c->k_payload != e->code.crc_deg) return fail(PB200_EINVAL, "K - k_payload must equal the CRC degree");
    }
    const int E = c->E == e->code.N ? (e->tb.E ? c->E : 0) : c->E;
    if (E != e->tb.E && !(E == 0 && e->tb.E == 0)) return fail(PB200_EINVAL, "cfg.E does not match pb200_set_rate_matching");
    if (c->n_frames < 0 || c->frame_begin < 0) return fail(PB200_EINVAL, "bad frame range");
    if (c->bit_error_span < 0 || c->bit_error_span > e->code.K) return fail(PB200_EINVAL, "bit_error_span must be in 0..K");
    return PB200_OK;
}

static __global__ void widen_beta_kernel(const float* __restrict__ in, double* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

// Shared driver: baseline launch + `retries` round launches over [frame_begin, frame_begin + n_frames).
static int run_sweep(pb200_engine* e, int M, SweepArgs a, cudaStream_t st) {
    const int MP = round_mp(M);
    a.beta64 = nullptr;
    if (a.beta != nullptr && a.retries > 0) {
        // flip.py:104-108 multiplies float64 |L0| by the float32 beta in float64: widen beta once per call (K^2 values)
        // instead of once per multiply-add inside the retry kernel
        const int kk = e->code.K * e->code.K;
        if (!e->d_beta64) CUDA_TRY(cudaMalloc((void**)&e->d_beta64, (size_t)kk * sizeof(double)));
        widen_beta_kernel<<<(kk + 255) / 256, 256, 0, st>>>(a.beta, e->d_beta64, kk);
        CUDA_TRY(cudaGetLastError());
        a.beta64 = e->d_beta64;
    }
    const bool big = e->code.n > 7;
    Code code = e->code;
    code.M = M;
    if (e->dl_binned < 0) { const char* env = getenv("PB200_DL_BINNED"); e->dl_binned = (env && env[0] == '0') ? 0 : 1; }
    const bool binned = e->dl_binned != 0;
    // Frame-per-group retry kernel: the baseline pass records the leaf-LLR trace and writes |L0| of every queued frame
    // (kind 2).  Binned retry kernel: the baseline pass is the plain one -- queued frames get their first |L0| row from an
    // "attempt 0" replay inside the retry kernel, so the frames that pass (most of them) never pay for a trace.
    // Which admission is faster depends on how many frames fail: tracing every frame costs 0.9 ms per Mi frames, replaying
    // the queued ones a decode each -- break-even near 12 % failures (B200, M = 4).  Both give identical results
    // (tests/test_gpu_dlbin.py), so the choice follows the failure fraction of this engine's previous DL-SCL piece, read
    // back asynchronously (unknown: replay, the better one over most of a FER curve).  PB200_DL_REPLAY=0/1 pins it.
    if (binned && e->queued_of > 0 && e->queued_ev && cudaEventQuery(e->queued_ev) == cudaSuccess) {
        e->dl_fail_frac = (double)e->h_queued[0] / (double)e->queued_of;
        e->queued_of = 0;
    }
    bool replay = binned && !(e->dl_fail_frac > 0.12);
    if (const char* env = getenv("PB200_DL_REPLAY")) replay = binned && env[0] != '0';
    const bool trace = a.retries > 0 && !replay;
    auto pick = [&](int kind) { return big ? pb_sweep_kernel_9(MP, kind) : code.n == 7 ? pb_sweep_kernel_7s(MP, kind) : pb_sweep_kernel_7(MP, kind); };
    const void* base = pick(trace ? 2 : 0);
    const void* round = pick(binned ? 3 : 1);
    KernelCfg kb, kr{};
    int rc = choose_cfg(e, base, MP, trace ? 6 : 4, warp_bytes(MP, code.N, 0, false, trace ? code.K : 0) + acc_bytes(MP), &kb);
    if (rc) return rc;
    int rgrid = 0;
    if (a.retries > 0) {
        rc = choose_cfg(e, round, MP, binned ? 7 : 5, warp_bytes(MP, code.N, code.K, true, code.K) + acc_bytes(MP), &kr);
        if (rc) return rc;
        if (binned) {
            // The warps of a CTA run in lock step behind one scheduler warp, so the CTA's serial section (claim, ring slots,
            // queue entries, LLR rows) idles all of them at once: split the resident warps over several CTAs per SM so that
            // one CTA's serial section overlaps another's decode (PB200_DL_WPC overrides the warps per CTA).
            auto it = e->dl_cfg.find(MP);
            if (it == e->dl_cfg.end()) {
                int want_wpc = std::max(4, kr.wpc / 2);
                bool forced = false;
                if (const char* env = getenv("PB200_DL_WPC")) { want_wpc = std::max(1, std::min(32, atoi(env))); forced = true; }
                if (want_wpc < kr.wpc) {
                    const size_t wb = warp_bytes(MP, code.N, code.K, true, code.K) + acc_bytes(MP);
                    int blocks = 0;
                    // (an explicit PB200_DL_WPC is taken even when it leaves a few warp slots of the SM unused)
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, round, want_wpc * 32, wb * want_wpc) == cudaSuccess &&
                        blocks > 0 && (forced || blocks * want_wpc >= kr.wpc * kr.ctas_per_sm)) {
                        kr.wpc = want_wpc; kr.ctas_per_sm = blocks; kr.smem = (int)(wb * want_wpc);
                    } else cudaGetLastError();
                }
                it = e->dl_cfg.emplace(MP, kr).first;
            }
            kr = it->second;
        }
        rgrid = std::max(1, e->sms * kr.ctas_per_sm);
    }
    const int fpw = 32 / MP;
    // Pieces of up to 4 Mi frames.  With retries every queued frame keeps its LLR row (N floats), its |L0| row (K floats)
    // and a queue entry; the stores are sized for the worst case of a piece (every frame fails), i.e. 2.4 GB for 4 Mi frames
    // of N = 128.  Every piece ends in a tail in which the last frames run their (sequential) retries on a mostly idle GPU,
    // so DL-SCL pieces are made as large as memory comfortably allows: up to 16 Mi frames when four times the stores fit
    // in free memory (a B200 has 180 GB), and DOWN to 64 Ki frames when they do not (small GPUs, MIG slices, a GPU shared
    // with another workload) -- the sweep then runs in more pieces instead of failing in cudaMalloc.
    long long piece_max = 1ll << 22;
    if (a.retries > 0) {
        // (decided once per engine: cudaMemGetInfo is a synchronous driver call that can take milliseconds on a busy
        //  host, and it would sit in front of every sweep's first launch)
        if (e->dl_piece_max == 0) {
            const size_t per_frame = (size_t)code.N * 4 + (size_t)code.K * 4 + entry_bytes(e);
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
            const size_t have = free_b + e->llr_store_bytes + e->abs_store_bytes + e->q_bytes;    // what we hold already counts
            while (piece_max < (1ll << 24) && (size_t)(piece_max * 2) * per_frame * 4 <= have) piece_max *= 2;
            while (piece_max > (1ll << 16) && (size_t)piece_max * per_frame * 2 > have) piece_max /= 2;
            e->dl_piece_max = piece_max;
        }
        piece_max = e->dl_piece_max;
    }
    const long long total = a.n_frames, begin0 = a.frame_begin;
    const long long out_base = a.frame_begin;   // per-frame outputs are indexed by frame - frame_begin of the whole call
    (void)out_base;
    for (long long off = 0; off < total; off += piece_max) {
        const long long nf = std::min(piece_max, total - off);
        SweepArgs p = a;
        // per-frame outputs / llr rows stay indexed relative to the whole call: shift the base pointers
        p.frame_begin = begin0 + off;
        p.n_frames = nf;
        if (p.llr) p.llr += (size_t)off * p.in_len;
        if (p.frame_bit_errors) p.frame_bit_errors += off;
        if (p.frame_work) p.frame_work += off;
        if (p.best_bits) p.best_bits += (size_t)off * code.K;
        if (p.best_words) p.best_words += (size_t)off * (code.N >= 32 ? code.N / 32 : 1);
        if (p.success) p.success += off;
        if (p.n_attempts) p.n_attempts += off;
        if (p.tried) p.tried += (size_t)off * p.R;
        if (p.flags) p.flags += off;
        if (a.retries > 0) {
            rc = ensure_queues(e, nf, a.retries, a.llr == nullptr);
            if (rc) return rc;
            p.llr_store = (a.llr == nullptr) ? e->d_llr_store : nullptr;
            p.abs_store = e->d_abs_store;
            CUDA_TRY(cudaMemsetAsync(e->d_q_counts, 0, sizeof(unsigned int) * (a.retries + 2), st));
            p.q_capacity = (unsigned int)nf;
            p.q_out = e->d_q[0];
            p.q_out_count = e->d_q_counts;
        } else {
            // retries <= 0: nothing is ever enqueued, but the kernel still takes a valid counter
            rc = ensure_queues(e, 1, 0, false);
            if (rc) return rc;
            CUDA_TRY(cudaMemsetAsync(e->d_q_counts, 0, sizeof(unsigned int) * 2, st));
            p.q_capacity = 1;
            p.q_out = e->d_q[0];
            p.q_out_count = e->d_q_counts;
        }
        const long long groups = (nf + fpw - 1) / fpw;
        const long long want = (groups + kb.wpc - 1) / kb.wpc;
        const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)e->sms * kb.ctas_per_sm));
        rc = ensure_scratch(e, st, (size_t)std::max(rgrid * kr.wpc, grid * kb.wpc), MP, &p.gscratch);
        if (rc) return rc;
        pin_scratch_in_l2(e, st, p.gscratch, (size_t)std::max(rgrid * kr.wpc, grid * kb.wpc) * warp_gbytes(MP, code.N, 0), MP >= 2);
        {
            void* args[3] = {(void*)&code, (void*)&e->tb, (void*)&p};
            CUDA_TRY(cudaLaunchKernel(base, dim3(grid), dim3(kb.wpc * 32), args, kb.smem, st));
        }
        if (a.retries > 0) {
            // one persistent retry launch: groups pull failed frames from the queue the baseline pass filled
            // (count at d_q_counts[0]) through the cursor d_q_counts[1] and keep each frame until it is done
            SweepArgs q = p;
            q.q_in = e->d_q[0];
            q.q_in_count = e->d_q_counts;
            q.q_out = nullptr;
            q.q_out_count = e->d_q_counts + 1;
            if (binned) {
                // rings: a frame waits in the ring of the index it flips next.  The kernel keeps about inflight_target frames
                // admitted (enough for every ring of a frequent index to hold full batches); the rings are sized beyond
                // that plus everything the resident warps can hold, so they never wrap onto an unread slot.
                const int K = code.K;
                // Everything is admitted up front when the rings can hold it (256 MB of rings: 1 Mi frames in flight at
                // K = 64): the frames then advance level by level and finish together, instead of a last generation that
                // runs its eight sequential attempts on a draining GPU.
                const size_t resident = (size_t)rgrid * kr.wpc * fpw;
                size_t cap = 1024;
                size_t cap_max = ((size_t)256 << 20) / ((size_t)K * sizeof(int));
                if (const char* env = getenv("PB200_DL_RING_MB")) cap_max = ((size_t)std::max(1, atoi(env)) << 20) / ((size_t)K * sizeof(int));
                while (cap < (size_t)nf && cap * 2 <= cap_max) cap *= 2;
                const size_t slack = 2 * resident + (size_t)K * fpw;
                while (cap < 2 * slack) cap *= 2;                              // (tiny budgets: at least the resident frames twice over)
                q.inflight_target = (unsigned int)(cap >= (size_t)nf ? (size_t)nf : cap - slack);   // (a frame enters a ring at most once)
                q.bin_cap = (unsigned int)cap;
                const size_t ring_bytes = (size_t)K * cap * sizeof(int);
                if (e->bin_ring_bytes < ring_bytes) {
                    cudaFree(e->d_bin_ring);
                    e->d_bin_ring = nullptr; e->bin_ring_bytes = 0;
                    CUDA_TRY(cudaMalloc((void**)&e->d_bin_ring, ring_bytes));
                    e->bin_ring_bytes = ring_bytes;
                    // all slots empty (-1).  Every pop resets its slot, so a finished launch leaves the rings empty again and
                    // they are initialised only here
                    CUDA_TRY(cudaMemsetAsync(e->d_bin_ring, 0xff, ring_bytes, st));
                }
                if (e->bin_ctrl_n < 96 + 32 * K) {
                    cudaFree(e->d_bin_ctrl);
                    e->d_bin_ctrl = nullptr; e->bin_ctrl_n = 0;
                    CUDA_TRY(cudaMalloc((void**)&e->d_bin_ctrl, sizeof(unsigned int) * (96 + 32 * K)));
                    e->bin_ctrl_n = 96 + 32 * K;
                }
                CUDA_TRY(cudaMemsetAsync(e->d_bin_ctrl, 0, sizeof(unsigned int) * (96 + 32 * K), st));
                q.replay_admission = replay ? 1 : 0;
                q.bin_ring = e->d_bin_ring;
                q.bin_ctrl = e->d_bin_ctrl;
            }
            void* args[3] = {(void*)&code, (void*)&e->tb, (void*)&q};
            CUDA_TRY(cudaLaunchKernel(round, dim3(rgrid), dim3(kr.wpc * 32), args, kr.smem, st));
            if (binned && e->queued_of == 0) {            // (one read-back in flight at a time)
                if (!e->h_queued) CUDA_TRY(cudaMallocHost((void**)&e->h_queued, sizeof(unsigned int)));
                if (!e->queued_ev) CUDA_TRY(cudaEventCreateWithFlags(&e->queued_ev, cudaEventDisableTiming));
                CUDA_TRY(cudaMemcpyAsync(e->h_queued, e->d_q_counts, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaEventRecord(e->queued_ev, st));
                e->queued_of = nf;
            }
        }
    }
    return PB200_OK;
}

static void span_mask(const pb200_engine* e, int span, uint32_t* mask) {
    for (int w = 0; w < kMaxWords; ++w) mask[w] = 0;
    for (int j = 0; j < span && j < e->code.K; ++j) mask[e->info_pos[j] >> 5] |= 1u << (e->info_pos[j] & 31);
}

extern "C" int pb200_sweep(pb200_engine* e, const pb200_sweep_cfg* c, const float* d_beta, int64_t* d_counters,
                           uint16_t* d_frame_bit_errors, uint16_t* d_frame_work, void* stream) {
    int rc = check_sweep_cfg(e, c);
    if (rc) return rc;
    if (!d_counters) return fail(PB200_EINVAL, "counters is NULL");
    if (c->n_frames == 0) return PB200_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    SweepArgs a{};
    a.llr = nullptr; a.in_len = 0;
    a.frame_begin = c->frame_begin; a.n_frames = c->n_frames;
    rc = fill_chan(e, c, &a.cc);
    if (rc) return rc;
    a.retries = c->retries; a.run_scl = c->run_scl; a.fe_mode = c->frame_error_mode;
    span_mask(e, c->bit_error_span, a.be_mask);
    a.beta = d_beta;
    a.counters = reinterpret_cast<unsigned long long*>(d_counters);
    a.frame_bit_errors = d_frame_bit_errors;
    a.frame_work = d_frame_work;
    a.R = std::max(c->retries, 1);
    return run_sweep(e, c->M, a, (cudaStream_t)stream);
}

extern "C" int pb200_dlscl_decode_batch(pb200_engine* e, const float* llr, int64_t B, int in_len, int M, int retries,
                                        const float* d_beta, const pb200_dl_out* out, void* stream) {
    int rc = check_decode_args(e, llr, B, in_len, M);
    if (rc) return rc;
    if (!out) return fail(PB200_EINVAL, "out is NULL");
    if (B == 0) return PB200_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    SweepArgs a{};
    a.llr = llr; a.in_len = in_len;
    a.frame_begin = 0; a.n_frames = B;
    a.cc = ChanCfg{};
    a.retries = std::max(retries, 0); a.run_scl = 0; a.fe_mode = 0;
    span_mask(e, e->code.K, a.be_mask);
    a.beta = d_beta;
    a.counters = nullptr;
    a.best_bits = out->best_bits; a.best_words = out->best_words; a.success = out->success;
    a.n_attempts = out->n_attempts; a.tried = out->tried; a.flags = out->flags;
    a.R = std::max(retries, 1);
    return run_sweep(e, M, a, (cudaStream_t)stream);
}

// Scheduler statistics of the last dl_bin_kernel launch of this engine (tuning aid): waits on empty rings, lost claims,
// batches, frame decodes, sum of the batches' start phases, batches mixing rings.  Synchronises the device.
extern "C" int pb200_debug_bin_stats(pb200_engine* e, unsigned int* out8) {
    if (!e || !out8) return fail(PB200_EINVAL, "NULL argument");
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    if (!e->d_bin_ctrl) return PB200_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out8, e->d_bin_ctrl + 64, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    return PB200_OK;
}

extern "C" int pb200_channel_batch(pb200_engine* e, const pb200_sweep_cfg* c, uint8_t* d_msg, float* d_llr, void* stream) {
    int rc = check_sweep_cfg(e, c);
    if (rc) return rc;
    if (c->n_frames == 0) return PB200_OK;
    if (!d_llr) return fail(PB200_EINVAL, "llr is NULL");
    CUDA_TRY(cudaSetDevice(e->device));
    SweepArgs a{};
    a.frame_begin = c->frame_begin; a.n_frames = c->n_frames;
    rc = fill_chan(e, c, &a.cc);
    if (rc) return rc;
    a.cc.include_uncoded = 0;
    const size_t wb = WarpMem<4, 5>::bytes(e->code.N);
    const int wpc = 4;
    const long long groups = (c->n_frames + 7) / 8;
    const int grid = (int)std::max<long long>(1, std::min<long long>((groups + wpc - 1) / wpc, (long long)e->sms * 8));
    rc = ensure_scratch(e, (cudaStream_t)stream, (size_t)grid * wpc, 4, &a.gscratch);
    if (rc) return rc;
    {
        const void* fn = pb_channel_kernel(e->code.n);
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(wb * wpc)));
        Code code = e->code;
        void* args[5] = {(void*)&code, (void*)&e->tb, (void*)&a, (void*)&d_msg, (void*)&d_llr};
        CUDA_TRY(cudaLaunchKernel(fn, dim3(grid), dim3(wpc * 32), args, wb * wpc, (cudaStream_t)stream));
    }
    CUDA_TRY(cudaGetLastError());
    return PB200_OK;
}

extern "C" int pb200_choose_flip_index_batch(const float* abs_l0, const float* beta, int32_t* idx, int64_t B, int K, void* stream) {
    if (K <= 0) return fail(PB200_EINVAL, "abs_l0 cannot be empty");
    if (B < 0) return fail(PB200_EINVAL, "bad B");
    if (B == 0) return PB200_OK;
    if (!abs_l0 || !idx) return fail(PB200_EINVAL, "NULL buffer");
    if (K > 4096) return fail(PB200_ENOSUP, "K > 4096");
    flip_index_kernel<<<(unsigned)B, 128, (size_t)K * sizeof(double), (cudaStream_t)stream>>>(abs_l0, beta, idx, K);
    CUDA_TRY(cudaGetLastError());
    return PB200_OK;
}

extern "C" int pb200_nr_encode_batch(pb200_engine* e, const uint8_t* payload, int8_t* tx, int64_t B, int E, void* stream) {
    if (!e) return fail(PB200_EINVAL, "engine is NULL");
    if (e->tb.E == 0 || E != e->tb.E) return fail(PB200_EINVAL, "call pb200_set_rate_matching(E) first");
    if (e->code.crc_deg == 0) return fail(PB200_EINVAL, "NR encode needs a CRC polynomial");
    if (B < 0) return fail(PB200_EINVAL, "bad B");
    if (B == 0) return PB200_OK;
    if (!payload || !tx) return fail(PB200_EINVAL, "NULL buffer");
    CUDA_TRY(cudaSetDevice(e->device));
    pb200_sweep_cfg c{};
    c.noise_var = 1.0; c.noise_var_uncoded = 1.0;
    c.k_payload = e->code.K - e->code.crc_deg;
    ChanCfg cc;
    int rc = fill_chan(e, &c, &cc);
    if (rc) return rc;
    nr_encode_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(e->code, e->tb, cc, payload, tx, B, E);
    CUDA_TRY(cudaGetLastError());
    return PB200_OK;
}
