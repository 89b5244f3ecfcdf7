// multi_device.inl -- ONE process, MANY GPUs behind the same C ABI (included by bsa_api.cu).
//
// The reference's caller is a single process (bin/cluster_sequences.rs:173-177 calls align_all_pairs
// once).  bsa_create_multi(device_ids, n_dev) returns a context whose entry points behave exactly
// like a single-device context's, but which owns, per GPU, one child context (BSA_MULTI_WORKERS_PER_GPU: up
// to 4) with one host worker thread each:
//   * sequence sets and scoring are replicated to every child (a set is a few MB to a few 100 MB);
//   * bsa_align_all_pairs cuts the template range into cell-balanced TILES -- by default one per GPU, each a
//     complete single-device call -- that the workers pull from one atomic counter (no collective); every
//     tile's results are copied by its child straight into the caller's buffer at the tile's own offset
//     (true DMA when the buffer comes from bsa_host_alloc_pinned): result k keeps its t-major position, so
//     the bytes are identical for any number of GPUs.  With two children per GPU the tiles are guided
//     (decreasing sizes) and the children take turns on their GPU (gpu_gate), so that one tile's copy and
//     the next tile's planning overlap the running tile's kernels;
//   * pair lists (bsa_align_pairs_paths / bsa_local_align_pairs) are cut into contiguous chunks of
//     equal cells, one per worker, and the paths are compacted back into pair order afterwards.
// There is still no CPU alignment path: the workers only call the single-device entry points.
#pragma once

namespace {

struct Worker {
    bsa_ctx* kid = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = true, quit = false;
    int rc = 0;

    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<int()> j = std::move(job);
            has_job = false;
            lk.unlock();
            const int r = j();
            lk.lock();
            rc = r;
            done = true;
            cv.notify_all();
        }
    }
    void post(std::function<int()> j) {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(j);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
};

}  // namespace

struct MultiState {
    std::vector<std::unique_ptr<Worker>> workers;
    std::vector<std::unique_ptr<std::mutex>> gates;     // one per GPU
    int n_dev = 0;
    std::mutex stats_mu;

    // run f(worker index, child) on every worker thread; first error wins and its text is kept
    int run_all(bsa_ctx* parent, const std::function<int(int, bsa_ctx*)>& f) {
        for (size_t i = 0; i < workers.size(); ++i) {
            Worker* w = workers[i].get();
            w->post([&f, i, w] { return f((int)i, w->kid); });
        }
        int rc = BSA_OK;
        for (auto& w : workers) {
            const int r = w->wait();
            if (r != BSA_OK && rc == BSA_OK) {
                rc = r;
                parent->err = w->kid->err;
            }
        }
        return rc;
    }
};

namespace {

constexpr int kWorkersPerGpuDefault = 1;

void multi_add_stats(bsa_ctx* parent, const bsa_stats& s) {
    std::lock_guard<std::mutex> lk(parent->multi->stats_mu);
    bsa_stats& d = parent->stats;
    d.pairs += s.pairs;
    d.cells += s.cells;
    d.padded_cells += s.padded_cells;
    d.launches += s.launches;
    d.items += s.items;
    d.h2d_bytes += s.h2d_bytes;
    d.d2h_bytes += s.d2h_bytes;
    d.fallback_pairs += s.fallback_pairs;
}

void multi_destroy(bsa_ctx* c) {
    MultiState* m = c->multi;
    for (auto& w : m->workers) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
            w->cv.notify_all();
        }
        if (w->th.joinable()) w->th.join();
        if (w->kid) bsa_destroy(w->kid);
    }
    delete m;
    delete c;
}

int multi_set_scoring(bsa_ctx* c, const int32_t* score, const uint8_t* aa, int32_t go, int32_t ge) {
    for (auto& w : c->multi->workers) {
        const int rc = bsa_set_scoring(w->kid, score, aa, go, ge);
        if (rc) { c->err = w->kid->err; return rc; }
    }
    c->go = go;
    c->ge = ge;
    c->have_scoring = true;
    return BSA_OK;
}

int multi_load_sequences(bsa_ctx* c, int set_id, const uint8_t* res, const uint64_t* offsets, uint32_t n) {
    if (set_id < 0 || set_id >= kMaxSets || !offsets) return fail(c, BSA_ERR_BAD_ARG, "bad set_id or null offsets");
    c->sets[set_id].loaded = false;
    const int rc = c->multi->run_all(c, [&](int, bsa_ctx* kid) { return bsa_load_sequences(kid, set_id, res, offsets, n); });
    if (rc) return rc;
    // host-side copy of the lengths: the parent plans the tiles
    SeqSet& S = c->sets[set_id];
    S.n = n;
    S.off.resize((size_t)n + 1);
    for (uint32_t i = 0; i <= n; ++i) S.off[i] = offsets[i] - offsets[0];
    S.total = S.off[n];
    S.loaded = true;
    return BSA_OK;
}

int multi_gather_sequences(bsa_ctx* c, int src_set, int dst_set, const uint32_t* idx, uint32_t n) {
    if (src_set < 0 || src_set >= kMaxSets || dst_set < 0 || dst_set >= kMaxSets || src_set == dst_set || !idx)
        return fail(c, BSA_ERR_BAD_ARG, "bad set id or null index list");
    const SeqSet& S = c->sets[src_set];
    if (!S.loaded) return fail(c, BSA_ERR_EMPTY, "sequence set not loaded");
    c->sets[dst_set].loaded = false;
    const int rc = c->multi->run_all(c, [&](int, bsa_ctx* kid) { return bsa_gather_sequences(kid, src_set, dst_set, idx, n); });
    if (rc) return rc;
    SeqSet& D = c->sets[dst_set];      // host-side copy of the lengths: the parent plans the tiles
    D.n = n;
    D.off.assign((size_t)n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) D.off[i + 1] = D.off[i] + S.len(idx[i]);
    D.total = D.off[n];
    D.loaded = true;
    return BSA_OK;
}

int multi_align_all_pairs(bsa_ctx* c, int q_set, int t_set, const uint32_t* q_counts, uint32_t t_begin,
                          uint32_t t_end, uint32_t flags, int32_t* scores, uint32_t* n_identical,
                          uint64_t* n_results) {
    const auto wall0 = std::chrono::steady_clock::now();
    if (q_set < 0 || q_set >= kMaxSets || t_set < 0 || t_set >= kMaxSets) return fail(c, BSA_ERR_BAD_ARG, "bad set id");
    const SeqSet &Q = c->sets[q_set], &T = c->sets[t_set];
    if (!Q.loaded || !T.loaded) return fail(c, BSA_ERR_EMPTY, "sequence set not loaded");
    if (t_begin > t_end || t_end > T.n) return fail(c, BSA_ERR_BAD_ARG, "bad template range");
    if (flags & BSA_OUT_DEVICE)
        return fail(c, BSA_ERR_BAD_ARG, "BSA_OUT_DEVICE needs a single-device context (results of a multi-device call land in host memory)");
    // result offsets and cell counts per template (the t-major layout of the single-device call)
    const size_t nt = (size_t)(t_end - t_begin);
    std::vector<uint64_t> first(nt + 1, 0);
    std::vector<double> pre(nt + 1, 0.0);
    for (size_t i = 0; i < nt; ++i) {
        const uint32_t t = t_begin + (uint32_t)i;
        const uint32_t cnt = q_counts ? q_counts[t] : Q.n;
        if (cnt > Q.n) return fail(c, BSA_ERR_BAD_ARG, "q_counts entry exceeds the query set size");
        first[i + 1] = first[i] + cnt;
        pre[i + 1] = pre[i] + (double)T.len(t) * (double)Q.off[cnt];
    }
    if (n_results) *n_results = first[nt];
    memset(&c->stats, 0, sizeof(c->stats));
    if (first[nt] == 0) return BSA_OK;
    // Tiles.  Default: ONE cell-balanced tile per GPU -- a tile is a complete single-device call (~100 kernel
    // groups whose tails only overlap inside one call), the GPUs are identical, and a static split of the
    // template range scales at 0.995 to 8 GPUs, so nothing is gained by cutting finer (measured: 6-7 guided
    // tiles per GPU cost 14 % on cfg2).  With more than one worker per GPU (BSA_MULTI_WORKERS_PER_GPU=2)
    // the tiles are guided instead -- each takes 1/(1.5 workers) of what is left, never less than kMinTile
    // cells -- and the two children of a GPU take turns on it (bsa_ctx::gpu_gate): one tile's kernels run
    // while the next tile is planned and the last one's results travel to the caller's buffer.
    const size_t nw = c->multi->workers.size();
    const size_t n_gpu = (size_t)c->multi->n_dev;
    double kMinTile = 2.5e11;
    if (const char* e = getenv("BSA_MULTI_MIN_TILE")) kMinTile = std::max(1e6, atof(e));
    std::vector<uint32_t> cut{0};
    if (nw == n_gpu) {
        const double total = pre[nt];
        for (size_t g = 1; g <= n_gpu && cut.back() < nt; ++g) {
            size_t e = g == n_gpu ? nt : (size_t)(std::lower_bound(pre.begin(), pre.end(), total * (double)g / (double)n_gpu) - pre.begin());
            e = std::min(std::max(e, (size_t)cut.back()), nt);
            if (e > cut.back()) cut.push_back((uint32_t)e);
        }
        if (cut.back() < nt) cut.push_back((uint32_t)nt);
    } else {
        double done = 0.0;
        const double total = pre[nt];
        while (cut.back() < nt) {
            const double chunk = std::max((total - done) / (1.5 * (double)nw), kMinTile);
            size_t e = (size_t)(std::upper_bound(pre.begin(), pre.end(), done + chunk) - pre.begin());
            e = std::min(std::max(e, (size_t)cut.back() + 1), nt);     // at least one template
            if (total - pre[e] < kMinTile * 0.25) e = nt;             // no crumbs at the end
            cut.push_back((uint32_t)e);
            done = pre[e];
        }
    }
    const size_t n_tiles = cut.size() - 1;
    std::atomic<size_t> next{0};
    const int rc = c->multi->run_all(c, [&](int, bsa_ctx* kid) {
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= n_tiles) return (int)BSA_OK;
            const uint32_t tb = t_begin + cut[k], te = t_begin + cut[k + 1];
            const uint64_t o = first[cut[k]];
            const int r = bsa_align_all_pairs(kid, q_set, t_set, q_counts, tb, te, flags, scores ? scores + o : nullptr,
                                              n_identical ? n_identical + o : nullptr, nullptr);
            if (r) { next.store(n_tiles); return r; }
            multi_add_stats(c, kid->stats);
        }
    });
    if (rc) return rc;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    c->stats.total_ms = ms;
    c->stats.kernel_ms = ms;     // devices do not share a clock: the call's wall time stands for both
    return BSA_OK;
}

// pair-list calls: contiguous chunks of equal cells, one per worker
template <class Call>
int multi_pair_list(bsa_ctx* c, int q_set, int t_set, const uint32_t* q_idx, const uint32_t* t_idx, uint64_t n_pairs,
                    uint8_t* path_buf, uint64_t* path_off, const Call& call) {
    const auto wall0 = std::chrono::steady_clock::now();
    if (q_set < 0 || q_set >= kMaxSets || t_set < 0 || t_set >= kMaxSets) return fail(c, BSA_ERR_BAD_ARG, "bad set id");
    const SeqSet &Q = c->sets[q_set], &T = c->sets[t_set];
    if (!Q.loaded || !T.loaded) return fail(c, BSA_ERR_EMPTY, "sequence set not loaded");
    if (n_pairs && (!q_idx || !t_idx)) return fail(c, BSA_ERR_BAD_ARG, "null pair list");
    if (path_buf && !path_off) return fail(c, BSA_ERR_BAD_ARG, "path_buf needs path_off");
    memset(&c->stats, 0, sizeof(c->stats));
    if (path_off) path_off[0] = 0;
    if (n_pairs == 0) return BSA_OK;
    std::vector<uint64_t> slot(n_pairs + 1, 0);
    std::vector<double> pre(n_pairs + 1, 0.0);
    for (uint64_t p = 0; p < n_pairs; ++p) {
        if (q_idx[p] >= Q.n || t_idx[p] >= T.n) return fail(c, BSA_ERR_BAD_ARG, "pair index out of range");
        const uint64_t n = Q.len(q_idx[p]), m = T.len(t_idx[p]);
        slot[p + 1] = slot[p] + n + m;
        pre[p + 1] = pre[p] + (double)n * (double)m + 1.0;
    }
    const size_t nw = std::min<size_t>(c->multi->workers.size(), (size_t)n_pairs);
    std::vector<uint64_t> cut(nw + 1, 0);
    for (size_t w = 1; w < nw; ++w) {
        uint64_t e = (uint64_t)(std::lower_bound(pre.begin(), pre.end(), pre[n_pairs] * (double)w / (double)nw) - pre.begin());
        cut[w] = std::min<uint64_t>(std::max(e, cut[w - 1]), n_pairs);
    }
    cut[nw] = n_pairs;
    std::vector<std::vector<uint64_t>> loff(nw);
    const int rc = c->multi->run_all(c, [&](int wi, bsa_ctx* kid) {
        if ((size_t)wi >= nw || cut[wi] == cut[wi + 1]) return (int)BSA_OK;
        const uint64_t a = cut[wi], cnt = cut[wi + 1] - a;
        loff[wi].assign(cnt + 1, 0);
        const int r = call(kid, a, cnt, path_buf ? path_buf + slot[a] : nullptr, path_off ? loff[wi].data() : nullptr);
        if (r == BSA_OK) multi_add_stats(c, kid->stats);
        return r;
    });
    if (rc) return rc;
    if (path_off) {
        uint64_t w = 0;
        for (size_t wi = 0; wi < nw; ++wi) {
            const uint64_t a = cut[wi], cnt = cut[wi + 1] - a;
            if (!cnt) continue;
            if (path_buf && loff[wi][cnt]) memmove(path_buf + w, path_buf + slot[a], loff[wi][cnt]);
            for (uint64_t i = 0; i < cnt; ++i) path_off[a + i + 1] = w + loff[wi][i + 1];
            w += loff[wi][cnt];
        }
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    c->stats.total_ms = ms;
    c->stats.kernel_ms = ms;
    return BSA_OK;
}

}  // namespace

extern "C" bsa_ctx* bsa_create_multi(const int* device_ids, int n_dev) {
    const int avail = bsa_device_count();
    if (avail <= 0) { fail(nullptr, BSA_ERR_CUDA, "no CUDA device is visible (there is no CPU fallback)"); return nullptr; }
    if (n_dev < 0 || (n_dev > 0 && !device_ids)) { fail(nullptr, BSA_ERR_BAD_ARG, "bad device list"); return nullptr; }
    std::vector<int> ids;
    if (n_dev == 0) for (int d = 0; d < avail; ++d) ids.push_back(d);     // every visible device
    else ids.assign(device_ids, device_ids + n_dev);
    for (size_t i = 0; i < ids.size(); ++i)
        for (size_t j = 0; j < i; ++j)
            if (ids[i] == ids[j]) { fail(nullptr, BSA_ERR_BAD_ARG, "device listed twice"); return nullptr; }
    int per_gpu = kWorkersPerGpuDefault;
    if (const char* e = getenv("BSA_MULTI_WORKERS_PER_GPU")) per_gpu = std::min(4, std::max(1, atoi(e)));
    bsa_ctx* c = new (std::nothrow) bsa_ctx();
    if (!c) return nullptr;
    for (int i = 0; i < 256; ++i) c->code_of[i] = -1;
    memset(&c->stats, 0, sizeof(c->stats));
    c->multi = new MultiState();
    c->multi->n_dev = (int)ids.size();
    c->device = ids[0];
    for (size_t g = 0; g < ids.size(); ++g) c->multi->gates.emplace_back(new std::mutex());
    for (int r = 0; r < per_gpu; ++r)           // worker order: one child per GPU first, then the second children
        for (size_t g = 0; g < ids.size(); ++g) {
            const int d = ids[g];
            bsa_ctx* kid = bsa_create(d);
            if (!kid) { multi_destroy(c); return nullptr; }   // bsa_create left the message
            if (per_gpu > 1) kid->gpu_gate = c->multi->gates[g].get();
            std::unique_ptr<Worker> w(new Worker());
            w->kid = kid;
            Worker* wp = w.get();
            w->th = std::thread([wp] { wp->loop(); });
            c->multi->workers.push_back(std::move(w));
        }
    c->sms = c->multi->workers[0]->kid->sms;
    return c;
}

extern "C" int bsa_context_devices(const bsa_ctx* ctx) {
    if (!ctx) return 0;
    return ctx->multi ? ctx->multi->n_dev : 1;
}
