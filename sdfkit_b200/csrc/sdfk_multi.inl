// sdfk_multi.inl -- the multi-GPU layer of libsdfk.so (textually included at the end of sdfk_api.cu).
//
// One process, N devices (sdfk_ctx_create_multi): the job context holds one sub-context per device -- its own stream,
// copy stream, block / pinned-buffer pools -- and a team of worker threads, one per device, that run every multi-GPU
// entry point concurrently (the calling thread drives device 0).  The grid shards into contiguous z-slabs of cell
// layers (the reference's analogue of "same job, more workers" is the Parallel.For of
// SdfKit/Voxels.cs:83-88), cut where a coarse probe pass of the same SDF says the COST is equal (plan_layers_native);
// halo slices are recomputed from the analytic SDF, never exchanged.  The one exchange step of the path -- every slab's
// (vertices, triangles) count, whose exclusive sums are the global ids -- happens in host memory behind a thread
// barrier: inside one process there is nothing to send.  Every device then writes its share of the mesh / image at its
// global offset of ONE host result over its own PCIe link.  The image shards by row bands (RayMarcher.cs:50-61).

// ------------------------------------------------------------------------------------------------
// worker team: one persistent thread per extra device
// ------------------------------------------------------------------------------------------------
struct Team {
    int n = 0;
    std::vector<int> devices;
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv;
    const std::function<int(int)>* job = nullptr;
    std::atomic<unsigned long long> gen{0};
    std::atomic<int> remaining{0};
    std::atomic<bool> quit{false};
    std::vector<int> rcs;
    std::vector<std::string> errs;
    std::atomic<int> bar_count{0};
    std::atomic<unsigned> bar_gen{0};

    explicit Team(const std::vector<int>& devs) : n((int)devs.size()), devices(devs), rcs(devs.size(), 0), errs(devs.size())
    {
        for (int r = 1; r < n; r++) threads.emplace_back([this, r] { worker(r); });
    }

    ~Team()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit.store(true);
        }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }

    void worker(int r)
    {
        cudaSetDevice(devices[(size_t)r]);
        unsigned long long seen = 0;
        for (;;) {
            // back-to-back calls find the worker still spinning; after ~0.3 ms of silence it sleeps on the condition variable
            const auto t0 = std::chrono::steady_clock::now();
            while (gen.load(std::memory_order_acquire) == seen && !quit.load(std::memory_order_acquire)) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return gen.load() != seen || quit.load(); });
                    break;
                }
                std::this_thread::yield();
            }
            if (quit.load()) return;
            seen = gen.load(std::memory_order_acquire);
            const int rc = (*job)(r);
            rcs[(size_t)r] = rc;
            if (rc) errs[(size_t)r] = g_err;
            remaining.fetch_sub(1, std::memory_order_release);
        }
    }

    // f(r) on every rank (rank 0 on the calling thread); the first failure's status and message are returned
    int run(const std::function<int(int)>& f)
    {
        job = &f;
        remaining.store(n - 1, std::memory_order_release);
        {
            std::lock_guard<std::mutex> lk(mu);
            gen.fetch_add(1, std::memory_order_release);
        }
        cv.notify_all();
        rcs[0] = f(0);
        if (rcs[0]) errs[0] = g_err;
        while (remaining.load(std::memory_order_acquire) > 0) std::this_thread::yield();
        for (int r = 0; r < n; r++)
            if (rcs[(size_t)r]) { g_err = errs[(size_t)r]; return rcs[(size_t)r]; }
        return SDFK_OK;
    }

    // every rank of a running job must call this the same number of times
    void barrier()
    {
        const unsigned g = bar_gen.load(std::memory_order_acquire);
        if (bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
            bar_count.store(0, std::memory_order_relaxed);
            bar_gen.fetch_add(1, std::memory_order_release);
        } else {
            while (bar_gen.load(std::memory_order_acquire) == g) std::this_thread::yield();
        }
    }
};

struct WallClock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int sdfk_ctx_create_multi(int ndev, const int* devices, sdfk_ctx** out)
{
    if (!out) return fail(SDFK_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (ndev < 1 || ndev > 64) return fail(SDFK_ERR_INVALID, "ndev = %d (1..64)", ndev);
    std::vector<int> ids((size_t)ndev);
    for (int r = 0; r < ndev; r++) ids[(size_t)r] = devices ? devices[r] : r;
    sdfk_ctx* c = nullptr;
    int rc = ctx_create(ids[0], nullptr, true, &c);
    if (rc) return rc;
    if (ndev > 1) {
        c->devs.push_back(c);
        for (int r = 1; r < ndev && rc == SDFK_OK; r++) {
            sdfk_ctx* d = nullptr;
            rc = ctx_create(ids[(size_t)r], nullptr, true, &d);
            if (rc == SDFK_OK) { d->parent = c; c->devs.push_back(d); }
        }
        if (rc) {
            const std::string keep = g_err;
            for (size_t r = 1; r < c->devs.size(); r++) sdfk_ctx_destroy(c->devs[r]);
            c->devs.clear();
            sdfk_ctx_destroy(c);
            g_err = keep;
            return rc;
        }
        c->team = new Team(ids);
        cudaSetDevice(ids[0]);
    }
    *out = c;
    return SDFK_OK;
}

extern "C" int sdfk_ctx_device_count(sdfk_ctx* c, int* ndev)
{
    if (!c || !ndev) return fail(SDFK_ERR_INVALID, "NULL argument");
    *ndev = is_multi(c) ? (int)c->devs.size() : 1;
    return SDFK_OK;
}

extern "C" int sdfk_ctx_last_wall_ms(sdfk_ctx* c, double* ms)
{
    if (!c || !ms) return fail(SDFK_ERR_INVALID, "NULL argument");
    *ms = c->last_wall_ms;
    return SDFK_OK;
}

static void multi_ctx_teardown(sdfk_ctx* c)   // called by sdfk_ctx_destroy before the primary itself goes
{
    delete c->team;
    c->team = nullptr;
    std::vector<sdfk_ctx*> subs(c->devs.begin() + (c->devs.empty() ? 0 : 1), c->devs.end());
    c->devs.clear();
    for (sdfk_ctx* d : subs) sdfk_ctx_destroy(d);
}

// ------------------------------------------------------------------------------------------------
// cost-balanced z-slabs (the native counterpart of sdfkit_b200/dist.py: plan_layers / weighted_partition)
// ------------------------------------------------------------------------------------------------
static void uniform_partition(int n, int parts, std::vector<std::pair<int, int>>& out)   // ceil(n/parts) items each, like
{                                                                                        // Vec3Data.PartitionVertically (VectorData.cs:512-526)
    out.clear();
    const int size = parts > 0 ? (n + parts - 1) / parts : n;
    int lo = 0;
    for (int k = 0; k < parts; k++) {
        const int hi = std::min(n, lo + size);
        out.emplace_back(lo, std::max(lo, hi));
        lo = std::max(lo, hi);
    }
}

// contiguous ranges of near-equal total weight: cut k goes where the running sum crosses k/parts of the total
static void weighted_partition(const std::vector<double>& w, int parts, std::vector<std::pair<int, int>>& out)
{
    const int n = (int)w.size();
    out.clear();
    if (parts <= 1 || n == 0) {
        out.emplace_back(0, n);
        for (int k = 1; k < parts; k++) out.emplace_back(n, n);
        return;
    }
    if (n <= parts) {
        for (int k = 0; k < n; k++) out.emplace_back(k, k + 1);
        for (int k = n; k < parts; k++) out.emplace_back(n, n);
        return;
    }
    std::vector<double> cum((size_t)n + 1, 0.0);
    for (int k = 0; k < n; k++) cum[(size_t)k + 1] = cum[(size_t)k] + w[(size_t)k];
    std::vector<int> cuts(1, 0);
    for (int k = 1; k < parts; k++) {
        const double target = cum[(size_t)n] * k / parts;
        int cpos = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        if (cpos > 0 && std::fabs(cum[(size_t)cpos - 1] - target) <= std::fabs(cum[(size_t)std::min(cpos, n)] - target)) cpos--;
        cpos = std::max(cpos, cuts.back() + 1);          // at least one layer per part ...
        cpos = std::min(cpos, n - (parts - k));          // ... and one left for each remaining part
        cuts.push_back(std::max(cpos, cuts.back()));
    }
    cuts.push_back(n);
    for (int k = 0; k < parts; k++) out.emplace_back(cuts[(size_t)k], cuts[(size_t)k + 1]);
}

static void slab_slices(int kb, int ke, int step, int nz, int& z0, int& z1)   // slices a device holds to mesh layers [kb, ke)
{
    const int ncz = cells_along(nz, step);
    if (ke <= kb) { z0 = 0; z1 = 1; return; }
    const int k0 = kb > 0 ? kb - 1 : 0, k1 = ke < ncz ? ke + 1 : ncz;
    z0 = k0 * step;
    z1 = k1 * step + 1;
}

// A z-slab job is as slow as its busiest device and a surface is rarely spread evenly in z (the README scene fills 18 % of
// the layers): cost(layer) = voxels(layer) + active_cell_cost * active_cells(layer), the active cells estimated from a
// coarse (<= 128^3) meshing pass of the same SDF on this context's device.  Caller holds the ctx lock.
static int plan_layers_native(sdfk_ctx* c, sdfk_sdf* s, const float mn[3], const float mx[3], int nx, int ny, int nz, int step, int clip,
                              int parts, double active_cell_cost, std::vector<std::pair<int, int>>& out)
{
    const int ncz = cells_along(nz, step);
    if (parts <= 1 || ncz <= parts) { uniform_partition(ncz, parts, out); return SDFK_OK; }
    const int cz = std::min(128, nz);
    if (cz < 8) { uniform_partition(ncz, parts, out); return SDFK_OK; }
    sdfk_sdf::PlanKey key;
    memset(&key, 0, sizeof(key));
    memcpy(key.mn, mn, 12); memcpy(key.mx, mx, 12);
    key.nx = nx; key.ny = ny; key.nz = nz; key.step = step; key.clip = clip ? 1 : 0; key.parts = parts; key.cost = active_cell_cost;
    static const bool use_cache = getenv("SDFK_NO_PLAN_CACHE") == nullptr;
    if (use_cache)
        for (auto& e : s->plans)
            if (memcmp(&e.first, &key, sizeof(key)) == 0) { out = e.second; return SDFK_OK; }
    const int cx = std::max(8, (int)std::lround((double)nx * cz / nz)), cy = std::max(8, (int)std::lround((double)ny * cz / nz));
    const int nkc = cells_along(cz, 1);
    std::vector<double> tri_per_layer((size_t)cz, 0.0);
    {
        sdfk_voxels* v = nullptr;
        sdfk_mesh* m = nullptr;
        int rc = voxels_alloc(c, mn, mx, cx, cy, cz, 0, cz, &v, false);
        if (rc == SDFK_OK) rc = voxels_sample_into(v, s, clip);
        if (rc == SDFK_OK) rc = mesh_stage_a(c, v, 0.0f, 1, 0, nkc, &m);
        if (rc == SDFK_OK) rc = mesh_stage_b(m);
        if (rc == SDFK_OK) rc = mesh_stage_b_finish(m);
        if (rc == SDFK_OK && m->compacted && m->ntris > 0) {
            // triangle prefixes at every coarse layer boundary, kernel-written to mapped page-locked memory
            if (!c->plan_scratch && cudaHostAlloc(&c->plan_scratch, 132 * sizeof(uint4), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess)
                rc = fail(SDFK_ERR_CUDA, "page-locked plan scratch allocation failed");
            if (rc == SDFK_OK) {
                uint4* pre = (uint4*)c->plan_scratch;
                cudaError_t e = mc_launch_layer_prefixes(m->base, m->abase, (size_t)m->g.ncy * m->g.cpr, (unsigned)m->g.nk, pre, m->ws);
                c->launches++;
                if (e == cudaSuccess) e = cudaStreamSynchronize(m->ws);
                if (e != cudaSuccess) rc = fail(SDFK_ERR_CUDA, "plan probe: %s", cudaGetErrorString(e));
                for (int k = 0; k < m->g.nk && rc == SDFK_OK; k++) {
                    const double hi = (k + 1 < m->g.nk) ? (double)pre[k + 1].z : (double)m->hs->tot2.ntris;
                    tri_per_layer[(size_t)k] = hi - (double)pre[k].z;
                }
            }
        }
        if (m) { m->vox = nullptr; mesh_free_device(m); delete m; }
        if (v) voxels_free(v);
        if (rc) return rc;
    }
    // a coarse layer covers nz/cz fine layers; active cells scale with the square of the refinement, ~2 triangles per cell
    const double scale = (double)nz / cz;
    std::vector<double> w((size_t)ncz);
    for (int k = 0; k < ncz; k++) {
        const int cl = (int)std::min<long long>(((long long)k * step * cz) / nz, cz - 1);
        const double active = 0.5 * tri_per_layer[(size_t)cl] * scale * step;
        w[(size_t)k] = (double)nx * ny * step + active_cell_cost * active;
    }
    weighted_partition(w, parts, out);
    if (s->plans.size() >= 16) s->plans.erase(s->plans.begin());
    s->plans.emplace_back(key, out);
    return SDFK_OK;
}

// cost of one active cell (compact + emit) in units of one sampled voxel, measured on B200 at 1024^3; with the mesh
// delivered to host memory an active cell also costs its ~60 bytes over PCIe (dist.py: ACTIVE_CELL_COST[_E2E])
static const double kActiveCellCost = 82.0, kActiveCellCostHost = 1200.0;

extern "C" int sdfk_plan_layers(sdfk_ctx* c, sdfk_sdf* s, const float mn[3], const float mx[3], int nx, int ny, int nz, int step, int clip,
                                int parts, double active_cell_cost, int* kb_ke)
{
    if (!c || !s || !mn || !mx || !kb_ke) return fail(SDFK_ERR_INVALID, "NULL argument");
    if (s->ctx != c) return fail(SDFK_ERR_INVALID, "sdf belongs to another context");
    if (nx < 1 || ny < 1 || nz < 1 || step < 1 || parts < 1) return fail(SDFK_ERR_INVALID, "bad argument to sdfk_plan_layers");
    std::vector<std::pair<int, int>> layers;
    {
        Lock l(c);
        int rc = plan_layers_native(c, s, mn, mx, nx, ny, nz, step, clip, parts, active_cell_cost > 0 ? active_cell_cost : kActiveCellCost, layers);
        if (rc) return rc;
    }
    for (int k = 0; k < parts; k++) { kb_ke[2 * k] = layers[(size_t)k].first; kb_ke[2 * k + 1] = layers[(size_t)k].second; }
    return SDFK_OK;
}

// ------------------------------------------------------------------------------------------------
// SDF: the same cubin loaded on every device
// ------------------------------------------------------------------------------------------------
static int multi_sdf_load(sdfk_ctx* c, sdfk_sdf* s0, const std::vector<char>& cubin, bool has_dist8)
{
    s0->parts.assign(1, s0);
    for (size_t r = 1; r < c->devs.size(); r++) {
        sdfk_sdf* sr = nullptr;
        int rc = sdf_load(c->devs[r], cubin, &sr, has_dist8);
        if (rc) {
            for (size_t q = 1; q < s0->parts.size(); q++) sdfk_sdf_destroy(s0->parts[q]);
            s0->parts.clear();
            return rc;
        }
        s0->parts.push_back(sr);
    }
    cudaSetDevice(c->device);
    return SDFK_OK;
}

// ------------------------------------------------------------------------------------------------
// Voxels sharded by z-slab
// ------------------------------------------------------------------------------------------------
static void multi_voxels_free(sdfk_voxels* shell)
{
    for (sdfk_voxels* p : shell->parts)
        if (p) sdfk_voxels_destroy(p);
    delete shell;
}

static int multi_voxels_sample(sdfk_ctx* c, sdfk_sdf* s, const float mn[3], const float mx[3], int nx, int ny, int nz, int clip,
                               bool colors, sdfk_voxels** out)
{
    if (s->parts.size() != c->devs.size()) return fail(SDFK_ERR_INVALID, "sdf was not compiled on this multi-GPU context");
    if (nx < 1 || ny < 1 || nz < 1) return fail(SDFK_ERR_INVALID, "bad grid %dx%dx%d", nx, ny, nz);
    std::lock_guard<std::mutex> job(c->multi_mu);
    WallClock wall;
    const int n = (int)c->devs.size();
    sdfk_voxels* shell = new sdfk_voxels();
    shell->ctx = c;
    shell->nx = nx; shell->ny = ny; shell->nz = nz; shell->z0 = 0; shell->nzl = nz;
    memcpy(shell->mn, mn, 12);
    memcpy(shell->mx, mx, 12);
    shell->parts.assign((size_t)n, nullptr);
    {
        Lock l(c);
        int rc = plan_layers_native(c, s, mn, mx, nx, ny, nz, 1, clip, n, kActiveCellCost, shell->layers);
        if (rc) { delete shell; return rc; }
    }
    int rc = c->team->run([&](int r) -> int {
        sdfk_ctx* d = c->devs[(size_t)r];
        const int kb = shell->layers[(size_t)r].first, ke = shell->layers[(size_t)r].second;
        if (ke <= kb && !(nz == 1 && r == 0)) return SDFK_OK;           // nothing to own (fewer cell layers than devices)
        int z0, z1;
        slab_slices(kb, ke, 1, nz, z0, z1);
        if (nz == 1) { z0 = 0; z1 = 1; }
        Lock l(d);
        sdfk_voxels* v = nullptr;
        int rr = voxels_alloc(d, mn, mx, nx, ny, nz, z0, z1, &v, colors);
        if (rr == SDFK_OK) rr = voxels_sample_into(v, s->parts[(size_t)r], clip);
        if (rr) { if (v) voxels_free(v); return rr; }
        shell->parts[(size_t)r] = v;
        return SDFK_OK;
    });
    if (rc) { const std::string keep = g_err; multi_voxels_free(shell); g_err = keep; return rc; }
    c->last_wall_ms = wall.ms();
    *out = shell;
    return SDFK_OK;
}

static int multi_voxels_resample(sdfk_voxels* shell, sdfk_sdf* s, int clip)
{
    sdfk_ctx* c = shell->ctx;
    if (s->parts.size() != c->devs.size()) return fail(SDFK_ERR_INVALID, "sdf was not compiled on this multi-GPU context");
    std::lock_guard<std::mutex> job(c->multi_mu);
    WallClock wall;
    int rc = c->team->run([&](int r) -> int {
        sdfk_voxels* v = shell->parts[(size_t)r];
        if (!v) return SDFK_OK;
        Lock l(v->ctx);
        return voxels_sample_into(v, s->parts[(size_t)r], clip);
    });
    c->last_wall_ms = wall.ms();
    return rc;
}

// slices [zlo, zhi) of the grid that device r OWNS (halo slices excluded): the planes of its cell layers, the last
// device also the top plane
static void owned_slices(const sdfk_voxels* shell, int r, int& zlo, int& zhi)
{
    const int ncz = cells_along(shell->nz, 1);
    const int kb = shell->layers[(size_t)r].first, ke = shell->layers[(size_t)r].second;
    zlo = kb;
    zhi = ke >= ncz ? shell->nz : ke;
    if (ke <= kb) { zlo = zhi = 0; if (shell->nz == 1 && r == 0) zhi = 1; }
}

static int multi_voxels_export(sdfk_voxels* shell, float* values, float* colors)
{
    sdfk_ctx* c = shell->ctx;
    std::lock_guard<std::mutex> job(c->multi_mu);
    WallClock wall;
    int rc = c->team->run([&](int r) -> int {
        sdfk_voxels* v = shell->parts[(size_t)r];
        if (!v) return SDFK_OK;
        int zlo, zhi;
        owned_slices(shell, r, zlo, zhi);
        if (zhi <= zlo) return SDFK_OK;
        Lock l(v->ctx);
        return voxels_export_range(v, zlo - v->z0, zhi - v->z0, values, colors, shell->nz, zlo);
    });
    c->last_wall_ms = wall.ms();
    return rc;
}

// ------------------------------------------------------------------------------------------------
// MarchingCubes.CreateMesh on sharded voxels -> a device-resident mesh in one part per device
// ------------------------------------------------------------------------------------------------
static void multi_mesh_free(sdfk_mesh* shell)
{
    for (sdfk_mesh* p : shell->parts)
        if (p) sdfk_mesh_destroy(p);
    shell->parts.clear();
}

static int multi_mesh_create(sdfk_ctx* c, sdfk_voxels* shell, float iso, int step, const float* M, const float* N,
                             sdfk_progress_fn progress, void* user, sdfk_mesh** out)
{
    if (step != 1) return fail(SDFK_ERR_UNSUPPORTED, "voxels sharded over %d GPUs hold one halo slice per slab and mesh at step 1 only "
                               "(Sdf.ToMesh / sdfk_sdf_to_mesh_host takes any step)", (int)c->devs.size());
    std::lock_guard<std::mutex> job(c->multi_mu);
    WallClock wall;
    const int n = (int)c->devs.size();
    sdfk_mesh* R = new sdfk_mesh();
    R->ctx = c;
    R->emitted = true;
    memset(&R->g, 0, sizeof(R->g));
    R->g.ncz = cells_along(shell->nz, step);
    R->parts.assign((size_t)n, nullptr);
    R->voff.assign((size_t)n, 0);
    R->toff.assign((size_t)n, 0);
    std::vector<int64_t> nv((size_t)n, 0), nt((size_t)n, 0);
    std::atomic<int> failed{0};
    Team* team = c->team;
    int rc = team->run([&](int r) -> int {
        sdfk_voxels* v = shell->parts[(size_t)r];
        sdfk_mesh* m = nullptr;
        int rr = SDFK_OK;
        if (v && shell->layers[(size_t)r].second > shell->layers[(size_t)r].first) {
            Lock l(v->ctx);
            rr = mesh_stage_a(v->ctx, v, iso, step, shell->layers[(size_t)r].first, shell->layers[(size_t)r].second, &m);
            if (rr == SDFK_OK) rr = mesh_stage_b(m);
            if (rr == SDFK_OK) rr = mesh_stage_b_finish(m);
            if (rr == SDFK_OK) { nv[(size_t)r] = m->nverts; nt[(size_t)r] = m->ntris; }
        }
        if (rr) failed.store(1);
        team->barrier();                                    // the exchange step: every slab's counts are now in host memory
        if (failed.load()) {
            if (m) { Lock l(m->ctx); mesh_free_device(m); delete m; }
            return rr;
        }
        int64_t vb = 0, tb = 0;
        for (int q = 0; q < r; q++) { vb += nv[(size_t)q]; tb += nt[(size_t)q]; }
        R->voff[(size_t)r] = vb;
        R->toff[(size_t)r] = tb;
        if (m) {
            Lock l(m->ctx);
            rr = mesh_emit(m, vb, tb, M, N);
            if (rr) { mesh_free_device(m); delete m; return rr; }
            R->parts[(size_t)r] = m;
        }
        return SDFK_OK;
    });
    if (rc) { const std::string keep = g_err; multi_mesh_free(R); delete R; g_err = keep; return rc; }
    bool any = false;
    for (int r = 0; r < n; r++) {
        sdfk_mesh* m = R->parts[(size_t)r];
        R->nverts += nv[(size_t)r];
        R->ntris += nt[(size_t)r];
        if (!m) continue;
        R->nact_total += m->rec_end - m->rec_begin;
        R->g.nchunks += m->g.nchunks;
        R->from_signs = m->from_signs;
        for (int k = 0; k < 4; k++) R->ms[k] = std::max(R->ms[k], m->ms[k]);
        if (m->nverts > 0) {
            for (int k = 0; k < 3; k++) {
                R->aabb[k] = any ? std::min(R->aabb[k], m->aabb[k]) : m->aabb[k];
                R->aabb[k + 3] = any ? std::max(R->aabb[k + 3], m->aabb[k + 3]) : m->aabb[k + 3];
            }
            any = true;
        }
    }
    R->rec_begin = 0;
    R->rec_end = (unsigned)std::min<int64_t>(R->nact_total, 0xFFFFFFFFll);
    c->last_wall_ms = wall.ms();
    if (progress) {
        const int nzb = shell->nz - 2 * step;
        for (int k = 0; k < R->g.ncz; k++) progress((float)(k * step) / (float)nzb, user);
    }
    *out = R;
    return SDFK_OK;
}

static int multi_mesh_export(sdfk_mesh* R, float* verts, float* cols, float* nrms, int32_t* tris, float aabb[6])
{
    sdfk_ctx* c = R->ctx;
    std::lock_guard<std::mutex> job(c->multi_mu);
    WallClock wall;
    int rc = c->team->run([&](int r) -> int {
        sdfk_mesh* m = R->parts[(size_t)r];
        if (!m) return SDFK_OK;
        const size_t vo = (size_t)R->voff[(size_t)r] * 3, to = (size_t)R->toff[(size_t)r] * 3;
        return sdfk_mesh_export(m, verts ? verts + vo : nullptr, cols ? cols + vo : nullptr, nrms ? nrms + vo : nullptr,
                                tris ? tris + to : nullptr, nullptr);
    });
    if (rc == SDFK_OK && aabb) memcpy(aabb, R->aabb, sizeof(R->aabb));
    c->last_wall_ms = wall.ms();
    return rc;
}

// ------------------------------------------------------------------------------------------------
// SdfEx.ToMesh on N devices, ONE host mesh (SdfKit/Sdf.cs:59-63)
// ------------------------------------------------------------------------------------------------
static int multi_to_mesh_host(sdfk_ctx* c, sdfk_sdf* s, const float mn[3], const float mx[3], int nx, int ny, int nz, int clip,
                              float iso, int step, const float* M, const float* N, sdfk_progress_fn progress, void* user, sdfk_mesh** out)
{
    if (s->parts.size() != c->devs.size()) return fail(SDFK_ERR_INVALID, "sdf was not compiled on this multi-GPU context");
    std::lock_guard<std::mutex> job(c->multi_mu);
    WallClock wall;
    const int n = (int)c->devs.size();
    const int ncz = cells_along(nz, step);
    std::vector<std::pair<int, int>> layers;
    {
        Lock l(c);
        int rc = plan_layers_native(c, s, mn, mx, nx, ny, nz, step, clip, n, kActiveCellCostHost, layers);
        if (rc) return rc;
    }
    sdfk_mesh* R = new sdfk_mesh();
    R->ctx = c;
    R->on_host = true;
    R->emitted = true;
    memset(&R->g, 0, sizeof(R->g));
    R->g.ncz = ncz;
    std::vector<int64_t> nv((size_t)n, 0), nt((size_t)n, 0), nact((size_t)n, 0);
    std::vector<unsigned> nchunks((size_t)n, 0);
    std::vector<std::array<float, 6>> boxes((size_t)n);
    std::vector<char> has_box((size_t)n, 0), from_signs((size_t)n, 0);
    std::vector<std::array<double, 4>> stage_ms((size_t)n, std::array<double, 4>{{0, 0, 0, 0}});
    std::atomic<int> failed{0};
    Team* team = c->team;
    if (g_trace) fprintf(stderr, "[sdfk] multi to_mesh_host: plan %.3f ms\n", wall.ms());
    // A device's layer range is meshed in sub-slabs of at most `cap` cell layers: 32-bit cell ids bound one meshing job to
    // 2^32 cells (2048^2 cells per layer: 1023 layers), and 256-layer blocks keep the allocations recyclable.
    const long long cells_per_layer = (long long)std::max(1, cells_along(nx, step)) * std::max(1, cells_along(ny, step));
    int cap = (int)std::max<long long>(1, std::min<long long>(256, 0xFFFFFFF0ll / cells_per_layer - 2));
    if (const char* env = getenv("SDFK_SUBSLAB_CAP")) cap = std::max(1, atoi(env));     // (tests: force several sub-slabs on small grids)
    struct Sub { sdfk_voxels* v = nullptr; sdfk_mesh* m = nullptr; int kb = 0, ke = 0; };
    int rc = team->run([&](int r) -> int {
        sdfk_ctx* d = c->devs[(size_t)r];
        const int kb = layers[(size_t)r].first, ke = layers[(size_t)r].second;
        std::vector<Sub> subs;
        for (int a = kb; a < ke; a += cap) { Sub sb; sb.kb = a; sb.ke = std::min(ke, a + cap); subs.push_back(sb); }
        int rr = SDFK_OK;
        WallClock wr;
        if (ke > kb) {
            Lock l(d);
            if (!d->copy_stream && cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
                rr = fail(SDFK_ERR_CUDA, "cudaStreamCreate failed");
            for (auto& sb : subs) {                              // stage A of every sub-slab, enqueued back to back
                int z0, z1;
                slab_slices(sb.kb, sb.ke, step, nz, z0, z1);
                if (rr == SDFK_OK) rr = voxels_alloc(d, mn, mx, nx, ny, nz, z0, z1, &sb.v, false);
                if (rr == SDFK_OK) rr = voxels_sample_into(sb.v, s->parts[(size_t)r], clip, iso);
                if (rr == SDFK_OK) rr = mesh_stage_a(d, sb.v, iso, step, sb.kb, sb.ke, &sb.m);
            }
            for (auto& sb : subs) {
                if (rr == SDFK_OK) rr = mesh_stage_b(sb.m);
                if (rr == SDFK_OK) rr = mesh_stage_b_finish(sb.m);
                if (rr == SDFK_OK) { nv[(size_t)r] += sb.m->nverts; nt[(size_t)r] += sb.m->ntris; }
            }
        }
        if (rr) failed.store(1);
        const double t_classified = wr.ms();
        team->barrier();                                    // counts of every slab are in host memory
        const double t_barrier = wr.ms();
        if (!failed.load() && r == 0) {                     // device 0's thread sizes the ONE host result (recycled page-locked buffers)
            Lock l(c);
            int64_t vt = 0, tt = 0;
            for (int q = 0; q < n; q++) { vt += nv[(size_t)q]; tt += nt[(size_t)q]; }
            cudaError_t e = cudaSuccess;
            if (vt > 0x7FFFFFFFll) { rr = fail(SDFK_ERR_UNSUPPORTED, "global vertex ids exceed int32 (reference Mesh.Triangles is int[])"); failed.store(1); }
            for (int a = 0; a < 3 && e == cudaSuccess && !rr; a++) e = host_ensure(c, R->host[a], (size_t)vt * 12, 0, c->host_mesh_hint[0]);
            if (e == cudaSuccess && !rr) e = host_ensure(c, R->host[3], (size_t)tt * 12, 0, c->host_mesh_hint[1]);
            if (e != cudaSuccess) { rr = fail(SDFK_ERR_CUDA, "page-locked mesh buffer: %s", cudaGetErrorString(e)); failed.store(1); }
            R->nverts = vt;
            R->ntris = tt;
        }
        team->barrier();
        if (!failed.load() && ke > kb) {
            Lock l(d);
            int64_t vb = 0, tb = 0;
            for (int q = 0; q < r; q++) { vb += nv[(size_t)q]; tb += nt[(size_t)q]; }
            for (auto& sb : subs) {                              // emit in sub-ranges, every finished part on its way to the host
                sdfk_mesh* m = sb.m;
                EmitPlan plan;
                plan.host = R->host;
                plan.copy = d->copy_stream;
                plan.host_voff = (size_t)vb * 12;
                plan.host_toff = (size_t)tb * 12;
                cudaError_t e = build_emit_plan(m, 0, plan);
                if (e != cudaSuccess && rr == SDFK_OK) rr = fail(SDFK_ERR_CUDA, "emit plan: %s", cudaGetErrorString(e));
                if (rr == SDFK_OK) rr = mesh_emit_async(m, vb, tb, M, N, nullptr, nullptr, &plan);
                vb += m->nverts;
                tb += m->ntris;
            }
            cudaError_t e = cudaStreamSynchronize(d->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(d->copy_stream);
            if (e != cudaSuccess && rr == SDFK_OK) rr = fail(SDFK_ERR_CUDA, "sdfk_sdf_to_mesh_host: %s", cudaGetErrorString(e));
            for (auto& sb : subs) {
                sdfk_mesh* m = sb.m;
                if (rr == SDFK_OK && m->hs->err) rr = fail(SDFK_ERR_INTERNAL, "marching-cubes emit: inconsistent vertex ownership (code %d)", m->hs->err);
                if (rr != SDFK_OK) break;
                mesh_stage_times(m);
                if (m->nverts > 0) {
                    float bx[6];
                    decode_aabb(m->hs->keys, bx);
                    for (int k = 0; k < 3; k++) {
                        boxes[(size_t)r][(size_t)k] = has_box[(size_t)r] ? std::min(boxes[(size_t)r][(size_t)k], bx[k]) : bx[k];
                        boxes[(size_t)r][(size_t)k + 3] = has_box[(size_t)r] ? std::max(boxes[(size_t)r][(size_t)k + 3], bx[k + 3]) : bx[k + 3];
                    }
                    has_box[(size_t)r] = 1;
                }
                nact[(size_t)r] += (int64_t)m->rec_end - (int64_t)m->rec_begin;
                nchunks[(size_t)r] += m->g.nchunks;
                if (m->from_signs) from_signs[(size_t)r] = 1;
                for (int k = 0; k < 4; k++) stage_ms[(size_t)r][(size_t)k] += m->ms[k];
            }
            if (g_trace) fprintf(stderr, "[sdfk] dev %d layers [%d,%d) in %d sub-slab(s): classified %.3f barrier %.3f done %.3f ms, %lld vertices\n", r, kb, ke,
                                 (int)subs.size(), t_classified, t_barrier, wr.ms(), (long long)nv[(size_t)r]);
        }
        {
            Lock l(d);
            for (auto& sb : subs) {
                if (sb.m) { sb.m->vox = nullptr; mesh_free_device(sb.m); delete sb.m; }
                if (sb.v) voxels_free(sb.v);
            }
        }
        return rr;
    });
    if (rc) {
        const std::string keep = g_err;
        { Lock l(c); mesh_free_device(R); }
        delete R;
        g_err = keep;
        return rc;
    }
    bool any = false;
    for (int r = 0; r < n; r++) {
        R->nact_total += nact[(size_t)r];
        R->g.nchunks += nchunks[(size_t)r];
        if (from_signs[(size_t)r]) R->from_signs = true;
        for (int k = 0; k < 4; k++) R->ms[k] = std::max(R->ms[k], stage_ms[(size_t)r][(size_t)k]);   // stage times: slowest device
        if (!has_box[(size_t)r]) continue;
        for (int k = 0; k < 3; k++) {
            R->aabb[k] = any ? std::min(R->aabb[k], boxes[(size_t)r][(size_t)k]) : boxes[(size_t)r][(size_t)k];
            R->aabb[k + 3] = any ? std::max(R->aabb[k + 3], boxes[(size_t)r][(size_t)k + 3]) : boxes[(size_t)r][(size_t)k + 3];
        }
        any = true;
    }
    R->rec_begin = 0;
    R->rec_end = (unsigned)std::min<int64_t>(R->nact_total, 0xFFFFFFFFll);
    c->host_mesh_hint[0] = (size_t)R->nverts * 12 + (size_t)R->nverts * 12 / 16;
    c->host_mesh_hint[1] = (size_t)R->ntris * 12 + (size_t)R->ntris * 12 / 16;
    c->last_wall_ms = wall.ms();
    if (progress) {   // the reference reports (float)z / nz_bound after every z layer (MarchingCubes.cs:81)
        const int nzb = nz - 2 * step;
        for (int k = 0; k < ncz; k++) progress((float)(k * step) / (float)nzb, user);
    }
    *out = R;
    return SDFK_OK;
}

// ------------------------------------------------------------------------------------------------
// RayMarcher: row bands, one per device (RayMarcher.cs:50-61, VectorData.cs:512-526)
// ------------------------------------------------------------------------------------------------
// kind: 0 = Render (3 floats / pixel), 1 = RenderDepth (1 float), 2 = Render + SaveTga bytes (3 bytes), 3 = RenderDepth + SaveDepthTga bytes (1 byte)
static int multi_render(sdfk_ctx* c, sdfk_sdf* s, int kind, int w, int h, const float cam[3], const float ivp[16], float nearp,
                        float farp, int iters, int r0, int r1, void* out, float tga_near)
{
    if (s->parts.size() != c->devs.size()) return fail(SDFK_ERR_INVALID, "sdf was not compiled on this multi-GPU context");
    std::lock_guard<std::mutex> job(c->multi_mu);
    WallClock wall;
    std::vector<std::pair<int, int>> bands;
    uniform_partition(r1 - r0, (int)c->devs.size(), bands);
    const size_t px_bytes = kind == 0 ? 12 : kind == 1 ? 4 : kind == 2 ? 3 : 1;
    int rc = c->team->run([&](int r) -> int {
        const int a = r0 + bands[(size_t)r].first, b = r0 + bands[(size_t)r].second;
        if (b <= a) return SDFK_OK;
        return render_single(c->devs[(size_t)r], s->parts[(size_t)r], kind, w, h, cam, ivp, nearp, farp, iters, a, b,
                             (char*)out + (size_t)(a - r0) * w * px_bytes, tga_near);
    });
    c->last_wall_ms = wall.ms();
    return rc;
}

// ------------------------------------------------------------------------------------------------
// introspection of sharded objects
// ------------------------------------------------------------------------------------------------
extern "C" int sdfk_voxels_layers(sdfk_voxels* v, int* kb_ke, int max_parts, int* nparts)
{
    if (!v || !nparts) return fail(SDFK_ERR_INVALID, "NULL argument");
    *nparts = (int)v->parts.size();
    for (int r = 0; r < *nparts && r < max_parts && kb_ke; r++) { kb_ke[2 * r] = v->layers[(size_t)r].first; kb_ke[2 * r + 1] = v->layers[(size_t)r].second; }
    return SDFK_OK;
}

extern "C" int sdfk_voxels_part(sdfk_voxels* v, int r, sdfk_voxels** part)
{
    if (!v || !part) return fail(SDFK_ERR_INVALID, "NULL argument");
    if (r < 0 || r >= (int)v->parts.size()) return fail(SDFK_ERR_INVALID, "part %d out of range (%d parts)", r, (int)v->parts.size());
    *part = v->parts[(size_t)r];
    return SDFK_OK;
}

extern "C" int sdfk_mesh_part(sdfk_mesh* m, int r, sdfk_mesh** part, int64_t* vertex_base, int64_t* triangle_base)
{
    if (!m || !part) return fail(SDFK_ERR_INVALID, "NULL argument");
    if (r < 0 || r >= (int)m->parts.size()) return fail(SDFK_ERR_INVALID, "part %d out of range (%d parts)", r, (int)m->parts.size());
    *part = m->parts[(size_t)r];
    if (vertex_base) *vertex_base = m->voff[(size_t)r];
    if (triangle_base) *triangle_base = m->toff[(size_t)r];
    return SDFK_OK;
}
