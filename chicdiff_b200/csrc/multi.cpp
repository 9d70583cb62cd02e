// multi.cpp -- several GPUs from ONE host process (cd_multi_*): what a single R session needs to reach the 8 B200s of a
// box (chicdiffPipeline runs in one process, chicdiff.R:301-347).
//
// A cd_multi owns one cd_ctx per device and drives them with one host thread per device for the duration of every
// call (the single-GPU entry points are synchronous, and the global steps of cd_region_test meet the other ranks inside
// their kernels, so the contexts must run concurrently).  Regions are cut into contiguous, bait-aligned, row-balanced
// shards (cd_plan_shards), shard k lives on device k, per-region results come back in region order.  The contexts are
// joined exactly like the ranks of a multi-process run (cd_comm_unique_id / cd_comm_init from every thread); because they
// share an address space, cd_comm_init maps the peer-memory mailboxes with cudaDeviceEnablePeerAccess and plain pointers
// instead of cudaIpc handles.  Built only on the public single-context ABI.
#include "../../include/chicdiff_b200.h"
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

struct cd_multi {
    std::vector<cd_ctx*> ctx;
    std::vector<int> device;
    std::string err;
    int S = 0, p = 0;
    int64_t n = 0, R = 0;
    std::vector<int64_t> bounds;        // region bounds per shard (n_gpus + 1)
    std::vector<int64_t> row_lo;        // first region row of every shard (n_gpus + 1)
    bool have_regions = false;

    int fail(int code, const std::string& m) { err = m; return code; }

    // runs fn(k) on one thread per device and returns the first failure (with that context's message)
    int each(const std::function<int(int)>& fn)
    {
        const int G = (int)ctx.size();
        std::vector<int> rc((size_t)G, CD_OK);
        std::vector<std::thread> th;
        th.reserve((size_t)G);
        for (int k = 0; k < G; k++) th.emplace_back([&, k]() { rc[(size_t)k] = fn(k); });
        for (auto& t : th) t.join();
        for (int k = 0; k < G; k++)
            if (rc[(size_t)k] != CD_OK) {
                const char* m = cd_last_error(ctx[(size_t)k]);
                err = "device " + std::to_string(device[(size_t)k]) + ": " + (m ? m : "");
                return rc[(size_t)k];
            }
        return CD_OK;
    }
};

// shard k's S x n_k (sample-major) block into the S x n matrix of the caller
template <typename T>
static void scatter_matrix(T* dst, const T* src, int rows, int64_t n, int64_t lo, int64_t nk)
{
    for (int j = 0; j < rows; j++) memcpy(dst + (size_t)j * (size_t)n + (size_t)lo, src + (size_t)j * (size_t)nk, sizeof(T) * (size_t)nk);
}

static thread_local std::string g_multi_create_error;

extern "C" {

const char* cd_multi_last_error(const cd_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

void cd_multi_destroy(cd_multi* m)
{
    if (!m) return;
    for (cd_ctx* c : m->ctx) cd_destroy(c);
    delete m;
}

int cd_multi_create(cd_multi** out, int n_gpus, const int* device_ids)
{
    if (!out) return CD_EINVAL;
    *out = nullptr;
    if (n_gpus < 1 || n_gpus > 16) { g_multi_create_error = "cd_multi_create: 1..16 GPUs"; return CD_EINVAL; }
    cd_multi* m = new (std::nothrow) cd_multi();
    if (!m) return CD_ENOMEM;
    for (int k = 0; k < n_gpus; k++) {
        const int dev = device_ids ? device_ids[k] : k;
        cd_ctx* c = nullptr;
        const int rc = cd_create(&c, dev);
        if (rc != CD_OK) {
            g_multi_create_error = std::string("cd_multi_create: device ") + std::to_string(dev) + ": " + cd_last_error(nullptr);
            cd_multi_destroy(m);
            return rc;
        }
        m->ctx.push_back(c);
        m->device.push_back(dev);
    }
    if (n_gpus > 1) {
        char id[128];
        int rc = cd_comm_unique_id(m->ctx[0], id);
        if (rc == CD_OK) rc = m->each([&](int k) { return cd_comm_init(m->ctx[(size_t)k], n_gpus, k, id); });
        else m->err = cd_last_error(m->ctx[0]);
        if (rc != CD_OK) {
            g_multi_create_error = "cd_multi_create: " + m->err;
            cd_multi_destroy(m);
            return rc;
        }
    }
    *out = m;
    return CD_OK;
}

int cd_multi_gpus(const cd_multi* m) { return m ? (int)m->ctx.size() : 0; }

int cd_multi_set_design(cd_multi* m, int S, int p, const double* X)
{
    if (!m) return CD_EINVAL;
    const int rc = m->each([&](int k) { return cd_set_design(m->ctx[(size_t)k], S, p, X); });
    if (rc == CD_OK) { m->S = S; m->p = p; m->have_regions = false; }
    return rc;
}

int cd_multi_set_regions(cd_multi* m, int64_t n, const int64_t* row_off, const int32_t* region_bait)
{
    if (!m) return CD_EINVAL;
    if (n < 1 || !row_off || !region_bait) return m->fail(CD_EINVAL, "cd_multi_set_regions: bad arguments");
    const int G = (int)m->ctx.size();
    m->bounds.assign((size_t)G + 1, 0);
    int rc = cd_plan_shards(n, region_bait, row_off, G, m->bounds.data());
    if (rc != CD_OK) return m->fail(rc, "cd_multi_set_regions: cd_plan_shards refused the regions");
    m->row_lo.assign((size_t)G + 1, 0);
    for (int k = 0; k <= G; k++) m->row_lo[(size_t)k] = row_off[m->bounds[(size_t)k]];
    rc = m->each([&](int k) {
        const int64_t lo = m->bounds[(size_t)k], hi = m->bounds[(size_t)k + 1];
        std::vector<int64_t> off((size_t)(hi - lo) + 1);
        for (int64_t i = lo; i <= hi; i++) off[(size_t)(i - lo)] = row_off[i] - row_off[lo];
        return cd_set_regions(m->ctx[(size_t)k], hi - lo, off.data());
    });
    if (rc == CD_OK) { m->n = n; m->R = row_off[n]; m->have_regions = true; }
    return rc;
}

int cd_multi_get_shards(const cd_multi* m, int64_t* bounds)
{
    if (!m || !bounds || !m->have_regions) return CD_EINVAL;
    for (size_t k = 0; k < m->bounds.size(); k++) bounds[k] = m->bounds[k];
    return CD_OK;
}

int cd_multi_set_sample_rows(cd_multi* m, int s, int64_t R, const int32_t* N, const double* fullmean)
{
    if (!m) return CD_EINVAL;
    if (!m->have_regions || R != m->R || !N || !fullmean) return m->fail(CD_EINVAL, "cd_multi_set_sample_rows: row count does not match the regions");
    return m->each([&](int k) {
        const int64_t r0 = m->row_lo[(size_t)k], r1 = m->row_lo[(size_t)k + 1];
        return cd_set_sample_rows(m->ctx[(size_t)k], s, r1 - r0, N + r0, fullmean + r0);
    });
}

int cd_multi_aggregate(cd_multi* m, int32_t* K_out, double* fullmean_out)
{
    if (!m) return CD_EINVAL;
    if (!m->have_regions) return m->fail(CD_EINVAL, "cd_multi_aggregate: call cd_multi_set_regions first");
    const int S = m->S;
    return m->each([&](int k) {
        const int64_t lo = m->bounds[(size_t)k], nk = m->bounds[(size_t)k + 1] - lo;
        std::vector<int32_t> Kk(K_out ? (size_t)S * (size_t)nk : 0);
        std::vector<double> Fk(fullmean_out ? (size_t)S * (size_t)nk : 0);
        const int rc = cd_aggregate(m->ctx[(size_t)k], K_out ? Kk.data() : nullptr, fullmean_out ? Fk.data() : nullptr);
        if (rc != CD_OK) return rc;
        if (K_out) scatter_matrix(K_out, Kk.data(), S, m->n, lo, nk);
        if (fullmean_out) scatter_matrix(fullmean_out, Fk.data(), S, m->n, lo, nk);
        return CD_OK;
    });
}

int cd_multi_region_test(cd_multi* m, const cd_options* opt, cd_results* out)
{
    if (!m) return CD_EINVAL;
    if (!opt || !out) return m->fail(CD_EINVAL, "cd_multi_region_test: null options / results");
    if (!m->have_regions) return m->fail(CD_EINVAL, "cd_multi_region_test: call cd_multi_set_regions first");
    const int G = (int)m->ctx.size(), S = m->S, p = m->p;
    std::vector<cd_results> res((size_t)G);
    const int rc = m->each([&](int k) {
        const int64_t lo = m->bounds[(size_t)k], nk = m->bounds[(size_t)k + 1] - lo;
        cd_results& r = res[(size_t)k];
        memset(&r, 0, sizeof(r));
        // per-region vectors: shards are contiguous region ranges, so the shard writes straight into the caller's arrays
#define CD_VEC(f) r.f = out->f ? out->f + lo : nullptr
        CD_VEC(baseMean); CD_VEC(baseVar); CD_VEC(dispGeneEst); CD_VEC(dispFit); CD_VEC(dispMAP); CD_VEC(dispersion);
        CD_VEC(log2FoldChange); CD_VEC(lfcSE); CD_VEC(stat); CD_VEC(pvalue); CD_VEC(deviance); CD_VEC(maxCooks);
        CD_VEC(dispGeneIter); CD_VEC(dispIter); CD_VEC(betaIter); CD_VEC(flags);
#undef CD_VEC
        // matrices (rows x n, sample-major): through a shard-sized buffer
        std::vector<double> beta(out->beta ? (size_t)p * (size_t)nk : 0), betaSE(out->betaSE ? (size_t)p * (size_t)nk : 0);
        std::vector<double> nf(out->normFactors ? (size_t)S * (size_t)nk : 0), mu(out->mu ? (size_t)S * (size_t)nk : 0);
        r.beta = out->beta ? beta.data() : nullptr; r.betaSE = out->betaSE ? betaSE.data() : nullptr;
        r.normFactors = out->normFactors ? nf.data() : nullptr; r.mu = out->mu ? mu.data() : nullptr;
        const int rc1 = cd_region_test(m->ctx[(size_t)k], opt, &r);
        if (rc1 != CD_OK) return rc1;
        if (out->beta) scatter_matrix(out->beta, beta.data(), p, m->n, lo, nk);
        if (out->betaSE) scatter_matrix(out->betaSE, betaSE.data(), p, m->n, lo, nk);
        if (out->normFactors) scatter_matrix(out->normFactors, nf.data(), S, m->n, lo, nk);
        if (out->mu) scatter_matrix(out->mu, mu.data(), S, m->n, lo, nk);
        return CD_OK;
    });
    if (rc != CD_OK) return rc;
    // the results of the global steps are identical on every shard; the counters add up
    const cd_results& r0 = res[0];
    memcpy(out->sizeFactors, r0.sizeFactors, sizeof(out->sizeFactors));
    memcpy(out->deviances, r0.deviances, sizeof(out->deviances));
    out->theta = r0.theta; out->n_deviances = r0.n_deviances;
    out->trend_a0 = r0.trend_a0; out->trend_a1 = r0.trend_a1;
    out->varLogDispEsts = r0.varLogDispEsts; out->dispPriorVar = r0.dispPriorVar;
    out->n_nonzero = out->n_gene_grid = out->n_map_grid = out->n_beta_noconv = 0;
    for (const cd_results& r : res) {
        out->n_nonzero += r.n_nonzero; out->n_gene_grid += r.n_gene_grid;
        out->n_map_grid += r.n_map_grid; out->n_beta_noconv += r.n_beta_noconv;
    }
    return CD_OK;
}

int cd_multi_last_timings(const cd_multi* m, double out_ms[8])
{
    if (!m || !out_ms) return CD_EINVAL;
    for (int k = 0; k < 8; k++) out_ms[k] = 0.0;
    for (cd_ctx* c : m->ctx) {
        double t[8];
        if (cd_last_timings(c, t) != CD_OK) return CD_EINVAL;
        for (int k = 0; k < 8; k++) out_ms[k] = std::fmax(out_ms[k], t[k]);      // the slowest device
    }
    return CD_OK;
}

}  // extern "C"
