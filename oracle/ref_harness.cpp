// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" harness around the UNMODIFIED reference engine: it #includes the
// vendored hnswlib headers where they lie under /root/reference (include path only,
// nothing is copied) and exposes them to ctypes so that tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs can run the real reference on the
// same inputs as the CUDA path.  Built by oracle/Makefile into oracle/_ref/ (git-ignored,
// travels to the GPU box as a binary).  Compile flags mirror the reference build
// (/root/reference/build/compile_commands.json:4: -O3 -DNDEBUG -std=gnu++20, no -march).
//
// Reference entry points exercised (file:line under /root/reference):
//   hnswlib/hnswalg.h:89-144   HierarchicalNSW build ctor
//   hnswlib/hnswalg.h:78-86    HierarchicalNSW load ctor -> loadIndex :716-822
//   hnswlib/hnswalg.h:954      addPoint
//   hnswlib/hnswalg.h:1270     searchKnn
//   hnswlib/hnswalg.h:685      saveIndex
//   hnswlib/bruteforce.h:106   BruteforceSearch::searchKnn
//   hnswlib/space_l2.h:207     L2Space, hnswlib/space_ip.h:343 InnerProductSpace
//   index_builder/build.cpp:124-138  data generator (mt19937_64(123) + normal_distribution<float>)
#include "hnswlib/hnswlib.h"

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local long t_dist_calls = 0;
thread_local std::string t_err;

struct CountCtx {
    hnswlib::DISTFUNC<float> fn;
    void *param;
    size_t dim;
};

float counting_dist(const void *a, const void *b, const void *ctxv) {
    const CountCtx *c = (const CountCtx *)ctxv;
    ++t_dist_calls;
    return c->fn(a, b, c->param);
}

// A SpaceInterface whose distance function is the reference one behind a call counter
// (the recipe of SURVEY.md appendix B): D = number of distance evaluations per query.
class CountingSpace : public hnswlib::SpaceInterface<float> {
 public:
    std::unique_ptr<hnswlib::SpaceInterface<float>> inner;
    CountCtx ctx;
    explicit CountingSpace(hnswlib::SpaceInterface<float> *s) : inner(s) {
        ctx.fn = s->get_dist_func();
        ctx.param = s->get_dist_func_param();
        ctx.dim = *(size_t *)ctx.param;
    }
    size_t get_data_size() override { return inner->get_data_size(); }
    hnswlib::DISTFUNC<float> get_dist_func() override { return counting_dist; }
    void *get_dist_func_param() override { return &ctx; }
};

hnswlib::SpaceInterface<float> *make_space(int metric, size_t dim) {
    if (metric == 0) return new hnswlib::L2Space(dim);
    return new hnswlib::InnerProductSpace(dim);
}

struct RefIndex {
    std::unique_ptr<hnswlib::SpaceInterface<float>> space;
    std::unique_ptr<hnswlib::HierarchicalNSW<float>> alg;
    size_t dim;
    bool counting;
};

struct RefBF {
    std::unique_ptr<hnswlib::SpaceInterface<float>> space;
    std::unique_ptr<hnswlib::BruteforceSearch<float>> alg;
    size_t dim;
};

template <class F>
void parallel_for(size_t n, int threads, F f) {
    if (threads <= 1) {
        for (size_t i = 0; i < n; i++) f(i, 0);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    std::atomic<bool> failed{false};
    std::string err;
    std::mutex errm;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&, t] {
            try {
                for (;;) {
                    size_t i = next.fetch_add(1);
                    if (i >= n || failed.load()) break;
                    f(i, t);
                }
            } catch (const std::exception &e) {
                std::lock_guard<std::mutex> g(errm);
                err = e.what();
                failed = true;
            }
        });
    }
    for (auto &th : pool) th.join();
    if (failed) throw std::runtime_error(err);
}

template <class F>
int guarded(F f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        t_err = e.what();
        return -1;
    }
}

}  // namespace

extern "C" {

const char *ref_last_error() { return t_err.c_str(); }

// "sse" / "avx" / "avx512": which distance kernels this build of the reference dispatches to
// (hnswlib/hnswlib.h:11-21, space_l2.h:214-225).
const char *ref_simd_level() {
#if defined(USE_AVX512)
    if (AVX512Capable()) return "avx512";
#endif
#if defined(USE_AVX)
    if (AVXCapable()) return "avx";
#endif
#if defined(USE_SSE)
    return "sse";
#else
    return "scalar";
#endif
}

// Row-major i.i.d. N(0,1) floats from one persistent distribution object, exactly the
// stream index_builder/build.cpp:124-138 and test.cpp:12-16 draw.
void ref_gen_gaussian(uint64_t seed, size_t n, size_t d, float *out) {
    std::mt19937_64 rng(seed);
    std::normal_distribution<float> nd(0.0f, 1.0f);
    for (size_t i = 0; i < n * d; i++) out[i] = nd(rng);
}

float ref_dist(int metric, size_t dim, const float *a, const float *b) {
    std::unique_ptr<hnswlib::SpaceInterface<float>> s(make_space(metric, dim));
    return s->get_dist_func()(a, b, s->get_dist_func_param());
}

void *ref_hnsw_new(int metric, size_t dim, size_t max_elements, size_t M, size_t efc, size_t seed,
                   int counting) {
    RefIndex *r = nullptr;
    int rc = guarded([&] {
        r = new RefIndex();
        r->dim = dim;
        r->counting = counting != 0;
        hnswlib::SpaceInterface<float> *s = make_space(metric, dim);
        r->space.reset(counting ? new CountingSpace(s) : s);
        r->alg.reset(new hnswlib::HierarchicalNSW<float>(r->space.get(), max_elements, M, efc, seed));
    });
    if (rc) { delete r; return nullptr; }
    return r;
}

void *ref_hnsw_load(int metric, size_t dim, const char *path, size_t max_elements, int counting) {
    RefIndex *r = nullptr;
    int rc = guarded([&] {
        r = new RefIndex();
        r->dim = dim;
        r->counting = counting != 0;
        hnswlib::SpaceInterface<float> *s = make_space(metric, dim);
        r->space.reset(counting ? new CountingSpace(s) : s);
        r->alg.reset(new hnswlib::HierarchicalNSW<float>(r->space.get(), std::string(path), false,
                                                        max_elements));
    });
    if (rc) { delete r; return nullptr; }
    return r;
}

void ref_hnsw_free(void *h) { delete (RefIndex *)h; }

// threads<=1: serial insertion in label order, the build.cpp:137-145 pattern (bit-reproducible).
// threads>1 : first point alone, then T threads pulling from an atomic counter (legal per the
//             locking in hnswalg.h:1158-1198; graph is not bit-reproducible).
int ref_hnsw_add(void *h, const float *X, const uint64_t *labels, size_t n, int threads,
                 double *seconds) {
    RefIndex *r = (RefIndex *)h;
    return guarded([&] {
        auto t0 = std::chrono::steady_clock::now();
        size_t start = 0;
        if (threads > 1 && r->alg->cur_element_count == 0 && n > 0) {
            r->alg->addPoint(X, labels ? labels[0] : 0);
            start = 1;
        }
        parallel_for(n - start, threads, [&](size_t j, int) {
            size_t i = j + start;
            r->alg->addPoint(X + i * r->dim, labels ? labels[i] : (hnswlib::labeltype)i);
        });
        if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
}

int ref_hnsw_save(void *h, const char *path) {
    RefIndex *r = (RefIndex *)h;
    return guarded([&] { r->alg->saveIndex(path); });
}

// info[0..7] = cur_element_count, max_elements, maxlevel, enterpoint, M, maxM0, ef_construction, size_data_per_element
void ref_hnsw_info(void *h, int64_t *info) {
    RefIndex *r = (RefIndex *)h;
    info[0] = (int64_t)r->alg->cur_element_count.load();
    info[1] = (int64_t)r->alg->max_elements_;
    info[2] = r->alg->maxlevel_;
    info[3] = (int64_t)r->alg->enterpoint_node_;
    info[4] = (int64_t)r->alg->M_;
    info[5] = (int64_t)r->alg->maxM0_;
    info[6] = (int64_t)r->alg->ef_construction_;
    info[7] = (int64_t)r->alg->size_data_per_element_;
}

int ref_hnsw_levels(void *h, int32_t *levels_out) {
    RefIndex *r = (RefIndex *)h;
    size_t n = r->alg->cur_element_count.load();
    for (size_t i = 0; i < n; i++) levels_out[i] = r->alg->element_levels_[i];
    return 0;
}

// Copies the link list of (internal id, level) into out (<= cap ids); returns the count or -1.
int ref_hnsw_links(void *h, uint32_t id, int level, uint32_t *out, int cap) {
    RefIndex *r = (RefIndex *)h;
    if (id >= r->alg->cur_element_count.load() || level > r->alg->element_levels_[id]) return -1;
    hnswlib::linklistsizeint *ll = r->alg->get_linklist_at_level(id, level);
    int cnt = r->alg->getListCount(ll);
    const hnswlib::tableint *d = (const hnswlib::tableint *)(ll + 1);
    for (int j = 0; j < cnt && j < cap; j++) out[j] = d[j];
    return cnt;
}

// Same as ref_hnsw_new with allow_replace_deleted (hnswalg.h:89-99).
void *ref_hnsw_new_replace(int metric, size_t dim, size_t max_elements, size_t M, size_t efc, size_t seed) {
    RefIndex *r = nullptr;
    int rc = guarded([&] {
        r = new RefIndex();
        r->dim = dim;
        r->counting = false;
        r->space.reset(make_space(metric, dim));
        r->alg.reset(new hnswlib::HierarchicalNSW<float>(r->space.get(), max_elements, M, efc, seed, true));
    });
    if (rc) { delete r; return nullptr; }
    return r;
}

// addPoint(data, label, replace_deleted = true), serial (hnswalg.h:954-992)
int ref_hnsw_add_replace_deleted(void *h, const float *X, const uint64_t *labels, size_t n) {
    RefIndex *r = (RefIndex *)h;
    return guarded([&] {
        for (size_t i = 0; i < n; i++) r->alg->addPoint(X + i * r->dim, labels ? labels[i] : (hnswlib::labeltype)i, true);
    });
}

int ref_hnsw_mark_delete(void *h, uint64_t label) {
    RefIndex *r = (RefIndex *)h;
    return guarded([&] { r->alg->markDelete(label); });
}

// Batched searchKnn, the hnsw_service/main.cpp:66-75 call pattern (setEf then searchKnn) over
// T threads striding the query array.  Output rows are closest-first, padded with
// label=UINT64_MAX / dist=+inf when fewer than k are returned.  If dcount != NULL (index
// created with counting=1) it receives per-query distance evaluations D; hops receives the
// upper-layer metric_hops delta only when threads<=1.
int ref_hnsw_search(void *h, const float *Q, size_t nq, size_t k, size_t ef, int threads,
                    uint64_t *labels, float *dists, uint32_t *counts, uint32_t *dcount,
                    uint32_t *hops_up, double *seconds) {
    RefIndex *r = (RefIndex *)h;
    return guarded([&] {
        r->alg->setEf(ef);
        auto t0 = std::chrono::steady_clock::now();
        parallel_for(nq, threads, [&](size_t i, int) {
            long d0 = t_dist_calls;
            long h0 = threads <= 1 ? r->alg->metric_hops.load() : 0;
            auto res = r->alg->searchKnn(Q + i * r->dim, k);
            if (dcount) dcount[i] = (uint32_t)(t_dist_calls - d0);
            if (hops_up && threads <= 1) hops_up[i] = (uint32_t)(r->alg->metric_hops.load() - h0);
            size_t sz = res.size();
            if (counts) counts[i] = (uint32_t)sz;
            for (size_t j = sz; j < k; j++) {
                labels[i * k + j] = UINT64_MAX;
                dists[i * k + j] = std::numeric_limits<float>::infinity();
            }
            while (!res.empty()) {
                --sz;
                labels[i * k + sz] = res.top().second;
                dists[i * k + sz] = res.top().first;
                res.pop();
            }
        });
        if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
}

// ---- BruteforceSearch (bruteforce.h) ----
void *ref_bf_new(int metric, size_t dim, size_t max_elements) {
    RefBF *r = nullptr;
    int rc = guarded([&] {
        r = new RefBF();
        r->dim = dim;
        r->space.reset(make_space(metric, dim));
        r->alg.reset(new hnswlib::BruteforceSearch<float>(r->space.get(), max_elements));
    });
    if (rc) { delete r; return nullptr; }
    return r;
}

void *ref_bf_load(int metric, size_t dim, const char *path) {
    RefBF *r = nullptr;
    int rc = guarded([&] {
        r = new RefBF();
        r->dim = dim;
        r->space.reset(make_space(metric, dim));
        r->alg.reset(new hnswlib::BruteforceSearch<float>(r->space.get(), std::string(path)));
    });
    if (rc) { delete r; return nullptr; }
    return r;
}

void ref_bf_free(void *h) { delete (RefBF *)h; }

int ref_bf_add(void *h, const float *X, const uint64_t *labels, size_t n) {
    RefBF *r = (RefBF *)h;
    return guarded([&] {
        for (size_t i = 0; i < n; i++) r->alg->addPoint(X + i * r->dim, labels ? labels[i] : i);
    });
}

int ref_bf_remove(void *h, uint64_t label) {
    RefBF *r = (RefBF *)h;
    return guarded([&] { r->alg->removePoint(label); });
}

int ref_bf_save(void *h, const char *path) {
    RefBF *r = (RefBF *)h;
    return guarded([&] { r->alg->saveIndex(path); });
}

int64_t ref_bf_count(void *h) { return (int64_t)((RefBF *)h)->alg->cur_element_count; }

int ref_bf_search(void *h, const float *Q, size_t nq, size_t k, int threads, uint64_t *labels,
                  float *dists, uint32_t *counts, double *seconds) {
    RefBF *r = (RefBF *)h;
    return guarded([&] {
        auto t0 = std::chrono::steady_clock::now();
        parallel_for(nq, threads, [&](size_t i, int) {
            auto res = r->alg->searchKnn(Q + i * r->dim, k);
            size_t sz = res.size();
            if (counts) counts[i] = (uint32_t)sz;
            for (size_t j = sz; j < k; j++) {
                labels[i * k + j] = UINT64_MAX;
                dists[i * k + j] = std::numeric_limits<float>::infinity();
            }
            while (!res.empty()) {
                --sz;
                labels[i * k + sz] = res.top().second;
                dists[i * k + sz] = res.top().first;
                res.pop();
            }
        });
        if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
}

}  // extern "C"
