// shim_api.cpp -- exercises, through the drop-in C++ header only, the parts of the vendored hnswlib API that none of the
// reference's three consumers touches (SURVEY.md 8(b) "full vendored API", 8(f) N4): BaseFilterFunctor on both index
// types, updatePoint, markDelete / replace_deleted, stop_condition.h, parallel addPoint / searchKnn.
// Built by research_new_hnsw_b200/hnswlib/Makefile, run by tests/test_consumers_gpu.py on the GPU box.
#include <cstdio>
#include <random>
#include <thread>

#include "hnswlib/hnswlib.h"
#include "hnswlib/stop_condition.h"

#define CHECK(c)                                                       \
    do {                                                               \
        if (!(c)) {                                                    \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                  \
        }                                                              \
    } while (0)

struct EvenOnly : hnswlib::BaseFilterFunctor {
    bool operator()(hnswlib::labeltype id) override { return id % 2 == 0; }
};

int main() {
    const size_t n = 4000, d = 32, k = 10;
    std::mt19937 rng(7);
    std::normal_distribution<float> g(0.f, 1.f);
    std::vector<float> X(n * d);
    for (float &v : X) v = g(rng);
    hnswlib::L2Space space(d);

    // ---- parallel addPoint (label-op locks in the reference, hnswalg.h:40-43), public fields stay consistent
    hnswlib::HierarchicalNSW<float> idx(&space, n, 12, 80, 100, /*allow_replace_deleted=*/true);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < 4; t++)
            th.emplace_back([&, t] {
                for (size_t i = t; i < n; i += 4) idx.addPoint(X.data() + i * d, i);
            });
        for (auto &x : th) x.join();
    }
    CHECK(idx.cur_element_count == n);
    CHECK(idx.element_levels_.size() >= n);
    int maxl = 0;
    for (size_t i = 0; i < n; i++) maxl = std::max(maxl, idx.element_levels_[i]);
    CHECK(maxl == idx.maxlevel_);   // no level lost by the concurrent mirror updates
    idx.setEf(64);

    // ---- parallel searchKnn (const and thread-safe in the reference, hnswalg.h:1270): concurrent one-query calls are
    // coalesced into shared launches and must return exactly what the same calls return one after the other
    {
        const size_t nqs = 800;
        std::vector<std::vector<std::pair<float, hnswlib::labeltype>>> par(nqs);
        std::vector<std::thread> th;
        for (int t = 0; t < 8; t++)
            th.emplace_back([&, t] {
                for (size_t i = t; i < nqs; i += 8) par[i] = idx.searchKnnCloserFirst(X.data() + i * d, k);
            });
        for (auto &x : th) x.join();
        size_t self = 0;
        for (size_t i = 0; i < nqs; i++) {
            auto seq = idx.searchKnnCloserFirst(X.data() + i * d, k);
            CHECK(seq == par[i]);
            CHECK(seq.size() == k);
            self += seq[0].second == i;
        }
        CHECK(self >= nqs * 95 / 100);  // a stored point is (almost always) its own nearest neighbour
    }

    // ---- BaseFilterFunctor on the graph index (hnswalg.h:1270,1306-1313)
    EvenOnly even;
    for (size_t i = 1; i < 50; i += 2) {
        auto r = idx.searchKnnCloserFirst(X.data() + i * d, k, &even);
        CHECK(r.size() == k);
        for (auto &p : r) CHECK(p.second % 2 == 0);
    }

    // ---- BruteforceSearch with the same functor (bruteforce.h:114,121) against a brute-force index of the even rows only
    hnswlib::BruteforceSearch<float> bf(&space, n), bf_even(&space, n);
    for (size_t i = 0; i < n; i++) {
        bf.addPoint(X.data() + i * d, i);
        if (i % 2 == 0) bf_even.addPoint(X.data() + i * d, i);
    }
    for (size_t i = 0; i < 20; i++) {
        auto a = bf.searchKnnCloserFirst(X.data() + i * d, k, &even);
        auto b = bf_even.searchKnnCloserFirst(X.data() + i * d, k);
        CHECK(a.size() == k && a == b);   // ids and distances bit-identical
    }

    // ---- updatePoint (hnswalg.h:995-1139): an element moved onto another one's vector is found there
    {
        std::vector<float> v(X.begin() + 10 * d, X.begin() + 11 * d);
        for (float &x : v) x += 1e-3f;
        const hnswlib::tableint internal = 5;
        const hnswlib::labeltype lab = idx.getExternalLabel(internal);
        idx.updatePoint(v.data(), internal, 1.0f);
        auto r = idx.searchKnnCloserFirst(v.data(), 2);
        CHECK(r.size() == 2 && r[0].second == lab);
        auto back = idx.getDataByLabel<float>(lab);
        CHECK(back.size() == d && back[0] == v[0]);
        CHECK(idx.cur_element_count == n);
    }

    // ---- markDelete / replace_deleted (hnswalg.h:853-883, 954-992)
    idx.markDelete(17);
    CHECK(idx.getDeletedCount() == 1);
    {
        auto r = idx.searchKnnCloserFirst(X.data() + 17 * d, k);
        for (auto &p : r) CHECK(p.second != 17);
    }
    std::vector<float> fresh(d, 0.25f);
    idx.addPoint(fresh.data(), 100000, /*replace_deleted=*/true);
    CHECK(idx.getDeletedCount() == 0 && idx.cur_element_count == n);
    {
        auto r = idx.searchKnnCloserFirst(fresh.data(), 1);
        CHECK(r.size() == 1 && r[0].second == 100000);
    }

    // ---- stop_condition.h: the epsilon condition runs, the multi-vector one has no GPU formulation and says so
    {
        auto nn = idx.searchKnnCloserFirst(X.data() + 3 * d, 20);
        const float eps = nn[9].first;  // distance of the 10th neighbour
        hnswlib::EpsilonSearchStopCondition<float> stop(eps, 5, 50);
        auto r = idx.searchStopConditionClosest(X.data() + 3 * d, stop);
        CHECK(r.size() >= 10 && r.size() <= 50);
        for (auto &p : r) CHECK(p.first <= eps);
        CHECK(r[0].second == nn[0].second);
        hnswlib::MultiVectorL2Space<unsigned> mv(d);
        CHECK(mv.get_data_size() == d * 4 + sizeof(unsigned));
        bool threw = false;
        try {
            hnswlib::HierarchicalNSW<float> bad(&mv, 10);
        } catch (const std::runtime_error &) {
            threw = true;
        }
        CHECK(threw);
        hnswlib::MultiVectorSearchStopCondition<unsigned, float> mstop(mv, 3);
        threw = false;
        try {
            idx.searchStopConditionClosest(X.data(), mstop);
        } catch (const std::runtime_error &) {
            threw = true;
        }
        CHECK(threw);
    }
    idx.checkIntegrity();
    std::printf("shim_api OK\n");
    return 0;
}
