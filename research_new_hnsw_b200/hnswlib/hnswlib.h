// hnswlib/hnswlib.h -- drop-in header for the hnswlib copy vendored in hiozings/Research-New-HNSW.
//
// Same names, same signatures, same public data members as /root/reference/hnswlib/{hnswlib.h,hnswalg.h,
// bruteforce.h,space_l2.h,space_ip.h}, so index_builder/build.cpp, hnsw_service/main.cpp and test.cpp compile
// UNCHANGED against it -- but every engine call forwards through the C ABI (include/b200hnsw.h)
// to hand-written sm_100a kernels in libb200hnsw.so.  Nothing here computes on the host except the DISTFUNC the
// space objects must hand out by contract (hnswlib.h:170-184); the engine itself never calls it, and there is no
// CPU fallback: without a GPU the constructors throw.
//
// Link with:  -L<repo>/research_new_hnsw_b200 -lb200hnsw -Wl,-rpath,<repo>/research_new_hnsw_b200
#pragma once

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <mutex>
#include <queue>
#include <random>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#if __has_include("b200hnsw.h")  // installed next to this header (the drop-in directory is self-contained)
#include "b200hnsw.h"
#else
#include "../../include/b200hnsw.h"
#endif

#ifndef HNSWLIB_ERR_OVERRIDE
#define HNSWERR std::cerr
#else
#define HNSWERR HNSWLIB_ERR_OVERRIDE
#endif

namespace hnswlib {
typedef size_t labeltype;          // hnswlib.h:125
typedef unsigned int tableint;     // hnswalg.h:14
typedef unsigned int linklistsizeint;  // hnswalg.h:15

// hnswlib.h:128-132
class BaseFilterFunctor {
 public:
    virtual bool operator()(hnswlib::labeltype id) { return true; }
    virtual ~BaseFilterFunctor() {}
};

// hnswlib.h:134-150.  Stop conditions are host callbacks; the GPU engine runs the ones that declare a device formulation
// through the non-reference hook at the end (stop_condition.h).
template <typename dist_t>
class BaseSearchStopCondition {
 public:
    virtual void add_point_to_result(labeltype label, const void *datapoint, dist_t dist) = 0;
    virtual void remove_point_from_result(labeltype label, const void *datapoint, dist_t dist) = 0;
    virtual bool should_stop_search(dist_t candidate_dist, dist_t lowerBound) = 0;
    virtual bool should_consider_candidate(dist_t candidate_dist, dist_t lowerBound) = 0;
    virtual bool should_remove_extra() = 0;
    virtual void filter_results(std::vector<std::pair<dist_t, labeltype>> &candidates) = 0;
    // true: "the max_candidates closest elements, cut at distance epsilon" describes this condition's result
    virtual bool b200_epsilon_form(float *epsilon, size_t *max_candidates) const { return false; }
    virtual ~BaseSearchStopCondition() {}
};

template <typename T>
class pairGreater {
 public:
    bool operator()(const T &p1, const T &p2) { return p1.first > p2.first; }
};

template <typename T>
static void writeBinaryPOD(std::ostream &out, const T &podRef) {
    out.write((char *)&podRef, sizeof(T));
}
template <typename T>
static void readBinaryPOD(std::istream &in, T &podRef) {
    in.read((char *)&podRef, sizeof(T));
}

template <typename MTYPE>
using DISTFUNC = MTYPE (*)(const void *, const void *, const void *);  // hnswlib.h:170-171

// hnswlib.h:173-184, plus one non-reference hook: which GPU metric the space stands for.
template <typename MTYPE>
class SpaceInterface {
 public:
    virtual size_t get_data_size() = 0;
    virtual DISTFUNC<MTYPE> get_dist_func() = 0;
    virtual void *get_dist_func_param() = 0;
    // -1: a user-defined space the GPU engine cannot run (it then throws; there is no CPU fallback)
    virtual int b200_metric() const { return -1; }
    virtual ~SpaceInterface() {}
};

// hnswlib.h:186-201
template <typename dist_t>
class AlgorithmInterface {
 public:
    virtual void addPoint(const void *datapoint, labeltype label, bool replace_deleted = false) = 0;
    virtual std::priority_queue<std::pair<dist_t, labeltype>> searchKnn(const void *, size_t,
                                                                        BaseFilterFunctor *isIdAllowed = nullptr) const = 0;
    virtual std::vector<std::pair<dist_t, labeltype>> searchKnnCloserFirst(const void *query_data, size_t k,
                                                                           BaseFilterFunctor *isIdAllowed = nullptr) const;
    virtual void saveIndex(const std::string &location) = 0;
    virtual ~AlgorithmInterface() {}
};

// hnswlib.h:203-225
template <typename dist_t>
std::vector<std::pair<dist_t, labeltype>> AlgorithmInterface<dist_t>::searchKnnCloserFirst(
    const void *query_data, size_t k, BaseFilterFunctor *isIdAllowed) const {
    std::vector<std::pair<dist_t, labeltype>> result;
    auto ret = searchKnn(query_data, k, isIdAllowed);
    size_t sz = ret.size();
    result.resize(sz);
    while (!ret.empty()) {
        result[--sz] = ret.top();
        ret.pop();
    }
    return result;
}

namespace b200detail {
inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(b200hnsw_last_error());
}
// Host DISTFUNCs: the contract of SpaceInterface::get_dist_func (callers may invoke it directly).  Same value as the
// reference's shipped SSE kernels: four lane accumulators over i = l mod 4, separate multiply and add,
// ((T0+T1)+T2)+T3, residual tail summed separately (space_l2.h:97-205, space_ip.h:211-339).
template <bool IP>
inline float lanes4(const float *a, const float *b, size_t n4) {
    volatile float s[4] = {0, 0, 0, 0};  // volatile: forbid FMA contraction / reassociation
    for (size_t i = 0; i < n4; i += 4)
        for (int l = 0; l < 4; l++) {
            float m;
            if (IP) {
                m = a[i + l] * b[i + l];
            } else {
                float t = a[i + l] - b[i + l];
                m = t * t;
            }
            volatile float mv = m;
            s[l] = s[l] + mv;
        }
    volatile float r = s[0] + s[1];
    r = r + s[2];
    r = r + s[3];
    return r;
}
template <bool IP>
inline float scalar_tail(const float *a, const float *b, size_t n) {
    volatile float res = 0;
    for (size_t i = 0; i < n; i++) {
        float m;
        if (IP) {
            m = a[i] * b[i];
        } else {
            float t = a[i] - b[i];
            m = t * t;
        }
        volatile float mv = m;
        res = res + mv;
    }
    return res;
}
template <bool IP>
inline float host_dist(const void *av, const void *bv, const void *qty_ptr) {
    const float *a = (const float *)av, *b = (const float *)bv;
    const size_t d = *(const size_t *)qty_ptr;
    float r;
    if (d % 4 == 0) {
        r = lanes4<IP>(a, b, d);
    } else if (d > 16) {
        const size_t q = d >> 4 << 4;
        volatile float m = lanes4<IP>(a, b, q), t = scalar_tail<IP>(a + q, b + q, d - q);
        r = m + t;
    } else if (d > 4) {
        const size_t q = d >> 2 << 2;
        volatile float m = lanes4<IP>(a, b, q), t = scalar_tail<IP>(a + q, b + q, d - q);
        r = m + t;
    } else {
        r = scalar_tail<IP>(a, b, d);
    }
    return IP ? 1.0f - r : r;
}
}  // namespace b200detail

// space_l2.h:207-253
class L2Space : public SpaceInterface<float> {
    DISTFUNC<float> fstdistfunc_;
    size_t data_size_;
    size_t dim_;

 public:
    L2Space(size_t dim) {
        fstdistfunc_ = b200detail::host_dist<false>;
        dim_ = dim;
        data_size_ = dim * sizeof(float);
    }
    size_t get_data_size() { return data_size_; }
    DISTFUNC<float> get_dist_func() { return fstdistfunc_; }
    void *get_dist_func_param() { return &dim_; }
    int b200_metric() const { return B200HNSW_L2; }
    ~L2Space() {}
};

// space_ip.h:343-398
class InnerProductSpace : public SpaceInterface<float> {
    DISTFUNC<float> fstdistfunc_;
    size_t data_size_;
    size_t dim_;

 public:
    InnerProductSpace(size_t dim) {
        fstdistfunc_ = b200detail::host_dist<true>;
        dim_ = dim;
        data_size_ = dim * sizeof(float);
    }
    size_t get_data_size() { return data_size_; }
    DISTFUNC<float> get_dist_func() { return fstdistfunc_; }
    void *get_dist_func_param() { return &dim_; }
    int b200_metric() const { return B200HNSW_IP; }
    ~InnerProductSpace() {}
};

namespace b200detail {
inline b200hnsw_params make_params(SpaceInterface<float> *s, size_t max_elements, size_t M, size_t efc, size_t seed,
                                   bool allow_replace_deleted) {
    if (!s) throw std::runtime_error("space is null");
    const int metric = s->b200_metric();
    if (metric < 0)
        throw std::runtime_error(
            "user-defined SpaceInterface cannot run on the GPU engine (only L2Space / InnerProductSpace; no CPU fallback)");
    b200hnsw_params p;
    memset(&p, 0, sizeof(p));
    p.metric = metric;
    p.storage = B200HNSW_F32;
    // the reference API has no storage knob: B200HNSW_STORAGE=bf16 selects the bf16 traversal + fp32 re-rank variant
    if (const char *e = getenv("B200HNSW_STORAGE"))
        if (!strcmp(e, "bf16") || !strcmp(e, "BF16")) p.storage = B200HNSW_BF16;
    p.device = -1;
    p.allow_replace_deleted = allow_replace_deleted ? 1 : 0;
    p.dim = *(size_t *)s->get_dist_func_param();
    p.max_elements = max_elements;
    p.M = M;
    p.ef_construction = efc;
    p.random_seed = seed;
    return p;
}
}  // namespace b200detail

// bruteforce.h:10-172
template <typename dist_t>
class BruteforceSearch : public AlgorithmInterface<dist_t> {
    static_assert(std::is_same<dist_t, float>::value, "the GPU engine implements BruteforceSearch<float>");
    b200bf_index *h_ = nullptr;

 public:
    size_t maxelements_ = 0;
    size_t cur_element_count = 0;
    size_t size_per_element_ = 0;
    size_t data_size_ = 0;
    DISTFUNC<dist_t> fstdistfunc_ = nullptr;
    void *dist_func_param_ = nullptr;

    BruteforceSearch(SpaceInterface<dist_t> *s) {}
    BruteforceSearch(SpaceInterface<dist_t> *s, const std::string &location) { loadIndex(location, s); }
    BruteforceSearch(SpaceInterface<dist_t> *s, size_t maxElements) {
        set_space(s);
        maxelements_ = maxElements;
        b200hnsw_params p = b200detail::make_params(s, maxElements, 0, 0, 0, false);
        b200detail::check(b200bf_create(&p, &h_));
    }
    ~BruteforceSearch() { b200bf_destroy(h_); }
    BruteforceSearch(const BruteforceSearch &) = delete;
    BruteforceSearch &operator=(const BruteforceSearch &) = delete;

    void addPoint(const void *datapoint, labeltype label, bool replace_deleted = false) {
        uint64_t lab = label;
        b200detail::check(b200bf_add_batch(h_, (const float *)datapoint, &lab, 1));
        sync();
    }
    // batched extension (not in the reference): n rows in one call
    void addPoints(const float *X, const labeltype *labels, size_t n) {
        static_assert(sizeof(labeltype) == sizeof(uint64_t), "labeltype must be 64-bit");
        b200detail::check(b200bf_add_batch(h_, X, (const uint64_t *)labels, n));
        sync();
    }
    void removePoint(labeltype cur_external) {
        b200detail::check(b200bf_remove(h_, cur_external));
        sync();
    }
    std::priority_queue<std::pair<dist_t, labeltype>> searchKnn(const void *query_data, size_t k,
                                                                BaseFilterFunctor *isIdAllowed = nullptr) const {
        std::priority_queue<std::pair<dist_t, labeltype>> res;
        if (cur_element_count == 0 || k == 0) return res;
        std::vector<uint64_t> labels(k);
        std::vector<float> dists(k);
        uint32_t cnt = 0;
        if (isIdAllowed) {
            // bruteforce.h:114,121: rows the functor rejects are skipped.  The functor is a host callback: it is evaluated
            // once per stored row, the scan kernels take the verdicts as a row mask.
            const size_t n = cur_element_count;
            std::vector<uint64_t> all(n);
            b200detail::check(b200bf_get_labels(h_, all.data(), n));
            std::vector<uint8_t> allowed(n);
            for (size_t i = 0; i < n; i++) allowed[i] = (*isIdAllowed)((labeltype)all[i]) ? 1 : 0;
            b200detail::check(b200bf_search_batch_filtered(h_, (const float *)query_data, 1, k, allowed.data(), labels.data(),
                                                           dists.data(), &cnt));
        } else {
            b200detail::check(b200bf_search_batch(h_, (const float *)query_data, 1, k, labels.data(), dists.data(), &cnt));
        }
        for (uint32_t j = 0; j < cnt; j++) res.emplace(dists[j], (labeltype)labels[j]);
        return res;
    }
    // batched extension: rows closest-first, padded with label = SIZE_MAX
    void searchKnnBatch(const float *Q, size_t nq, size_t k, labeltype *labels_out, dist_t *dists_out,
                        uint32_t *counts_out = nullptr) const {
        b200detail::check(b200bf_search_batch(h_, Q, nq, k, (uint64_t *)labels_out, dists_out, counts_out));
    }
    void saveIndex(const std::string &location) { b200detail::check(b200bf_save(h_, location.c_str())); }
    void loadIndex(const std::string &location, SpaceInterface<dist_t> *s) {
        set_space(s);
        b200bf_destroy(h_);
        h_ = nullptr;
        b200hnsw_params p = b200detail::make_params(s, 0, 0, 0, 0, false);
        b200detail::check(b200bf_load(location.c_str(), &p, &h_));
        sync();
    }

 private:
    void set_space(SpaceInterface<dist_t> *s) {
        data_size_ = s->get_data_size();
        fstdistfunc_ = s->get_dist_func();
        dist_func_param_ = s->get_dist_func_param();
        size_per_element_ = data_size_ + sizeof(labeltype);
    }
    void sync() {
        uint64_t c = 0;
        b200detail::check(b200bf_count(h_, &c));
        cur_element_count = c;
    }
};

// hnswalg.h:17-1411
template <typename dist_t>
class HierarchicalNSW : public AlgorithmInterface<dist_t> {
    static_assert(std::is_same<dist_t, float>::value, "the GPU engine implements HierarchicalNSW<float>");
    b200hnsw_index *h_ = nullptr;

 public:
    static const tableint MAX_LABEL_OPERATION_LOCKS = 65536;
    static const unsigned char DELETE_MARK = 0x01;

    // ---- public data members of the reference class (hnswalg.h:23-71); refreshed after every mutating call ----
    size_t max_elements_{0};
    mutable std::atomic<size_t> cur_element_count{0};
    size_t size_data_per_element_{0};
    size_t size_links_per_element_{0};
    mutable std::atomic<size_t> num_deleted_{0};
    size_t M_{0};
    size_t maxM_{0};
    size_t maxM0_{0};
    size_t ef_construction_{0};
    size_t ef_{0};
    double mult_{0.0}, revSize_{0.0};
    int maxlevel_{0};
    tableint enterpoint_node_{0};
    size_t size_links_level0_{0};
    size_t offsetData_{0}, offsetLevel0_{0}, label_offset_{0};
    std::vector<int> element_levels_;
    size_t data_size_{0};
    DISTFUNC<dist_t> fstdistfunc_ = nullptr;
    void *dist_func_param_{nullptr};
    mutable std::atomic<long> metric_distance_computations{0};
    mutable std::atomic<long> metric_hops{0};
    bool allow_replace_deleted_ = false;

    HierarchicalNSW(SpaceInterface<dist_t> *s) {}

    HierarchicalNSW(SpaceInterface<dist_t> *s, const std::string &location, bool nmslib = false, size_t max_elements = 0,
                    bool allow_replace_deleted = false)
        : allow_replace_deleted_(allow_replace_deleted) {
        loadIndex(location, s, max_elements);
    }

    HierarchicalNSW(SpaceInterface<dist_t> *s, size_t max_elements, size_t M = 16, size_t ef_construction = 200,
                    size_t random_seed = 100, bool allow_replace_deleted = false)
        : allow_replace_deleted_(allow_replace_deleted) {
        if (M > 10000) {
            HNSWERR << "warning: M parameter exceeds 10000 which may lead to adverse effects." << std::endl;
            HNSWERR << "         Cap to 10000 will be applied for the rest of the processing." << std::endl;
        }
        set_space(s);
        b200hnsw_params p = b200detail::make_params(s, max_elements, M, ef_construction, random_seed, allow_replace_deleted);
        b200detail::check(b200hnsw_create(&p, &h_));
        element_levels_.assign(max_elements, 0);
        sync_fields();
        // initializations for special treatment of the first node (hnswalg.h:134-136)
        enterpoint_node_ = (tableint)-1;
        maxlevel_ = -1;
    }

    ~HierarchicalNSW() { clear(); }
    HierarchicalNSW(const HierarchicalNSW &) = delete;
    HierarchicalNSW &operator=(const HierarchicalNSW &) = delete;

    void clear() {
        b200hnsw_destroy(h_);
        h_ = nullptr;
        cur_element_count = 0;
    }

    struct CompareByFirst {
        constexpr bool operator()(std::pair<dist_t, tableint> const &a, std::pair<dist_t, tableint> const &b) const noexcept {
            return a.first < b.first;
        }
    };

    void setEf(size_t ef) {
        ef_ = ef;
        b200detail::check(b200hnsw_set_ef(h_, ef));
    }

    inline labeltype getExternalLabel(tableint internal_id) const {
        uint64_t l = 0;
        b200detail::check(b200hnsw_get_label(h_, internal_id, &l));
        return (labeltype)l;
    }

    inline char *getDataByInternalId(tableint internal_id) const {
        const float *v = nullptr;
        b200detail::check(b200hnsw_get_data(h_, internal_id, &v));
        return (char *)v;
    }

    size_t getMaxElements() { return max_elements_; }
    size_t getCurrentElementCount() { return cur_element_count; }
    size_t getDeletedCount() { return num_deleted_; }

    linklistsizeint *get_linklist0(tableint internal_id) const { return get_linklist_at_level(internal_id, 0); }
    linklistsizeint *get_linklist(tableint internal_id, int level) const { return get_linklist_at_level(internal_id, level); }
    // Pointer into the reference-layout host mirror (hnswalg.h:486-503): u16 count in the low half-word, neighbour
    // ids from ptr + 1.  Links any staged insertions first.
    linklistsizeint *get_linklist_at_level(tableint internal_id, int level) const {
        const uint32_t *p = nullptr;
        b200detail::check(b200hnsw_get_linklist(h_, internal_id, level, &p));
        return (linklistsizeint *)p;
    }
    unsigned short int getListCount(linklistsizeint *ptr) const { return *((unsigned short int *)ptr); }

    bool isMarkedDeleted(tableint internalId) const {
        unsigned char *ll_cur = ((unsigned char *)get_linklist0(internalId)) + 2;
        return *ll_cur & DELETE_MARK;
    }

    void resizeIndex(size_t new_max_elements) {
        b200detail::check(b200hnsw_resize(h_, new_max_elements));
        element_levels_.resize(new_max_elements);
        sync_fields();
    }

    size_t indexFileSize() const {
        uint64_t b = 0;
        b200detail::check(b200hnsw_index_file_size(h_, &b));
        return (size_t)b;
    }

    void saveIndex(const std::string &location) { b200detail::check(b200hnsw_save(h_, location.c_str())); }

    void loadIndex(const std::string &location, SpaceInterface<dist_t> *s, size_t max_elements_i = 0) {
        set_space(s);
        b200hnsw_destroy(h_);
        h_ = nullptr;
        b200hnsw_params p = b200detail::make_params(s, max_elements_i, 0, 0, 0, allow_replace_deleted_);
        b200detail::check(b200hnsw_load(location.c_str(), &p, &h_));
        sync_fields();
        const int32_t *lv = nullptr;
        b200detail::check(b200hnsw_get_levels(h_, &lv));
        element_levels_.assign(max_elements_, 0);
        for (size_t i = 0; i < cur_element_count; i++) element_levels_[i] = lv[i];
        levels_mirrored_ = cur_element_count;
    }

    template <typename data_t>
    std::vector<data_t> getDataByLabel(labeltype label) const {
        std::vector<float> tmp(data_size_ / sizeof(float));
        b200detail::check(b200hnsw_get_data_by_label(h_, label, tmp.data()));
        return std::vector<data_t>(tmp.begin(), tmp.end());
    }

    void markDelete(labeltype label) {
        b200detail::check(b200hnsw_mark_delete(h_, label));
        sync_fields();
    }
    void unmarkDelete(labeltype label) {
        b200detail::check(b200hnsw_unmark_delete(h_, label));
        sync_fields();
    }

    // addPoint (hnswalg.h:954-992): the point is staged; cur_element_count, element_levels_, enterpoint_node_ and
    // maxlevel_ are updated immediately (they depend only on the level generator), graph links are built on the GPU
    // in batches at the next searchKnn / saveIndex / get_linklist* / flush().
    void addPoint(const void *data_point, labeltype label, bool replace_deleted = false) {
        uint64_t lab = label;
        if (replace_deleted) {
            if (!allow_replace_deleted_)
                throw std::runtime_error("Replacement of deleted elements is disabled in constructor");
            b200detail::check(b200hnsw_add_batch_replace_deleted(h_, (const float *)data_point, &lab, 1));
        } else {
            b200detail::check(b200hnsw_add_batch(h_, (const float *)data_point, &lab, 1));
        }
        after_add(1);
    }
    // batched extension (not in the reference): n rows in one call
    void addPoints(const float *X, const labeltype *labels, size_t n) {
        b200detail::check(b200hnsw_add_batch(h_, X, (const uint64_t *)labels, n));
        after_add(n);
    }
    void flush() { b200detail::check(b200hnsw_flush(h_)); }

    std::priority_queue<std::pair<dist_t, labeltype>> searchKnn(const void *query_data, size_t k,
                                                                BaseFilterFunctor *isIdAllowed = nullptr) const {
        std::priority_queue<std::pair<dist_t, labeltype>> result;
        if (cur_element_count == 0 || k == 0) return result;
        std::vector<uint64_t> labels(k);
        std::vector<float> dists(k);
        uint32_t cnt = 0;
        if (isIdAllowed) {
            // the functor is a host callback: evaluate it once per stored label, the kernel treats "not allowed" like a
            // delete mark (hnswalg.h:406-407)
            // (the label of every internal id is cached until the next mutating call; the library skips the upload of a
            // mask identical to the one it already holds, so repeated calls with the same functor pay only the callbacks)
            std::vector<uint8_t> allowed;
            {
                std::lock_guard<std::mutex> g(fields_mu_);
                const size_t n = cur_element_count;
                if (filter_labels_.size() != n) {
                    filter_labels_.resize(n);
                    b200detail::check(b200hnsw_get_labels(h_, filter_labels_.data(), n));
                }
                allowed.resize(n);
                for (size_t i = 0; i < n; i++) allowed[i] = (*isIdAllowed)((labeltype)filter_labels_[i]) ? 1 : 0;
            }
            b200detail::check(b200hnsw_search_batch_filtered(h_, (const float *)query_data, 1, k, 0, allowed.data(),
                                                             labels.data(), dists.data(), &cnt));
            for (uint32_t j = 0; j < cnt; j++) result.emplace(dists[j], (labeltype)labels[j]);
            return result;
        }
        uint32_t work[4] = {0, 0, 0, 0};
        b200detail::check(b200hnsw_search_batch(h_, (const float *)query_data, 1, k, 0, labels.data(), dists.data(), &cnt, work));
        metric_distance_computations += work[0];
        metric_hops += work[1] + work[2];
        for (uint32_t j = 0; j < cnt; j++) result.emplace(dists[j], (labeltype)labels[j]);
        return result;
    }
    // hnswalg.h:1327-1378.  Only stop conditions with a device formulation run (stop_condition.h); the others throw.
    std::vector<std::pair<dist_t, labeltype>> searchStopConditionClosest(const void *query_data,
                                                                         BaseSearchStopCondition<dist_t> &stop_condition,
                                                                         BaseFilterFunctor *isIdAllowed = nullptr) const {
        std::vector<std::pair<dist_t, labeltype>> result;
        if (cur_element_count == 0) return result;
        float epsilon = 0.f;
        size_t kmax = 0;
        if (!stop_condition.b200_epsilon_form(&epsilon, &kmax))
            throw std::runtime_error(
                "searchStopConditionClosest: this stop condition is a host callback without a GPU formulation (only "
                "EpsilonSearchStopCondition is supported; there is no CPU fallback)");
        if (kmax == 0) return result;
        auto top = [&]() {
            // ef = k = max_num_candidates; searchKnn uses max(ef, k)
            return isIdAllowed ? searchKnn(query_data, kmax, isIdAllowed) : searchKnn(query_data, kmax);
        }();
        result.resize(top.size());
        size_t sz = top.size();
        while (!top.empty()) {  // closest first, as the reference returns (:1369-1376)
            result[--sz] = top.top();
            top.pop();
        }
        stop_condition.filter_results(result);
        return result;
    }

    // hnswalg.h:995-1072: new vector for an element that is already stored, both phases on the GPU (the
    // updateNeighborProbability of the reference is 1.0 from every caller in the repository and is taken as 1.0 here).
    void updatePoint(const void *dataPoint, tableint internalId, float updateNeighborProbability) {
        (void)updateNeighborProbability;
        uint64_t lab = getExternalLabel(internalId);
        b200detail::check(b200hnsw_add_batch(h_, (const float *)dataPoint, &lab, 1));
        after_add(1);
    }

    // batched extension: one kernel launch for nq queries; rows closest-first, padded with label = SIZE_MAX
    void searchKnnBatch(const float *Q, size_t nq, size_t k, labeltype *labels_out, dist_t *dists_out,
                        uint32_t *counts_out = nullptr, size_t ef = 0) const {
        b200detail::check(b200hnsw_search_batch(h_, Q, nq, k, ef, (uint64_t *)labels_out, dists_out, counts_out, nullptr));
    }

    // hnswalg.h:1381-1410 over the host mirror
    void checkIntegrity() {
        int connections_checked = 0;
        const size_t n = cur_element_count;
        std::vector<int> inbound(n, 0);
        for (size_t i = 0; i < n; i++) {
            for (int l = 0; l <= element_levels_[i]; l++) {
                linklistsizeint *ll_cur = get_linklist_at_level((tableint)i, l);
                int size = getListCount(ll_cur);
                tableint *data = (tableint *)(ll_cur + 1);
                std::unordered_set<tableint> s;
                for (int j = 0; j < size; j++) {
                    if (data[j] >= n || data[j] == i) throw std::runtime_error("integrity: bad link");
                    inbound[data[j]]++;
                    s.insert(data[j]);
                    connections_checked++;
                }
                if ((int)s.size() != size) throw std::runtime_error("integrity: duplicate link");
            }
        }
        if (n > 1) {
            int min1 = inbound[0], max1 = inbound[0];
            for (size_t i = 0; i < n; i++) {
                min1 = std::min(inbound[i], min1);
                max1 = std::max(inbound[i], max1);
            }
            std::cout << "Min inbound: " << min1 << ", Max inbound:" << max1 << "\n";
        }
        std::cout << "integrity ok, checked " << connections_checked << " connections\n";
    }

 private:
    void set_space(SpaceInterface<dist_t> *s) {
        data_size_ = s->get_data_size();
        fstdistfunc_ = s->get_dist_func();
        dist_func_param_ = s->get_dist_func_param();
    }
    void sync_fields() {
        b200hnsw_info o;
        b200detail::check(b200hnsw_get_info(h_, &o));
        max_elements_ = o.max_elements;
        cur_element_count = o.cur_element_count;
        num_deleted_ = o.num_deleted;
        size_data_per_element_ = o.size_data_per_element;
        size_links_per_element_ = o.size_links_per_element;
        size_links_level0_ = o.size_links_level0;
        M_ = o.M; maxM_ = o.maxM; maxM0_ = o.maxM0;
        ef_construction_ = o.ef_construction;
        ef_ = o.ef;
        mult_ = o.mult;
        revSize_ = o.mult != 0.0 ? 1.0 / o.mult : 0.0;
        maxlevel_ = o.maxlevel;
        enterpoint_node_ = o.enterpoint_node;
        offsetData_ = o.offset_data;
        offsetLevel0_ = 0;
        label_offset_ = o.label_offset;
    }
    // Parallel addPoint callers (the reference allows them, hnswalg.h:40-43): whoever gets here copies every level that
    // has not been mirrored yet, so no window between "count before" and "count after" can lose one.
    mutable std::mutex fields_mu_;
    size_t levels_mirrored_ = 0;
    mutable std::vector<uint64_t> filter_labels_;  // getExternalLabel of every internal id, for filter functors
    void after_add(size_t) {
        std::lock_guard<std::mutex> g(fields_mu_);
        filter_labels_.clear();
        sync_fields();
        const int32_t *lv = nullptr;
        b200detail::check(b200hnsw_get_levels(h_, &lv));
        if (element_levels_.size() < max_elements_) element_levels_.resize(max_elements_, 0);
        const size_t cur = cur_element_count;
        for (size_t i = levels_mirrored_; i < cur; i++) element_levels_[i] = lv[i];
        if (cur > levels_mirrored_) levels_mirrored_ = cur;
    }
};

}  // namespace hnswlib
