// hnswlib/stop_condition.h -- drop-in for /root/reference/hnswlib/stop_condition.h (names and signatures only).
//
// The reference drives searchStopConditionClosest (hnswalg.h:1327-1378) through HOST callbacks invoked per candidate
// (BaseSearchStopCondition, hnswlib.h:134-150).  A GPU kernel cannot call back into the host, so each stop condition
// needs a device formulation:
//   * EpsilonSearchStopCondition (stop_condition.h:218-275) has one: it is a searchKnn with ef = k = max_num_candidates
//     whose result is cut at distance epsilon (the callbacks only END the traversal earlier; the GPU traversal explores at
//     least as much, so every element the reference returns inside the epsilon ball is found, possibly more);
//   * MultiVectorSearchStopCondition and the MultiVector*Space classes (rows that carry a document id behind the
//     vector, :9-216) have none yet: the classes exist so that code naming them compiles, constructing an index over
//     such a space or searching with that condition throws std::runtime_error -- there is no CPU fallback.
#pragma once
#include "hnswlib.h"

namespace hnswlib {

// stop_condition.h:9-16
template <typename DOCIDTYPE>
class BaseMultiVectorSpace : public SpaceInterface<float> {
 public:
    virtual DOCIDTYPE get_doc_id(const void *datapoint) = 0;
    virtual void set_doc_id(void *datapoint, DOCIDTYPE doc_id) = 0;
};

namespace b200detail {
template <typename DOCIDTYPE, bool IP>
class MultiVectorSpaceImpl : public BaseMultiVectorSpace<DOCIDTYPE> {
    size_t data_size_, vector_size_, dim_;

 public:
    explicit MultiVectorSpaceImpl(size_t dim) : data_size_(dim * sizeof(float) + sizeof(DOCIDTYPE)), vector_size_(dim * sizeof(float)), dim_(dim) {}
    size_t get_data_size() override { return data_size_; }                       // vector + document id
    DISTFUNC<float> get_dist_func() override { return host_dist<IP>; }           // usable on the host, as the contract says
    void *get_dist_func_param() override { return &dim_; }
    DOCIDTYPE get_doc_id(const void *datapoint) override { return *(const DOCIDTYPE *)((const char *)datapoint + vector_size_); }
    void set_doc_id(void *datapoint, DOCIDTYPE doc_id) override { *(DOCIDTYPE *)((char *)datapoint + vector_size_) = doc_id; }
    // b200_metric() stays -1: rows with a trailing document id are not a layout the GPU engine stores
};
}  // namespace b200detail

template <typename DOCIDTYPE>
class MultiVectorL2Space : public b200detail::MultiVectorSpaceImpl<DOCIDTYPE, false> {  // stop_condition.h:18-74
 public:
    explicit MultiVectorL2Space(size_t dim) : b200detail::MultiVectorSpaceImpl<DOCIDTYPE, false>(dim) {}
};

template <typename DOCIDTYPE>
class MultiVectorInnerProductSpace : public b200detail::MultiVectorSpaceImpl<DOCIDTYPE, true> {  // stop_condition.h:77-143
 public:
    explicit MultiVectorInnerProductSpace(size_t dim) : b200detail::MultiVectorSpaceImpl<DOCIDTYPE, true>(dim) {}
};

// stop_condition.h:146-216: the bookkeeping is the reference's contract (callers may drive it themselves); the GPU search
// does not accept it.
template <typename DOCIDTYPE, typename dist_t>
class MultiVectorSearchStopCondition : public BaseSearchStopCondition<dist_t> {
    size_t curr_num_docs_ = 0, num_docs_to_search_, ef_collection_;
    std::unordered_map<DOCIDTYPE, size_t> doc_counter_;
    std::priority_queue<std::pair<dist_t, DOCIDTYPE>> search_results_;
    BaseMultiVectorSpace<DOCIDTYPE> &space_;

 public:
    MultiVectorSearchStopCondition(BaseMultiVectorSpace<DOCIDTYPE> &space, size_t num_docs_to_search, size_t ef_collection = 10)
        : num_docs_to_search_(num_docs_to_search), ef_collection_(std::max(ef_collection, num_docs_to_search)), space_(space) {}
    void add_point_to_result(labeltype, const void *datapoint, dist_t dist) override {
        const DOCIDTYPE doc = space_.get_doc_id(datapoint);
        if (doc_counter_[doc]++ == 0) curr_num_docs_++;
        search_results_.emplace(dist, doc);
    }
    void remove_point_from_result(labeltype, const void *datapoint, dist_t) override {
        const DOCIDTYPE doc = space_.get_doc_id(datapoint);
        if (--doc_counter_[doc] == 0) curr_num_docs_--;
        search_results_.pop();
    }
    bool should_stop_search(dist_t candidate_dist, dist_t lowerBound) override {
        return candidate_dist > lowerBound && curr_num_docs_ == ef_collection_;
    }
    bool should_consider_candidate(dist_t candidate_dist, dist_t lowerBound) override {
        return curr_num_docs_ < ef_collection_ || lowerBound > candidate_dist;
    }
    bool should_remove_extra() override { return curr_num_docs_ > ef_collection_; }
    void filter_results(std::vector<std::pair<dist_t, labeltype>> &candidates) override {
        while (curr_num_docs_ > num_docs_to_search_) {
            const DOCIDTYPE doc = search_results_.top().second;
            if (--doc_counter_[doc] == 0) curr_num_docs_--;
            search_results_.pop();
            candidates.pop_back();
        }
    }
};

// stop_condition.h:218-275
template <typename dist_t>
class EpsilonSearchStopCondition : public BaseSearchStopCondition<dist_t> {
    float epsilon_;
    size_t min_num_candidates_, max_num_candidates_, curr_num_items_ = 0;

 public:
    EpsilonSearchStopCondition(float epsilon, size_t min_num_candidates, size_t max_num_candidates)
        : epsilon_(epsilon), min_num_candidates_(min_num_candidates), max_num_candidates_(max_num_candidates) {}
    void add_point_to_result(labeltype, const void *, dist_t) override { curr_num_items_++; }
    void remove_point_from_result(labeltype, const void *, dist_t) override { curr_num_items_--; }
    bool should_stop_search(dist_t candidate_dist, dist_t lowerBound) override {
        if (candidate_dist > lowerBound && curr_num_items_ == max_num_candidates_) return true;
        return candidate_dist > epsilon_ && curr_num_items_ >= min_num_candidates_;
    }
    bool should_consider_candidate(dist_t candidate_dist, dist_t lowerBound) override {
        return curr_num_items_ < max_num_candidates_ || lowerBound > candidate_dist;
    }
    bool should_remove_extra() override { return curr_num_items_ > max_num_candidates_; }
    void filter_results(std::vector<std::pair<dist_t, labeltype>> &candidates) override {
        while (!candidates.empty() && candidates.back().first > epsilon_) candidates.pop_back();
        while (candidates.size() > max_num_candidates_) candidates.pop_back();
    }
    // the device formulation (see the header comment)
    bool b200_epsilon_form(float *epsilon, size_t *max_candidates) const override {
        *epsilon = epsilon_;
        *max_candidates = max_num_candidates_;
        return true;
    }
};

}  // namespace hnswlib
