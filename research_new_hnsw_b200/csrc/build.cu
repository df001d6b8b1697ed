// build.cu -- batched GPU graph construction behind addPoint.
//
// Reference path replaced (file:line under /root/reference/hnswlib/hnswalg.h):
//   :1153-1267  addPoint(data, label, level)      -> HnswIndex::add_batch (staging, eager public fields) + flush()
//   :1214-1239  greedy descent above the new level -> build_search_kernel prologue (greedy_level)
//   :225-305    searchBaseLayer (ef_construction)  -> build_search_kernel (beam_level on every level <= the new one)
//   :443-483    getNeighborsByHeuristic2           -> heuristic_prune (device function)
//   :506-630    mutuallyConnectNewElement          -> build_link_kernel (forward lists + reverse-edge staging) and
//                                                     build_reverse_kernel (append, or re-prune to Mcurmax when full)
//
// Batching.  The reference inserts one point at a time; here points are linked in insertion order in batches that
// see the graph as of the start of their batch.  A batch is never larger than (linked points) / build_ratio, so at
// most a 1/build_ratio (default 1/32) fraction of a point's true neighbours is invisible to it, and a point that raises the
// maximum level always ends its batch (it becomes the entry point of everything after it, :1262-1265).  Levels,
// element count, entry point and max level are assigned at add time with the reference's generator and are
// bit-identical to the reference for the same insertion order (:207-211,1187-1198).
//
// Reverse links are lock-free: build_link_kernel appends (dist, new id) to a per-list incoming buffer with one
// atomicAdd per edge and records each touched list once; build_reverse_kernel then gives every touched list to one
// CTA which appends while there is room (:586-588) and otherwise re-runs the heuristic over existing + incoming
// (:590-612), once per batch instead of once per edge.
//
// No host round trip inside flush(): batch boundaries depend only on the element levels, so the whole plan (list
// slots of every batch) is computed up front and uploaded ONCE; per batch the host only enqueues three kernels and
// a 4-byte memset.  build_reverse_kernel is persistent (grid-stride over the device-side count of touched lists).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_set>
#include <vector>

#include "build_kernels.cuh"

namespace b200 {

// ---- host side ---------------------------------------------------------------------------------------------

BuildProfile *build_profile() {
    static BuildProfile prof;
    static const bool on = getenv("B200HNSW_BUILD_PROFILE") && atoi(getenv("B200HNSW_BUILD_PROFILE")) != 0;
    return on ? &prof : nullptr;
}

static size_t env_size(const char *name, size_t dflt) {
    if (const char *e = getenv(name)) {
        const long long v = atoll(e);
        if (v > 0) return (size_t)v;
    }
    return dflt;
}

// addPoint staging (hnswalg.h:1153-1211,1255-1265): everything that does not need a distance.
int HnswIndex::add_batch(const float *X, const uint64_t *labels, size_t n, bool replace_deleted) {
    std::unique_lock<std::shared_mutex> g(rw);
    drain_async();
    struct StagedFlag {  // whatever path leaves this function, has_staged reflects the host image
        HnswIndex &ix;
        ~StagedFlag() { ix.has_staged = ix.linked < ix.host.cur; }
    } staged_flag{*this};
    HostImage &m = host;
    std::vector<uint32_t> updates, revived;  // revived: slots that were marked deleted until this call
    // Pass 1 (sequential: label map, level generator, slot assignment) records which record each new row goes to;
    // pass 2 writes the records of the new points with several threads (first touch of ~0.8 KB per point dominates a
    // million-point addPoints otherwise); rows that overwrite a point staged by this very call are applied after it,
    // in call order.
    const size_t cur0 = m.cur;
    std::vector<size_t> fresh_row;  // row of X behind slot cur0 + j
    fresh_row.reserve(n);
    std::vector<std::pair<size_t, uint32_t>> later;  // (row of X, slot) for labels repeated inside this call
    auto write_rows = [&]() {  // also on the error returns below: the rows accepted so far stay added
        stage_records(X, labels, cur0, fresh_row);
        for (const auto &u : later) memcpy(m.rec(u.second) + m.off_data, X + u.first * m.dim, m.dim * 4);
    };
    for (size_t i = 0; i < n; i++) {
        const uint64_t lab = labels ? labels[i] : (uint64_t)m.cur;
        auto known = m.label_lookup.find(lab);
        if (known != m.label_lookup.end()) {
            // existing label: update instead of insert (hnswalg.h:1157-1174)
            const uint32_t c = known->second;
            if (c < cur0 && m.deleted(c)) {  // (records of this call's own new points are written in pass 2)
                if (prm.allow_replace_deleted) {
                    set_error("Can't use addPoint to update deleted elements if replacement of deleted elements is enabled.");
                    write_rows();
                    return B200HNSW_E_STATE;
                }
                *((unsigned char *)m.rec(c) + 2) &= (unsigned char)~1;  // unmarkDeletedInternal
                m.num_deleted--;
                flags_dirty = true;
                if (c < linked) revived.push_back(c);
            }
            if (c >= cur0) later.emplace_back(i, c);
            else memcpy(m.rec(c) + m.off_data, X + i * m.dim, m.dim * 4);
            if (c < linked) updates.push_back(c);  // a staged point is simply linked with its new vector later
            continue;
        }
        if (replace_deleted && m.num_deleted > 0) {
            // vacant place: a deleted element takes the new label and vector (hnswalg.h:965-992)
            size_t c = replace_scan < m.cur ? replace_scan : 0;
            while (c >= cur0 || !m.deleted(c)) c = c + 1 < cur0 ? c + 1 : 0;  // num_deleted > 0: terminates
            replace_scan = c + 1;
            uint64_t old_label;
            memcpy(&old_label, m.rec(c) + m.off_label, 8);
            m.label_lookup.erase(old_label);
            m.label_lookup[lab] = (uint32_t)c;
            memcpy(m.rec(c) + m.off_label, &lab, 8);
            *((unsigned char *)m.rec(c) + 2) &= (unsigned char)~1;
            m.num_deleted--;
            flags_dirty = true;
            memcpy(m.rec(c) + m.off_data, X + i * m.dim, m.dim * 4);
            if (c < linked) { updates.push_back((uint32_t)c); revived.push_back((uint32_t)c); }
            continue;
        }
        if (m.cur >= m.max_elements) {
            set_error("The number of elements exceeds the specified limit");
            write_rows();
            return B200HNSW_E_CAPACITY;
        }
        const size_t c = m.cur++;
        m.label_lookup[lab] = (uint32_t)c;
        const int level = m.random_level();
        m.levels[c] = level;
        fresh_row.push_back(i);
        if (level > 0) m.upper[c].assign((size_t)level * m.size_links, 0);
        else m.upper[c].clear();
        if (c == 0) {
            m.enterpoint = 0;
            m.maxlevel = level;
        } else if (level > m.maxlevel) {
            m.enterpoint = (uint32_t)c;
            m.maxlevel = level;
        }
    }
    write_rows();
    return updates.empty() ? 0 : relink_points(std::move(updates), revived);
}

// Level-0 records of the new points [cur0, cur0 + rows.size()): zeroed header and link list, vector, label.
void HnswIndex::stage_records(const float *X, const uint64_t *labels, size_t cur0, const std::vector<size_t> &rows) {
    HostImage &m = host;
    const size_t cnt = rows.size();
    auto fill = [&](size_t a, size_t b) {
        for (size_t j = a; j < b; j++) {
            const size_t c = cur0 + j, i = rows[j];
            const uint64_t lab = labels ? labels[i] : (uint64_t)c;
            memset(m.rec(c), 0, m.off_data);
            memcpy(m.rec(c) + m.off_data, X + i * m.dim, m.dim * 4);
            memcpy(m.rec(c) + m.off_label, &lab, 8);
        }
    };
    size_t nt = std::min<size_t>(std::min<size_t>(std::thread::hardware_concurrency(), 8), cnt / 16384);
    if (nt <= 1) {
        fill(0, cnt);
        return;
    }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) th.emplace_back(fill, cnt * t / nt, cnt * (t + 1) / nt);
    for (auto &t : th) t.join();
}

// updatePoint (hnswalg.h:995-1139) for points that are already part of the device graph.  Per group of points: the new
// vectors (and labels: replace_deleted relabels) are staged and scattered into the device rows, the first phase
// re-prunes every old neighbour over the 1-hop + 2-hop set (update_*_kernel, :1009-1069), then
// repairConnectionsForUpdate runs with the construction kernels in update mode: search from the entry point, the point
// itself dropped from its candidates, heuristic selection, forward list REPLACED, reverse edges added unless already
// present (mutuallyConnectNewElement with isUpdate, :485-639).
//
// The reference updates one point at a time, each seeing the graph the previous update left.  Up to
// B200HNSW_UPDATE_SEQ (default 2048) points per call are therefore processed ONE PER GROUP, in call order -- the same
// sequence of graph states as the reference; larger calls are processed in groups that see the graph as of the start of
// their group (the batching rule of the build), a neighbour shared by two points of a group being re-pruned for the
// later one.
int HnswIndex::relink_points(std::vector<uint32_t> ids, const std::vector<uint32_t> &revived) {
    HostImage &m = host;
    {   // call order, one entry per point (a label given twice: the host image already holds the last vector)
        std::vector<uint32_t> uniq;
        uniq.reserve(ids.size());
        std::unordered_set<uint32_t> seen;
        for (uint32_t id : ids)
            if (seen.insert(id).second) uniq.push_back(id);
        ids.swap(uniq);
    }
    B200_CUDA_OK(cudaSetDevice(dev.device));
    const size_t build_ratio = env_size("B200HNSW_BUILD_RATIO", 32);
    const size_t max_batch = env_size("B200HNSW_BUILD_BATCH", 16384);
    const size_t seq_limit = env_size("B200HNSW_UPDATE_SEQ", 2048);
    const bool sequential = ids.size() <= seq_limit;
    if (!bld.plevel) B200_CUDA_OK(cudaMalloc(&bld.plevel, std::max<size_t>(dev.cap, 1) * 4));
    B200_CUDA_OK(cudaMemcpy(bld.plevel, m.levels.data(), linked * 4, cudaMemcpyHostToDevice));
    // A slot that was marked deleted until this call holds its OLD vector on the device until its own turn: the
    // kernels must keep seeing it as deleted until then (the reference unmarks it right before updatePoint, :983-986,
    // :1168-1171).  The device marks start from that state; scatter_rows_kernel clears the mark with the new row.
    revived_on_device = !revived.empty();
    if (revived_on_device) {
        const int rcf = upload_flags(nullptr, revived.data(), revived.size());
        if (rcf) { revived_on_device = false; return rcf; }
    }
    BuildArgs a{};
    size_t max_lists = 0, smem_search = 0, smem_link = 0;
    int rc = prepare_build(&a, max_batch, &max_lists, &smem_search, &smem_link);
    revived_on_device = false;
    if (rc) return rc;
    const bool nb = a.flags != nullptr;
    if (!bld.batch_ids) B200_CUDA_OK(cudaMalloc(&bld.batch_ids, max_batch * 4));
    // scratch of the first phase: one re-pruned list per (list of the group, neighbour slot); bounded to 64 MB
    const size_t per_list = m.maxM0 * m.maxM0 * 4;
    const size_t upd_lists = std::max<size_t>(64, std::min<size_t>(max_lists, ((size_t)64 << 20) / per_list));
    if (!bld.newlists || bld.newlists_cap < upd_lists) {
        cudaFree(bld.newlists);
        bld.newlists = nullptr;
        B200_CUDA_OK(cudaMalloc(&bld.newlists, upd_lists * per_list));
        bld.newlists_cap = upd_lists;
    }
    const size_t rowf = dev.d4 * 4;
    const size_t stage_rows = sequential ? 1 : std::min(max_batch, ids.size());
    if (bld.stage_cap < stage_rows) {
        cudaFree(bld.stage_rows); cudaFree(bld.stage_labels);
        bld.stage_rows = nullptr; bld.stage_labels = nullptr;
        B200_CUDA_OK(cudaMalloc(&bld.stage_rows, stage_rows * rowf * 4));
        B200_CUDA_OK(cudaMalloc(&bld.stage_labels, stage_rows * 8));
        bld.stage_cap = stage_rows;
    }
    std::vector<uint32_t> off, lp, ll;
    std::vector<float> hrows;
    std::vector<uint64_t> hlabels;
    uint64_t launches = 0;
    for (size_t b0 = 0; b0 < ids.size() && rc == 0;) {
        size_t B = sequential ? 1 : std::max<size_t>(1, std::min(max_batch, linked / build_ratio));
        B = std::min(B, ids.size() - b0);
        off.clear(); lp.clear(); ll.clear();
        size_t used = 0;
        for (; used < B; used++) {
            const uint32_t id = ids[b0 + used];
            const int top = std::min(m.levels[id], dev_maxlevel);
            if (used > 0 && lp.size() + (size_t)top + 1 > std::min(max_lists, upd_lists)) break;
            off.push_back((uint32_t)lp.size());
            for (int l = 0; l <= top; l++) { lp.push_back(id); ll.push_back((uint32_t)l); }
        }
        B = used;
        // new vectors / labels of the group -> device rows (one staged copy + one scatter launch per group)
        hrows.assign(B * rowf, 0.f);
        hlabels.resize(B);
        for (size_t i = 0; i < B; i++) {
            const uint32_t id = ids[b0 + i];
            memcpy(hrows.data() + i * rowf, m.rec(id) + m.off_data, m.dim * 4);
            memcpy(&hlabels[i], m.rec(id) + m.off_label, 8);
        }
        B200_CUDA_OK(cudaMemcpyAsync(bld.stage_rows, hrows.data(), B * rowf * 4, cudaMemcpyHostToDevice, stream));
        B200_CUDA_OK(cudaMemcpyAsync(bld.stage_labels, hlabels.data(), B * 8, cudaMemcpyHostToDevice, stream));
        B200_CUDA_OK(cudaMemcpyAsync(bld.batch_ids, ids.data() + b0, B * 4, cudaMemcpyHostToDevice, stream));
        scatter_rows_kernel<<<(unsigned)((B * 32 + 127) / 128), 128, 0, stream>>>(
            (const float4 *)bld.stage_rows, bld.stage_labels, bld.batch_ids, (uint32_t)B, (uint32_t)dev.d4, (uint32_t)dev.d16,
            dev.vec, dev.labels, prm.storage == B200HNSW_BF16 ? dev.vec16 : nullptr, nb ? dev.flags : nullptr);
        if (nb) flags_on_device_valid = false;  // the kernel clears marks on the device
        launches += 1;
        b0 += B;
        if (linked <= 1) continue;  // a single element has nothing to connect to (hnswalg.h:1001-1003)
        B200_CUDA_OK(cudaMemcpyAsync(bld.list_off, off.data(), B * 4, cudaMemcpyHostToDevice, stream));
        B200_CUDA_OK(cudaMemcpyAsync(bld.list_point, lp.data(), lp.size() * 4, cudaMemcpyHostToDevice, stream));
        B200_CUDA_OK(cudaMemcpyAsync(bld.list_level, ll.data(), ll.size() * 4, cudaMemcpyHostToDevice, stream));
        a.first = 0; a.batch = (uint32_t)B; a.lists = (uint32_t)lp.size();
        a.entry = dev_entry; a.maxlevel = dev_maxlevel;
        a.batch_ids = bld.batch_ids;
        rc = build_run_update_phase1(prm.metric, a, bld.newlists, stream);
        if (rc) break;
        rc = nb ? build_run_batch_update_nb(prm.metric, a, smem_search, smem_link, stream)
                : build_run_batch_update(prm.metric, a, smem_search, smem_link, stream);
        launches += 6;
    }
    {
        const cudaError_t e = cudaStreamSynchronize(stream);
        if (rc == 0 && e != cudaSuccess) {
            set_error(std::string("CUDA error in updatePoint: ") + cudaGetErrorString(e));
            rc = B200HNSW_E_CUDA;
        }
    }
    {
        std::lock_guard<std::mutex> sg(stats_mu);
        stats.kernel_launches += launches;
    }
    mirror_dirty = true;
    return rc;
}

// Device graph -> reference-layout host mirror (needed by saveIndex and get_linklist*).
int HnswIndex::sync_host_mirror() {
    if (!mirror_dirty) return 0;
    B200_CUDA_OK(cudaSetDevice(dev.device));
    HostImage &m = host;
    const size_t n = m.cur;
    const size_t chunk = std::max<size_t>(1, (size_t)(64u << 20) / (m.maxM0 * 4));
    std::vector<uint32_t> tmp(chunk * m.maxM0);
    for (size_t s = 0; s < n; s += chunk) {
        const size_t c = std::min(chunk, n - s);
        B200_CUDA_OK(cudaMemcpy(tmp.data(), dev.links0 + s * m.maxM0, c * m.maxM0 * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < c; i++) {
            const uint32_t *src = tmp.data() + i * m.maxM0;
            uint32_t *dst = m.list(s + i, 0);
            unsigned cnt = 0;
            while (cnt < m.maxM0 && src[cnt] != kEmpty) cnt++;
            HostImage::set_count(dst, (unsigned short)cnt);
            memcpy(dst + 1, src, cnt * 4);
            memset(dst + 1 + cnt, 0, (m.maxM0 - cnt) * 4);
        }
    }
    if (dev.up_lists) {
        std::vector<uint32_t> up(dev.up_lists * m.maxM);
        B200_CUDA_OK(cudaMemcpy(up.data(), dev.links_up, up.size() * 4, cudaMemcpyDeviceToHost));
        size_t li = 0;
        for (size_t i = 0; i < n; i++)
            for (int l = 1; l <= m.levels[i]; l++, li++) {
                const uint32_t *src = up.data() + li * m.maxM;
                uint32_t *dst = m.list(i, l);
                unsigned cnt = 0;
                while (cnt < m.maxM && src[cnt] != kEmpty) cnt++;
                dst[0] = 0;
                HostImage::set_count(dst, (unsigned short)cnt);
                memcpy(dst + 1, src, cnt * 4);
                memset(dst + 1 + cnt, 0, (m.maxM - cnt) * 4);
            }
    }
    mirror_dirty = false;
    return 0;
}

// Scratch of the build kernels (sized for batches of max_batch points) and the kernel arguments that do not depend on
// the batch; shared by flush() (new points) and relink_points() (updatePoint).
int HnswIndex::prepare_build(void *args, size_t max_batch, size_t *max_lists_out, size_t *smem_search, size_t *smem_link) {
    HostImage &m = host;
    const size_t nl = dev.cap + dev.up_lists_cap;
    if (!bld.plevel) B200_CUDA_OK(cudaMalloc(&bld.plevel, std::max<size_t>(dev.cap, 1) * 4));
    if (!bld.incnt) {
        B200_CUDA_OK(cudaMalloc(&bld.incnt, nl * 4));
        B200_CUDA_OK(cudaMemset(bld.incnt, 0, nl * 4));
        B200_CUDA_OK(cudaMalloc(&bld.incoming, nl * kCapIn * 8));
    }
    const size_t max_lists = max_batch + max_batch / 2 + 64;  // level-0 list per point + the rare upper lists
    *max_lists_out = max_lists;
    if (!bld.cand || bld.cand_efc != m.efc) {
        cudaFree(bld.cand); cudaFree(bld.cand_cnt); cudaFree(bld.list_off); cudaFree(bld.list_point);
        cudaFree(bld.list_level); cudaFree(bld.aff_node); cudaFree(bld.aff_level); cudaFree(bld.aff_count);
        cudaFree(bld.work);
        B200_CUDA_OK(cudaMalloc(&bld.cand, max_lists * m.efc * 8));
        B200_CUDA_OK(cudaMalloc(&bld.cand_cnt, max_lists * 4));
        B200_CUDA_OK(cudaMalloc(&bld.list_off, max_batch * 4));
        B200_CUDA_OK(cudaMalloc(&bld.list_point, max_lists * 4));
        B200_CUDA_OK(cudaMalloc(&bld.list_level, max_lists * 4));
        B200_CUDA_OK(cudaMalloc(&bld.aff_node, max_lists * m.M * 4));
        B200_CUDA_OK(cudaMalloc(&bld.aff_level, max_lists * m.M * 4));
        B200_CUDA_OK(cudaMalloc(&bld.aff_count, 4));
        B200_CUDA_OK(cudaMemset(bld.aff_count, 0, 4));
        B200_CUDA_OK(cudaMalloc(&bld.work, 64));
        bld.cand_efc = m.efc;
    }
    B200_CUDA_OK(cudaMemset(bld.work, 0, 64));

    const size_t list_cap = std::max(m.maxM, m.maxM0);
    BuildArgs &a = *(BuildArgs *)args;
    a = BuildArgs{};
    a.vec = dev.vec; a.links0 = dev.links0; a.up_base = dev.up_base; a.links_up = dev.links_up;
    a.plevel = bld.plevel; a.cand = bld.cand; a.cand_cnt = bld.cand_cnt; a.list_off = bld.list_off;
    a.list_point = bld.list_point; a.list_level = bld.list_level; a.incnt = bld.incnt; a.incoming = bld.incoming;
    a.aff_node = bld.aff_node; a.aff_level = bld.aff_level; a.aff_count = bld.aff_count; a.work = bld.work;
    a.cap = (uint32_t)dev.cap; a.d4 = (uint32_t)dev.d4; a.maxM = (uint32_t)m.maxM; a.maxM0 = (uint32_t)m.maxM0;
    a.M = (uint32_t)m.M; a.efc = (uint32_t)m.efc;
    // elements marked deleted: the construction search keeps them out of the candidates (NB kernels, buffer of 2 * efc)
    const bool nb = m.num_deleted != 0 || revived_on_device;
    if (nb) {
        if (linked >= (1u << 30)) {
            set_error("build with deleted elements supports at most 2^30 elements");
            return B200HNSW_E_UNSUPPORTED;
        }
        if (!revived_on_device) {  // relink_points has already uploaded the marks it wants the kernels to see
            const int rcf = upload_flags();
            if (rcf) return rcf;
        }
    }
    a.flags = nb ? dev.flags : nullptr;
    {
        static const int pf_env = getenv("B200HNSW_BUILD_PF") ? atoi(getenv("B200HNSW_BUILD_PF")) : -1;
        a.pf = pf_env >= 0 ? (uint32_t)pf_env : (kPfRows | kPfGreedy | kPfRound1);
    }
    const size_t bufcap = nb ? 2 * m.efc : m.efc;
    // construction searches evaluate ~40 * efc nodes; a table smaller than that is rebuilt from the candidate buffer
    // when it fills (re-evaluations only, same graph) and lets more CTAs share an SM: measured at C5 (efc = 200) with
    // 128-thread CTAs, construction-search time 1470 ms with 8192 slots (6 CTAs per SM) vs 1271 ms with 4096; with
    // 64-thread CTAs 1203 ms (4096) / 924 ms (2048, 16 CTAs per SM, +5 % evaluations) / 976 ms (1024, +8 %)
    {
        size_t want = std::min<size_t>(env_size("B200HNSW_BUILD_HASH", 2048), 32 * m.efc + 1024);
        want = std::max(want, 8 * (bufcap + list_cap) / 3 + 64);  // a hop must fit above the 5/8 rebuild mark
        a.hash_bits = 10;
        while ((1ull << a.hash_bits) < want) a.hash_bits++;
    }
    const SearchSmem SL((uint32_t)bufcap, (uint32_t)list_cap, a.d4, a.hash_bits);
    const LinkSmem LL((uint32_t)list_cap + kCapIn);
    if (SL.total > 226 * 1024) {
        set_error("ef_construction too large for the build kernel's shared memory");
        return B200HNSW_E_UNSUPPORTED;
    }
    *smem_search = SL.total;
    *smem_link = LL.total;
    return 0;
}


// Link every staged point (ids [linked, host.cur)) into the device graph.
int HnswIndex::flush() {
    if (!has_staged) return 0;  // the common case of a search: nothing staged, no exclusive lock
    std::unique_lock<std::shared_mutex> g(rw);
    drain_async();
    return flush_locked();
}

int HnswIndex::flush_locked() {
    struct StagedFlag {
        HnswIndex &ix;
        ~StagedFlag() { ix.has_staged = ix.linked < ix.host.cur; }
    } staged_flag{*this};
    HostImage &m = host;
    if (linked >= m.cur) return 0;
    B200_CUDA_OK(cudaSetDevice(dev.device));
    const size_t n_new = m.cur - linked;
    const size_t rec = m.size_data;
    const size_t build_ratio = env_size("B200HNSW_BUILD_RATIO", 32);
    const size_t max_batch = env_size("B200HNSW_BUILD_BATCH", 16384);
    // measured at C5 (gpurun_out/r02_build_ramp.log): ramp ratio 32 / 16 / 8 / 4 -> 1185 / 780 / 555 / 432 launches,
    // 1634 / 1540 / 1502 / 1480 ms of kernels, recall@10 at ef=28 0.9486 / 0.9485 / 0.9486 / 0.9507 (reference-built: 0.950)
    const size_t ramp_ratio = env_size("B200HNSW_BUILD_RAMP_RATIO", std::min<size_t>(build_ratio, 8));
    const size_t ramp_until = env_size("B200HNSW_BUILD_RAMP_UNTIL", 65536);

    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    // (the rows themselves are uploaded further down, chunk by chunk, while the first batches are already being linked)
    // ---- upper-level list slots of the staged points (appended after the existing ones) ----
    {
        std::vector<uint32_t> base(n_new, kEmpty);
        size_t lists = dev.up_lists;
        for (size_t i = 0; i < n_new; i++)
            if (m.levels[linked + i] > 0) {
                base[i] = (uint32_t)lists;
                lists += (size_t)m.levels[linked + i];
            }
        if (lists > dev.up_lists_cap) {
            const size_t ncap = std::max(lists, dev.up_lists_cap * 2);
            uint32_t *nu = nullptr;
            B200_CUDA_OK(cudaMalloc(&nu, ncap * m.maxM * 4));
            B200_CUDA_OK(cudaMemset(nu, 0xFF, ncap * m.maxM * 4));
            if (dev.up_lists)
                B200_CUDA_OK(cudaMemcpy(nu, dev.links_up, dev.up_lists * m.maxM * 4, cudaMemcpyDeviceToDevice));
            cudaFree(dev.links_up);
            dev.links_up = nu;
            dev.up_lists_cap = ncap;
            cudaFree(bld.incnt); cudaFree(bld.incoming);  // sized by cap + up_lists_cap
            bld.incnt = nullptr; bld.incoming = nullptr;
        } else if (lists > dev.up_lists) {
            B200_CUDA_OK(cudaMemset(dev.links_up + dev.up_lists * m.maxM, 0xFF, (lists - dev.up_lists) * m.maxM * 4));
        }
        dev.up_lists = lists;
        B200_CUDA_OK(cudaMemcpy(dev.up_base + linked, base.data(), n_new * 4, cudaMemcpyHostToDevice));
    }
    // ---- build scratch + kernel arguments ----
    if (!bld.plevel) B200_CUDA_OK(cudaMalloc(&bld.plevel, std::max<size_t>(dev.cap, 1) * 4));
    B200_CUDA_OK(cudaMemcpy(bld.plevel + linked, m.levels.data() + linked, n_new * 4, cudaMemcpyHostToDevice));
    BuildArgs a{};
    size_t max_lists = 0, smem_search = 0, smem_link = 0;
    {
        const int rcp = prepare_build(&a, max_batch, &max_lists, &smem_search, &smem_link);
        if (rcp) return rcp;
    }

    const bool nb = a.flags != nullptr;
    // ---- the whole batch plan, from host-side information only (levels, number of linked points) ----
    struct BatchPlan { uint32_t first, batch, lists_off, lists, entry; int32_t maxlevel; };
    std::vector<BatchPlan> plan;
    const size_t linked0 = linked;
    std::vector<uint32_t> off_all(n_new), lp_all, ll_all;
    lp_all.reserve(n_new + n_new / 8 + 16);
    ll_all.reserve(n_new + n_new / 8 + 16);
    {
        size_t lk = linked;
        uint32_t ent = dev_entry;
        int ml = dev_maxlevel;
        if (lk == 0) {  // first element: nothing to link (hnswalg.h:1255-1259)
            ent = 0;
            ml = m.levels[0];
            lk = 1;
        }
        while (lk < m.cur) {
            // while fewer points are linked than a batch needs to fill the GPU, a batch may be a larger fraction of them
            // (only on builds that go far beyond the ramp: a small index is built at the validated ratio throughout)
            const size_t ratio = (lk < ramp_until && m.cur >= 4 * ramp_until) ? ramp_ratio : build_ratio;
            size_t B = std::max<size_t>(1, std::min(max_batch, lk / ratio));
            B = std::min(B, m.cur - lk);
            BatchPlan bp{(uint32_t)lk, 0, (uint32_t)lp_all.size(), 0, ent, ml};
            size_t used = 0, nl = 0;
            while (used < B) {
                const int lev = m.levels[lk + used];
                const int top = std::min(lev, ml);
                if (used > 0 && nl + (size_t)top + 1 > max_lists) break;  // the candidate pool holds max_lists lists
                off_all[lk + used - linked0] = (uint32_t)nl;
                for (int l = 0; l <= top; l++) { lp_all.push_back((uint32_t)(lk + used)); ll_all.push_back((uint32_t)l); }
                nl += (size_t)top + 1;
                used++;
                if (lev > ml) break;  // a new top level ends the batch: it is the entry point of everything after it
            }
            bp.batch = (uint32_t)used;
            bp.lists = (uint32_t)nl;
            plan.push_back(bp);
            const size_t last = lk + used - 1;
            if (m.levels[last] > ml) { ent = (uint32_t)last; ml = m.levels[last]; }
            lk += used;
        }
    }
    if (bld.plan_off_cap < n_new) {
        cudaFree(bld.plan_off);
        bld.plan_off = nullptr; bld.plan_off_cap = 0;
        B200_CUDA_OK(cudaMalloc(&bld.plan_off, std::max<size_t>(n_new, 1) * 4));
        bld.plan_off_cap = std::max<size_t>(n_new, 1);
    }
    if (bld.plan_lists_cap < lp_all.size()) {
        cudaFree(bld.plan_lp); cudaFree(bld.plan_ll);
        bld.plan_lp = bld.plan_ll = nullptr; bld.plan_lists_cap = 0;
        B200_CUDA_OK(cudaMalloc(&bld.plan_lp, std::max<size_t>(lp_all.size(), 1) * 4));
        B200_CUDA_OK(cudaMalloc(&bld.plan_ll, std::max<size_t>(lp_all.size(), 1) * 4));
        bld.plan_lists_cap = std::max<size_t>(lp_all.size(), 1);
    }
    uint32_t *d_off = bld.plan_off, *d_lp = bld.plan_lp, *d_ll = bld.plan_ll;
    B200_CUDA_OK(cudaMemcpyAsync(d_off, off_all.data(), n_new * 4, cudaMemcpyHostToDevice, stream));
    B200_CUDA_OK(cudaMemcpyAsync(d_lp, lp_all.data(), lp_all.size() * 4, cudaMemcpyHostToDevice, stream));
    B200_CUDA_OK(cudaMemcpyAsync(d_ll, ll_all.data(), ll_all.size() * 4, cudaMemcpyHostToDevice, stream));

    const double ms_plan = since(t_start);
    cudaEvent_t e0, e1;
    B200_CUDA_OK(cudaEventCreate(&e0));
    B200_CUDA_OK(cudaEventCreate(&e1));
    uint64_t launches = 0;
    int rc = 0;
    // ---- rows + labels of the staged points (records carry empty lists), overlapped with the build ----
    // The records go up in 32 MB chunks on their own stream (two staging buffers; a pageable source makes every copy
    // block the host until it is staged, which is the time the chunks are cut for); the build stream waits for the event
    // of the chunk that completes a batch's rows, so the first batches are linked while the later rows are still on
    // their way -- a million 128-d rows are ~0.1 s of H2D that used to precede the first kernel.
    struct { uint32_t *raw[2]; cudaEvent_t ev[2]; cudaStream_t up; } us;
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(n_new, env_size("B200HNSW_UPLOAD_CHUNK", (size_t)32 << 20) / rec + 1));
    if (!bld.up_stream) {
        B200_CUDA_OK(cudaStreamCreateWithFlags(&bld.up_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) B200_CUDA_OK(cudaEventCreateWithFlags(&bld.up_ev[b], cudaEventDisableTiming));
    }
    if (bld.up_raw_bytes < chunk * rec) {
        for (int b = 0; b < 2; b++) { cudaFree(bld.up_raw[b]); bld.up_raw[b] = nullptr; }
        bld.up_raw_bytes = 0;
        for (int b = 0; b < 2; b++) B200_CUDA_OK(cudaMalloc(&bld.up_raw[b], chunk * rec));
        bld.up_raw_bytes = chunk * rec;
    }
    for (int b = 0; b < 2; b++) { us.raw[b] = bld.up_raw[b]; us.ev[b] = bld.up_ev[b]; }
    us.up = bld.up_stream;
    double ms_upload = 0.0;
    size_t uploaded = linked0;  // rows [0, uploaded) are on the device (or on their way, ordered before the build stream)
    size_t n_chunks = 0;
    auto upload_next = [&]() -> int {
        const auto t_u = std::chrono::steady_clock::now();
        const size_t cnt = std::min(chunk, m.cur - uploaded);
        const int b = (int)(n_chunks & 1);
        if (n_chunks >= 2) B200_CUDA_OK(cudaEventSynchronize(us.ev[b]));  // the kernel that last read this buffer is done
        B200_CUDA_OK(cudaMemcpyAsync(us.raw[b], m.level0 + uploaded * rec, cnt * rec, cudaMemcpyHostToDevice, us.up));
        deinterleave_kernel<<<(unsigned)((cnt * 32 + 255) / 256), 256, 0, us.up>>>(
            us.raw[b], rec / 4, (uint32_t)uploaded, (uint32_t)cnt, (uint32_t)m.maxM0, (uint32_t)m.dim, (uint32_t)dev.d4,
            (uint32_t)m.cur, (float *)dev.vec, dev.links0, dev.labels, dev.err_flag);
        B200_CUDA_OK(cudaGetLastError());
        {
            const int rc16 = sync_bf16(uploaded, cnt, us.up);
            if (rc16) return rc16;
        }
        B200_CUDA_OK(cudaEventRecord(us.ev[b], us.up));
        B200_CUDA_OK(cudaStreamWaitEvent(stream, us.ev[b], 0));
        if (n_chunks == 0) B200_CUDA_OK(cudaEventRecord(e0, stream));  // kernel time is counted from the first rows on
        uploaded += cnt;
        n_chunks++;
        ms_upload += since(t_u);
        return 0;
    };
    if (linked == 0) {
        dev_entry = 0;
        dev_maxlevel = m.levels[0];
        linked = 1;
    }
    size_t next_batch = 0;
    while (rc == 0 && (uploaded < m.cur || next_batch < plan.size())) {
        if (uploaded < m.cur) {
            rc = upload_next();
            if (rc) break;
        }
        // every batch whose rows are complete
        while (next_batch < plan.size() && (size_t)plan[next_batch].first + plan[next_batch].batch <= uploaded) {
            const BatchPlan &bp = plan[next_batch++];
            a.first = bp.first; a.batch = bp.batch; a.lists = bp.lists;
            a.entry = bp.entry; a.maxlevel = bp.maxlevel;
            a.list_off = d_off + (bp.first - linked0);
            a.list_point = d_lp + bp.lists_off;
            a.list_level = d_ll + bp.lists_off;
            rc = nb ? build_run_batch_insert_nb(prm.metric, a, smem_search, smem_link, stream)
                    : build_run_batch_insert(prm.metric, a, smem_search, smem_link, stream);
            if (rc) break;
            launches += 3;
            const size_t last = (size_t)bp.first + bp.batch - 1;
            if (m.levels[last] > dev_maxlevel) {
                dev_entry = (uint32_t)last;
                dev_maxlevel = m.levels[last];
            }
            linked = (size_t)bp.first + bp.batch;
        }
    }
    if (rc == 0) {
        if (n_chunks == 0) B200_CUDA_OK(cudaEventRecord(e0, stream));
        B200_CUDA_OK(cudaEventRecord(e1, stream));
        B200_CUDA_OK(cudaStreamSynchronize(stream));
        if (BuildProfile *prof = build_profile()) {
            float t[3] = {0, 0, 0};
            for (size_t i = 0; i + 3 < prof->ev.size(); i += 4)
                for (int kx = 0; kx < 3; kx++) {
                    float ms = 0;
                    cudaEventElapsedTime(&ms, prof->ev[i + kx], prof->ev[i + kx + 1]);
                    t[kx] += ms;
                }
            {   // the last batch alone (full-size graph) and the ramp-up (batches below the maximum size)
                const size_t nb_ = prof->ev.size() / 4;
                float tl[3] = {0, 0, 0}, ramp = 0;
                for (int kx = 0; kx < 3 && nb_; kx++) cudaEventElapsedTime(&tl[kx], prof->ev[(nb_ - 1) * 4 + kx], prof->ev[(nb_ - 1) * 4 + kx + 1]);
                for (size_t i = 0; i < nb_ && i < plan.size(); i++)
                    if (plan[i].batch < max_batch) {
                        float ms = 0;
                        cudaEventElapsedTime(&ms, prof->ev[i * 4], prof->ev[i * 4 + 3]);
                        ramp += ms;
                    }
                fprintf(stderr, "[b200hnsw build profile] last batch (%u points): search %.2f ms, link %.2f ms, reverse %.2f ms; "
                        "batches below %zu points: %.1f ms\n", plan.empty() ? 0u : plan.back().batch, tl[0], tl[1], tl[2], max_batch, ramp);
            }
            fprintf(stderr, "[b200hnsw build profile] %zu batches: search %.1f ms, link %.1f ms, reverse %.1f ms; host: upload "
                    "%.1f ms, plan %.1f ms, flush so far %.1f ms\n",
                    prof->ev.size() / 4, t[0], t[1], t[2], ms_upload, ms_plan, since(t_start));
            for (cudaEvent_t e : prof->ev) cudaEventDestroy(e);
            prof->ev.clear();
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        B200_CUDA_OK(cudaMemcpy(w, bld.work, 64, cudaMemcpyDeviceToHost));
        std::lock_guard<std::mutex> sg(stats_mu);
        stats.queries = n_new;
        stats.dist_evals = w[0]; stats.hops_base = w[1]; stats.hops_upper = w[2]; stats.visited_resets = w[3];
        stats.dropped_reverse_edges = w[4];
        stats.kernel_launches += launches;
        stats.last_kernel_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) {  // nothing of this flush may still be running when the buffers are reused
        cudaStreamSynchronize(stream);
        cudaStreamSynchronize(bld.up_stream);
    }
    if (build_profile()) fprintf(stderr, "[b200hnsw build profile] flush total %.1f ms\n", since(t_start));
    dev.n = linked;
    mirror_dirty = true;
    flags_dirty = true;
    return rc;
}

void BuildScratch::release() {
    cudaFree(plevel); cudaFree(cand); cudaFree(cand_cnt); cudaFree(list_off); cudaFree(list_point);
    cudaFree(list_level); cudaFree(incnt); cudaFree(incoming); cudaFree(aff_node); cudaFree(aff_level);
    cudaFree(aff_count); cudaFree(work); cudaFree(batch_ids); cudaFree(newlists); cudaFree(stage_rows);
    cudaFree(stage_labels); cudaFree(plan_off); cudaFree(plan_lp); cudaFree(plan_ll);
    for (int b = 0; b < 2; b++) { cudaFree(up_raw[b]); if (up_ev[b]) cudaEventDestroy(up_ev[b]); }
    if (up_stream) cudaStreamDestroy(up_stream);
    *this = BuildScratch();
}

}  // namespace b200
