// oracle/hnsw_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (own code, written from the algorithm, not copied) of the reference hot
// path: distance spaces, saveIndex/loadIndex byte format, searchKnn, serial addPoint (including
// updatePoint for an existing label and replace_deleted) and BruteforceSearch.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
// load this library; the product (libb200hnsw.so) never links or calls it.
//
// PARITY PINNED: tests/test_oracle.py checks this restatement against (a) the golden
// fixtures under tests/golden/ (generated from the unmodified reference by
// tests/golden/make_golden.py) and SURVEY.md section 4's known answers (index sha256,
// entry point / max level for test.cpp's N=10000 run), and (b) oracle/_ref (the real
// reference compiled from /root/reference) whenever that binary is present.
//
// Each function cites the reference file:line (under /root/reference) it follows.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <queue>
#include <random>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace orc {

typedef uint32_t tableint;
typedef std::pair<float, tableint> cand_t;

// ---------------------------------------------------------------------------------------------
// Distances.  The shipped reference build has no -march, so USE_SSE only (hnswlib.h:11-21) and
// the 4-lane SSE kernels run: lane l accumulates elements i == l (mod 4) in index order with a
// separate multiply and add (no FMA), the horizontal sum is ((T0+T1)+T2)+T3
// (space_l2.h:97-143,165-190; space_ip.h:211-303).  Dispatch ladder: space_l2.h:214-238,
// space_ip.h:348-383.  Compile with -ffp-contract=off so the scalar code below rounds the same.
// ---------------------------------------------------------------------------------------------
static float l2_scalar(const float *a, const float *b, size_t n) {  // space_l2.h:6-20
    float res = 0;
    for (size_t i = 0; i < n; i++) {
        float t = a[i] - b[i];
        res += t * t;
    }
    return res;
}

static float ip_scalar(const float *a, const float *b, size_t n) {  // space_ip.h:6-14
    float res = 0;
    for (size_t i = 0; i < n; i++) res += a[i] * b[i];
    return res;
}

static float l2_lanes4(const float *a, const float *b, size_t n4) {  // n4 = multiple of 4
    float s[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < n4; i += 4)
        for (int l = 0; l < 4; l++) {
            float t = a[i + l] - b[i + l];
            float m = t * t;
            s[l] = s[l] + m;
        }
    return s[0] + s[1] + s[2] + s[3];
}

static float ip_lanes4(const float *a, const float *b, size_t n4) {
    float s[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < n4; i += 4)
        for (int l = 0; l < 4; l++) {
            float m = a[i + l] * b[i + l];
            s[l] = s[l] + m;
        }
    return s[0] + s[1] + s[2] + s[3];
}

float dist_l2(const float *a, const float *b, size_t d) {
    if (d % 16 == 0 || d % 4 == 0) return l2_lanes4(a, b, d);
    if (d > 16) {  // L2SqrSIMD16ExtResiduals, space_l2.h:149-160
        size_t q = d >> 4 << 4;
        float r = l2_lanes4(a, b, q);
        float t = l2_scalar(a + q, b + q, d - q);
        return r + t;
    }
    if (d > 4) {  // L2SqrSIMD4ExtResiduals, space_l2.h:192-205
        size_t q = d >> 2 << 2;
        float r = l2_lanes4(a, b, q);
        float t = l2_scalar(a + q, b + q, d - q);
        return r + t;
    }
    return l2_scalar(a, b, d);
}

float dist_ip(const float *a, const float *b, size_t d) {
    if (d % 16 == 0 || d % 4 == 0) return 1.0f - ip_lanes4(a, b, d);
    if (d > 16) {  // space_ip.h:313-325
        size_t q = d >> 4 << 4;
        float r = ip_lanes4(a, b, q);
        float t = ip_scalar(a + q, b + q, d - q);
        return 1.0f - (r + t);
    }
    if (d > 4) {  // space_ip.h:327-339
        size_t q = d >> 2 << 2;
        float r = ip_lanes4(a, b, q);
        float t = ip_scalar(a + q, b + q, d - q);
        return 1.0f - (r + t);
    }
    return 1.0f - ip_scalar(a, b, d);  // space_ip.h:16-19
}

struct CompareByFirst {  // hnswalg.h:165-170
    bool operator()(const cand_t &a, const cand_t &b) const { return a.first < b.first; }
};
typedef std::priority_queue<cand_t, std::vector<cand_t>, CompareByFirst> heap_t;

// ---------------------------------------------------------------------------------------------
// Index in the reference's memory layout (hnswalg.h:112-142): level-0 block of
// size_data_per_element records [u16 count|u8 flags|u8 0][maxM0 x u32][d x f32][u64 label],
// upper lists per element as level x [u32 count][maxM x u32].
// ---------------------------------------------------------------------------------------------
struct Index {
    int metric = 0;  // 0 = L2, 1 = inner product
    size_t dim = 0;
    size_t max_elements = 0, cur = 0;
    size_t M = 0, maxM = 0, maxM0 = 0, efc = 0, ef = 10;
    double mult = 0;
    int maxlevel = -1;
    tableint enterpoint = (tableint)-1;
    size_t size_links0 = 0, size_data = 0, size_links = 0, off_data = 0, off_label = 0, off_level0 = 0;
    std::vector<char> level0;
    std::vector<std::vector<char>> upper;
    std::vector<int> levels;
    std::unordered_map<uint64_t, tableint> label_lookup;
    std::default_random_engine level_rng;  // hnswalg.h:62,117
    size_t num_deleted = 0;
    bool allow_replace_deleted = false;               // hnswalg.h:73,93
    std::unordered_set<tableint> deleted_elements;    // hnswalg.h:75 (filled only with allow_replace_deleted)
    // counters (hnswalg.h:65-66 metric_* analogues)
    mutable uint64_t c_dist = 0, c_hops0 = 0, c_hops_up = 0;

    float dist(const float *a, const float *b) const {
        ++c_dist;
        return metric == 0 ? dist_l2(a, b, dim) : dist_ip(a, b, dim);
    }
    char *rec(tableint i) { return level0.data() + (size_t)i * size_data; }
    const char *rec(tableint i) const { return level0.data() + (size_t)i * size_data; }
    const float *vec(tableint i) const { return (const float *)(rec(i) + off_data); }
    uint64_t label(tableint i) const {
        uint64_t l;
        memcpy(&l, rec(i) + off_label, 8);
        return l;
    }
    bool deleted(tableint i) const { return ((const unsigned char *)rec(i))[2] & 1; }  // hnswalg.h:934-937
    uint32_t *list(tableint i, int level) {
        return level == 0 ? (uint32_t *)rec(i) : (uint32_t *)(upper[i].data() + (size_t)(level - 1) * size_links);
    }
    const uint32_t *list(tableint i, int level) const { return const_cast<Index *>(this)->list(i, level); }
    static unsigned count_of(const uint32_t *l) { return *(const unsigned short *)l; }  // hnswalg.h:940-942
    static void set_count(uint32_t *l, unsigned short c) { *(unsigned short *)l = c; }  // hnswalg.h:945-947

    void init(int metric_, size_t dim_, size_t max_el, size_t M_, size_t efc_, size_t seed) {  // hnswalg.h:89-144
        metric = metric_;
        dim = dim_;
        max_elements = max_el;
        M = M_ <= 10000 ? M_ : 10000;
        maxM = M;
        maxM0 = 2 * M;
        efc = std::max(efc_, M);
        ef = 10;
        level_rng.seed(seed);
        size_links0 = maxM0 * 4 + 4;
        size_data = size_links0 + dim * 4 + 8;
        off_data = size_links0;
        off_label = size_links0 + dim * 4;
        off_level0 = 0;
        size_links = maxM * 4 + 4;
        level0.assign(max_elements * size_data, 0);
        upper.assign(max_elements, {});
        levels.assign(max_elements, 0);
        mult = 1 / log(1.0 * M);
        cur = 0;
        maxlevel = -1;
        enterpoint = (tableint)-1;
    }

    int random_level() {  // hnswalg.h:207-211
        std::uniform_real_distribution<double> distribution(0.0, 1.0);
        double r = -log(distribution(level_rng)) * mult;
        return (int)r;
    }

    // hnswalg.h:225-305 (construction search on one layer, ef_construction wide)
    heap_t search_layer_build(tableint ep, const float *q, int layer, std::vector<uint32_t> &visited, uint32_t tag) {
        heap_t top, cset;
        float lower;
        if (!deleted(ep)) {
            float d = dist(q, vec(ep));
            top.emplace(d, ep);
            lower = d;
            cset.emplace(-d, ep);
        } else {
            lower = std::numeric_limits<float>::max();
            cset.emplace(-lower, ep);
        }
        visited[ep] = tag;
        while (!cset.empty()) {
            cand_t cur_pair = cset.top();
            if ((-cur_pair.first) > lower && top.size() == efc) break;
            cset.pop();
            const uint32_t *l = list(cur_pair.second, layer);
            unsigned n = count_of(l);
            for (unsigned j = 1; j <= n; j++) {
                tableint c = l[j];
                if (visited[c] == tag) continue;
                visited[c] = tag;
                float d1 = dist(q, vec(c));
                if (top.size() < efc || lower > d1) {
                    cset.emplace(-d1, c);
                    if (!deleted(c)) top.emplace(d1, c);
                    if (top.size() > efc) top.pop();
                    if (!top.empty()) lower = top.top().first;
                }
            }
        }
        return top;
    }

    // hnswalg.h:443-483
    void heuristic(heap_t &top, size_t Mlim) {
        if (top.size() < Mlim) return;
        std::priority_queue<cand_t> closest;  // default pair order on (-dist, id)
        std::vector<cand_t> keep;
        while (!top.empty()) {
            closest.emplace(-top.top().first, top.top().second);
            top.pop();
        }
        while (!closest.empty()) {
            if (keep.size() >= Mlim) break;
            cand_t c = closest.top();
            float dq = -c.first;
            closest.pop();
            bool good = true;
            for (const cand_t &r : keep) {
                float dd = dist(vec(r.second), vec(c.second));
                if (dd < dq) {
                    good = false;
                    break;
                }
            }
            if (good) keep.push_back(c);
        }
        for (const cand_t &c : keep) top.emplace(-c.first, c.second);
    }

    // hnswalg.h:506-630; is_update: the list of cur_c is replaced, a neighbour that already lists cur_c is left alone
    tableint connect(tableint cur_c, heap_t &top, int level, bool is_update = false) {
        size_t Mcur = level ? maxM : maxM0;
        heuristic(top, M);
        std::vector<tableint> sel;
        while (!top.empty()) {
            sel.push_back(top.top().second);
            top.pop();
        }
        tableint next_ep = sel.back();
        uint32_t *lc = list(cur_c, level);
        set_count(lc, (unsigned short)sel.size());
        for (size_t i = 0; i < sel.size(); i++) lc[1 + i] = sel[i];
        for (size_t i = 0; i < sel.size(); i++) {
            uint32_t *lo = list(sel[i], level);
            size_t sz = count_of(lo);
            uint32_t *data = lo + 1;
            if (is_update) {  // is_cur_c_present, hnswalg.h:566-580
                bool present = false;
                for (size_t j = 0; j < sz; j++) present |= data[j] == cur_c;
                if (present) continue;
            }
            if (sz < Mcur) {
                data[sz] = cur_c;
                set_count(lo, (unsigned short)(sz + 1));
            } else {
                float dmax = dist(vec(cur_c), vec(sel[i]));
                heap_t cands;
                cands.emplace(dmax, cur_c);
                for (size_t j = 0; j < sz; j++) cands.emplace(dist(vec(data[j]), vec(sel[i])), data[j]);
                heuristic(cands, Mcur);
                int idx = 0;
                while (!cands.empty()) {
                    data[idx++] = cands.top().second;
                    cands.pop();
                }
                set_count(lo, (unsigned short)idx);
            }
        }
        return next_ep;
    }

    std::vector<uint32_t> visited_build;
    uint32_t visited_tag = 0;

    void unmark_deleted(tableint id) {  // hnswalg.h:903-917
        ((unsigned char *)rec(id))[2] &= (unsigned char)~1;
        num_deleted--;
        if (allow_replace_deleted) deleted_elements.erase(id);
    }

    // updatePoint, hnswalg.h:995-1072, updateNeighborProbability = 1.0 (the value addPoint passes).  The containers are
    // the reference's (unordered sets iterated in their own order, max-heaps compared by distance only) so that ties
    // resolve the same way.
    void update_point(const float *x, tableint id) {
        memcpy(rec(id) + off_data, x, dim * 4);
        int max_level_copy = maxlevel;
        tableint entry_copy = enterpoint;
        if (entry_copy == id && cur == 1) return;
        int elem_level = levels[id];
        for (int layer = 0; layer <= elem_level; layer++) {
            std::unordered_set<tableint> s_cand, s_neigh;
            const uint32_t *l1 = list(id, layer);
            std::vector<tableint> one_hop(l1 + 1, l1 + 1 + count_of(l1));
            if (one_hop.empty()) continue;
            s_cand.insert(id);
            for (tableint el : one_hop) {
                s_cand.insert(el);
                s_neigh.insert(el);
                const uint32_t *l2 = list(el, layer);
                unsigned n2 = count_of(l2);
                for (unsigned j = 1; j <= n2; j++) s_cand.insert(l2[j]);
            }
            for (tableint neigh : s_neigh) {
                heap_t cands;
                size_t size = s_cand.find(neigh) == s_cand.end() ? s_cand.size() : s_cand.size() - 1;
                size_t keep = std::min(efc, size);
                for (tableint c : s_cand) {
                    if (c == neigh) continue;
                    float d = dist(vec(neigh), vec(c));
                    if (cands.size() < keep) {
                        cands.emplace(d, c);
                    } else if (d < cands.top().first) {
                        cands.pop();
                        cands.emplace(d, c);
                    }
                }
                heuristic(cands, layer == 0 ? maxM0 : maxM);
                uint32_t *ln = list(neigh, layer);
                size_t cs = cands.size();
                set_count(ln, (unsigned short)cs);
                for (size_t idx = 0; idx < cs; idx++) {
                    ln[1 + idx] = cands.top().second;
                    cands.pop();
                }
            }
        }
        repair_connections(x, entry_copy, id, elem_level, max_level_copy);
    }

    // repairConnectionsForUpdate, hnswalg.h:1075-1139
    void repair_connections(const float *x, tableint entry, tableint id, int elem_level, int max_level) {
        tableint cur_obj = entry;
        if (elem_level < max_level) {
            float curdist = dist(x, vec(cur_obj));
            for (int level = max_level; level > elem_level; level--) {
                bool changed = true;
                while (changed) {
                    changed = false;
                    const uint32_t *l = list(cur_obj, level);
                    unsigned n = count_of(l);
                    for (unsigned i = 1; i <= n; i++) {
                        float d = dist(x, vec(l[i]));
                        if (d < curdist) {
                            curdist = d;
                            cur_obj = l[i];
                            changed = true;
                        }
                    }
                }
            }
        }
        if (visited_build.size() != max_elements) visited_build.assign(max_elements, 0);
        for (int level = elem_level; level >= 0; level--) {
            ++visited_tag;
            heap_t top = search_layer_build(cur_obj, x, level, visited_build, visited_tag);
            heap_t filtered;
            while (!top.empty()) {
                if (top.top().second != id) filtered.push(top.top());
                top.pop();
            }
            if (!filtered.empty()) {
                if (deleted(entry)) {
                    filtered.emplace(dist(x, vec(entry)), entry);
                    if (filtered.size() > efc) filtered.pop();
                }
                cur_obj = connect(id, filtered, level, true);
            }
        }
    }

    // addPoint(data, label, replace_deleted = true), hnswalg.h:954-992.  Returns -3 when replacement is disabled.
    int add_point_replace(const float *x, uint64_t lab) {
        if (!allow_replace_deleted) return -3;
        if (deleted_elements.empty()) return add_point(x, lab);
        tableint id = *deleted_elements.begin();
        deleted_elements.erase(id);
        uint64_t old = label(id);
        memcpy(rec(id) + off_label, &lab, 8);
        label_lookup.erase(old);
        label_lookup[lab] = id;
        unmark_deleted(id);
        update_point(x, id);
        return 0;
    }

    // hnswalg.h:1153-1267; an existing label is updated (:1157-1174)
    int add_point(const float *x, uint64_t lab) {
        auto known = label_lookup.find(lab);
        if (known != label_lookup.end()) {
            tableint id = known->second;
            if (allow_replace_deleted && deleted(id)) return -4;  // "Can't use addPoint to update deleted elements ..."
            if (deleted(id)) unmark_deleted(id);
            update_point(x, id);
            return 0;
        }
        if (cur >= max_elements) return -1;      // "The number of elements exceeds the specified limit"
        tableint cur_c = (tableint)cur++;
        label_lookup[lab] = cur_c;
        int curlevel = random_level();
        levels[cur_c] = curlevel;
        int maxlevelcopy = maxlevel;
        tableint cur_obj = enterpoint;
        memset(rec(cur_c), 0, size_data);
        memcpy(rec(cur_c) + off_label, &lab, 8);
        memcpy(rec(cur_c) + off_data, x, dim * 4);
        if (curlevel) upper[cur_c].assign(size_links * curlevel, 0);
        if (visited_build.size() != max_elements) visited_build.assign(max_elements, 0);
        if ((int)cur_obj != -1) {
            if (curlevel < maxlevelcopy) {
                float curdist = dist(x, vec(cur_obj));
                for (int level = maxlevelcopy; level > curlevel; level--) {
                    bool changed = true;
                    while (changed) {
                        changed = false;
                        const uint32_t *l = list(cur_obj, level);
                        unsigned n = count_of(l);
                        for (unsigned i = 1; i <= n; i++) {
                            float d = dist(x, vec(l[i]));
                            if (d < curdist) {
                                curdist = d;
                                cur_obj = l[i];
                                changed = true;
                            }
                        }
                    }
                }
            }
            for (int level = std::min(curlevel, maxlevelcopy); level >= 0; level--) {
                ++visited_tag;
                heap_t top = search_layer_build(cur_obj, x, level, visited_build, visited_tag);
                cur_obj = connect(cur_c, top, level);
            }
        } else {
            enterpoint = 0;
            maxlevel = curlevel;
        }
        if (curlevel > maxlevelcopy) {
            enterpoint = cur_c;
            maxlevel = curlevel;
        }
        return 0;
    }

    // hnswalg.h:685-713
    int save(const char *path) const {
        FILE *f = fopen(path, "wb");
        if (!f) return -1;
        uint64_t u;
        auto w8 = [&](uint64_t v) { u = v; fwrite(&u, 8, 1, f); };
        w8(off_level0); w8(max_elements); w8(cur); w8(size_data); w8(off_label); w8(off_data);
        int32_t ml = maxlevel; fwrite(&ml, 4, 1, f);
        uint32_t ep = enterpoint; fwrite(&ep, 4, 1, f);
        w8(maxM); w8(maxM0); w8(M);
        fwrite(&mult, 8, 1, f);
        w8(efc);
        fwrite(level0.data(), 1, cur * size_data, f);
        for (size_t i = 0; i < cur; i++) {
            uint32_t sz = levels[i] > 0 ? (uint32_t)(size_links * levels[i]) : 0;
            fwrite(&sz, 4, 1, f);
            if (sz) fwrite(upper[i].data(), 1, sz, f);
        }
        fclose(f);
        return 0;
    }

    // hnswalg.h:716-822. Returns 0, -1 "Cannot open file", -2 "Index seems to be corrupted or unsupported".
    int load(const char *path, int metric_, size_t dim_, size_t max_el_arg) {
        FILE *f = fopen(path, "rb");
        if (!f) return -1;
        fseek(f, 0, SEEK_END);
        long total = ftell(f);
        fseek(f, 0, SEEK_SET);
        metric = metric_;
        dim = dim_;
        uint64_t h[6];
        if (fread(h, 8, 6, f) != 6) { fclose(f); return -2; }
        off_level0 = h[0]; max_elements = h[1]; cur = h[2]; size_data = h[3]; off_label = h[4]; off_data = h[5];
        size_t max_el = max_el_arg;
        if (max_el < cur) max_el = max_elements;
        max_elements = max_el;
        int32_t ml; uint32_t ep; uint64_t t[3]; uint64_t e;
        if (fread(&ml, 4, 1, f) != 1 || fread(&ep, 4, 1, f) != 1 || fread(t, 8, 3, f) != 3 ||
            fread(&mult, 8, 1, f) != 1 || fread(&e, 8, 1, f) != 1) { fclose(f); return -2; }
        maxlevel = ml; enterpoint = ep; maxM = t[0]; maxM0 = t[1]; M = t[2]; efc = e;
        size_links = maxM * 4 + 4;
        size_links0 = maxM0 * 4 + 4;
        long pos = ftell(f);
        // integrity walk, hnswalg.h:754-770
        long p = pos + (long)(cur * size_data);
        for (size_t i = 0; i < cur; i++) {
            if (p < 0 || p >= total) { fclose(f); return -2; }
            fseek(f, p, SEEK_SET);
            uint32_t sz;
            if (fread(&sz, 4, 1, f) != 1) { fclose(f); return -2; }
            p += 4 + sz;
        }
        if (p != total) { fclose(f); return -2; }
        fseek(f, pos, SEEK_SET);
        level0.assign(max_elements * size_data, 0);
        if (fread(level0.data(), 1, cur * size_data, f) != cur * size_data) { fclose(f); return -2; }
        upper.assign(max_elements, {});
        levels.assign(max_elements, 0);
        ef = 10;
        label_lookup.clear();
        num_deleted = 0;
        for (size_t i = 0; i < cur; i++) {
            label_lookup[label((tableint)i)] = (tableint)i;
            uint32_t sz;
            if (fread(&sz, 4, 1, f) != 1) { fclose(f); return -2; }
            if (sz) {
                levels[i] = (int)(sz / size_links);
                upper[i].resize(sz);
                if (fread(upper[i].data(), 1, sz, f) != sz) { fclose(f); return -2; }
            }
        }
        for (size_t i = 0; i < cur; i++) if (deleted((tableint)i)) num_deleted++;
        fclose(f);
        return 0;
    }

    // hnswalg.h:309-440 (bare_bone_search=true when nothing is deleted, else the non-bare branch
    // without filter / stop condition)
    heap_t search_layer0(tableint ep, const float *q, size_t ef_, bool bare, std::vector<uint32_t> &visited, uint32_t tag) const {
        heap_t top, cset;
        float lower;
        if (bare || !deleted(ep)) {
            float d = dist(q, vec(ep));
            lower = d;
            top.emplace(d, ep);
            cset.emplace(-d, ep);
        } else {
            lower = std::numeric_limits<float>::max();
            cset.emplace(-lower, ep);
        }
        visited[ep] = tag;
        while (!cset.empty()) {
            cand_t cp = cset.top();
            float cd = -cp.first;
            bool stop = bare ? (cd > lower) : (cd > lower && top.size() == ef_);
            if (stop) break;
            cset.pop();
            const uint32_t *l = list(cp.second, 0);
            unsigned n = count_of(l);
            ++c_hops0;
            for (unsigned j = 1; j <= n; j++) {
                tableint c = l[j];
                if (visited[c] == tag) continue;
                visited[c] = tag;
                float d = dist(q, vec(c));
                if (top.size() < ef_ || lower > d) {
                    cset.emplace(-d, c);
                    if (bare || !deleted(c)) top.emplace(d, c);
                    while (top.size() > ef_) top.pop();
                    if (!top.empty()) lower = top.top().first;
                }
            }
        }
        return top;
    }

    // hnswalg.h:1270-1324. out rows closest-first; returns the number of results.
    size_t search_knn(const float *q, size_t k, std::vector<uint32_t> &visited, uint32_t tag,
                      uint64_t *labels, float *dists) const {
        if (cur == 0) return 0;
        tableint cur_obj = enterpoint;
        float curdist = dist(q, vec(enterpoint));
        for (int level = maxlevel; level > 0; level--) {
            bool changed = true;
            while (changed) {
                changed = false;
                const uint32_t *l = list(cur_obj, level);
                unsigned n = count_of(l);
                ++c_hops_up;
                for (unsigned i = 1; i <= n; i++) {
                    float d = dist(q, vec(l[i]));
                    if (d < curdist) {
                        curdist = d;
                        cur_obj = l[i];
                        changed = true;
                    }
                }
            }
        }
        bool bare = num_deleted == 0;
        heap_t top = search_layer0(cur_obj, q, std::max(ef, k), bare, visited, tag);
        while (top.size() > k) top.pop();
        std::priority_queue<std::pair<float, uint64_t>> result;
        while (!top.empty()) {
            result.push(std::make_pair(top.top().first, label(top.top().second)));
            top.pop();
        }
        size_t n = result.size(), i = n;
        while (!result.empty()) {
            --i;
            labels[i] = result.top().second;
            dists[i] = result.top().first;
            result.pop();
        }
        return n;
    }
};

// bruteforce.h:10-172 restated: rows [vector][u64 label], linear scan with a max-heap of pairs.
struct Brute {
    int metric = 0;
    size_t dim = 0, maxel = 0, cur = 0, row = 0;
    std::vector<char> data;
    std::unordered_map<uint64_t, size_t> lookup;
    float dist(const float *a, const float *b) const { return metric == 0 ? dist_l2(a, b, dim) : dist_ip(a, b, dim); }
    void init(int m, size_t d, size_t n) {
        metric = m; dim = d; maxel = n; cur = 0; row = d * 4 + 8;
        data.assign(n * row, 0);
    }
    int add(const float *x, uint64_t lab) {  // bruteforce.h:64-83
        size_t idx;
        auto it = lookup.find(lab);
        if (it != lookup.end()) idx = it->second;
        else {
            if (cur >= maxel) return -1;
            idx = cur++;
            lookup[lab] = idx;
        }
        memcpy(data.data() + row * idx + dim * 4, &lab, 8);
        memcpy(data.data() + row * idx, x, dim * 4);
        return 0;
    }
    void remove(uint64_t lab) {  // bruteforce.h:86-103
        auto it = lookup.find(lab);
        if (it == lookup.end()) return;
        size_t c = it->second;
        lookup.erase(it);
        uint64_t last;
        memcpy(&last, data.data() + row * (cur - 1) + dim * 4, 8);
        lookup[last] = c;
        memmove(data.data() + row * c, data.data() + row * (cur - 1), row);
        cur--;
    }
    size_t search(const float *q, size_t k, uint64_t *labels, float *dists) const {  // bruteforce.h:106-135
        std::priority_queue<std::pair<float, uint64_t>> top;
        if (cur == 0) return 0;
        for (size_t i = 0; i < k && i < cur; i++) {
            uint64_t lab;
            memcpy(&lab, data.data() + row * i + dim * 4, 8);
            top.emplace(dist(q, (const float *)(data.data() + row * i)), lab);
        }
        float last = top.empty() ? std::numeric_limits<float>::max() : top.top().first;
        for (size_t i = k; i < cur; i++) {
            float d = dist(q, (const float *)(data.data() + row * i));
            if (d <= last) {
                uint64_t lab;
                memcpy(&lab, data.data() + row * i + dim * 4, 8);
                top.emplace(d, lab);
                if (top.size() > k) top.pop();
                if (!top.empty()) last = top.top().first;
            }
        }
        size_t n = top.size(), i = n;
        while (!top.empty()) {
            --i;
            labels[i] = top.top().second;
            dists[i] = top.top().first;
            top.pop();
        }
        return n;
    }
    int save(const char *path) const {  // bruteforce.h:138-149
        FILE *f = fopen(path, "wb");
        if (!f) return -1;
        uint64_t h[3] = {maxel, row, cur};
        fwrite(h, 8, 3, f);
        fwrite(data.data(), 1, maxel * row, f);
        fclose(f);
        return 0;
    }
};

}  // namespace orc

extern "C" {

float orc_dist(int metric, size_t dim, const float *a, const float *b) {
    return metric == 0 ? orc::dist_l2(a, b, dim) : orc::dist_ip(a, b, dim);
}

void *orc_hnsw_new(int metric, size_t dim, size_t max_elements, size_t M, size_t efc, size_t seed) {
    orc::Index *ix = new orc::Index();
    ix->init(metric, dim, max_elements, M, efc, seed);
    return ix;
}

void *orc_hnsw_load(int metric, size_t dim, const char *path, size_t max_elements, int *rc) {
    orc::Index *ix = new orc::Index();
    int r = ix->load(path, metric, dim, max_elements);
    if (rc) *rc = r;
    if (r) { delete ix; return nullptr; }
    return ix;
}

void orc_hnsw_free(void *h) { delete (orc::Index *)h; }

int orc_hnsw_add(void *h, const float *X, const uint64_t *labels, size_t n) {
    orc::Index *ix = (orc::Index *)h;
    for (size_t i = 0; i < n; i++) {
        int r = ix->add_point(X + i * ix->dim, labels ? labels[i] : i);
        if (r) return r;
    }
    return 0;
}

int orc_hnsw_save(void *h, const char *path) { return ((orc::Index *)h)->save(path); }

void orc_hnsw_info(void *h, int64_t *info) {
    orc::Index *ix = (orc::Index *)h;
    info[0] = (int64_t)ix->cur; info[1] = (int64_t)ix->max_elements; info[2] = ix->maxlevel;
    info[3] = (int64_t)ix->enterpoint; info[4] = (int64_t)ix->M; info[5] = (int64_t)ix->maxM0;
    info[6] = (int64_t)ix->efc; info[7] = (int64_t)ix->size_data;
}

int orc_hnsw_levels(void *h, int32_t *out) {
    orc::Index *ix = (orc::Index *)h;
    for (size_t i = 0; i < ix->cur; i++) out[i] = ix->levels[i];
    return 0;
}

int orc_hnsw_links(void *h, uint32_t id, int level, uint32_t *out, int cap) {
    orc::Index *ix = (orc::Index *)h;
    if (id >= ix->cur || level > ix->levels[id]) return -1;
    const uint32_t *l = ix->list(id, level);
    int cnt = (int)orc::Index::count_of(l);
    for (int j = 0; j < cnt && j < cap; j++) out[j] = l[1 + j];
    return cnt;
}

int orc_hnsw_mark_delete(void *h, uint64_t label) {  // hnswalg.h:853-883
    orc::Index *ix = (orc::Index *)h;
    auto it = ix->label_lookup.find(label);
    if (it == ix->label_lookup.end()) return -1;
    unsigned char *p = (unsigned char *)ix->rec(it->second) + 2;
    if (*p & 1) return -2;
    *p |= 1;
    ix->num_deleted++;
    if (ix->allow_replace_deleted) ix->deleted_elements.insert(it->second);  // hnswalg.h:876-879
    return 0;
}

void orc_hnsw_allow_replace_deleted(void *h, int on) { ((orc::Index *)h)->allow_replace_deleted = on != 0; }

int orc_hnsw_add_replace_deleted(void *h, const float *X, const uint64_t *labels, size_t n) {
    orc::Index *ix = (orc::Index *)h;
    for (size_t i = 0; i < n; i++) {
        int r = ix->add_point_replace(X + i * ix->dim, labels ? labels[i] : i);
        if (r) return r;
    }
    return 0;
}

// Serial batched searchKnn with work counters per query: D = distance evaluations (all
// layers), H0 = base-layer expansions, Hup = upper-layer list scans.
int orc_hnsw_search(void *h, const float *Q, size_t nq, size_t k, size_t ef, uint64_t *labels,
                    float *dists, uint32_t *counts, uint32_t *D, uint32_t *H0, uint32_t *Hup,
                    double *seconds) {
    orc::Index *ix = (orc::Index *)h;
    ix->ef = ef;
    std::vector<uint32_t> visited(ix->max_elements ? ix->max_elements : 1, 0);
    auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < nq; i++) {
        uint64_t d0 = ix->c_dist, a0 = ix->c_hops0, u0 = ix->c_hops_up;
        for (size_t j = 0; j < k; j++) {
            labels[i * k + j] = UINT64_MAX;
            dists[i * k + j] = std::numeric_limits<float>::infinity();
        }
        size_t n = ix->search_knn(Q + i * ix->dim, k, visited, (uint32_t)(i + 1), labels + i * k, dists + i * k);
        if (counts) counts[i] = (uint32_t)n;
        if (D) D[i] = (uint32_t)(ix->c_dist - d0);
        if (H0) H0[i] = (uint32_t)(ix->c_hops0 - a0);
        if (Hup) Hup[i] = (uint32_t)(ix->c_hops_up - u0);
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

void *orc_bf_new(int metric, size_t dim, size_t max_elements) {
    orc::Brute *b = new orc::Brute();
    b->init(metric, dim, max_elements);
    return b;
}
void orc_bf_free(void *h) { delete (orc::Brute *)h; }
int orc_bf_add(void *h, const float *X, const uint64_t *labels, size_t n) {
    orc::Brute *b = (orc::Brute *)h;
    for (size_t i = 0; i < n; i++) {
        int r = b->add(X + i * b->dim, labels ? labels[i] : i);
        if (r) return r;
    }
    return 0;
}
int orc_bf_remove(void *h, uint64_t label) { ((orc::Brute *)h)->remove(label); return 0; }
int orc_bf_save(void *h, const char *path) { return ((orc::Brute *)h)->save(path); }
int64_t orc_bf_count(void *h) { return (int64_t)((orc::Brute *)h)->cur; }
int orc_bf_search(void *h, const float *Q, size_t nq, size_t k, uint64_t *labels, float *dists,
                  uint32_t *counts, double *seconds) {
    orc::Brute *b = (orc::Brute *)h;
    auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < nq; i++) {
        for (size_t j = 0; j < k; j++) {
            labels[i * k + j] = UINT64_MAX;
            dists[i * k + j] = std::numeric_limits<float>::infinity();
        }
        size_t n = b->search(Q + i * b->dim, k, labels + i * k, dists + i * k);
        if (counts) counts[i] = (uint32_t)n;
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

}  // extern "C"
