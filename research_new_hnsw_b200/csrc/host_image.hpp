// host_image.hpp -- host mirror of the index in the reference's memory layout.
//
// The reference's consumers read public fields and walk raw link-list pointers (SURVEY.md 8(b);
// build.cpp:24-36,51-99), and saveIndex/loadIndex must stay byte-compatible both ways (hnswalg.h:685-822).
// The mirror therefore keeps exactly the reference layout (hnswalg.h:112-142):
//   level-0 record  = [u16 count | u8 flags | u8 0][maxM0 x u32 nbrs][dim x f32 vector][u64 label]
//   upper lists     = per element, levels x ([u32 count][maxM x u32 nbrs])
// The device works on a de-interleaved SoA copy (device_index.cuh); this file is I/O and bookkeeping only --
// no distance is ever evaluated on the host.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

namespace b200 {

struct HostImage {
    size_t dim = 0;
    size_t max_elements = 0, cur = 0, num_deleted = 0;
    size_t M = 0, maxM = 0, maxM0 = 0, efc = 0;
    double mult = 0.0;
    int maxlevel = -1;
    uint32_t enterpoint = (uint32_t)-1;
    size_t size_links0 = 0, size_data = 0, size_links = 0, off_data = 0, off_label = 0, off_level0 = 0;
    char *level0 = nullptr;                 // max_elements * size_data bytes
    std::vector<std::vector<char>> upper;   // [max_elements], levels[i] * size_links bytes each
    std::vector<int> levels;                // element_levels_ (hnswalg.h:52)
    std::unordered_map<uint64_t, uint32_t> label_lookup;  // hnswalg.h:60
    std::default_random_engine level_rng;   // hnswalg.h:62 (std::minstd_rand0)

    ~HostImage() { free(level0); }
    HostImage() = default;
    HostImage(const HostImage &) = delete;
    HostImage &operator=(const HostImage &) = delete;

    char *rec(size_t i) const { return level0 + i * size_data; }
    float *vec(size_t i) const { return (float *)(rec(i) + off_data); }
    uint64_t label(size_t i) const {
        uint64_t l;
        memcpy(&l, rec(i) + off_label, 8);
        return l;
    }
    bool deleted(size_t i) const { return ((const unsigned char *)rec(i))[2] & 1; }  // hnswalg.h:934-937
    uint32_t *list(size_t i, int level) const {  // hnswalg.h:486-503
        return level == 0 ? (uint32_t *)rec(i) : (uint32_t *)(upper[i].data() + (size_t)(level - 1) * size_links);
    }
    static unsigned short count_of(const uint32_t *l) { return *(const unsigned short *)l; }
    static void set_count(uint32_t *l, unsigned short c) { *(unsigned short *)l = c; }

    // hnswalg.h:89-144
    bool init(size_t dim_, size_t max_el, size_t M_, size_t efc_, size_t seed) {
        dim = dim_;
        max_elements = max_el;
        M = M_ <= 10000 ? M_ : 10000;
        maxM = M;
        maxM0 = M * 2;
        efc = efc_ > M ? efc_ : M;
        level_rng.seed(seed);
        size_links0 = maxM0 * 4 + 4;
        size_data = size_links0 + dim * 4 + 8;
        off_data = size_links0;
        off_label = size_links0 + dim * 4;
        off_level0 = 0;
        size_links = maxM * 4 + 4;
        mult = 1 / log(1.0 * (double)M);
        cur = 0;
        num_deleted = 0;
        maxlevel = -1;
        enterpoint = (uint32_t)-1;
        free(level0);
        level0 = (char *)malloc(max_elements * size_data + 1);
        if (!level0) return false;
        upper.assign(max_elements, {});
        levels.assign(max_elements, 0);
        label_lookup.clear();
        label_lookup.reserve(max_elements);  // addPoint of a million labels: no rehashing on the way
        return true;
    }

    int random_level() {  // hnswalg.h:207-211
        std::uniform_real_distribution<double> distribution(0.0, 1.0);
        double r = -log(distribution(level_rng)) * mult;
        return (int)r;
    }

    // hnswalg.h:658-683
    uint64_t file_size() const {
        uint64_t s = 96 + (uint64_t)cur * size_data;
        for (size_t i = 0; i < cur; i++) s += 4 + (levels[i] > 0 ? size_links * levels[i] : 0);
        return s;
    }

    // hnswalg.h:685-713
    int save(const char *path) const {
        FILE *f = fopen(path, "wb");
        if (!f) return -1;
        auto w8 = [&](uint64_t v) { fwrite(&v, 8, 1, f); };
        w8(off_level0); w8(max_elements); w8(cur); w8(size_data); w8(off_label); w8(off_data);
        int32_t ml = maxlevel;
        uint32_t ep = enterpoint;
        fwrite(&ml, 4, 1, f);
        fwrite(&ep, 4, 1, f);
        w8(maxM); w8(maxM0); w8(M);
        fwrite(&mult, 8, 1, f);
        w8(efc);
        if (cur) fwrite(level0, 1, cur * size_data, f);
        for (size_t i = 0; i < cur; i++) {
            uint32_t sz = levels[i] > 0 ? (uint32_t)(size_links * levels[i]) : 0;
            fwrite(&sz, 4, 1, f);
            if (sz) fwrite(upper[i].data(), 1, sz, f);
        }
        int bad = ferror(f);
        fclose(f);
        return bad ? -1 : 0;
    }

    // hnswalg.h:716-822.  0 ok, -1 cannot open, -2 corrupted/unsupported, -3 out of memory.
    int load(const char *path, size_t dim_, size_t max_el_arg) {
        FILE *f = fopen(path, "rb");
        if (!f) return -1;
        fseek(f, 0, SEEK_END);
        long long total = ftell(f);
        fseek(f, 0, SEEK_SET);
        dim = dim_;
        uint64_t h6[6], t3[3], e;
        int32_t ml;
        uint32_t ep;
        bool ok = fread(h6, 8, 6, f) == 6 && fread(&ml, 4, 1, f) == 1 && fread(&ep, 4, 1, f) == 1 &&
                  fread(t3, 8, 3, f) == 3 && fread(&mult, 8, 1, f) == 1 && fread(&e, 8, 1, f) == 1;
        if (!ok) { fclose(f); return -2; }
        off_level0 = h6[0]; max_elements = h6[1]; cur = h6[2]; size_data = h6[3]; off_label = h6[4]; off_data = h6[5];
        maxlevel = ml; enterpoint = ep; maxM = t3[0]; maxM0 = t3[1]; M = t3[2]; efc = e;
        size_t max_el = max_el_arg;
        if (max_el < cur) max_el = max_elements;  // hnswalg.h:732-735
        max_elements = max_el;
        size_links = maxM * 4 + 4;
        size_links0 = maxM0 * 4 + 4;
        // The GPU layout needs the record to be what the space implies; anything else is "unsupported".
        if (off_level0 != 0 || off_data != size_links0 || off_label != size_links0 + dim * 4 ||
            size_data != size_links0 + dim * 4 + 8 || cur > max_elements || maxM0 > 65535) {
            fclose(f);
            return -2;
        }
        long long pos = 96;
        long long p = pos + (long long)(cur * size_data);  // integrity walk, hnswalg.h:754-770
        for (size_t i = 0; i < cur; i++) {
            if (p < 0 || p >= total) { fclose(f); return -2; }
            fseek(f, p, SEEK_SET);
            uint32_t sz;
            if (fread(&sz, 4, 1, f) != 1) { fclose(f); return -2; }
            p += 4 + (long long)sz;
        }
        if (p != total) { fclose(f); return -2; }
        fseek(f, pos, SEEK_SET);
        free(level0);
        level0 = (char *)malloc(max_elements * size_data + 1);
        if (!level0) { fclose(f); return -3; }
        if (cur && fread(level0, 1, cur * size_data, f) != cur * size_data) { fclose(f); return -2; }
        upper.assign(max_elements, {});
        levels.assign(max_elements, 0);
        label_lookup.clear();
        label_lookup.reserve(cur * 2);
        num_deleted = 0;
        for (size_t i = 0; i < cur; i++) {
            label_lookup[label(i)] = (uint32_t)i;
            uint32_t sz;
            if (fread(&sz, 4, 1, f) != 1) { fclose(f); return -2; }
            if (sz) {
                if (sz % size_links) { fclose(f); return -2; }
                levels[i] = (int)(sz / size_links);
                upper[i].resize(sz);
                if (fread(upper[i].data(), 1, sz, f) != sz) { fclose(f); return -2; }
            }
        }
        for (size_t i = 0; i < cur; i++)
            if (deleted(i)) num_deleted++;
        fclose(f);
        return 0;
    }

    // hnswalg.h:633-656
    bool resize(size_t new_max) {
        char *p = (char *)realloc(level0, new_max * size_data + 1);
        if (!p) return false;
        level0 = p;
        upper.resize(new_max);
        levels.resize(new_max, 0);
        max_elements = new_max;
        return true;
    }
};

// bruteforce.h:10-172 host mirror: rows [vector][u64 label], same bytes as the reference's data_.
struct HostBrute {
    size_t dim = 0, maxel = 0, cur = 0, row = 0;
    std::vector<char> data;
    std::unordered_map<uint64_t, size_t> lookup;
    void init(size_t d, size_t n) {
        dim = d; maxel = n; cur = 0; row = d * 4 + 8;
        data.assign(n * row, 0);
        lookup.clear();
    }
    uint64_t label(size_t i) const {  // bruteforce.h:51-52,81-82: the label sits behind the vector
        uint64_t l;
        memcpy(&l, data.data() + i * row + dim * 4, 8);
        return l;
    }
    int save(const char *path) const {  // bruteforce.h:138-149
        FILE *f = fopen(path, "wb");
        if (!f) return -1;
        uint64_t h[3] = {maxel, row, cur};
        fwrite(h, 8, 3, f);
        fwrite(data.data(), 1, maxel * row, f);
        int bad = ferror(f);
        fclose(f);
        return bad ? -1 : 0;
    }
    int load(const char *path, size_t d) {  // bruteforce.h:152-171
        FILE *f = fopen(path, "rb");
        if (!f) return -1;
        uint64_t h[3];
        if (fread(h, 8, 3, f) != 3) { fclose(f); return -2; }
        dim = d; maxel = h[0]; cur = h[2]; row = d * 4 + 8;
        if (h[1] != row || cur > maxel) { fclose(f); return -2; }
        data.assign(maxel * row, 0);
        if (fread(data.data(), 1, maxel * row, f) != maxel * row) { fclose(f); return -2; }
        fclose(f);
        lookup.clear();
        for (size_t i = 0; i < cur; i++) {
            uint64_t lab;
            memcpy(&lab, data.data() + i * row + dim * 4, 8);
            lookup[lab] = i;
        }
        return 0;
    }
};

}  // namespace b200
