// TEST DOUBLE (never linked into the product): the handful of liborbb200 entry points that orb_slam3_ros_b200/host/ORBmatcherGPU.cc calls,
// implemented on the CPU, so that the HOST logic of the compiled matcher replacements -- query construction, the in-order decision loops,
// rescans and fallbacks -- can be compared with the reference's own function bodies where there is no GPU (tests/test_matcher_host_cpu.py).
// Semantics as declared in include/orbb200.h: orbb_search_area_topk walks the 64 x 48 grid the way Frame::GetFeaturesInArea does
// (Frame.cc:657-723: cell columns, then rows, then key point index) and keeps the k nearest in (distance, visit order); orbb_best2_csr keeps
// best / second best with strict '<' in list order.  The GPU tests run the same comparisons through the real library.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "orbb200.h"

struct orbb_matcher {
    std::string err;
    struct Slot { std::vector<float> xy, ur; std::vector<int32_t> oct; std::vector<uint8_t> desc; } slot[ORBB_FRAME_SLOTS];
};

extern "C" {

int orbb_hamming_distance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

int orbb_matcher_create(int, orbb_matcher** out) { *out = new orbb_matcher(); return ORBB_OK; }
void orbb_matcher_destroy(orbb_matcher* m) { delete m; }
const char* orbb_matcher_last_error(const orbb_matcher* m) { return m ? m->err.c_str() : "fake"; }

int orbb_frame_upload(orbb_matcher* m, int slot, const orbb_frame_view* h, orbb_frame_view* dev) {
    if (!m || slot < 0 || slot >= ORBB_FRAME_SLOTS || !h || !dev) return ORBB_ERR_ARG;
    orbb_matcher::Slot& s = m->slot[slot];
    const int n = h->n;
    s.xy.resize((size_t)n * 2); s.oct.resize(n); s.desc.assign(h->desc, h->desc + (size_t)n * 32);
    for (int i = 0; i < n; i++) {
        memcpy(&s.xy[2 * i], (const char*)h->kps_xy + (size_t)i * h->kps_stride, 8);
        memcpy(&s.oct[i], (const char*)h->octaves + (size_t)i * h->oct_stride, 4);
    }
    if (h->u_right) s.ur.assign(h->u_right, h->u_right + n); else s.ur.clear();
    dev->kps_xy = s.xy.data(); dev->kps_stride = 8; dev->octaves = s.oct.data(); dev->oct_stride = 4;
    dev->desc = s.desc.data(); dev->u_right = h->u_right ? s.ur.data() : nullptr; dev->n = n; dev->on_device = 1;
    return ORBB_OK;
}

int orbb_search_area_topk(orbb_matcher* m, const orbb_frame_view* f, const float* grid4, const float* queries, const int32_t* qlev, const uint8_t* qdesc,
                          int nq, const uint8_t* skip, int init, int k, int32_t* out) {
    if (!m || !f || !grid4 || !out || (k != 1 && k != 2 && k != 4 && k != 8)) return ORBB_ERR_ARG;
    const int COLS = 64, ROWS = 48, n = f->n;
    const float minX = grid4[0], minY = grid4[1], invW = grid4[2], invH = grid4[3];
    auto X = [&](int i) { float v; memcpy(&v, (const char*)f->kps_xy + (size_t)i * f->kps_stride, 4); return v; };
    auto Y = [&](int i) { float v; memcpy(&v, (const char*)f->kps_xy + (size_t)i * f->kps_stride + 4, 4); return v; };
    auto O = [&](int i) { int32_t v; memcpy(&v, (const char*)f->octaves + (size_t)i * f->oct_stride, 4); return v; };
    std::vector<std::vector<int> > cell((size_t)COLS * ROWS);
    for (int i = 0; i < n; i++) {                                        // Frame::AssignFeaturesToGrid / PosInGrid (Frame.cc:385-416, :725-735)
        const int px = (int)std::round((X(i) - minX) * invW), py = (int)std::round((Y(i) - minY) * invH);
        if (px < 0 || px >= COLS || py < 0 || py >= ROWS) continue;
        cell[(size_t)px * ROWS + py].push_back(i);
    }
    for (int q = 0; q < nq; q++) {
        const float x = queries[4 * q], y = queries[4 * q + 1], r = queries[4 * q + 2], uR = queries[4 * q + 3];
        const int minLevel = qlev[2 * q], maxLevel = qlev[2 * q + 1];
        int32_t* o = out + (size_t)q * k * 2;
        for (int t = 0; t < k; t++) { o[2 * t] = init; o[2 * t + 1] = -1; }
        const int cx0 = std::max(0, (int)std::floor((x - minX - r) * invW)), cx1 = std::min(COLS - 1, (int)std::ceil((x - minX + r) * invW));
        const int cy0 = std::max(0, (int)std::floor((y - minY - r) * invH)), cy1 = std::min(ROWS - 1, (int)std::ceil((y - minY + r) * invH));
        if (cx0 >= COLS || cx1 < 0 || cy0 >= ROWS || cy1 < 0) continue;
        const bool levels = minLevel > 0 || maxLevel >= 0;
        std::vector<std::pair<int, int> > best;                          // (dist, idx), sorted by distance, ties in visit order
        for (int ix = cx0; ix <= cx1; ix++)
            for (int iy = cy0; iy <= cy1; iy++)
                for (int i : cell[(size_t)ix * ROWS + iy]) {
                    const int oc = O(i);
                    if (levels && (oc < minLevel || (maxLevel >= 0 && oc > maxLevel))) continue;
                    if (!(std::fabs(X(i) - x) < r && std::fabs(Y(i) - y) < r)) continue;
                    if (skip && skip[i]) continue;
                    if (f->u_right && f->u_right[i] > 0) {               // ORBmatcher.cc:94-100 / :1755-1761
                        const float er = std::fabs(uR - f->u_right[i]);
                        if (er > r) continue;
                    }
                    const int d = orbb_hamming_distance(qdesc + (size_t)32 * q, f->desc + (size_t)32 * i);
                    if (d >= init) continue;
                    size_t pos = best.size();
                    while (pos > 0 && best[pos - 1].first > d) pos--;    // behind every entry with distance <= d
                    best.insert(best.begin() + pos, std::make_pair(d, i));
                    if ((int)best.size() > k) best.pop_back();
                }
        for (size_t t = 0; t < best.size(); t++) { o[2 * t] = best[t].first; o[2 * t + 1] = best[t].second; }
    }
    return ORBB_OK;
}

int orbb_best2_csr(orbb_matcher* m, const uint8_t* q, int nq, const uint8_t* train, int ntrain, const int32_t* cand, const int32_t* rowptr, int init,
                   int32_t* out4) {
    if (!m || !out4) return ORBB_ERR_ARG;
    for (int j = 0; j < nq; j++) {
        int b1 = init, i1 = -1, b2 = init, i2 = -1;
        for (int c = rowptr[j]; c < rowptr[j + 1]; c++) {
            if (cand[c] < 0 || cand[c] >= ntrain) return ORBB_ERR_ARG;
            const int d = orbb_hamming_distance(q + (size_t)32 * j, train + (size_t)32 * cand[c]);
            if (d < b1) { b2 = b1; i2 = i1; b1 = d; i1 = cand[c]; }
            else if (d < b2) { b2 = d; i2 = cand[c]; }
        }
        out4[4 * j] = b1; out4[4 * j + 1] = i1; out4[4 * j + 2] = b2; out4[4 * j + 3] = i2;
    }
    return ORBB_OK;
}

int orbb_best2_csr_dev(orbb_matcher* m, const uint8_t* q, int nq, const uint8_t* train_dev, int ntrain, const int32_t* cand, const int32_t* rowptr,
                       int init, int32_t* out4) {
    return orbb_best2_csr(m, q, nq, train_dev, ntrain, cand, rowptr, init, out4);
}

}  // extern "C"
