// Sequential model of the ARRAY-REBUILD formulation of DistributeOctTree that the CUDA kernel
// (orb_slam3_ros_b200/csrc/orbb_extract.cu: k_octree) implements.  It replaces the reference's std::list with
// per-pass array rebuilds and per-node key segments in ping-pong buffers; tests/test_octree_model.py checks it
// against the std::list oracle (oracle/orb_port.cpp) so the reformulation is validated on the CPU before the
// GPU ever runs it.  Reference: orb_slam3/src/ORBextractor.cc:555-779.
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../orb_slam3_ros_b200/csrc/introsort.cuh"

namespace {
struct Node { int x0, y0, x1, y1, start, count, buf; };
struct Key { int x, y, resp, orig; };

struct Tree {
    std::vector<Key> keys[2];
    std::vector<Node> nodes;
    int cnt4[4];

    // stable 4-way partition of node p's segment into the other buffer; child segment sizes -> cnt4
    void split(const Node& p) {
        const int mx = p.x0 + ((p.x1 - p.x0 + 1) >> 1), my = p.y0 + ((p.y1 - p.y0 + 1) >> 1);
        const std::vector<Key>& src = keys[p.buf];
        std::vector<Key>& dst = keys[p.buf ^ 1];
        int c[4] = {0, 0, 0, 0};
        auto quad = [&](const Key& k) { return k.x < mx ? (k.y < my ? 0 : 2) : (k.y < my ? 1 : 3); };
        for (int i = 0; i < p.count; i++) c[quad(src[p.start + i])]++;
        int base[4] = {p.start, p.start + c[0], p.start + c[0] + c[1], p.start + c[0] + c[1] + c[2]};
        for (int i = 0; i < p.count; i++) { const Key& k = src[p.start + i]; dst[base[quad(k)]++] = k; }
        for (int q = 0; q < 4; q++) cnt4[q] = c[q];
    }
    // children of p in creation order n1..n4 (empty ones have count 0)
    void children(const Node& p, Node out[4]) const {
        const int mx = p.x0 + ((p.x1 - p.x0 + 1) >> 1), my = p.y0 + ((p.y1 - p.y0 + 1) >> 1);
        const int s1 = p.start + cnt4[0], s2 = s1 + cnt4[1], s3 = s2 + cnt4[2];
        out[0] = {p.x0, p.y0, mx, my, p.start, cnt4[0], p.buf ^ 1};
        out[1] = {mx, p.y0, p.x1, my, s1, cnt4[1], p.buf ^ 1};
        out[2] = {p.x0, my, mx, p.y1, s2, cnt4[2], p.buf ^ 1};
        out[3] = {mx, my, p.x1, p.y1, s3, cnt4[3], p.buf ^ 1};
    }
};
}  // namespace

extern "C" int model_distribute(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int* outIdx, int cap) {
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni <= 0) return -2;
    const float hX = (float)(maxX - minX) / nIni;
    Tree t;
    t.keys[0].resize(n); t.keys[1].resize(n);
    for (int i = 0; i < n; i++) t.keys[0][i] = {(int)xyr[3 * i], (int)xyr[3 * i + 1], (int)xyr[3 * i + 2], i};
    // roots: stable bucket by slot into buffer 1
    std::vector<int> slotCount(nIni, 0);
    for (int i = 0; i < n; i++) { int s = (int)((float)t.keys[0][i].x / hX); if (s >= nIni) return -2; slotCount[s]++; }
    std::vector<int> slotStart(nIni, 0);
    for (int s = 1; s < nIni; s++) slotStart[s] = slotStart[s - 1] + slotCount[s - 1];
    { std::vector<int> fill = slotStart;
      for (int i = 0; i < n; i++) { int s = (int)((float)t.keys[0][i].x / hX); t.keys[1][fill[s]++] = t.keys[0][i]; } }
    for (int s = 0; s < nIni; s++)
        if (slotCount[s] > 0)
            t.nodes.push_back({(int)(hX * (float)s), 0, (int)(hX * (float)(s + 1)), maxY - minY, slotStart[s], slotCount[s], 1});

    std::vector<int> pending;   // node indices (into t.nodes) of expandable children in creation order
    bool finish = false;
    while (!finish) {
        // ---- phase-1 pass: split every node with >1 keys, list order ----
        const int prevSize = (int)t.nodes.size();
        std::vector<Node> childBlocks;          // creation order: e_1:n1..n4, e_2:n1..n4 ...
        std::vector<int> blockLen;
        std::vector<Node> keep;
        for (const Node& p : t.nodes) {
            if (p.count == 1) { keep.push_back(p); continue; }
            t.split(p);
            Node c[4]; t.children(p, c);
            int len = 0;
            for (int q = 0; q < 4; q++) if (c[q].count > 0) { childBlocks.push_back(c[q]); len++; }
            blockLen.push_back(len);
        }
        // new list = children blocks in REVERSE creation order (push_front), then the untouched single-key nodes
        std::vector<Node> next;
        for (int i = (int)childBlocks.size() - 1; i >= 0; i--) next.push_back(childBlocks[i]);
        const int T = (int)childBlocks.size();
        for (const Node& k : keep) next.push_back(k);
        // pending = children with >1 keys in creation order; creation index j sits at list position T-1-j
        pending.clear();
        int nToExpand = 0;
        for (int j = 0; j < T; j++) if (childBlocks[j].count > 1) { pending.push_back(T - 1 - j); nToExpand++; }
        t.nodes.swap(next);
        const int size = (int)t.nodes.size();
        if (size >= N || size == prevSize) { finish = true; break; }
        if (size + nToExpand * 3 > N) {
            while (!finish) {
                const int prev2 = (int)t.nodes.size();
                // sort pending by (count, x0) with the libstdc++ emulation
                std::vector<orbb_rec_t> rec(pending.size());
                for (size_t i = 0; i < pending.size(); i++) {
                    const Node& p = t.nodes[pending[i]];
                    rec[i] = ((orbb_rec_t)(uint32_t)p.count << 40) | ((orbb_rec_t)(uint16_t)p.x0 << 24) | (orbb_rec_t)pending[i];
                }
                orbb::std_sort_emul(rec.data(), (int)rec.size());
                std::vector<char> erased(t.nodes.size(), 0);
                std::vector<Node> created;      // creation order
                int cur = prev2;
                for (int j = (int)rec.size() - 1; j >= 0; j--) {
                    const int idx = (int)(rec[j] & 0xFFFFFF);
                    const Node p = t.nodes[idx];
                    t.split(p);
                    Node c[4]; t.children(p, c);
                    for (int q = 0; q < 4; q++) if (c[q].count > 0) { created.push_back(c[q]); cur++; }
                    erased[idx] = 1; cur--;
                    if (cur >= N) break;
                }
                std::vector<Node> nx;
                const int T2 = (int)created.size();
                for (int i = T2 - 1; i >= 0; i--) nx.push_back(created[i]);
                for (size_t i = 0; i < t.nodes.size(); i++) if (!erased[i]) nx.push_back(t.nodes[i]);
                pending.clear();
                for (int j = 0; j < T2; j++) if (created[j].count > 1) pending.push_back(T2 - 1 - j);
                t.nodes.swap(nx);
                if ((int)t.nodes.size() >= N || (int)t.nodes.size() == prev2) finish = true;
            }
        }
    }
    if ((int)t.nodes.size() > cap) return -3;
    int m = 0;
    for (const Node& p : t.nodes) {
        const std::vector<Key>& k = t.keys[p.buf];
        int best = p.start;
        for (int i = 1; i < p.count; i++) if (k[p.start + i].resp > k[best].resp) best = p.start + i;
        outIdx[m++] = k[best].orig;
    }
    return m;
}
