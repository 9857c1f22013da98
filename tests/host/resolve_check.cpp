// CPU check of ORBmatcherGPU::ResolveInOrder (pure host code): random candidate lists with heavy competition for the same key
// points; the batched best-two + in-order resolution must equal the plain sequential loop of ORBmatcher.cc:77-141.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "ORBmatcherGPU.h"

using ORB_SLAM3::ORBmatcherGPU;
typedef ORBmatcherGPU::BestTwo BestTwo;

static BestTwo scan(const std::vector<int>& cand, const std::vector<int>& dist, const std::vector<unsigned char>& taken) {
    BestTwo b{256, -1, 256, -1};
    for (size_t c = 0; c < cand.size(); c++) {
        if (taken[cand[c]]) continue;
        if (dist[c] < b.bestDist) { b.secondDist = b.bestDist; b.secondIdx = b.bestIdx; b.bestDist = dist[c]; b.bestIdx = cand[c]; }
        else if (dist[c] < b.secondDist) { b.secondDist = dist[c]; b.secondIdx = cand[c]; }
    }
    return b;
}

int main() {
    std::mt19937 rng(7);
    int rescans = 0, total = 0;
    for (int trial = 0; trial < 200; trial++) {
        const int nkp = 50 + rng() % 200, nq = 20 + rng() % 400;
        std::vector<int> octave(nkp);
        for (int& o : octave) o = rng() % 4;
        std::vector<unsigned char> taken0(nkp);
        for (auto& t : taken0) t = rng() % 5 == 0;
        std::vector<std::vector<int>> cand(nq), dist(nq);
        for (int j = 0; j < nq; j++) {
            const int nc = rng() % 12;
            for (int c = 0; c < nc; c++) { cand[j].push_back(rng() % nkp); dist[j].push_back(rng() % 140); }
        }
        const float nnratio = 0.8f;
        // the reference's order: scan, decide, update -- one query after the other
        std::vector<unsigned char> takenSeq = taken0;
        std::vector<int> matchSeq(nkp, -1);
        int nSeq = 0;
        for (int j = 0; j < nq; j++) {
            const BestTwo b = scan(cand[j], dist[j], takenSeq);
            if (b.bestIdx < 0 || b.bestDist > ORBmatcherGPU::TH_HIGH) continue;
            const int l1 = octave[b.bestIdx], l2 = b.secondIdx >= 0 ? octave[b.secondIdx] : -1;
            if (l1 == l2 && b.bestDist > nnratio * b.secondDist) continue;
            matchSeq[b.bestIdx] = j; takenSeq[b.bestIdx] = 1; nSeq++;
        }
        // batched: every query scanned against the state before the call, then resolved in order
        std::vector<BestTwo> best(nq);
        for (int j = 0; j < nq; j++) best[j] = scan(cand[j], dist[j], taken0);
        std::vector<unsigned char> taken = taken0;
        std::vector<int> match;
        const int n = ORBmatcherGPU::ResolveInOrder(best, octave, taken, nnratio,
                                                    [&](int j, const std::vector<unsigned char>& t) { rescans++; return scan(cand[j], dist[j], t); }, match);
        total += nq;
        if (n != nSeq || match != matchSeq || taken != takenSeq) { printf("MISMATCH trial %d: %d vs %d\n", trial, n, nSeq); return 1; }
    }
    printf("resolve_check OK: %d queries, %d rescans\n", total, rescans);
    return rescans > 0 ? 0 : 2;
}
