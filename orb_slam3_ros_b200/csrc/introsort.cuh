// Exact re-statement of libstdc++'s std::sort (bits/stl_algo.h: __introsort_loop + __final_insertion_sort,
// bits/stl_heap.h for the depth-limit fallback) for arrays of 64-bit records.
//
// Why: DistributeOctTree (reference orb_slam3/src/ORBextractor.cc:700) sorts (nKeys, node) pairs with a
// comparator that only orders by (nKeys, UL.x) (:538-553).  std::sort is unstable, ties are frequent, and the
// loop that follows stops as soon as enough nodes exist (:746) -- so WHICH of several equivalent nodes is split
// depends on the exact element moves of libstdc++'s algorithm.  Reproducing the reference bit-for-bit therefore
// needs the same sequence of comparisons and moves, not just "a" sort.  tests/test_introsort_model.py checks this
// file (compiled for the host) against the real std::sort on tie-heavy and adversarial inputs.
//
// Record layout: key = rec >> ORBB_SORT_PAYLOAD_BITS (compared), low bits = payload (moved along, never compared).
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#include <algorithm>
#include <vector>
#endif
typedef unsigned long long orbb_rec_t;

#ifdef __CUDACC__
#define ORBB_HD __host__ __device__ __forceinline__
#else
#define ORBB_HD inline
#endif

#define ORBB_SORT_PAYLOAD_BITS 24

namespace orbb {

#ifdef ORBB_SORT_STATS
static int g_heap_fallbacks = 0;   // host-only test instrumentation
#endif

ORBB_HD bool rec_less(orbb_rec_t a, orbb_rec_t b) { return (a >> ORBB_SORT_PAYLOAD_BITS) < (b >> ORBB_SORT_PAYLOAD_BITS); }

ORBB_HD void rec_swap(orbb_rec_t* a, int i, int j) {
    orbb_rec_t t = a[i];
    a[i] = a[j];
    a[j] = t;
}

// std::__push_heap
ORBB_HD void heap_push(orbb_rec_t* a, int first, int hole, int top, orbb_rec_t value) {
    int parent = (hole - 1) / 2;
    while (hole > top && rec_less(a[first + parent], value)) {
        a[first + hole] = a[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[first + hole] = value;
}

// std::__adjust_heap
ORBB_HD void heap_adjust(orbb_rec_t* a, int first, int hole, int len, orbb_rec_t value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (rec_less(a[first + child], a[first + child - 1])) child--;
        a[first + hole] = a[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[first + hole] = a[first + child - 1];
        hole = child - 1;
    }
    heap_push(a, first, hole, top, value);
}

// std::__partial_sort(first, last, last) == __make_heap + __sort_heap (the __heap_select loop is empty)
ORBB_HD void heap_sort_range(orbb_rec_t* a, int first, int last) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            orbb_rec_t v = a[first + parent];
            heap_adjust(a, first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int end = last;
    while (end - first > 1) {
        --end;
        orbb_rec_t v = a[end];           // std::__pop_heap(first, end, end)
        a[end] = a[first];
        heap_adjust(a, first, 0, end - first, v);
    }
}

// std::__unguarded_linear_insert
ORBB_HD void unguarded_linear_insert(orbb_rec_t* a, int last) {
    orbb_rec_t val = a[last];
    int next = last - 1;
    while (rec_less(val, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}

// std::__insertion_sort
ORBB_HD void insertion_sort(orbb_rec_t* a, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (rec_less(a[i], a[first])) {
            orbb_rec_t val = a[i];
            for (int j = i; j > first; --j) a[j] = a[j - 1];   // std::move_backward(first, i, i + 1)
            a[first] = val;
        } else {
            unguarded_linear_insert(a, i);
        }
    }
}

// std::sort(a, a + n, rec_less)
ORBB_HD void std_sort_emul(orbb_rec_t* a, int n) {
    if (n <= 1) return;
    const int kThreshold = 16;
    // std::__lg(n) * 2
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    // explicit stack replaces the recursion on the right-hand part; ranges are disjoint so the order in which
    // they are finished does not change any element move inside them.
    int stFirst[64], stLast[64], stDepth[64];
    int sp = 0;
    int first = 0, last = n, depth = 2 * lg;
    while (true) {
        while (last - first > kThreshold) {
            if (depth == 0) {
#ifdef ORBB_SORT_STATS
                ++g_heap_fallbacks;
#endif
                heap_sort_range(a, first, last);
                break;
            }
            --depth;
            // std::__unguarded_partition_pivot
            const int mid = first + (last - first) / 2;
            {   // std::__move_median_to_first(first, first+1, mid, last-1)
                const int x = first + 1, y = mid, z = last - 1;
                if (rec_less(a[x], a[y])) {
                    if (rec_less(a[y], a[z])) rec_swap(a, first, y);
                    else if (rec_less(a[x], a[z])) rec_swap(a, first, z);
                    else rec_swap(a, first, x);
                } else if (rec_less(a[x], a[z])) rec_swap(a, first, x);
                else if (rec_less(a[y], a[z])) rec_swap(a, first, z);
                else rec_swap(a, first, y);
            }
            int lo = first + 1, hi = last;
            const orbb_rec_t pivot_pos = first;   // pivot stays at a[first] during __unguarded_partition
            while (true) {
                while (rec_less(a[lo], a[pivot_pos])) ++lo;
                --hi;
                while (rec_less(a[pivot_pos], a[hi])) --hi;
                if (!(lo < hi)) break;
                rec_swap(a, lo, hi);
                ++lo;
            }
            const int cut = lo;
            // recurse on [cut, last) later, continue with [first, cut)
            stFirst[sp] = cut; stLast[sp] = last; stDepth[sp] = depth; sp++;
            last = cut;
        }
        if (sp == 0) break;
        sp--;
        first = stFirst[sp]; last = stLast[sp]; depth = stDepth[sp];
    }
    // std::__final_insertion_sort
    if (n > kThreshold) {
        insertion_sort(a, 0, kThreshold);
        for (int i = kThreshold; i != n; ++i) unguarded_linear_insert(a, i);
    } else {
        insertion_sort(a, 0, n);
    }
}


// ---- parallel formulation ------------------------------------------------------------------------------------------
// The same arrangement as std_sort_emul, restated so that the work inside one step is order-free (the device version,
// sort_emul_cta in orbb_extract.cu, executes each step with a warp / the CTA; this host restatement is what
// tests/test_introsort_model.py checks against the real std::sort):
//  * __unguarded_partition(first+1, last, pivot at first): let l_1 < l_2 < ... be the positions in [first+1, last) whose
//    element is NOT less than the pivot and r_1 > r_2 > ... the positions in [first, last) whose element is NOT greater.
//    The sequential two-pointer loop swaps exactly the pairs (l_k, r_k) with l_k < r_k, k = 1..K -- positions between
//    l_k and r_k are untouched when the pointers pass them -- and returns cut = min(l_{K+1}, r_K): after K swaps the
//    upward scan stops at the next original l or at r_K, which now holds an element >= pivot.
//  * the ranges left by __introsort_loop are disjoint, so the order in which they are partitioned is irrelevant.
//  * __final_insertion_sort is a stable insertion sort of the whole array; since everything left of a cut is <= everything
//    right of it and an element only moves past STRICTLY greater ones, it permutes only inside the leaf ranges
//    (<= 16 elements) -- one independent insertion sort per leaf range.  Ranges finished by the heapsort fallback are
//    already sorted and stay as they are.
#ifndef __CUDACC__
inline void std_sort_emul_pf(orbb_rec_t* a, int n) {
    if (n <= 1) return;
    const int kThreshold = 16, INF = 0x7fffffff;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    std::vector<int> stF, stL, stD, leafF, leafL, L, R;
    stF.push_back(0); stL.push_back(n); stD.push_back(2 * lg);
    while (!stF.empty()) {
        int first = stF.back(), last = stL.back(), depth = stD.back();
        stF.pop_back(); stL.pop_back(); stD.pop_back();
        bool heap = false;
        while (last - first > kThreshold) {
            if (depth == 0) { heap_sort_range(a, first, last); heap = true; break; }
            --depth;
            const int mid = first + (last - first) / 2;
            {
                const int x = first + 1, y = mid, z = last - 1;
                if (rec_less(a[x], a[y])) {
                    if (rec_less(a[y], a[z])) rec_swap(a, first, y);
                    else if (rec_less(a[x], a[z])) rec_swap(a, first, z);
                    else rec_swap(a, first, x);
                } else if (rec_less(a[x], a[z])) rec_swap(a, first, x);
                else if (rec_less(a[y], a[z])) rec_swap(a, first, z);
                else rec_swap(a, first, y);
            }
            const orbb_rec_t pivot = a[first];
            L.clear(); R.clear();
            for (int p = first + 1; p < last; p++) if (!rec_less(a[p], pivot)) L.push_back(p);
            for (int p = last - 1; p >= first; p--) if (!rec_less(pivot, a[p])) R.push_back(p);
            int K = 0;
            for (size_t k = 0; k < L.size() && k < R.size(); k++) K += L[k] < R[k];
            for (int k = 0; k < K; k++) rec_swap(a, L[k], R[k]);            // simultaneous: all positions distinct
            const int cut = std::min(K < (int)L.size() ? L[K] : INF, K > 0 ? R[K - 1] : INF);
            stF.push_back(cut); stL.push_back(last); stD.push_back(depth);
            last = cut;
        }
        if (!heap) { leafF.push_back(first); leafL.push_back(last); }
    }
    for (size_t s = 0; s < leafF.size(); s++) insertion_sort(a, leafF[s], leafL[s]);      // independent of each other
}
#endif

}  // namespace orbb
