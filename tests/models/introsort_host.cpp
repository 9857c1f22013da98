// Host build of orb_slam3_ros_b200/csrc/introsort.cuh for tests/test_introsort_model.py (CPU only).
#include <algorithm>
#include <cstdint>
#include <vector>
#define ORBB_SORT_STATS 1
#include "../../orb_slam3_ros_b200/csrc/introsort.cuh"

extern "C" {

int model_heap_fallbacks() { return orbb::g_heap_fallbacks; }

// sort (size, x0) pairs with the product's emulation; perm[i] = original index now at position i
void model_sort_nodes(const int* sizes, const int* x0s, int n, int* perm) {
    std::vector<orbb_rec_t> v(n);
    for (int i = 0; i < n; i++)
        v[i] = ((orbb_rec_t)(uint32_t)sizes[i] << 40) | ((orbb_rec_t)(uint16_t)x0s[i] << 24) | (orbb_rec_t)i;
    orbb::std_sort_emul(v.data(), n);
    for (int i = 0; i < n; i++) perm[i] = (int)(v[i] & 0xFFFFFF);
}

// the same through the parallel formulation (std_sort_emul_pf) that the CUDA kernel follows
void model_sort_nodes_pf(const int* sizes, const int* x0s, int n, int* perm) {
    std::vector<orbb_rec_t> v(n);
    for (int i = 0; i < n; i++)
        v[i] = ((orbb_rec_t)(uint32_t)sizes[i] << 40) | ((orbb_rec_t)(uint16_t)x0s[i] << 24) | (orbb_rec_t)i;
    orbb::std_sort_emul_pf(v.data(), n);
    for (int i = 0; i < n; i++) perm[i] = (int)(v[i] & 0xFFFFFF);
}

// McIlroy's "A Killer Adversary for Quicksort": builds the key sequence that drives THIS libstdc++ std::sort
// into its depth-limit (heapsort) fallback.  Output: keys[n] (distinct ints).
static int g_nsolid, g_candidate, g_gas;
static std::vector<int>* g_val;
static bool adv_less(int x, int y) {
    std::vector<int>& val = *g_val;
    if (val[x] == g_gas && val[y] == g_gas) {
        if (x == g_candidate) val[x] = g_nsolid++;
        else val[y] = g_nsolid++;
    }
    if (val[x] == g_gas) g_candidate = x;
    else if (val[y] == g_gas) g_candidate = y;
    return val[x] < val[y];
}
void model_killer_sequence(int n, int* keys) {
    std::vector<int> val(n), ptr(n);
    g_val = &val;
    g_gas = n - 1;
    g_nsolid = g_candidate = 0;
    for (int i = 0; i < n; i++) { ptr[i] = i; val[i] = g_gas; }
    std::sort(ptr.begin(), ptr.end(), adv_less);
    for (int i = 0; i < n; i++) keys[i] = val[i];
}
}
