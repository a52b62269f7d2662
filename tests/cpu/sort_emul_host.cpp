// Host build of csrc/sort_emul.h for CPU unit tests (tests/test_sort_emul.py): the emulation's sequential
// model against the real libstdc++ std::sort / std::__introsort_loop, including forced depth-limit exhaustion.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../structure-from-motion-3d-reconstruction_b200/csrc/sort_emul.h"

namespace {
struct KI {
  sfm_key_t k;
  uint32_t i;
};
struct Seg {
  int f, l, d;
};
}  // namespace

extern "C" {

// Emulated permutation: perm[pos] = original index.  depth < 0 -> libstdc++'s 2*lg(n).
int emul_sort_perm(const double* keys, int n, int* perm, int depth) {
  std::vector<sfm_key_t> key(n);
  std::vector<uint32_t> idx(n), lpos(n + 1), rpos(n + 1);
  for (int i = 0; i < n; i++) {
    std::memcpy(&key[i], &keys[i], 8);
    idx[i] = (uint32_t)i;
  }
  if (n > 0) {
    std::vector<Seg> st;
    st.push_back({0, n, depth < 0 ? 2 * sfm_lg2((unsigned)n) : depth});
    while (!st.empty()) {
      Seg s = st.back();
      st.pop_back();
      if (s.l - s.f <= SFM_SORT_THRESHOLD) {
        sfm_leaf_sort(key.data(), idx.data(), s.f, s.l);
        continue;
      }
      if (s.d == 0) {
        sfm_heap_sort(key.data(), idx.data(), s.f, s.l);
        continue;
      }
      const int cut = sfm_partition_model(key.data(), idx.data(), s.f, s.l, lpos.data(), rpos.data());
      st.push_back({cut, s.l, s.d - 1});  // right part later
      st.push_back({s.f, cut, s.d - 1});  // left part first (lazy, left-to-right)
    }
  }
  for (int i = 0; i < n; i++) perm[i] = (int)idx[i];
  return 0;
}

// The real thing.  depth < 0: std::sort; otherwise libstdc++'s own introsort loop with that depth limit
// followed by its final insertion sort (exactly what std::sort does after computing the limit).
int std_sort_perm(const double* keys, int n, int* perm, int depth) {
  std::vector<KI> v(n);
  for (int i = 0; i < n; i++) {
    std::memcpy(&v[i].k, &keys[i], 8);
    v[i].i = (uint32_t)i;
  }
  auto comp = [](const KI& a, const KI& b) { return a.k > b.k; };
  if (depth < 0) {
    std::sort(v.begin(), v.end(), comp);
  } else if (n > 0) {
    auto c = __gnu_cxx::__ops::__iter_comp_iter(comp);
    std::__introsort_loop(v.begin(), v.end(), (long)depth, c);
    std::__final_insertion_sort(v.begin(), v.end(), c);
  }
  for (int i = 0; i < n; i++) perm[i] = (int)v[i].i;
  return 0;
}
}
