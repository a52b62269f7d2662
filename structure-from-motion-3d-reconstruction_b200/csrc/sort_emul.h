// sort_emul.h — the sequential building blocks of the exact libstdc++ std::sort emulation.
//
// The reference orders corner candidates with std::sort (cpp/src/templering_sfm.cpp:286), an UNSTABLE sort:
// with tied scores the resulting permutation is defined by libstdc++'s introsort (bits/stl_algo.h,
// __introsort_loop / __move_median_to_first / __unguarded_partition / __partial_sort / __final_insertion_sort).
// Corner order decides track ids and RANSAC samples, so the permutation must be reproduced bit for bit.
//
// Facts used (DESIGN.md §corner-select gives the derivations):
//  (1) Hoare partition == pair the k-th LEFT MISFIT (key <= pivot, ascending position) with the k-th RIGHT
//      MISFIT (key >= pivot, descending position) while Lpos[k] < Rpos[k]; m = number of such k; swap the
//      pairs; cut = min(Lpos[m], Rpos[m-1]) (whichever exist).  Two compactions + one gather: parallel.
//  (2) introsort recursion is a tree of disjoint segments: processing order is irrelevant, so segments can be
//      handled left-first, lazily, and in parallel.
//  (3) __final_insertion_sort is a stable sort; segment s precedes segment s+1 key-wise, so stably sorting
//      every leaf (<= 16 elements) on its own gives the same array.
//  (4) depth-limit exhaustion falls back to __partial_sort == make_heap + sort_heap (sequential; restated here).
//
// Keys are the bit patterns of non-negative doubles (bit order == value order); the order is DESCENDING:
// comp(a, b) := a > b.  Everything here compiles for host and device so the logic is unit-tested on the CPU
// (tests/test_sort_emul.py builds tests/cpu/sort_emul_host.cpp against this header).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SFM_HD __host__ __device__ __forceinline__
#else
#define SFM_HD inline
#endif

typedef unsigned long long sfm_key_t;

#define SFM_SORT_THRESHOLD 16

SFM_HD int sfm_lg2(unsigned n) {  // floor(log2 n), n > 0  (std::__lg)
  int k = 0;
  while (n >>= 1) k++;
  return k;
}

SFM_HD void sfm_swap_elem(sfm_key_t* key, uint32_t* idx, int a, int b) {
  const sfm_key_t k = key[a];
  key[a] = key[b];
  key[b] = k;
  const uint32_t i = idx[a];
  idx[a] = idx[b];
  idx[b] = i;
}

// __move_median_to_first(result=f, a=f+1, b=mid, c=l-1) with comp = '>'.
SFM_HD void sfm_median_to_first(sfm_key_t* key, uint32_t* idx, int f, int l) {
  const int a = f + 1, b = f + (l - f) / 2, c = l - 1;
  const sfm_key_t ka = key[a], kb = key[b], kc = key[c];
  int pick;
  if (ka > kb) {
    if (kb > kc)
      pick = b;
    else if (ka > kc)
      pick = c;
    else
      pick = a;
  } else if (ka > kc)
    pick = a;
  else if (kb > kc)
    pick = c;
  else
    pick = b;
  sfm_swap_elem(key, idx, f, pick);
}

// ---- heap fallback: std::__partial_sort(first, last, last) with comp = '>' --------------------------------
SFM_HD void sfm_push_heap(sfm_key_t* key, uint32_t* idx, int first, int hole, int top, sfm_key_t vk, uint32_t vi) {
  int parent = (hole - 1) / 2;
  while (hole > top && key[first + parent] > vk) {
    key[first + hole] = key[first + parent];
    idx[first + hole] = idx[first + parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  key[first + hole] = vk;
  idx[first + hole] = vi;
}

SFM_HD void sfm_adjust_heap(sfm_key_t* key, uint32_t* idx, int first, int hole, int len, sfm_key_t vk, uint32_t vi) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (key[first + child] > key[first + child - 1]) child--;
    key[first + hole] = key[first + child];
    idx[first + hole] = idx[first + child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    key[first + hole] = key[first + child - 1];
    idx[first + hole] = idx[first + child - 1];
    hole = child - 1;
  }
  sfm_push_heap(key, idx, first, hole, top, vk, vi);
}

SFM_HD void sfm_heap_sort(sfm_key_t* key, uint32_t* idx, int first, int last) {
  const int len = last - first;
  if (len >= 2) {  // __make_heap
    int parent = (len - 2) / 2;
    while (true) {
      sfm_adjust_heap(key, idx, first, parent, len, key[first + parent], idx[first + parent]);
      if (parent == 0) break;
      parent--;
    }
  }
  int end = last;  // __sort_heap
  while (end - first > 1) {
    --end;
    const sfm_key_t vk = key[end];
    const uint32_t vi = idx[end];
    key[end] = key[first];
    idx[end] = idx[first];
    sfm_adjust_heap(key, idx, first, 0, end - first, vk, vi);
  }
}

// ---- leaf: stable insertion sort of [f, l), l - f <= 16, descending ----------------------------------------
SFM_HD void sfm_leaf_sort(sfm_key_t* key, uint32_t* idx, int f, int l) {
  for (int i = f + 1; i < l; i++) {
    const sfm_key_t vk = key[i];
    const uint32_t vi = idx[i];
    int j = i;
    while (j > f && vk > key[j - 1]) {
      key[j] = key[j - 1];
      idx[j] = idx[j - 1];
      j--;
    }
    key[j] = vk;
    idx[j] = vi;
  }
}

// ---- sequential model of the parallel Hoare partition (fact 1).  lpos / rpos are scratch of >= l-f ints. ----
// Returns cut.  The CUDA kernels implement exactly these steps with group-wide scans.
SFM_HD int sfm_partition_model(sfm_key_t* key, uint32_t* idx, int f, int l, uint32_t* lpos, uint32_t* rpos) {
  sfm_median_to_first(key, idx, f, l);
  const sfm_key_t p = key[f];
  int nl = 0, nr = 0;
  for (int i = f + 1; i < l; i++)
    if (!(key[i] > p)) lpos[nl++] = (uint32_t)i;  // left misfit: would stop the left scan
  for (int j = l - 1; j > f; j--)
    if (!(p > key[j])) rpos[nr++] = (uint32_t)j;  // right misfit: would stop the right scan
  const int lim = nl < nr ? nl : nr;
  int m = 0;
  while (m < lim && lpos[m] < rpos[m]) m++;
  for (int k = 0; k < m; k++) sfm_swap_elem(key, idx, (int)lpos[k], (int)rpos[k]);
  uint32_t cut = 0xFFFFFFFFu;
  if (m < nl) cut = lpos[m];
  if (m > 0 && rpos[m - 1] < cut) cut = rpos[m - 1];
  return (int)cut;
}
