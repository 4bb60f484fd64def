// Deterministic re-statement of libstdc++'s std::sort (introsort: median-of-3 quicksort with a
// 2*floor(log2 n) depth limit -> heap sort, threshold 16, final insertion sort), usable on host
// and device.  PCL's nearestKSearch calls std::sort on its candidate list with a comparator
// that only looks at the distance (octree_search.hpp, prioPointQueueEntry::operator<); for
// more than 16 candidates that sort is not stable, so reproducing the reference's neighbour
// order under exact distance ties needs the very same permutation.  tests/test_stdsort.py
// checks this file against the real std::sort of the toolchain on tie-heavy inputs.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define PGMO_HD __host__ __device__
#else
#define PGMO_HD
#endif

namespace pgmo {
namespace stdsort {

template <typename T, typename Less>
PGMO_HD inline void unguardedLinearInsert(T* last, Less less) {
  T val = *last;
  T* next = last - 1;
  while (less(val, *next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

template <typename T, typename Less>
PGMO_HD inline void insertionSort(T* first, T* last, Less less) {
  if (first == last) return;
  for (T* i = first + 1; i != last; ++i) {
    if (less(*i, *first)) {
      T val = *i;
      for (T* p = i; p != first; --p) *p = *(p - 1);
      *first = val;
    } else {
      unguardedLinearInsert(i, less);
    }
  }
}

template <typename T, typename Less>
PGMO_HD inline void unguardedInsertionSort(T* first, T* last, Less less) {
  for (T* i = first; i != last; ++i) unguardedLinearInsert(i, less);
}

template <typename T, typename Less>
PGMO_HD inline void pushHeap(T* first, int64_t hole, int64_t top, T value, Less less) {
  int64_t parent = (hole - 1) / 2;
  while (hole > top && less(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

template <typename T, typename Less>
PGMO_HD inline void adjustHeap(T* first, int64_t hole, int64_t len, T value, Less less) {
  const int64_t top = hole;
  int64_t second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (less(first[second], first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  pushHeap(first, hole, top, value, less);
}

template <typename T, typename Less>
PGMO_HD inline void heapSortAll(T* first, T* last, Less less) {
  // std::__partial_sort(first, last, last): __heap_select with middle == last is make_heap,
  // followed by sort_heap
  const int64_t len = last - first;
  if (len >= 2) {
    int64_t parent = (len - 2) / 2;
    while (true) {
      T value = first[parent];
      adjustHeap(first, parent, len, value, less);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    T value = *last;
    *last = *first;
    adjustHeap(first, int64_t(0), int64_t(last - first), value, less);
  }
}

template <typename T, typename Less>
PGMO_HD inline void moveMedianToFirst(T* result, T* a, T* b, T* c, Less less) {
  T* pick;
  if (less(*a, *b)) {
    if (less(*b, *c)) pick = b;
    else if (less(*a, *c)) pick = c;
    else pick = a;
  } else if (less(*a, *c)) pick = a;
  else if (less(*b, *c)) pick = c;
  else pick = b;
  T tmp = *result;
  *result = *pick;
  *pick = tmp;
}

template <typename T, typename Less>
PGMO_HD inline T* unguardedPartition(T* first, T* last, T* pivot, Less less) {
  while (true) {
    while (less(*first, *pivot)) ++first;
    --last;
    while (less(*pivot, *last)) --last;
    if (!(first < last)) return first;
    T tmp = *first;
    *first = *last;
    *last = tmp;
    ++first;
  }
}

// std::sort(first, last, less)
template <typename T, typename Less>
PGMO_HD inline void sort(T* first, T* last, Less less) {
  if (first == last) return;
  const int64_t n = last - first;
  // __lg(n) * 2
  int lg = 0;
  for (int64_t m = n; m > 1; m >>= 1) ++lg;
  // __introsort_loop recurses on the right part and iterates on the left part; an explicit
  // stack of (first, last, depth) replays that exact order.
  struct Frame {
    T* first;
    T* last;
    int depth;
  };
  Frame stack[64];
  int sp = 0;
  stack[sp++] = {first, last, lg * 2};
  while (sp > 0) {
    Frame f = stack[--sp];
    while (f.last - f.first > 16) {
      if (f.depth == 0) {
        heapSortAll(f.first, f.last, less);
        break;
      }
      --f.depth;
      T* mid = f.first + (f.last - f.first) / 2;
      moveMedianToFirst(f.first, f.first + 1, mid, f.last - 1, less);
      T* cut = unguardedPartition(f.first + 1, f.last, f.first, less);
      // the reference recurses into [cut, last) first, then continues with [first, cut):
      // both sub-ranges are independent, so deferring the right part keeps the result identical
      if (sp < 64) stack[sp++] = {cut, f.last, f.depth};
      f.last = cut;
    }
  }
  if (n > 16) {
    insertionSort(first, first + 16, less);
    unguardedInsertionSort(first + 16, last, less);
  } else {
    insertionSort(first, last, less);
  }
}

}  // namespace stdsort
}  // namespace pgmo
