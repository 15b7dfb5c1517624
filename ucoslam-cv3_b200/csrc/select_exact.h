// select_exact.h — exact emulation of the keypoint selection the reference performs with
//   cv::KeyPointsFilter::retainBest(kps, n) ; kps.resize(n)        (src/featureextractors/ORBextractor.cpp:1053-1055, 1069-1073)
// i.e. std::nth_element(begin, begin+n-1, end, response-greater) as implemented by libstdc++ (GCC 13, bits/stl_algo.h
// __introselect / __unguarded_partition_pivot / __insertion_sort / __heap_select), followed by truncation to n.
// The std::partition of retainBest only touches elements at positions >= n, which the reference then drops, so it
// is not needed.  FAST responses are small integers with massive ties, so WHICH tied keypoints survive and in what
// order is decided by this exact sequence of swaps; it is therefore restated operation by operation.
//
// Elements are packed keypoints  score << 24 | y << 12 | x ; the comparator looks at the score only.
// Usable from device code (one thread runs one selection) and from host code (CPU-side unit tests of the restatement).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define UCO_HD __host__ __device__ __forceinline__
#else
#define UCO_HD inline
#endif

namespace uco_sel {
typedef uint32_t T;
UCO_HD bool gt(T a, T b) { return (a >> 24) > (b >> 24); }  // KeypointResponseGreater
UCO_HD void iswap(T* a, T* b) { T t = *a; *a = *b; *b = t; }

UCO_HD void move_median_to_first(T* result, T* a, T* b, T* c) {
    if (gt(*a, *b)) {
        if (gt(*b, *c)) iswap(result, b);
        else if (gt(*a, *c)) iswap(result, c);
        else iswap(result, a);
    } else if (gt(*a, *c)) iswap(result, a);
    else if (gt(*b, *c)) iswap(result, c);
    else iswap(result, b);
}
UCO_HD T* unguarded_partition(T* first, T* last, T* pivot) {
    for (;;) {
        while (gt(*first, *pivot)) ++first;
        --last;
        while (gt(*pivot, *last)) --last;
        if (!(first < last)) return first;
        iswap(first, last);
        ++first;
    }
}
UCO_HD void insertion_sort(T* first, T* last) {
    if (first == last) return;
    for (T* i = first + 1; i != last; ++i) {
        T val = *i;
        if (gt(val, *first)) {
            for (T* p = i; p != first; --p) *p = *(p - 1);  // move_backward(first, i, i+1)
            *first = val;
        } else {
            T* hole = i;
            T* next = i - 1;
            while (gt(val, *next)) {
                *hole = *next;
                hole = next;
                --next;
            }
            *hole = val;
        }
    }
}
UCO_HD void push_heap_(T* first, long hole, long top, T value) {
    long parent = (hole - 1) / 2;
    while (hole > top && gt(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
UCO_HD void adjust_heap(T* first, long hole, long len, T value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (gt(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value);
}
UCO_HD void heap_select(T* first, T* middle, T* last) {
    long len = middle - first;
    if (len >= 2) {  // make_heap
        long parent = (len - 2) / 2;
        for (;;) {
            T v = first[parent];
            adjust_heap(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    for (T* i = middle; i < last; ++i)
        if (gt(*i, *first)) {  // pop_heap(first, middle, i)
            T v = *i;
            *i = *first;
            adjust_heap(first, 0, len, v);
        }
}
UCO_HD int lg2(long n) {  // std::__lg
    int k = 0;
    while (n > 1) { n >>= 1; k++; }
    return k;
}
UCO_HD void nth_element_(T* first, T* nth, T* last) {
    if (first == last || nth == last) return;
    int depth = lg2(last - first) * 2;
    while (last - first > 3) {
        if (depth == 0) {
            heap_select(first, nth + 1, last);
            iswap(first, nth);
            return;
        }
        --depth;
        T* mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        T* cut = unguarded_partition(first + 1, last, first);
        if (cut <= nth) first = cut;
        else last = cut;
    }
    insertion_sort(first, last);
}
// retainBest(v, n) followed by resize(n): permutes v in place, returns the new element count
UCO_HD int retain_best_truncate(T* v, int count, int n) {
    if (n >= 0 && count > n) {
        if (n == 0) return 0;
        nth_element_(v, v + n - 1, v + count);
        return n;
    }
    return count;
}
}  // namespace uco_sel
