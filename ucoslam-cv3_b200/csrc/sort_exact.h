// sort_exact.h — exact emulation of libstdc++'s std::sort (GCC 13 bits/stl_algo.h: __sort -> __introsort_loop with
// __unguarded_partition_pivot / __move_median_to_first, depth limit 2*lg(n) falling back to heap sort, then
// __final_insertion_sort with the 16-element threshold).  std::sort is not stable, so WHICH of several equal keys ends up
// where is decided by this exact sequence of moves.  The reference's kd-tree build (src/basictypes/picoflann.h:310-318) sorts
// the point indices of a node by one coordinate when the mean cut is degenerate, and the order of equal coordinates then
// decides leaf contents and therefore the radius search's visit order, which the order-dependent best / second-best
// bookkeeping of the projection matchers observes (src/map.cpp:722-737).  Hence the operation-by-operation restatement.
//
// Usable from device code (one thread sorts one node's range) and from host code (CPU unit tests of the restatement
// against the real std::sort).  Elements are uint32_t indices; `Less` is a functor on two elements.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define UCO_SHD __host__ __device__ __forceinline__
#else
#define UCO_SHD inline
#endif

namespace uco_sort {
typedef uint32_t T;

template <class Less>
UCO_SHD void move_median_to_first(T* result, T* a, T* b, T* c, Less lt) {
    T t;
#define UCO_ISWAP(x, y) { t = *(x); *(x) = *(y); *(y) = t; }
    if (lt(*a, *b)) {
        if (lt(*b, *c)) UCO_ISWAP(result, b)
        else if (lt(*a, *c)) UCO_ISWAP(result, c)
        else UCO_ISWAP(result, a)
    } else if (lt(*a, *c)) UCO_ISWAP(result, a)
    else if (lt(*b, *c)) UCO_ISWAP(result, c)
    else UCO_ISWAP(result, b)
}
template <class Less>
UCO_SHD T* unguarded_partition(T* first, T* last, T* pivot, Less lt) {
    T t;
    for (;;) {
        while (lt(*first, *pivot)) ++first;
        --last;
        while (lt(*pivot, *last)) --last;
        if (!(first < last)) return first;
        UCO_ISWAP(first, last)
        ++first;
    }
}
#undef UCO_ISWAP
template <class Less>
UCO_SHD void unguarded_linear_insert(T* last, Less lt) {
    T val = *last;
    T* next = last - 1;
    while (lt(val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}
template <class Less>
UCO_SHD void insertion_sort(T* first, T* last, Less lt) {
    if (first == last) return;
    for (T* i = first + 1; i != last; ++i) {
        if (lt(*i, *first)) {
            T val = *i;
            for (T* p = i; p != first; --p) *p = *(p - 1);  // move_backward(first, i, i + 1)
            *first = val;
        } else unguarded_linear_insert(i, lt);
    }
}
template <class Less>
UCO_SHD void push_heap_(T* first, long hole, long top, T value, Less lt) {
    long parent = (hole - 1) / 2;
    while (hole > top && lt(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
template <class Less>
UCO_SHD void adjust_heap(T* first, long hole, long len, T value, Less lt) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (lt(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value, lt);
}
// __partial_sort(first, last, last): make_heap + sort_heap
template <class Less>
UCO_SHD void heap_sort(T* first, T* last, Less lt) {
    const long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            T v = first[parent];
            adjust_heap(first, parent, len, v, lt);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        T v = *last;
        *last = *first;
        adjust_heap(first, 0, last - first, v, lt);
    }
}
UCO_SHD int lg2(long n) {
    int k = 0;
    while (n > 1) { n >>= 1; k++; }
    return k;
}
// std::sort(first, last, lt).  __introsort_loop recurses on the right part and loops on the left one: the explicit stack holds
// the deferred (cut, last, depth) triples; at most 2*lg(n) + 1 are live.
template <class Less>
UCO_SHD void sort_(T* first, T* last, Less lt) {
    if (first == last) return;
    struct Frame { T* first; T* last; int depth; };
    Frame stack[72];
    int sp = 0;
    stack[sp++] = Frame{first, last, lg2(last - first) * 2};
    while (sp > 0) {
        Frame f = stack[--sp];
        // one activation of __introsort_loop(f.first, f.last, f.depth): the recursive calls it makes run BEFORE its own loop
        // continues, so the right parts are processed first (depth first); since the ranges are disjoint the final arrangement
        // does not depend on that interleaving, only on each range's own sequence of operations.
        while (f.last - f.first > 16) {
            if (f.depth == 0) {
                heap_sort(f.first, f.last, lt);
                break;
            }
            --f.depth;
            T* mid = f.first + (f.last - f.first) / 2;
            move_median_to_first(f.first, f.first + 1, mid, f.last - 1, lt);
            T* cut = unguarded_partition(f.first + 1, f.last, f.first, lt);
            if (sp < 72) stack[sp++] = Frame{cut, f.last, f.depth};
            f.last = cut;
        }
    }
    if (last - first > 16) {  // __final_insertion_sort
        insertion_sort(first, first + 16, lt);
        for (T* i = first + 16; i != last; ++i) unguarded_linear_insert(i, lt);
    } else insertion_sort(first, last, lt);
}
}  // namespace uco_sort
