// Host helpers of the set-up steps: chunked thread teams and a multi-threaded std::sort that keeps std::sort's
// exact permutation (the BVH shape depends on how the library orders triangles whose keys tie, mesh_host.cpp).
#pragma once
#include <algorithm>
#include <cstdint>
#include <system_error>
#include <thread>
#include <utility>
#include <vector>

namespace sdfb200 {

// Starts `work` on a new thread appended to `team`; false (nothing started) when the system refuses another thread.
template <class W> bool tryFork(std::vector<std::thread>& team, W&& work) {
    try {
        team.emplace_back(std::forward<W>(work));
        return true;
    } catch (const std::system_error&) {
        return false;
    }
}

// Runs fn(chunkBegin, chunkEnd) over [0, n) on `threads` plain threads (the caller takes the last chunk).
template <class F> void forChunks(int32_t n, int threads, F&& fn) {
    threads = std::max(1, std::min(threads, n / 16384));
    if (threads == 1) { fn(0, n); return; }
    auto chunk = [&fn, n, threads](int c) { fn(int32_t(int64_t(n) * c / threads), int32_t(int64_t(n) * (c + 1) / threads)); };
    std::vector<std::thread> team;
    team.reserve(size_t(threads));
    for (int c = 0; c + 1 < threads; c++)
        if (!tryFork(team, [chunk, c] { chunk(c); })) chunk(c);   // no thread to be had: this one does the chunk
    chunk(threads - 1);
    for (std::thread& t : team) t.join();
}

#if defined(__GLIBCXX__)
// std::sort's permutation on several threads. libstdc++'s introsort partitions around a median-of-three pivot,
// recurses into the right part, loops on the left part, and finishes with ONE insertion pass over the whole array
// (bits/stl_algo.h, __introsort_loop / __final_insertion_sort). After a partition everything on the left is <= the
// pivot <= everything on the right, so (a) the right part is an independent range that another thread can take, and
// (b) the final pass — a stable insertion sort whose elements never cross a partition boundary — gives the same
// result when it is run range by range. The partition, heap-sort fallback and insertion routines called here are
// the library's own, with the library's depth limit, so every comparison that decides a tie is the one std::sort
// would make.
template <class It, class Cmp> void sortPiece(It first, It last, long depthLimit, Cmp comp, long grain) {
    std::vector<std::thread> forks;
    while (last - first > 16) {
        if (depthLimit == 0) { std::__partial_sort(first, last, last, comp); break; }
        --depthLimit;
        It cut = std::__unguarded_partition_pivot(first, last, comp);
        if (!(last - cut > grain && tryFork(forks, [=] { sortPiece(cut, last, depthLimit, comp, grain); }))) {
            std::__introsort_loop(cut, last, depthLimit, comp);
            std::__insertion_sort(cut, last, comp);
        }
        last = cut;
    }
    std::__insertion_sort(first, last, comp);
    for (std::thread& t : forks) t.join();
}
template <class It, class Less> void sortLikeStd(It first, It last, Less less, int threads) {
    const long n = long(last - first);
    if (threads <= 1 || n < 65536) { std::sort(first, last, less); return; }
    sortPiece(first, last, long(std::__lg(n)) * 2, __gnu_cxx::__ops::__iter_comp_iter(less), std::max<long>(8192, n / (4L * threads)));
}
#else
template <class It, class Less> void sortLikeStd(It first, It last, Less less, int) { std::sort(first, last, less); }
#endif

}  // namespace sdfb200
