// Host-side helper of the stage-2 planners: cover an ascending lag grid by arithmetic progressions.
//
// kinisi's default grid, np.linspace(min_dt, max_dt, n_steps, dtype=int) (kinisi/parser.py:150-168), is an arithmetic
// progression up to integer truncation; np.unique(np.geomspace(..., dtype=int)) (:165) starts with a stretch of
// consecutive integers.  The moment kernels reuse samples along such progressions.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace kb2 {

struct LagSegment {
    int k0, d, len;               // lags k0 + i*d, i < len (d is irrelevant for len == 1)
    std::vector<int32_t> slots;   // position of lag i in the caller's grid
};

// Greedy cover: every still-uncovered lag, smallest first, starts the longest progression of uncovered lags among the
// differences to its next `candidates` distinct successors; progressions shorter than `min_len` are dropped and their
// first lag becomes a segment of one.  Segments come out in ascending order of their first lag.  Duplicate lags are
// legal (the reference's integer grids repeat when n_steps exceeds the trajectory): each copy is covered once.
inline std::vector<LagSegment> cover_by_progressions(const int32_t* lags, int K, int n_frames, int min_len,
                                                     int candidates) {
    std::vector<LagSegment> segs;
    std::vector<char> used(K, 0);
    auto find = [&](long long value, int from) -> int {
        const int32_t* lo = std::lower_bound(lags + from, lags + K, (int32_t)value);
        for (; lo < lags + K && *lo == value; ++lo)
            if (!used[lo - lags]) return (int)(lo - lags);
        return -1;
    };
    for (int a = 0; a < K; ++a) {
        if (used[a]) continue;
        int best_len = 1, best_d = 0, tried = 0;
        for (int b = a + 1; b < K && tried < candidates; ++b) {
            if (used[b] || lags[b] == lags[a]) continue;
            ++tried;
            const int d = lags[b] - lags[a];
            int len = 1, at = a;
            for (;;) {
                const long long nxt = (long long)lags[a] + (long long)len * d;
                if (nxt > n_frames) break;
                const int idx = find(nxt, at + 1);
                if (idx < 0) break;
                at = idx;
                ++len;
            }
            if (len > best_len) {
                best_len = len;
                best_d = d;
            }
        }
        LagSegment sg{lags[a], 1, 1, {a}};
        used[a] = 1;
        if (best_len >= min_len) {
            sg.d = best_d;
            sg.len = best_len;
            int at = a;
            for (int i = 1; i < best_len; ++i) {
                at = find((long long)lags[a] + (long long)i * best_d, at + 1);
                used[at] = 1;
                sg.slots.push_back(at);
            }
        }
        segs.push_back(std::move(sg));
    }
    return segs;
}

}  // namespace kb2
