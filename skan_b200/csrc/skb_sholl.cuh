// skan_b200 -- Sholl analysis (reference csr.py:1332-1390), the N- and E-proportional part: distance of every
// node from the centre, shell index of every node (np.digitize), and the histogram of edges whose endpoints
// lie in different shells.  Included by skb_items.cuh; the entry points are in skb_api.inl.
//
// Arithmetic follows the reference bit by bit so that the shell indices (integers) are identical:
//   scaled = coordinates * spacing            int64 * int64 -> int64, or int64 -> fp64 times fp64   (csr.py:1364)
//   d = sqrt(sum_k |c_k - scaled_k|^2)        scipy.spatial.distance_matrix, p = 2: fp64, summed left to right,
//                                             squares and the root correctly rounded                (csr.py:1376-1377)
//   bin = np.digitize(d, radii)               number of radii <= d for increasing radii; for decreasing radii
//                                             the number of radii > d                                (csr.py:1378-1379)
// (compiled without FMA contraction like the rest of the fp64 statistics).
#pragma once

struct SkbShollGeom {
    double center[3];
    double spacing_f[3];
    int64_t spacing_i[3];
    int32_t ndim;
    int32_t spacing_is_int;
};

SKB_HD double skb_sholl_distance(const SkbShollGeom& g, const int64_t* c) {
    double acc = 0.0;
    for (int k = 0; k < g.ndim; ++k) {
        double x = g.spacing_is_int ? (double)(c[k] * g.spacing_i[k]) : (double)c[k] * g.spacing_f[k];
        double diff = g.center[k] - x;
        if (diff < 0.0) diff = -diff;
        double sq = diff * diff;
        acc = k == 0 ? sq : acc + sq;
    }
    return skb_sqrt(acc);
}

// np.digitize(d, radii) for monotonic radii (n >= 1)
SKB_HD int32_t skb_digitize(const double* radii, int32_t n, int32_t decreasing, double d) {
    int32_t lo = 0, hi = n;  // first index whose radius is > d (increasing) / <= d (decreasing)
    while (lo < hi) {
        int32_t mid = (lo + hi) >> 1;
        bool right = decreasing ? radii[mid] > d : radii[mid] <= d;
        if (right) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

struct SkbShollBinItem {  // over nodes
    SkbShollGeom g;
    const int64_t* coords;  // (N, ndim)
    const double* radii;    // device, n_radii entries (may be null: distances only)
    int32_t n_radii;
    int32_t decreasing;
    int32_t* bins;          // may be null
    double* dist;           // may be null
    uint64_t* max_bits;     // may be null: maximum distance, as the bit pattern of a non-negative double
    SKB_HD void operator()(int64_t i) const {
        double d = skb_sholl_distance(g, coords + i * g.ndim);
        if (dist) dist[i] = d;
        if (max_bits) skb_atomic_max_u64_lazy(max_bits, skb_f64_bits(d));
        if (bins) bins[i] = skb_digitize(radii, n_radii, decreasing, d);
    }
};

struct SkbShollCountItem {  // over nodes: the directed edges of row i whose endpoints lie in different shells
    const int32_t* indptr;
    const int32_t* indices;
    const int32_t* bins;
    unsigned long long* hist;  // n_radii + 1 counters, zeroed; every crossing is counted from both ends
    SKB_HD void operator()(int64_t i) const {
        const int32_t b = bins[i];
        for (int32_t e = indptr[i]; e < indptr[i + 1]; ++e) {
            const int32_t c = bins[indices[e]];
            if (c != b) skb_atomic_add_u64(hist + (c < b ? c : b), 1ull);
        }
    }
};
