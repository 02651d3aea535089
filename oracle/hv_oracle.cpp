// hv_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the HighVoronoi.jl raycast vertex search (the hot path named by
// BASELINE.json.north_star) for generators in general position.  It exists so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg have
// a checker / CPU baseline.  NOTHING in the product path (highvoronoi.jl_b200/) may
// include, link or call this file.
//
// PARITY STATUS: "parity unpinned" by the reference's own tests -- the reference ships no
// golden vectors for this path (all tests draw unseeded rand()) and
// Julia is not installed in the build container, so this restatement could not be run
// against the real package.  It is pinned instead against an independent Qhull oracle
// (oracle/qhull_oracle.py) and the reference's own validity predicates (verify_vertex,
// raycast.jl:477-502), see tests/test_oracle.py.  The one known-answer test the reference's
// tests do hold for this path -- the cell volumes add up to the volume of the domain
// (test/rcmethods.jl:8, test/multithread.jl:8, test/basics.jl:46) -- is run on this oracle's
// rows in tests/test_hostsim.py::test_cell_volume_formula_matches_qhull and on the GPU result
// in tests/test_gpu_volumes.py.  The reference also PUBLISHES outputs of this path: docs/src/index.md:93,96 print the matrices
// its own harness (statistics.jl:98-143) produced with the real package -- vertices, boundary vertices, walks and nn-searches per
// walk for N uniform points in the unit cube, d = 4 and 5, means over 4 unseeded clouds.  tests/golden/ref_published holds them and
// tests/test_oracle.py::test_published_statistics_of_the_reference / test_published_vertex_count_at_30000_nodes hold this
// restatement to them statistically (841 395.0 published vertices at d = 4, N = 30 000: +0.012 % here; 2 687 943.75 at d = 5,
// N = 20 000: -0.042 %; descents 36.0 against 39.5; nn-searches per walk of RCOriginal 2.60 against 2.61).  Bit-level values
// (signature by signature) stay unpinned until oracle/make_reference_fixtures.jl has been run by someone who has Julia.
//
// Methods: the default RCNonGeneralHP (raycast.jl:794-970) and, selected with hvo_set_method, RCOriginal (:972-1012),
// RCCombined (:504-528 with the nested KD traversal of extended.jl:155-176, kd_tree.jl:210-336, searchtrees.jl:50-194) and
// RCNonGeneralFast (:542-631) -- the four that test/rcmethods.jl:10-13 runs.  tests/test_oracle.py checks that they return the
// same mesh.
//
// Every function cites the reference file:line (relative to /root/reference/src) it follows.
// Deviations from the reference, all irrelevant for general-position input:
//  * Double64 "full_mode" re-orthogonalised correction (raycast.jl:633-707) is not restated;
//    it only polishes r at the 1e-13 level.
//  * IterativeSolvers.cg! inside _correct_vertex (raycast.jl:287-318) is replaced by a direct
//    solve of the same normal equations (the package is a third-party dependency that is not
//    vendored; compat floor IterativeSolvers >= 0.9.2, Project.toml).
//  * StaticArrays.qr (tools.jl:787) is restated as Householder QR.
//  * Degenerate vertices (> d+1 cospherical generators, FastEdgeIterator, edgeiterate.jl) are
//    detected, counted and reported, not enumerated.
//  * The KD-tree is a balanced median-split tree with leafsize 10 and per-leaf reordered
//    points like NearestNeighborModified/kd_tree.jl:27-158, but stored with explicit child
//    ranges instead of the implicit heap numbering of tree_ops.jl:11-81 (same query results).
//
// Build: see oracle/Makefile  (g++ -O2 -pthread -shared -fPIC).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <mutex>
#include <random>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <thread>
#include <chrono>

namespace {

typedef int64_t i64;
const int MAXD = 8;
const double INF = std::numeric_limits<double>::infinity();

inline double dot(const double* a, const double* b, int d) {
    double s = 0;
    for (int i = 0; i < d; ++i) s += a[i] * b[i];
    return s;
}
inline double dist2(const double* a, const double* b, int d) {
    double s = 0;
    for (int i = 0; i < d; ++i) { double t = a[i] - b[i]; s += t * t; }
    return s;
}

// ---------------------------------------------------------------------------------------------
// tools.jl:4-28  fnv1a_hash over Int64 ids (64-bit variant; the reference also derives a
// UInt128 variant, used here only as a hash for std containers).
// ---------------------------------------------------------------------------------------------
struct SigHash {
    size_t operator()(const std::vector<i64>& v) const {
        uint64_t h = 0xcbf29ce484222325ULL;
        for (i64 x : v) {
            uint64_t u = (uint64_t)x;
            for (int b = 0; b < 8; ++b) { h ^= (u >> (8 * b)) & 0xff; h *= 0x100000001b3ULL; }
        }
        return (size_t)h;
    }
};

// ---------------------------------------------------------------------------------------------
// NearestNeighborModified/kd_tree.jl:27-158 (build), :497-538 (knn kernel with skip predicate),
// :683-743 (inrange kernel), tree_ops.jl:134-154,183-195 (leaf scans).
// ---------------------------------------------------------------------------------------------
struct KDTree {
    static const int MAXD_ = 6;
    int d = 0;
    i64 n = 0;
    static const int LEAF = 10;                 // kd_tree.jl:29 leafsize = 10
    std::vector<double> pts;                    // reordered points (tree_ops.jl:66-81)
    std::vector<i64> idx;                       // reordered -> original
    struct Node { i64 lo_i, hi_i; int split_dim; double lo, hi, split_val; i64 left, right; };
    std::vector<Node> nodes;
    std::vector<double> bmin, bmax;

    void build(const double* x, i64 n_, int d_) {
        d = d_; n = n_;
        idx.resize(n);
        for (i64 i = 0; i < n; ++i) idx[i] = i;
        bmin.assign(d, INF); bmax.assign(d, -INF);
        for (i64 i = 0; i < n; ++i)
            for (int k = 0; k < d; ++k) {
                bmin[k] = std::min(bmin[k], x[i * d + k]);
                bmax[k] = std::max(bmax[k], x[i * d + k]);
            }
        nodes.clear();
        nodes.reserve(2 * (n / LEAF + 2));
        std::vector<double> mn = bmin, mx = bmax;
        rec(x, 0, n, mn, mx);
        pts.resize(n * d);
        for (i64 i = 0; i < n; ++i) std::memcpy(&pts[i * d], &x[idx[i] * d], sizeof(double) * d);
    }
    i64 rec(const double* x, i64 lo, i64 hi, std::vector<double>& mn, std::vector<double>& mx) {
        i64 me = (i64)nodes.size();
        nodes.push_back(Node());
        nodes[me].lo_i = lo; nodes[me].hi_i = hi; nodes[me].left = nodes[me].right = -1;
        if (hi - lo <= LEAF) return me;
        int sd = 0; double sp = 0;                                   // kd_tree.jl:117-126 max spread
        for (int k = 0; k < d; ++k) if (mx[k] - mn[k] > sp) { sp = mx[k] - mn[k]; sd = k; }
        i64 mid = lo + (hi - lo) / 2;                                // median (find_split analogue)
        std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                         [&](i64 a, i64 b) { return x[a * d + sd] < x[b * d + sd]; });
        double sv = x[idx[mid] * d + sd];
        nodes[me].split_dim = sd; nodes[me].split_val = sv; nodes[me].lo = mn[sd]; nodes[me].hi = mx[sd];
        double old = mx[sd]; mx[sd] = sv;
        i64 l = rec(x, lo, mid, mn, mx); mx[sd] = old;
        old = mn[sd]; mn[sd] = sv;
        i64 r = rec(x, mid, hi, mn, mx); mn[sd] = old;
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }
    double min_dist2_box(const double* p) const {                    // hyperrectangles.jl get_min_distance
        double s = 0;
        for (int k = 0; k < d; ++k) {
            double dd = 0;
            if (p[k] < bmin[k]) dd = bmin[k] - p[k]; else if (p[k] > bmax[k]) dd = p[k] - bmax[k];
            s += dd * dd;
        }
        return s;
    }
    template <class Skip>
    void nn_rec(i64 ni, const double* p, double min_d2, i64& best, double& best_d2, Skip& skip, i64& visits) const {
        const Node& nd = nodes[ni];
        if (nd.left < 0) {                                           // tree_ops.jl:134-154
            for (i64 i = nd.lo_i; i < nd.hi_i; ++i) {
                ++visits;
                double d2 = dist2(&pts[i * d], p, d);
                if (d2 < best_d2) { if (skip(idx[i])) continue; best_d2 = d2; best = idx[i]; }
            }
            return;
        }
        double pd = p[nd.split_dim], sdiff = pd - nd.split_val, ddiff;   // kd_tree.jl:510-537
        i64 close, far;
        if (sdiff > 0) { close = nd.right; far = nd.left; ddiff = std::max(0.0, pd - nd.hi); }
        else { close = nd.left; far = nd.right; ddiff = std::max(0.0, nd.lo - pd); }
        nn_rec(close, p, min_d2, best, best_d2, skip, visits);
        double new_min = min_d2 + sdiff * sdiff - ddiff * ddiff;
        if (new_min < best_d2) nn_rec(far, p, new_min, best, best_d2, skip, visits);
    }
    template <class Skip>
    std::pair<i64, double> nn(const double* p, Skip skip, i64& visits) const {   // searchtrees.jl:238-242
        i64 best = -1; double best_d2 = INF;
        if (n > 0) nn_rec(0, p, min_dist2_box(p), best, best_d2, skip, visits);
        return std::make_pair(best, best < 0 ? INF : std::sqrt(best_d2));
    }
    // kd_tree.jl:210-336 _knn_flex / knn_kernel_flex!: the nested traversal of RCCombined.  `data` is the NNSearchData of
    // searchtrees.jl:50-108 (see CombinedData below); a leaf that improves the candidate moves the query point and ends the pass
    // (returns false), the next pass skips the nodes already marked.
    template <class Data>
    bool flex_rec(i64 ni, double min_d2, Data& dt, i64& visits) const {
        const Node& nd = nodes[ni];
        if (nd.left < 0) {
            double old_r[MAXD_];
            std::memcpy(old_r, dt.new_r, sizeof(double) * d);
            dt.visited_nodes[ni] = 1;
            for (i64 i = nd.lo_i; i < nd.hi_i; ++i) {                 // tree_ops.jl:113-128 add_points_knn_flex!
                ++visits;
                const double dd = dist2(&pts[i * d], dt.new_r, d);
                const double correction = dt.dist_new_r_x0_2 * 1000 * dt.plane_tol;
                if (dd <= dt.dist_new_r_x0_2 + correction) dt.offer(&pts[i * d], idx[i], dd);
            }
            if (std::memcmp(old_r, dt.new_r, sizeof(double) * d) != 0) {
                std::memcpy(dt.r, dt.new_r, sizeof(double) * d);
                dt.bestdist = dist2(dt.x0, dt.new_r, d) * (1 + 1000 * dt.plane_tol);
                dt.dist_r_x0_2 = dt.bestdist;
                return false;
            }
            return true;
        }
        const int sd = nd.split_dim;
        const double pd = dt.r[sd], sdiff = pd - nd.split_val;
        i64 close, far; double ddiff; bool right;
        if (sdiff > 0) { close = nd.right; far = nd.left; ddiff = std::max(0.0, pd - nd.hi); right = true; }
        else { close = nd.left; far = nd.right; ddiff = std::max(0.0, nd.lo - pd); right = false; }
        auto valid_leaf = [&](bool rgt, double& old_val) {            // kd_tree.jl:241-267 switch_leaf / valid_leaf
            double& slot = rgt ? dt.mins[sd] : dt.maxs[sd];
            old_val = slot; slot = nd.split_val;
            double max_dot = 0;
            for (int k = 0; k < d; ++k) max_dot += (dt.u[k] > 0 ? dt.maxs[k] : dt.mins[k]) * dt.u[k];
            return max_dot > dt.c;
        };
        double old_val;
        bool valid = valid_leaf(right, old_val);
        if (!valid) dt.visited_nodes[close] = 1;
        if (!dt.visited_nodes[close]) {
            if (!flex_rec(close, min_d2, dt, visits)) return false;
        }
        (right ? dt.mins[sd] : dt.maxs[sd]) = old_val;
        const double new_min = min_d2 + sdiff * sdiff - ddiff * ddiff;
        valid = valid_leaf(!right, old_val);
        if (!valid) dt.visited_nodes[far] = 1;
        if (new_min < dt.bestdist && !dt.visited_nodes[far]) {
            if (!flex_rec(far, new_min, dt, visits)) return false;
        }
        (!right ? dt.mins[sd] : dt.maxs[sd]) = old_val;
        dt.visited_nodes[ni] = 1;
        return true;
    }
    template <class Data>
    void knn_flex(Data& dt, i64& visits) const {
        dt.visited_nodes.assign(nodes.size(), 0);
        const double init_min = min_dist2_box(dt.r);
        for (;;) {
            for (int k = 0; k < d; ++k) { dt.mins[k] = bmin[k]; dt.maxs[k] = bmax[k]; }
            if (n == 0 || flex_rec(0, init_min, dt, visits)) break;
        }
    }
    void inrange_rec(i64 ni, const double* p, double r2, double min_d2, std::vector<i64>& out, i64& visits) const {
        if (min_d2 > r2) return;                                     // kd_tree.jl:700-703
        const Node& nd = nodes[ni];
        if (nd.left < 0) {                                           // tree_ops.jl:183-195
            for (i64 i = nd.lo_i; i < nd.hi_i; ++i) {
                ++visits;
                if (dist2(&pts[i * d], p, d) <= r2) out.push_back(idx[i]);
            }
            return;
        }
        double pd = p[nd.split_dim], sdiff = pd - nd.split_val, ddiff;
        i64 close, far;
        if (sdiff > 0) { close = nd.right; far = nd.left; ddiff = std::max(0.0, pd - nd.hi); }
        else { close = nd.left; far = nd.right; ddiff = std::max(0.0, nd.lo - pd); }
        inrange_rec(close, p, r2, min_d2, out, visits);
        inrange_rec(far, p, r2, min_d2 + sdiff * sdiff - ddiff * ddiff, out, visits);
    }
    void inrange(const double* p, double r, std::vector<i64>& out, i64& visits) const {
        if (n > 0) inrange_rec(0, p, r * r, min_dist2_box(p), out, visits);
    }
};

struct Vertex { std::vector<i64> sig; double r[MAXD]; };
struct Ray { std::vector<i64> edge; double r[MAXD], u[MAXD]; i64 cell; };

struct Stats {
    i64 raycasts = 0, nn_calls = 0, inrange_calls = 0, points_visited = 0, descents = 0,
        corrections = 0, degenerate = 0, duplicates = 0, rejected = 0, search_us = 0, build_us = 0;
};

// Shared problem description (read-only during the search).
struct Problem {
    int d = 0; i64 N = 0; int P = 0;
    std::vector<double> xs;                     // N*d
    std::vector<double> pbase, pnormal;         // P*d
    KDTree tree;
    double variance_tol = 1e-15, break_tol = 1e-5, b_nodes_tol = 1e-7, plane_tol = 1e-12;   // raycast-types.jl:226-230
    // search_settings = (method = ...,): 0 RCNonGeneralHP (RCStandard), 1 RCOriginal, 2 RCCombined, 3 RCNonGeneralFast -- the numbering of
    // hvb_params.method (include/hvb200.h)
    int method = 0;
};

// Shared vertex store: hvdatabase.jl:94-116 (heap + key set), vdbdatabaseref.jl:118-143 (per-cell
// index lists), abstractmesh.jl:111-125 (push! + push_ref!).  One mutex stands in for the
// reference's read/write lock (hvdatabase.jl:42,97,112).
struct Store {
    // vertex records live in fixed blocks so that readers never see a reallocation (the reference's blocked heap,
    // hvdatabase.jl:94-110); the key set and the per-cell lists are sharded so that slab threads rarely collide
    // (the reference uses one read/write lock, hvdatabase.jl:42: this restatement is at least as scalable)
    static const int BLK = 4096, NSH = 64, NCL = 4096;
    std::vector<Vertex*> blocks;
    i64 nverts = 0;
    std::mutex vmtx;
    std::vector<std::vector<i64> > cell_lists;  // per real cell: indices of vertex records (owned + refs)
    std::unordered_set<std::vector<i64>, SigHash> keys[NSH];
    std::mutex kmtx[NSH], cmtx[NCL], rmtx;
    std::vector<Ray> rays;
    std::vector<char> dirty;                    // stands in for searcher.positions (sysvoronoi.jl:190-204)
    bool threaded = false;
    Store() : blocks(1 << 18, (Vertex*)0) {}
    ~Store() { for (Vertex* b : blocks) delete[] b; }
    Vertex& at(i64 i) { return blocks[i / BLK][i % BLK]; }
};

// Per-thread searcher state: raycast-types.jl:373-431 (RaycastIncircleSkip) + extended.jl:9-63,78-142.
struct Searcher {
    const Problem& pb;
    Store& st;
    int d; i64 N; int P;
    std::vector<double> mirror;                 // P*d  extended_xs[N+p] for the current cell
    std::vector<char> active;                   // extended.jl:81 active mirrors
    Stats stats;
    std::mt19937_64 rng;
    Searcher(const Problem& p, Store& s, uint64_t seed) : pb(p), st(s), d(p.d), N(p.N), P(p.P),
        mirror((size_t)p.P * p.d), active(p.P, 0), rng(seed) {}

    const double* X(i64 i) const { return i < N ? &pb.xs[i * d] : &mirror[(i - N) * d]; }

    // raycast.jl:354-375 activate_cell/activate_mirror, boundary.jl:206-211 reflect
    void activate_cell(i64 cell) {
        for (int p = 0; p < P; ++p) {
            const double* n = &pb.pnormal[p * d]; const double* b = &pb.pbase[p * d]; const double* x = &pb.xs[cell * d];
            double s = 0;
            for (int k = 0; k < d; ++k) s += n[k] * (b[k] - x[k]);
            for (int k = 0; k < d; ++k) mirror[p * d + k] = x[k] + n[k] * (2 * s);
            active[p] = 1;
        }
    }
    // extended.jl:108-142 nn(::ExtendedTree): tree query, then linear scan of active mirrors
    template <class Skip>
    std::pair<i64, double> nn(const double* p, Skip skip) {
        ++stats.nn_calls;
        std::pair<i64, double> res = pb.tree.nn(p, skip, stats.points_visited);
        for (int m = 0; m < P; ++m) {
            if (!active[m] || skip(N + m)) continue;
            double dd = std::sqrt(dist2(p, &mirror[m * d], d));
            if (dd < res.second) { res.first = N + m; res.second = dd; }
        }
        return res;
    }
    // extended.jl:265-277 _inrange(::ExtendedTree)
    void inrange(const double* p, double r, std::vector<i64>& out) {
        ++stats.inrange_calls;
        out.clear();
        pb.tree.inrange(p, r, out, stats.points_visited);
        for (int m = 0; m < P; ++m) {
            if (!active[m]) continue;
            if (std::sqrt(dist2(p, &mirror[m * d], d)) < r) out.push_back(N + m);
        }
    }

    // tools.jl:773-790 u_qr: last column of Q of the d x d matrix
    // [x_e1 - x_ed, ..., x_e(d-1) - x_ed, x_drop - x_ed], sign from R[d,d].
    void u_qr(const std::vector<i64>& sig, int drop_pos, double* u) {
        double A[MAXD][MAXD];                   // A[row][col]
        std::vector<i64> e;
        for (int j = 0; j < (int)sig.size(); ++j) if (j != drop_pos) e.push_back(sig[j]);
        const double* origin = X(e[d - 1]);
        for (int c = 0; c < d - 1; ++c) { const double* x = X(e[c]); for (int k = 0; k < d; ++k) A[k][c] = x[k] - origin[k]; }
        { const double* x = X(sig[drop_pos]); for (int k = 0; k < d; ++k) A[k][d - 1] = x[k] - origin[k]; }
        double V[MAXD][MAXD]; double beta[MAXD];
        for (int c = 0; c < d; ++c) {            // Householder
            double nrm = 0;
            for (int k = c; k < d; ++k) nrm += A[k][c] * A[k][c];
            nrm = std::sqrt(nrm);
            for (int k = 0; k < d; ++k) V[k][c] = 0;
            if (nrm == 0) { beta[c] = 0; continue; }
            double alpha = A[c][c] >= 0 ? -nrm : nrm;
            double vn = 0;
            for (int k = c; k < d; ++k) { V[k][c] = A[k][c]; if (k == c) V[k][c] -= alpha; vn += V[k][c] * V[k][c]; }
            beta[c] = vn > 0 ? 2.0 / vn : 0.0;
            for (int cc = c; cc < d; ++cc) {
                double s = 0;
                for (int k = c; k < d; ++k) s += V[k][c] * A[k][cc];
                s *= beta[c];
                for (int k = c; k < d; ++k) A[k][cc] -= s * V[k][c];
            }
        }
        double q[MAXD];
        for (int k = 0; k < d; ++k) q[k] = (k == d - 1) ? 1.0 : 0.0;     // Q e_d = H_1 ... H_d e_d
        for (int c = d - 1; c >= 0; --c) {
            double s = 0;
            for (int k = c; k < d; ++k) s += V[k][c] * q[k];
            s *= beta[c];
            for (int k = c; k < d; ++k) q[k] -= s * V[k][c];
        }
        double sgn = A[d - 1][d - 1] > 0 ? 1.0 : (A[d - 1][d - 1] < 0 ? -1.0 : 0.0);
        for (int k = 0; k < d; ++k) u[k] = -q[k] * sgn;
    }

    // raycast.jl:427-432 get_t_hp
    double get_t_hp(const double* r, const double* u, const double* x0, const double* xn) const {
        double num = 0, den = 0;
        for (int k = 0; k < d; ++k) { double Dx = xn[k] - x0[k]; num += Dx * (x0[k] + xn[k] - 2 * r[k]); den += u[k] * Dx; }
        return num / (2 * den);
    }
    // raycast.jl:399-409 get_t_hp_ (normalised Dx, returns value and error estimate)
    void get_t_hp_(const double* r, const double* u, const double* x0, const double* xn, double du, double& value, double& err) const {
        double Dx[MAXD], nrm = 0;
        for (int k = 0; k < d; ++k) { Dx[k] = xn[k] - x0[k]; nrm += Dx[k] * Dx[k]; }
        nrm = std::sqrt(nrm);
        double den = 0, num = 0, rn = 0;
        for (int k = 0; k < d; ++k) { Dx[k] /= nrm; den += u[k] * Dx[k]; num += Dx[k] * (x0[k] + xn[k] - 2 * r[k]); rn += r[k] * r[k]; }
        value = num / (2 * den);
        err = (value * du + std::sqrt(rn) * 1e-15) / den;
    }
    // raycast.jl:530-540 get_scale
    double get_scale(const double* u, const double* x0, const double* r_) const {
        double delta[MAXD], ref = 0, v = 0;
        for (int k = 0; k < d; ++k) { delta[k] = r_[k] - x0[k]; ref += delta[k] * delta[k]; v += u[k] * delta[k]; }
        v = v * v;
        double hori = v > ref ? 0.0 : ref - v;
        return std::sqrt(hori / ref);
    }

    // raycast.jl:320-329 vertex_variance
    double vertex_variance(const std::vector<i64>& sig, const double* r) const {
        int n = (int)sig.size();
        double dist[MAXD + 2], mean = 0;
        for (int k = 0; k < n; ++k) { dist[k] = dist2(X(sig[k]), r, d); mean += dist[k]; }
        mean /= n;
        double s = 0;
        for (int k = 0; k < n; ++k) s += (dist[k] - mean) * (dist[k] - mean);
        return s / (mean * mean);
    }
    // raycast.jl:287-318 _correct_vertex: normal equations of the circumcentre system,
    // restated with a direct solve instead of cg!.
    void correct_vertex(const std::vector<i64>& sig, double* r) const {
        double V[MAXD][MAXD], rhs[MAXD], S[MAXD][MAXD + 1];
        const double* xl = X(sig[d]);
        double diff = dot(xl, xl, d);
        for (int k = 0; k < d; ++k) rhs[k] = 0;
        for (int i = 0; i < d; ++i) {
            const double* xi = X(sig[i]);
            double h = 0.5 * (dot(xi, xi, d) - diff);
            for (int k = 0; k < d; ++k) { V[k][i] = xi[k] - xl[k]; rhs[k] += h * V[k][i]; }
        }
        for (int i = 0; i < d; ++i) {
            for (int j = 0; j < d; ++j) { double s = 0; for (int k = 0; k < d; ++k) s += V[i][k] * V[j][k]; S[i][j] = s; }
            S[i][d] = rhs[i];
        }
        for (int c = 0; c < d; ++c) {            // Gaussian elimination, partial pivoting
            int pv = c;
            for (int k = c + 1; k < d; ++k) if (std::fabs(S[k][c]) > std::fabs(S[pv][c])) pv = k;
            if (S[pv][c] == 0) return;
            if (pv != c) for (int j = 0; j <= d; ++j) std::swap(S[c][j], S[pv][j]);
            for (int k = c + 1; k < d; ++k) {
                double f = S[k][c] / S[c][c];
                for (int j = c; j <= d; ++j) S[k][j] -= f * S[c][j];
            }
        }
        for (int c = d - 1; c >= 0; --c) {
            double s = S[c][d];
            for (int j = c + 1; j < d; ++j) s -= S[c][j] * r[j];
            r[c] = s / S[c][c];
        }
    }
    // raycast.jl:242-279 walkray_correct_vertex
    bool walkray_correct_vertex(double* r, const std::vector<i64>& edge, i64 generator) {
        std::vector<i64> sig(edge.begin(), edge.begin() + d);
        sig.push_back(generator);
        double vv = vertex_variance(sig, r);
        int i = 0;
        while (i < 3 && vv > 0.0001 * pb.variance_tol) { ++i; ++stats.corrections; correct_vertex(sig, r); vv = vertex_variance(sig, r); }
        if (vv > pb.variance_tol && vv < pb.break_tol) { correct_vertex(sig, r); vv = vertex_variance(sig, r); }
        if (vv > pb.break_tol) { ++stats.rejected; return false; }
        return true;                            // adjust_boundary_vertex (boundary.jl:444) is the identity
    }

    // raycast.jl:719-773 get__r : at most two more predicate-nn refinements
    template <class Skip>
    void get__r(double* _r, const std::vector<i64>& edge, Skip& skip, const double* u, const double* r, const double* x0,
                double first_t, double& t_out, i64& ret_i) {
        i64 i0 = edge[0];
        double my_dist = std::sqrt(dist2(x0, _r, d));
        double t2buf = first_t;
        ret_i = -1;
        int ii = 1;
        while (true) {
            ++ii; if (ii == 4) break;
            std::pair<i64, double> res = nn(_r, skip);
            i64 i = res.first;
            if (i == i0) break;
            ret_i = i;
            if (i < 0) break;
            const double* x = X(i);
            if (std::sqrt(dist2(x, _r, d)) >= my_dist) break;
            double t2 = get_t_hp(r, u, x0, x);
            t2buf = t2;
            double rr[MAXD];
            for (int k = 0; k < d; ++k) rr[k] = r[k] + t2 * u[k];
            t2 += get_t_hp(rr, u, x0, x);
            for (int k = 0; k < d; ++k) _r[k] = r[k] + t2 * u[k];
            double nd = std::sqrt(dist2(x, _r, d));
            if (nd >= my_dist) break;
            my_dist = nd;
        }
        t_out = t2buf;
    }

    // raycast.jl:794-970 raycast_des2(::HPUnion) -- the default method RCNonGeneralHP.
    // sig: generators the ray is equidistant to (full_edge, or the partial simplex in descent);
    // origin: ids that may not be returned (the origin vertex's sig).  On success the new
    // generator(s) are appended to sig (sorted).  Returns generator or -1; t = INF if none.
    // raycast.jl:394 get_t: the plain (cancellation-prone) form the FP64-only methods use
    double get_t(const double* r, const double* u, const double* x0, const double* xn) const {
        double den = 0;
        for (int k = 0; k < d; ++k) den += u[k] * (xn[k] - x0[k]);
        return (dist2(r, xn, d) - dist2(r, x0, d)) / (2 * den);
    }

    // searchtrees.jl:50-108 NNSearchData / reset! and :125-194 skip_nodes_on_search: the state of RCCombined's nested search
    struct CombinedData {
        int d; double plane_tol;
        std::vector<i64> sigma, taboo;
        double bestdist; i64 bestnode;
        double c, main_c, current_c, dist_r_x0_2, dist_new_r_x0_2;
        double u[MAXD], r[MAXD], x0[MAXD], new_r[MAXD], mins[MAXD], maxs[MAXD];
        i64 lt, visited;
        std::vector<char> visited_nodes;
        void offer(const double* x_new, i64 i, double dist) {           // skip_nodes_on_search
            if (visited < lt && std::binary_search(taboo.begin(), taboo.end(), i)) { ++visited; return; }
            const double dn = dist_new_r_x0_2, correction = dn * 10 * plane_tol;
            if (dist > dn + correction) return;                           // by no means a better candidate
            double c_new = 0, abs_dx = 0;
            for (int k = 0; k < d; ++k) { c_new += x_new[k] * u[k]; abs_dx += (x_new[k] - x0[k]) * (x_new[k] - x0[k]); }
            if (c_new <= c) return;                                       // original raycast exclusion principle
            if (std::fabs(dist - dn) < correction) {                      // as good as the current candidate
                if (abs_dx / dist < 100 * correction) return;
                sigma.push_back(i);
                if (c_new > current_c) { bestnode = i; current_c = c_new; }
                return;
            }
            double rx2 = 0;
            for (int k = 0; k < d; ++k) rx2 += (r[k] - x_new[k]) * (r[k] - x_new[k]);
            const double new_t = (rx2 - dist_r_x0_2) / (2 * (c_new - main_c));
            if (abs_dx / (new_t * new_t) < plane_tol) return;
            double nr[MAXD], a = 0, b = 0, den = 0;
            for (int k = 0; k < d; ++k) nr[k] = r[k] + new_t * u[k];
            for (int k = 0; k < d; ++k) { a += (nr[k] - x_new[k]) * (nr[k] - x_new[k]); b += (nr[k] - x0[k]) * (nr[k] - x0[k]); den += u[k] * (x_new[k] - x0[k]); }
            const double t2 = (a - b) / (2 * den);                        // second order correction: get_t(new_r, u, x0, x_new)
            double dr = 0, dx0 = 0;
            for (int k = 0; k < d; ++k) { nr[k] += t2 * u[k]; new_r[k] = nr[k]; dr += (r[k] - nr[k]) * (r[k] - nr[k]); dx0 += (nr[k] - x0[k]) * (nr[k] - x0[k]); }
            dist_new_r_x0_2 = dx0;
            bestdist = (std::sqrt(dr) + std::sqrt(dx0)) * (std::sqrt(dr) + std::sqrt(dx0)) * (1 + 10000 * plane_tol);
            bestnode = i; current_c = c_new;
            sigma.assign(1, i);
        }
    };

    // raycast.jl:504-528 raycast_des2(::Raycast_Combined) with search_vertex2 (extended.jl:155-176): ONE nested traversal of the
    // tree in which every point inside the current candidate ball either replaces the candidate (the ball shrinks, the query
    // point moves to the new centre and the traversal restarts) or joins it (a tie).  The reference returns the dummy t = 1.0.
    i64 raycast_combined(std::vector<i64>& sig, const double* old_r, const double* u, const std::vector<i64>& edge,
                         const std::vector<i64>& origin, double& t, double* r2) {
        const double* x0 = X(edge[0]);
        CombinedData dt;
        dt.d = d; dt.plane_tol = pb.plane_tol;
        double a = 0, rr[MAXD];
        for (int k = 0; k < d; ++k) a += u[k] * (x0[k] - old_r[k]);
        for (int k = 0; k < d; ++k) rr[k] = old_r[k] + u[k] * a;                             // :508
        a = 0;
        for (int k = 0; k < d; ++k) a += u[k] * (x0[k] - rr[k]);
        for (int k = 0; k < d; ++k) rr[k] += u[k] * a;                                       // :509
        dt.taboo = origin; std::sort(dt.taboo.begin(), dt.taboo.end());                      // reset!(data, origin, r, x0, u, ...)
        dt.lt = (i64)dt.taboo.size(); dt.visited = 0;
        dt.bestdist = INF; dt.bestnode = -1;
        double c1 = -INF;
        for (i64 g : origin) c1 = std::max(c1, dot(X(g), u, d));
        dt.c = c1 + std::fabs(c1) * pb.plane_tol;
        dt.main_c = dot(x0, u, d); dt.current_c = dt.main_c;
        dt.dist_r_x0_2 = dist2(rr, x0, d); dt.dist_new_r_x0_2 = INF;
        for (int k = 0; k < d; ++k) { dt.u[k] = u[k]; dt.r[k] = rr[k]; dt.new_r[k] = rr[k]; dt.x0[k] = x0[k]; }
        for (int m = 0; m < P; ++m) {                                                        // search_vertex2: the mirrors first
            if (!active[m]) continue;
            dt.offer(&mirror[m * d], N + m, dist2(dt.new_r, &mirror[m * d], d));
        }
        ++stats.nn_calls;
        pb.tree.knn_flex(dt, stats.points_visited);
        t = INF;
        if (dt.bestnode < 0) { std::memcpy(r2, old_r, sizeof(double) * d); return -1; }
        const size_t before = sig.size();
        for (i64 g : dt.sigma) sig.push_back(g);
        std::sort(sig.begin(), sig.end());
        sig.erase(std::unique(sig.begin(), sig.end()), sig.end());
        if (sig.size() > before + 1) ++stats.degenerate;
        std::memcpy(r2, dt.new_r, sizeof(double) * d);
        t = 1.0;                                                                             // :517
        return dt.bestnode;
    }

    // raycast.jl:972-1012 raycast_des2(::Raycast_Original): the classic incircle iteration -- nearest neighbour of the foot
    // point under the half-space predicate, then nearest neighbours of r + t u WITHOUT predicate until the answer is a
    // generator of the edge or repeats.  General position only (one generator is appended).  `full_mode` (correct_cast on
    // ill-conditioned rays, :990,999) re-solves the candidate centre; it does not change which generator wins and is left out.
    i64 raycast_original(std::vector<i64>& sig, const double* r, const double* u, const std::vector<i64>& edge, double& t, double* r2) {
        const double* x0 = X(edge[0]);
        double c1 = -INF;
        for (i64 g : sig) c1 = std::max(c1, dot(X(g), u, d));
        const double c = c1 + std::fabs(c1) * pb.plane_tol;
        auto skip = [&](i64 i) { return dot(X(i), u, d) <= c; };
        auto noskip = [](i64) { return false; };
        double vvv[MAXD], a = 0;
        for (int k = 0; k < d; ++k) a += u[k] * (x0[k] - r[k]);
        for (int k = 0; k < d; ++k) vvv[k] = r[k] + u[k] * a;
        std::pair<i64, double> res = nn(vvv, skip);
        t = INF;
        if (res.first < 0) { std::memcpy(r2, r, sizeof(double) * d); return -1; }
        i64 i = res.first, i2 = res.first;
        double current_t = INF;
        while (i >= 0) {
            const double tt = get_t_hp(r, u, x0, X(i));
            if (tt >= current_t) break;
            current_t = tt;
            for (int k = 0; k < d; ++k) vvv[k] = r[k] + tt * u[k];
            i = nn(vvv, noskip).first;
            if (i == i2 || std::find(sig.begin(), sig.end(), i) != sig.end()) i = -1;
            if (i >= 0) i2 = i;
        }
        t = get_t_hp(r, u, x0, X(i2));
        for (int k = 0; k < d; ++k) r2[k] = r[k] + t * u[k];
        sig.push_back(i2);
        std::sort(sig.begin(), sig.end());
        return i2;
    }

    // raycast.jl:542-631 raycast_des2(::Raycast_Non_General): two predicate nearest-neighbour steps, one in-range ball, the
    // smallest t with the tie rule (largest u.(x - x0) within 1e-7), everything in plain FP64 (get_t)
    i64 raycast_nongeneral_fast(std::vector<i64>& sig, const double* r, const double* u, const std::vector<i64>& edge,
                                const std::vector<i64>& origin, double& t, double* r2) {
        const double* x0 = X(edge[0]);
        double c1 = -INF;
        for (i64 g : sig) c1 = std::max(c1, dot(X(g), u, d));
        const double c = c1 + std::fabs(c1) * pb.plane_tol;
        auto skip = [&](i64 i) { return dot(X(i), u, d) <= c; };
        double vvv[MAXD], _vvv[MAXD], _r[MAXD], a = 0;
        for (int k = 0; k < d; ++k) a += u[k] * (x0[k] - r[k]);
        for (int k = 0; k < d; ++k) vvv[k] = r[k] + u[k] * a;                               // :552
        std::pair<i64, double> res = nn(vvv, skip);
        t = INF;
        if (res.first < 0) { std::memcpy(r2, r, sizeof(double) * d); return -1; }
        double tt = get_t(r, u, x0, X(res.first));                                          // :557
        for (int k = 0; k < d; ++k) _vvv[k] = r[k] + tt * u[k];
        tt = get_t(_vvv, u, x0, X(res.first));
        for (int k = 0; k < d; ++k) vvv[k] = _vvv[k] + tt * u[k];
        res = nn(vvv, skip);                                                                 // :561
        if (res.first >= 0) tt = get_t(r, u, x0, X(res.first));
        for (int k = 0; k < d; ++k) _r[k] = r[k] + tt * u[k];
        double measure = 0;
        for (i64 g : sig) measure = std::max(measure, std::sqrt(dist2(X(g), _r, d)));
        double upper_t = tt + 2 * measure;
        const double scale = get_scale(u, x0, _r);
        std::vector<i64> idss;
        inrange(_r, (1 + std::max(1e-12, pb.b_nodes_tol * 100 * scale)) * measure, idss);   // :573
        if (idss.empty()) { std::memcpy(r2, r, sizeof(double) * d); return -1; }
        const i64 MAXI = std::numeric_limits<i64>::max();
        std::vector<double> ts(idss.size());
        for (size_t k = 0; k < idss.size(); ++k) {
            const bool in_origin = std::find(origin.begin(), origin.end(), idss[k]) != origin.end();
            ts[k] = in_origin ? 0.0 : get_t(r, u, x0, X(idss[k]));                          // :579
            if (ts[k] < pb.plane_tol) { idss[k] = MAXI; ts[k] = 0.0; }
            else if (ts[k] < upper_t) upper_t = ts[k];
        }
        upper_t += 10e-8;                                                                    // :594
        double max_dist = 0; i64 generator = -1;
        for (size_t k = 0; k < idss.size(); ++k) {
            if (ts[k] > upper_t) ts[k] = 0.0;
            else if (idss[k] < MAXI) {
                double v = 0; const double* xk = X(idss[k]);
                for (int q = 0; q < d; ++q) v += u[q] * (xk[q] - x0[q]);
                ts[k] = v;
                if (v > max_dist) { max_dist = v; generator = idss[k]; }
            }
        }
        if (generator < 0) { std::memcpy(r2, r, sizeof(double) * d); return -1; }
        t = get_t(r, u, x0, X(generator));                                                   // :611
        for (int k = 0; k < d; ++k) r2[k] = r[k] + t * u[k];
        // correct_cast(::Raycast_By_Walkray) (:443-447) = walkray_correct_vertex, which walkray applies again to the result
        double measure2 = std::sqrt(dist2(X(generator), r2, d));
        for (i64 g : sig) measure2 = std::max(measure2, std::sqrt(dist2(X(g), r2, d)));
        measure2 *= (1 + 0.1 * pb.b_nodes_tol);                                              // :616
        const size_t before = sig.size();
        for (size_t k = 0; k < idss.size(); ++k) {
            if (idss[k] == MAXI || std::sqrt(dist2(X(idss[k]), r2, d)) > measure2) continue;
            if (std::find(sig.begin(), sig.end(), idss[k]) == sig.end()) sig.push_back(idss[k]);
        }
        if (std::find(sig.begin(), sig.end(), generator) == sig.end()) sig.push_back(generator);
        if (sig.size() > before + 1) ++stats.degenerate;
        std::sort(sig.begin(), sig.end());
        return generator;
    }

    i64 raycast(std::vector<i64>& sig, const double* r, const double* u, const std::vector<i64>& edge,
                const std::vector<i64>& origin, double& t, double* r2) {
        ++stats.raycasts;
        if (pb.method == 1) return raycast_original(sig, r, u, edge, t, r2);
        if (pb.method == 2) return raycast_combined(sig, r, u, edge, origin, t, r2);
        if (pb.method == 3) return raycast_nongeneral_fast(sig, r, u, edge, origin, t, r2);
        const double* x0 = X(edge[0]);
        double c1 = -INF;
        for (i64 g : sig) c1 = std::max(c1, dot(X(g), u, d));             // :802-804
        double c = c1 + std::fabs(c1) * pb.plane_tol;
        auto skip = [&](i64 i) { return dot(X(i), u, d) <= c; };          // myskips :385
        double vvv[MAXD];
        double a = 0;
        for (int k = 0; k < d; ++k) a += u[k] * (x0[k] - r[k]);
        for (int k = 0; k < d; ++k) vvv[k] = r[k] + u[k] * a;             // :806
        std::pair<i64, double> res = nn(vvv, skip);
        t = INF;
        if (res.first < 0) { std::memcpy(r2, r, sizeof(double) * d); return -1; }
        const double* x = X(res.first);
        double tt, full_error;
        const double du = 1e-14;
        get_t_hp_(r, u, x0, x, du, tt, full_error);                       // :813
        double first_t = tt;
        double _vvv[MAXD];
        for (int k = 0; k < d; ++k) _vvv[k] = r[k] + tt * u[k];
        double scale = get_scale(u, x0, _vvv);
        double relative_error = full_error / std::sqrt(dist2(_vvv, r, d));
        double t2, e2;
        get_t_hp_(_vvv, u, x0, x, du, t2, e2);
        tt += t2;                                                          // :823
        for (int k = 0; k < d; ++k) vvv[k] = _vvv[k] + tt * u[k];         // :824 (sic: centre at ~2 t)
        double _r[MAXD];
        std::memcpy(_r, vvv, sizeof(double) * d);
        i64 new_i; double tcur;
        get__r(_r, edge, skip, u, r, x0, first_t, tcur, new_i);           // :827
        if (new_i < 0) { std::memcpy(r2, r, sizeof(double) * d); return -1; }
        double measure = 0;
        for (int k = 0; k < d && k < (int)edge.size(); ++k) measure = std::max(measure, std::sqrt(dist2(X(edge[k]), _r, d)));
        double upper_t = tcur * 1.0000000001;                              // :869
        std::vector<i64> idss;
        inrange(_r, (1 + std::max(relative_error, pb.b_nodes_tol * 10 * scale)) * measure, idss);   // :870
        if (idss.empty()) { std::memcpy(r2, r, sizeof(double) * d); return -1; }
        std::vector<double> ts(idss.size());
        const i64 MAXI = std::numeric_limits<i64>::max();
        for (size_t k = 0; k < idss.size(); ++k) {
            bool in_origin = std::find(origin.begin(), origin.end(), idss[k]) != origin.end();
            ts[k] = in_origin ? 0.0 : get_t_hp(r, u, x0, X(idss[k]));     // :877
            if (in_origin) { idss[k] = MAXI; continue; }
            if (ts[k] < pb.plane_tol || ts[k] > upper_t) ts[k] = 0.0;     // :887-892
            else if (ts[k] < upper_t) upper_t = ts[k];
        }
        upper_t += std::min(10e-8, (10 + d) * full_error);                // :902
        double max_dist = 0; i64 generator = -1;
        for (size_t k = 0; k < idss.size(); ++k) {
            if (idss[k] == MAXI) continue;
            if (ts[k] > upper_t) ts[k] = 0.0;
            else if (ts[k] > 0.0) {                                        // tie-break: max u.(x-x0), :907-913
                double v = 0; const double* xk = X(idss[k]);
                for (int q = 0; q < d; ++q) v += u[q] * (xk[q] - x0[q]);
                ts[k] = v;
                if (v > max_dist) { max_dist = v; generator = idss[k]; }
            }
        }
        if (generator < 0) { std::memcpy(r2, r, sizeof(double) * d); return -1; }
        t = get_t_hp(r, u, x0, X(generator));                              // :924
        for (int k = 0; k < d; ++k) r2[k] = r[k] + t * u[k];
        double minm, maxm; minm = maxm = std::sqrt(dist2(X(generator), r2, d));
        for (i64 s : sig) { double n2 = std::sqrt(dist2(X(s), r2, d)); minm = std::min(minm, n2); maxm = std::max(maxm, n2); }
        double measure2 = maxm + scale * (maxm - minm);                    // :934
        size_t before = sig.size();
        for (size_t k = 0; k < idss.size(); ++k) {                         // cospherical capture :936-949
            if (idss[k] == MAXI) continue;
            if (std::sqrt(dist2(X(idss[k]), r2, d)) > measure2) continue;
            if (std::find(sig.begin(), sig.end(), idss[k]) == sig.end()) sig.push_back(idss[k]);
        }
        if (std::find(sig.begin(), sig.end(), generator) == sig.end()) sig.push_back(generator);
        if (sig.size() > before + 1) ++stats.degenerate;
        std::sort(sig.begin(), sig.end());
        return generator;
    }

    // raycast.jl:211-239 randray: random unit vector orthogonal to span{x_g - x_last}
    void randray(const std::vector<i64>& gens, double* u) {
        int k = (int)gens.size();
        double v[MAXD][MAXD];
        const double* xl = X(gens[k - 1]);
        for (int i = 0; i < k - 1; ++i) {
            const double* xi = X(gens[i]);
            for (int q = 0; q < d; ++q) v[i][q] = xi[q] - xl[q];
            for (int rep = 0; rep < 2; ++rep) {
                for (int j = 0; j < i; ++j) { double s = dot(v[i], v[j], d); for (int q = 0; q < d; ++q) v[i][q] -= s * v[j][q]; }
                double nr = std::sqrt(dot(v[i], v[i], d));
                for (int q = 0; q < d; ++q) v[i][q] /= nr;
            }
        }
        std::normal_distribution<double> nd(0.0, 1.0);
        for (int q = 0; q < d; ++q) u[q] = nd(rng);
        for (int rep = 0; rep < 2; ++rep) {
            for (int i = 0; i < k - 1; ++i) { double s = dot(u, v[i], d); for (int q = 0; q < d; ++q) u[q] -= s * v[i][q]; }
            double nr = std::sqrt(dot(u, u, d));
            for (int q = 0; q < d; ++q) u[q] /= nr;
        }
    }

    // raycast.jl:45-109 descent
    bool descent(i64 start, std::vector<i64>& sig_out, double* r_out) {
        ++stats.descents;
        for (int attempt = 0; attempt < 10; ++attempt) {
            std::vector<i64> sig(1, start);
            std::vector<i64> minimal(1, start);
            double r[MAXD];
            std::memcpy(r, X(start), sizeof(double) * d);
            bool ok = true;
            int loop_counter = 0;
            for (int k = 1; k <= d && ok;) {
                ++loop_counter;
                double u[MAXD], t, r2[MAXD];
                randray(minimal, u);
                std::vector<i64> s2 = sig;
                i64 g = raycast(s2, r, u, sig, sig, t, r2);
                if (t == INF) {
                    for (int q = 0; q < d; ++q) u[q] = -u[q];
                    s2 = sig;
                    g = raycast(s2, r, u, sig, sig, t, r2);
                }
                if (t == INF) { if (loop_counter <= 100) continue; ok = false; break; }
                sig = s2;
                std::memcpy(r, r2, sizeof(double) * d);
                minimal.push_back(g);
                if (vertex_variance(minimal, r) > pb.variance_tol) { ok = false; break; }
                ++k; loop_counter = 0;
            }
            if (!ok || (int)sig.size() != d + 1) continue;
            std::sort(sig.begin(), sig.end());
            // project(r, domain) boundary.jl:461-473
            for (int p = 0; p < P; ++p) {
                double dd = 0;
                for (int q = 0; q < d; ++q) dd += (pb.pbase[p * d + q] - r[q]) * pb.pnormal[p * d + q];
                if (dd < 0) for (int q = 0; q < d; ++q) r[q] += dd * pb.pnormal[p * d + q];
            }
            std::vector<i64> e(minimal.begin(), minimal.begin() + d);
            walkray_correct_vertex(r, e, minimal[d]);
            sig_out = sig;
            std::memcpy(r_out, r, sizeof(double) * d);
            return true;
        }
        return false;
    }

    // ---- store access (abstractmesh.jl:111-153) -------------------------------------------
    // returns false if the vertex was already present (never in single-thread mode).
    bool push_vertex(const std::vector<i64>& sig, const double* r, i64 cell) {
        int sh = (int)(SigHash()(sig) % Store::NSH);
        {
            std::unique_lock<std::mutex> lk(st.kmtx[sh], std::defer_lock);
            if (st.threaded) lk.lock();
            if (!st.keys[sh].insert(sig).second) { ++stats.duplicates; return false; }
        }
        i64 id;
        {
            std::unique_lock<std::mutex> lk(st.vmtx, std::defer_lock);
            if (st.threaded) lk.lock();
            id = st.nverts++;
            if (!st.blocks[id / Store::BLK]) st.blocks[id / Store::BLK] = new Vertex[Store::BLK];
        }
        Vertex& v = st.at(id);
        v.sig = sig; std::memcpy(v.r, r, sizeof(double) * d);
        for (i64 g : sig) {
            if (g >= N) continue;                                        // planes hold no list (voronoi_mesh.jl:222)
            std::unique_lock<std::mutex> lk(st.cmtx[g % Store::NCL], std::defer_lock);
            if (st.threaded) lk.lock();
            st.cell_lists[g].push_back(id);
            if (g != cell) st.dirty[g] = 1;
        }
        return true;
    }
    void cell_vertices(i64 cell, std::vector<i64>& out) {
        std::unique_lock<std::mutex> lk(st.cmtx[cell % Store::NCL], std::defer_lock);
        if (st.threaded) lk.lock();
        st.dirty[cell] = 0;
        out = st.cell_lists[cell];
    }
    void get_vertex(i64 id, std::vector<i64>& sig, double* r) {
        // the record was completed before its index was published under the cell's lock
        const Vertex& v = st.at(id);
        sig = v.sig; std::memcpy(r, v.r, sizeof(double) * d);
    }
    void push_ray(const std::vector<i64>& edge, const double* r, const double* u, i64 cell) {
        std::unique_lock<std::mutex> lk(st.rmtx, std::defer_lock);
        if (st.threaded) lk.lock();
        Ray ry; ry.edge = edge; std::memcpy(ry.r, r, sizeof(double) * d); std::memcpy(ry.u, u, sizeof(double) * d); ry.cell = cell;
        st.rays.push_back(ry);
    }

    // ---- per-cell edge table: edgehashing.jl:66-111 pushedge! semantics ---------------------
    typedef std::unordered_map<std::vector<i64>, std::pair<i64, i64>, SigHash> EdgeTable;
    EdgeTable E;
    bool pushedge(const std::vector<i64>& key, i64 cell, bool mode) {
        EdgeTable::iterator it = E.find(key);
        if (it == E.end()) { E.emplace(key, std::make_pair(cell, (i64)-1)); return false; }
        if (it->second.second != -1) return true;
        if (mode || it->second.first != cell) it->second.second = cell;
        return false;
    }
    // edgeiteratebase.jl:37-53,76-82 General_EdgeIterator: which sub-facets are visited in cell i
    void edge_range(const std::vector<i64>& sig, i64 cell, int& a, int& b) const {
        if (sig[0] == cell) { a = 0; b = (int)sig.size() - 1; }
        else if (sig[1] == cell) { a = 0; b = 0; }
        else { a = 1; b = 0; }
    }
    // edgeiteratebase.jl:128-140 queue_edges_general_position
    bool register_edges(const std::vector<i64>& sig, i64 cell) {
        if ((int)sig.size() == d + 1 && sig[1] < cell) return true;
        bool all = true;
        int a, b; edge_range(sig, cell, a, b);
        for (int k = a; k <= b; ++k) {
            std::vector<i64> e; e.reserve(d);
            for (int j = 0; j < (int)sig.size(); ++j) if (j != k) e.push_back(sig[j]);
            all &= pushedge(e, sig[k], false);
        }
        return all;
    }

    // sysvoronoi.jl:490-525 systematic_explore_vertex (+ walkray raycast.jl:125-164)
    struct QItem { std::vector<i64> sig; double r[MAXD]; };
    i64 explore_vertex(const std::vector<i64>& sig, const double* r, i64 cell, std::vector<QItem>& queue) {
        i64 found = 0;
        int a, b; edge_range(sig, cell, a, b);
        for (int k = a; k <= b; ++k) {
            std::vector<i64> e; e.reserve(d + 1);
            for (int j = 0; j < (int)sig.size(); ++j) if (j != k) e.push_back(sig[j]);
            bool closed = pushedge(e, cell, true);
            if (e[0] != cell || closed) continue;
            double u[MAXD];
            u_qr(sig, k, u);                                                // get_full_edge sysvoronoi.jl:444-452
            std::vector<i64> sig2 = e;
            double t, r2[MAXD];
            i64 g = raycast(sig2, r, u, e, sig, t, r2);
            if (g < 0 || t == INF) { push_ray(e, r, u, cell); continue; }    // sysvoronoi.jl:504-511
            if ((int)sig2.size() > d + 1) continue;                        // degenerate: counted, not enumerated
            if (!walkray_correct_vertex(r2, e, g)) continue;
            bool isnew = push_vertex(sig2, r2, cell);
            if (isnew) ++found;
            else if (!st.threaded) continue;                               // MT: a peer found it first (parallelmesh.jl:202-235); keep walking it in this cell
            if (register_edges(sig2, cell)) continue;                      // :520
            QItem q; q.sig = sig2; std::memcpy(q.r, r2, sizeof(double) * d);
            queue.push_back(q);
        }
        return found;
    }

    // sysvoronoi.jl:384-442 systematic_explore_cell
    i64 explore_cell(i64 cell) {
        i64 found = 0;
        activate_cell(cell);
        E.clear();
        std::vector<QItem> queue;
        std::vector<i64> known;
        cell_vertices(cell, known);
        std::vector<QItem> seeds(known.size());
        for (size_t i = 0; i < known.size(); ++i) get_vertex(known[i], seeds[i].sig, seeds[i].r);
        for (size_t i = 0; i < seeds.size(); ++i) register_edges(seeds[i].sig, cell);          // :398-407
        for (size_t i = 0; i < seeds.size(); ++i) found += explore_vertex(seeds[i].sig, seeds[i].r, cell, queue);   // :409-414
        if (queue.empty() && seeds.empty()) {                                                   // :416-429
            QItem q;
            if (descent(cell, q.sig, q.r)) {
                if (push_vertex(q.sig, q.r, cell)) ++found;
                register_edges(q.sig, cell);
                queue.push_back(q);
            }
        }
        while (!queue.empty()) {                                                                // :430-434 (LIFO, queues.jl:41-60)
            QItem q = queue.back(); queue.pop_back();
            found += explore_vertex(q.sig, q.r, cell, queue);
        }
        return found;
    }
};

void add_stats(Stats& a, const Stats& b);

struct Result {
    int d; i64 N; int P;
    std::vector<Vertex> verts;      // sorted lexicographically by sig
    std::vector<Ray> rays;
    std::vector<i64> nb_off, nb_ids;
    Stats stats;
    std::string error;
};

void add_stats(Stats& a, const Stats& b) {
    a.raycasts += b.raycasts; a.nn_calls += b.nn_calls; a.inrange_calls += b.inrange_calls; a.points_visited += b.points_visited;
    a.descents += b.descents; a.corrections += b.corrections; a.degenerate += b.degenerate; a.duplicates += b.duplicates; a.rejected += b.rejected;
}

}  // namespace

extern "C" {

// Runs voronoi(mesh; searcher=Raycast(xs; domain)) (sysvoronoi.jl:21-39,152-215).
// nthreads == 1: SingleThread path (:41).  nthreads > 1: MultiThread(nthreads,1) path (:50-82):
// contiguous index slabs (parallelmesh.jl:52-87), one searcher per thread, one shared store.
static int g_method = 0;
// search_settings = (method = ...,) of the next hvo_run calls: 0 default (RCNonGeneralHP), 1 RCOriginal, 2 RCCombined, 3 RCNonGeneralFast
int hvo_set_method(int method) {
    if (method < 0 || method > 3) return -1;
    g_method = method;
    return 0;
}
void* hvo_run(int dim, int64_t n, const double* xs, int nplanes, const double* plane_base, const double* plane_normal,
              int nthreads, uint64_t seed) {
    Result* res = new Result();
    res->d = dim; res->N = n; res->P = nplanes;
    if (dim < 2 || dim > 6 || n <= dim) { res->error = "not enough points / bad dimension"; return res; }
    Problem pb;
    pb.d = dim; pb.N = n; pb.P = nplanes; pb.method = g_method;
    pb.xs.assign(xs, xs + (size_t)n * dim);
    pb.pbase.assign(plane_base, plane_base + (size_t)nplanes * dim);
    pb.pnormal.assign(plane_normal, plane_normal + (size_t)nplanes * dim);
    for (int p = 0; p < nplanes; ++p) {
        double nr = std::sqrt(dot(&pb.pnormal[p * dim], &pb.pnormal[p * dim], dim));
        for (int k = 0; k < dim; ++k) pb.pnormal[p * dim + k] /= nr;
    }
    auto t0 = std::chrono::steady_clock::now();
    pb.tree.build(pb.xs.data(), n, dim);
    auto t1 = std::chrono::steady_clock::now();
    Store st;
    st.cell_lists.resize(n);
    st.dirty.assign(n, 1);
    if (nthreads < 1) nthreads = 1;
    st.threaded = nthreads > 1;
    std::vector<Stats> tstats(nthreads);
    if (nthreads == 1) {
        Searcher s(pb, st, seed);
        for (i64 i = 0; i < n; ++i) s.explore_cell(i);
        tstats[0] = s.stats;
    } else {
        // repeat loop of __voronoi (sysvoronoi.jl:163-206): cells that received a vertex from a peer
        // after they were processed are visited again (iteration_count < 6).
        std::vector<i64> todo(n);
        for (i64 i = 0; i < n; ++i) todo[i] = i;
        for (int pass = 0; pass < 6 && !todo.empty(); ++pass) {
            i64 m = (i64)todo.size();
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; ++t)
                pool.emplace_back([&, t]() {
                    i64 lo = m * t / nthreads, hi = m * (t + 1) / nthreads;   // partition_indices parallelmesh.jl:52-87
                    Searcher s(pb, st, seed + 7919u * (uint64_t)t + 104729u * (uint64_t)pass);
                    for (i64 i = lo; i < hi; ++i) s.explore_cell(todo[i]);
                    add_stats(tstats[t], s.stats);
                });
            for (auto& th : pool) th.join();
            todo.clear();
            for (i64 i = 0; i < n; ++i) if (st.dirty[i]) todo.push_back(i);
        }
    }
    auto t2 = std::chrono::steady_clock::now();
    for (int t = 0; t < nthreads; ++t) add_stats(res->stats, tstats[t]);
    res->stats.build_us = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
    res->stats.search_us = std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
    res->verts.reserve(st.nverts);
    for (i64 i = 0; i < st.nverts; ++i) res->verts.push_back(st.at(i));
    std::sort(res->verts.begin(), res->verts.end(), [](const Vertex& a, const Vertex& b) { return a.sig < b.sig; });
    res->rays.swap(st.rays);
    std::sort(res->rays.begin(), res->rays.end(), [](const Ray& a, const Ray& b) { return a.edge < b.edge; });
    // neighbors.jl:219-262 neighbors_of_cell_new: sorted unique union of sig entries != i
    std::vector<std::vector<i64> > nb(n);
    for (const Vertex& v : res->verts)
        for (i64 g : v.sig) if (g < n) for (i64 h : v.sig) if (h != g) nb[g].push_back(h);
    res->nb_off.assign(n + 1, 0);
    for (i64 i = 0; i < n; ++i) {
        std::sort(nb[i].begin(), nb[i].end());
        nb[i].erase(std::unique(nb[i].begin(), nb[i].end()), nb[i].end());
        res->nb_off[i + 1] = res->nb_off[i] + (i64)nb[i].size();
    }
    res->nb_ids.reserve(res->nb_off[n]);
    for (i64 i = 0; i < n; ++i) res->nb_ids.insert(res->nb_ids.end(), nb[i].begin(), nb[i].end());
    return res;
}

const char* hvo_error(void* h) { return ((Result*)h)->error.c_str(); }

void hvo_counts(void* h, int64_t* nvert, int64_t* nrays, int64_t* nneigh) {
    Result* r = (Result*)h;
    *nvert = (int64_t)r->verts.size(); *nrays = (int64_t)r->rays.size(); *nneigh = r->nb_off.empty() ? 0 : r->nb_off.back();
}
// ids are written 1-based; boundary plane p (1-based) appears as n+p (docs/src/man/short.md:41-42)
void hvo_fetch_vertices(void* h, int64_t* sig, double* r) {
    Result* R = (Result*)h; int d = R->d;
    for (size_t i = 0; i < R->verts.size(); ++i) {
        for (int k = 0; k <= d; ++k) sig[i * (d + 1) + k] = R->verts[i].sig[k] + 1;
        for (int k = 0; k < d; ++k) r[i * d + k] = R->verts[i].r[k];
    }
}
void hvo_fetch_rays(void* h, int64_t* edge, double* base, double* dir, int64_t* node) {
    Result* R = (Result*)h; int d = R->d;
    for (size_t i = 0; i < R->rays.size(); ++i) {
        for (int k = 0; k < d; ++k) { edge[i * d + k] = R->rays[i].edge[k] + 1; base[i * d + k] = R->rays[i].r[k]; dir[i * d + k] = R->rays[i].u[k]; }
        node[i] = R->rays[i].cell + 1;
    }
}
void hvo_fetch_neighbors(void* h, int64_t* offsets, int64_t* ids) {
    Result* R = (Result*)h;
    for (size_t i = 0; i < R->nb_off.size(); ++i) offsets[i] = R->nb_off[i];
    for (size_t i = 0; i < R->nb_ids.size(); ++i) ids[i] = R->nb_ids[i] + 1;
}
// stats[0..10] = raycasts, nn_calls, inrange_calls, points_visited, descents, corrections, degenerate, duplicates, rejected,
//                 search_us (voronoi() proper: the cell loop), build_us (KD-tree build = Raycast(xs))
void hvo_stats(void* h, int64_t* s) {
    Result* R = (Result*)h;
    s[0] = R->stats.raycasts; s[1] = R->stats.nn_calls; s[2] = R->stats.inrange_calls; s[3] = R->stats.points_visited;
    s[4] = R->stats.descents; s[5] = R->stats.corrections; s[6] = R->stats.degenerate; s[7] = R->stats.duplicates; s[8] = R->stats.rejected;
    s[9] = R->stats.search_us; s[10] = R->stats.build_us;
}
void hvo_free(void* h) { delete (Result*)h; }

}  // extern "C"
