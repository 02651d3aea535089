// hvb_coop.cuh -- warp-cooperative min-t query and the persistent walk kernel built on it (device only, sm_100a).
//
// Why: with one lane per ray (k_walk<D,1>) the query -- staged probe balls, rows of the cell box, FP32 scan, survivor
// bookkeeping -- executes with 5-11 of 32 threads active (profiles/r1_srcprofile_k_walk.md): rays differ in stages,
// rows per stage, points per row and survivors.  Here a warp still owns 32 rays (setup and commit stay lane-per-ray,
// they are uniform), but the rows of ALL its rays go into one pool in shared memory and the lanes pull row tasks from
// it, whoever the ray belongs to.  A lane that scans a row reads the ray's FP32 description from shared memory,
// tightens the ray's bound with an atomicMin and appends survivors to the ray's list; afterwards the owner lane
// verifies its survivors in FP64 (get_t_hp, raycast.jl:427).  Stages, rows and points are balanced across the warp,
// and lanes without a ray help.  The decision rules (FP32 filter with explicit error bounds, FP64 verification of
// everything that could be the winner or tie with it) are those of min_t_query; rays the FP32 geometry cannot
// describe (unbounded edges, huge balls) and rays whose survivor list overflows twice fall back to min_t_query itself.
#pragma once
#include "hvb_core.cuh"

namespace hvb {

template <int D>
struct CoopCfg {
    static const int TA = (D <= 3) ? D - 1 : D - 2;   // leading axes enumerated by the task index (d >= 4: a task loops over axis D-2)
    static const int CAP = 16;                        // survivor list entries per ray and stage
    static const int MAXT = 384;                      // chunk tasks (of <= 4 points) a round can hold; the rest is scanned by the emitting lane
};

template <int D>
struct __align__(16) CoopShared {      // one warp's 32 rays, structure of arrays: [field][lane]
    u64 surv[32][CoopCfg<D>::CAP];     // {FP32 lower bound of 2t : 32, bounded : 1, generator : 31}
    float uf[D][32], w2f[D][32], x0f[D][32], r32[D][32];
    float a32[32], perp2[32], ts0[32], m[32], en[32], ed[32];
    float re[D][32];                   // 1 / extent of the cell box along axis k (k < D-1)
    unsigned tb2[32];                  // FP32 upper bound of 2t of the winner (float bits, atomicMin)
    int clo[D][32], ext[D][32];        // cell box of the stage's ball: lower corner, extent (k < D-1)
    int excl[D + 1][32];
    int tighten[32];
    int ntask[32];                     // row tasks of the ray's current stage (0: none)
    int nexttask[32];                  // next task of the ray (atomicAdd; >= ntask: exhausted)
    int cnt[32];                       // survivors appended (may exceed CAP: overflow)
    // ---- the statically scheduled pool (pool_scan) ----
    int rbase[32], rcnt[32];           // this round's offer of the ray: first row task, number of row tasks
    unsigned chunk[CoopCfg<D>::MAXT];  // chunk tasks of the round: {owner : 5, points - 1 : 2, first point : 25}
};

// point range of the row (leading cell coordinates folded into base / d2 / umax / uabs) along the last axis:
// the cells that can hold a generator inside the ball and on the positive side of the edge's hyperplane
template <int D>
__device__ __forceinline__ bool coop_zrange(const Dev<D>& dv, const CoopShared<D>& sh, int o, float Ts, float rho2, float m,
                                            float d2, float umax, float uabs, int base, int& pa, int& pb) {
    if (!(d2 <= rho2)) return false;
    const int L = D - 1;
    const float s = sqrtf(rho2 - d2) * 1.000001f + m;
    const float ul = sh.uf[L][o], x0L = sh.x0f[L][o];
    const float cenL = fmaf(Ts, ul, sh.r32[L][o]);
    float zlo = cenL - s, zhi = cenL + s;
    const float slack = m * (uabs + fabsf(ul)) + 1e-5f * fabsf(umax) + 1e-6f * dv.ext32;
    if (ul > 1e-3f) zlo = fmaxf(zlo, x0L - (umax + slack) / ul * 1.00001f - m);
    else if (ul < -1e-3f) zhi = fminf(zhi, x0L - (umax + slack) / ul * 1.00001f + m);
    else if (umax + slack + fabsf(ul) * dv.ext32 * 2.f <= 0.f) return false;
    const float ihl = dv.inv_h32[L];
    const float gl = (float)dv.g[L];
    const float vlo = fminf(fmaxf(zlo * ihl - 2e-3f, 0.f), gl - 1.f);
    const float vhi = fminf(fmaxf(zhi * ihl + 2e-3f, -1.f), gl - 1.f);
    const int z0 = (int)floorf(vlo), z1 = (int)floorf(vhi);
    if (z1 < z0) return false;
    const int* cs = dv.cell_start + (size_t)base * dv.g[L];
    pa = __ldg(cs + z0); pb = __ldg(cs + z1 + 1);         // not waited for here: the caller scans another row first
    return true;
}

// contribution of leading axis k (cell coordinate c) to the row geometry
template <int D>
__device__ __forceinline__ void coop_axis(const Dev<D>& dv, const CoopShared<D>& sh, int o, int k, int c, float Ts, float m,
                                          float& d2, float& umax, float& uabs, int& base) {
    const float hk = dv.h32[k];
    const float blo = (float)c * hk - m;
    const float bhi = blo + hk + 2.f * m;
    const float ufk = sh.uf[k][o], x0k = sh.x0f[k][o];
    const float cenk = fmaf(Ts, ufk, sh.r32[k][o]);
    const float dd = fmaxf(0.f, fmaxf(blo - cenk, cenk - bhi));
    d2 = fmaf(dd, dd, d2);
    umax += fmaxf(ufk * (blo - x0k), ufk * (bhi - x0k));
    uabs += fabsf(ufk);
    base = base * dv.g[k] + c;
}

// The pooled scan of one stage.  Every lane takes the row tasks of its OWN ray first, one at a time (so the ray's
// bound tightens from row to row as in the one-lane query), then steals tasks of the other rays of the warp (lanes
// without a ray steal from the start).  The next row's range is requested before the current row is scanned.
template <int D>
__device__ __forceinline__ void coop_scan(const Dev<D>& dv, CoopShared<D>& sh, LocalStats& ls) {
    const int TA = CoopCfg<D>::TA, CAP = CoopCfg<D>::CAP;
    int o = threadIdx.x & 31, probes = 0;
    // d >= 4: rows c2 .. c2hi along axis D-2 of the current task
    int c2 = 1, c2hi = 0, base_p = 0;
    float d2_p = 0.f, umax_p = 0.f, uabs_p = 0.f, Ts_c = 0.f, rho2_c = 0.f, m_c = 0.f;

    // next row with work: (ray, point range); false when every ray of the warp is exhausted
    auto next_row = [&](int& ro, int& rpa, int& rpb) -> bool {
        for (;;) {
            if (D >= 4 && c2 <= c2hi) {
                float d2 = d2_p, umax = umax_p, uabs = uabs_p; int base = base_p;
                coop_axis<D>(dv, sh, o, D - 2, c2, Ts_c, m_c, d2, umax, uabs, base);
                ++c2;
                if (coop_zrange<D>(dv, sh, o, Ts_c, rho2_c, m_c, d2, umax, uabs, base, rpa, rpb)) { ro = o; return true; }
                continue;
            }
            if (probes >= 32) return false;
            int rem = atomicAdd(&sh.nexttask[o], 1);
            if (rem >= sh.ntask[o]) { o = (o + 1) & 31; ++probes; continue; }
            // the ray's current ball: the bound may have tightened since the stage was set up
            const float tb2 = __uint_as_float(*(volatile unsigned*)&sh.tb2[o]);
            const float Ts = fminf(sh.ts0[o], 0.5f * tb2 * 1.000001f);
            const float a = sh.a32[o];
            const float dT = fabsf(Ts - a) + 4e-7f * (fabsf(a) + Ts);
            const float rho2 = fmaf(dT, dT, sh.perp2[o]) * 1.00001f;
            const float m = sh.m[o];
            int cc[TA];
#pragma unroll
            for (int k = TA - 1; k >= 0; --k) {
                const int e = sh.ext[k][o];
                const int qd = (int)(((float)rem + 0.5f) * sh.re[k][o]);     // rem / e without an integer division (rem < 2^22)
                cc[k] = sh.clo[k][o] + (rem - qd * e);
                rem = qd;
            }
            float d2 = 0.f, umax = 0.f, uabs = 0.f; int base = 0;
#pragma unroll
            for (int k = 0; k < TA; ++k) coop_axis<D>(dv, sh, o, k, cc[k], Ts, m, d2, umax, uabs, base);
            if (D <= 3) {
                if (coop_zrange<D>(dv, sh, o, Ts, rho2, m, d2, umax, uabs, base, rpa, rpb)) { ro = o; return true; }
            } else if (d2 <= rho2) {
                // cells of axis D-2 the ball's slice reaches, inside the stage's box
                const int A = D - 2;
                const float s = sqrtf(rho2 - d2) * 1.000001f + m;
                const float cenA = fmaf(Ts, sh.uf[A][o], sh.r32[A][o]);
                const float iha = dv.inv_h32[A];
                const int blo = sh.clo[A][o], bhi = blo + sh.ext[A][o] - 1;
                const float vlo = fminf(fmaxf((cenA - s) * iha - 2e-3f, (float)blo), (float)bhi);
                const float vhi = fminf(fmaxf((cenA + s) * iha + 2e-3f, (float)blo - 1.f), (float)bhi);
                c2 = (int)floorf(vlo); c2hi = (int)floorf(vhi);
                d2_p = d2; umax_p = umax; uabs_p = uabs; base_p = base; Ts_c = Ts; rho2_c = rho2; m_c = m;
            }
        }
    };

    int co = 0, cpa = 0, cpb = 0, no = 0, npa = 0, npb = 0;
    bool have = next_row(co, cpa, cpb);
    int po = -1;
    float uf[D], w2f[D], x0f[D], en = 0.f, ed = 0.f;
    while (have) {
        const bool have_n = next_row(no, npa, npb);      // its cell_start loads fly while the current row is scanned
        if (cpa < cpb) { ls.rows++; ls.cand32 += (u32)(cpb - cpa); }
        if (co != po) {
#pragma unroll
            for (int k = 0; k < D; ++k) { uf[k] = sh.uf[k][co]; w2f[k] = sh.w2f[k][co]; x0f[k] = sh.x0f[k][co]; }
            en = sh.en[co]; ed = sh.ed[co];
            po = co;
        }
        for (int pa = cpa; pa < cpb; pa += 4) {
            // ---- FP32 pass over up to four points of the row: all loads before any arithmetic ----------------
            const int U = 4;
            float x[U][D];
#pragma unroll
            for (int i = 0; i < U; ++i) load_x32<D>(dv.x32, (pa + i < cpb) ? (pa + i) : (cpb - 1), x[i]);
            float tb2 = __uint_as_float(*(volatile unsigned*)&sh.tb2[co]);
            float nm[U], den[U];
            unsigned pmask = 0;
#pragma unroll
            for (int i = 0; i < U; ++i) {
                float d_ = 0.f, n_ = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float qk = x[i][k] - x0f[k];
                    d_ = fmaf(uf[k], qk, d_);
                    n_ = fmaf(qk, qk - w2f[k], n_);
                }
                nm[i] = n_; den[i] = d_;
                const bool pass = (pa + i < cpb) && (d_ + ed > 0.f) && (n_ - en <= tb2 * (d_ + ed));
                pmask |= pass ? (1u << i) : 0u;
            }
            while (pmask) {
                int i = 0;
#pragma unroll
                for (int b = U - 1; b >= 0; --b) i = ((pmask >> b) & 1u) ? b : i;       // lowest set bit
                pmask &= pmask - 1u;
                float nm_i = nm[0], den_i = den[0];
#pragma unroll
                for (int b = 1; b < U; ++b) { nm_i = (i == b) ? nm[b] : nm_i; den_i = (i == b) ? den[b] : den_i; }
                const int id = pa + i;
                bool excluded = false;
#pragma unroll
                for (int e = 0; e < D + 1; ++e) excluded |= (sh.excl[e][co] == id);
                if (excluded) continue;
                const float nlo = nm_i - en, nhi = nm_i + en, dh = den_i + ed, dl = den_i - ed;
                // an FP32 upper bound exists when the denominator is safely positive and t is safely > 0: such a
                // candidate is valid in FP64 as well (u.x > c, den > 0, t >= plane_tol)
                const bool bounded = sh.tighten[co] && dl > ed && nlo > 0.f;
                float lo = 0.f;
                if (bounded) {
                    lo = nlo / dh; lo -= fabsf(lo) * 4e-7f;
                    const float hi = nhi / dl * 1.000001f;
                    tb2 = __uint_as_float(*(volatile unsigned*)&sh.tb2[co]);
                    if (!(lo <= tb2)) continue;                                      // the bound tightened meanwhile
                    if (hi < tb2) atomicMin(&sh.tb2[co], __float_as_uint(hi));
                }
                const int pos = atomicAdd(&sh.cnt[co], 1);
                if (pos < CAP) sh.surv[co][pos] = ((u64)__float_as_uint(lo) << 32) | (bounded ? 0x80000000ULL : 0ULL) | (u64)(u32)id;
            }
        }
        co = no; cpa = npa; cpb = npb; have = have_n;
    }
}

// ------------------------------------------------------------------------------------------------------------
// pool_scan: the same stage scan with STATIC scheduling.  coop_scan hands out whole rows through shared-memory tickets:
// the rows of a warp's rays differ in length, so its lanes diverge again inside the rows, and the ticket traffic costs
// more than the row geometry it distributes.  Here nothing is fetched dynamically:
//   round:  every ray that still has row tasks offers its next Kr of them (Kr a power of two chosen so that the offers
//           of all active rays fill one or two warp steps);
//   rows:   slot s of the round belongs to (s / Kr)-th active ray, task s % Kr: the lanes run the row geometry for the
//           slots, whoever the ray belongs to, and turn every point range into chunk tasks of <= 4 points
//           (a warp prefix sum places them in the round's task array);
//   chunks: the lanes run the FP32 filter over the chunk tasks, 32 chunks per step, all lanes on the same instruction;
//           survivors tighten the ray's bound (atomicMin in shared memory) and are appended to the ray's list.
// The next round's rows see the tightened bounds, as the rows of the one-lane query do (in-flight shrink of the ball).
// ------------------------------------------------------------------------------------------------------------
#ifndef HVB_POOL_SLOTS
#define HVB_POOL_SLOTS 64          // row slots a round aims at (d <= 3); d >= 4 uses half: its tasks loop over axis D-2
#endif
template <int D>
__device__ __forceinline__ void pool_eval_chunk(const Dev<D>& dv, CoopShared<D>& sh, int o, int p0, int cnt) {
    const int CAP = CoopCfg<D>::CAP;
    const int U = 4;
    float uf[D], w2f[D], x0f[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { uf[k] = sh.uf[k][o]; w2f[k] = sh.w2f[k][o]; x0f[k] = sh.x0f[k][o]; }
    const float en = sh.en[o], ed = sh.ed[o];
    float x[U][D];
#pragma unroll
    for (int i = 0; i < U; ++i) load_x32<D>(dv.x32, p0 + (i < cnt ? i : cnt - 1), x[i]);
    float tb2 = __uint_as_float(*(volatile unsigned*)&sh.tb2[o]);
    float nm[U], den[U];
    unsigned pmask = 0;
#pragma unroll
    for (int i = 0; i < U; ++i) {
        float d_ = 0.f, n_ = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float qk = x[i][k] - x0f[k];
            d_ = fmaf(uf[k], qk, d_);
            n_ = fmaf(qk, qk - w2f[k], n_);
        }
        nm[i] = n_; den[i] = d_;
        const bool pass = (i < cnt) && (d_ + ed > 0.f) && (n_ - en <= tb2 * (d_ + ed));
        pmask |= pass ? (1u << i) : 0u;
    }
    while (pmask) {
        const int i = __ffs((int)pmask) - 1;
        pmask &= pmask - 1u;
        float nm_i = nm[0], den_i = den[0];
#pragma unroll
        for (int b = 1; b < U; ++b) { nm_i = (i == b) ? nm[b] : nm_i; den_i = (i == b) ? den[b] : den_i; }
        const int id = p0 + i;
        bool excluded = false;
#pragma unroll
        for (int e = 0; e < D + 1; ++e) excluded |= (sh.excl[e][o] == id);
        if (excluded) continue;
        const float nlo = nm_i - en, nhi = nm_i + en, dh = den_i + ed, dl = den_i - ed;
        // an FP32 upper bound exists when the denominator is safely positive and t is safely > 0: such a
        // candidate is valid in FP64 as well (u.x > c, den > 0, t >= plane_tol)
        const bool bounded = sh.tighten[o] && dl > ed && nlo > 0.f;
        float lo = 0.f;
        if (bounded) {
            lo = nlo / dh; lo -= fabsf(lo) * 4e-7f;
            const float hi = nhi / dl * 1.000001f;
            tb2 = __uint_as_float(*(volatile unsigned*)&sh.tb2[o]);
            if (!(lo <= tb2)) continue;                                      // the bound tightened meanwhile
            if (hi < tb2) atomicMin(&sh.tb2[o], __float_as_uint(hi));
        }
        const int pos = atomicAdd(&sh.cnt[o], 1);
        if (pos < CAP) sh.surv[o][pos] = ((u64)__float_as_uint(lo) << 32) | (bounded ? 0x80000000ULL : 0ULL) | (u64)(u32)id;
    }
}

template <int D>
__device__ __forceinline__ void pool_scan(const Dev<D>& dv, CoopShared<D>& sh, LocalStats& ls) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int TA = CoopCfg<D>::TA, MAXT = CoopCfg<D>::MAXT;
    const int mytasks = sh.ntask[lane];
    int done = 0;                                   // row tasks of MY ray handed out so far
    for (;;) {
        const unsigned amask = __ballot_sync(FULL, done < mytasks);
        if (!amask) break;
        const int nact = __popc(amask);
        // row tasks per ray and round: a power of two, so that the slots decode with shifts
        const int want = ((D <= 3) ? HVB_POOL_SLOTS : HVB_POOL_SLOTS / 2) / nact;
        int ke = 0;
        while ((2 << ke) <= want && ke < 4) ++ke;
        const int Kr = 1 << ke;
        const int left = mytasks - done;
        const int k = left > Kr ? Kr : (left > 0 ? left : 0);
        sh.rbase[lane] = done; sh.rcnt[lane] = k;
        done += k;
        __syncwarp();
        const int nslots = nact << ke;
        int total = 0;                              // chunk tasks emitted in this round (warp-uniform)
        for (int s0 = 0; s0 < nslots; s0 += 32) {
            const int s = s0 + lane;
            bool valid = s < nslots;
            int o = 0, idx = 0;
            if (valid) {
                o = __fns(amask, 0, (s >> ke) + 1);           // lane of the (s >> ke)-th active ray
                idx = s & (Kr - 1);
                valid = idx < sh.rcnt[o];
            }
            // ---- row geometry of task (o, rbase + idx) ------------------------------------------------------
            int c2 = 1, c2hi = 0, base_p = 0, pa = 0, pb = 0;
            float d2_p = 0.f, umax_p = 0.f, uabs_p = 0.f, Ts = 0.f, rho2 = 0.f, m = 0.f;
            bool have = false;
            if (valid) {
                int rem = sh.rbase[o] + idx;
                // the ray's current ball: the bound may have tightened since the stage was set up
                const float tb2 = __uint_as_float(*(volatile unsigned*)&sh.tb2[o]);
                Ts = fminf(sh.ts0[o], 0.5f * tb2 * 1.000001f);
                const float a = sh.a32[o];
                const float dT = fabsf(Ts - a) + 4e-7f * (fabsf(a) + Ts);
                rho2 = fmaf(dT, dT, sh.perp2[o]) * 1.00001f;
                m = sh.m[o];
                int cc[TA];
#pragma unroll
                for (int kk = TA - 1; kk >= 0; --kk) {
                    const int e = sh.ext[kk][o];
                    const int qd = (int)(((float)rem + 0.5f) * sh.re[kk][o]);     // rem / e without an integer division (rem < 2^22)
                    cc[kk] = sh.clo[kk][o] + (rem - qd * e);
                    rem = qd;
                }
                float d2 = 0.f, umax = 0.f, uabs = 0.f; int base = 0;
#pragma unroll
                for (int kk = 0; kk < TA; ++kk) coop_axis<D>(dv, sh, o, kk, cc[kk], Ts, m, d2, umax, uabs, base);
                if (D <= 3) have = coop_zrange<D>(dv, sh, o, Ts, rho2, m, d2, umax, uabs, base, pa, pb);
                else if (d2 <= rho2) {
                    // cells of axis D-2 the ball's slice reaches, inside the stage's box
                    const int A = D - 2;
                    const float sq = sqrtf(rho2 - d2) * 1.000001f + m;
                    const float cenA = fmaf(Ts, sh.uf[A][o], sh.r32[A][o]);
                    const float iha = dv.inv_h32[A];
                    const int blo = sh.clo[A][o], bhi = blo + sh.ext[A][o] - 1;
                    const float vlo = fminf(fmaxf((cenA - sq) * iha - 2e-3f, (float)blo), (float)bhi);
                    const float vhi = fminf(fmaxf((cenA + sq) * iha + 2e-3f, (float)blo - 1.f), (float)bhi);
                    c2 = (int)floorf(vlo); c2hi = (int)floorf(vhi);
                    d2_p = d2; umax_p = umax; uabs_p = uabs; base_p = base;
                }
            }
            // ---- rows -> chunk tasks: d <= 3 one row per slot; d >= 4 the rows of the task along axis D-2 ------------
            int span = (D <= 3) ? 1 : ((valid && c2hi >= c2) ? c2hi - c2 + 1 : 0);
            if (D >= 4) {
#pragma unroll
                for (int mm = 16; mm >= 1; mm >>= 1) span = max(span, __shfl_xor_sync(FULL, span, mm));
            }
            for (int j = 0; j < span; ++j) {
                if (D >= 4) {
                    have = false;
                    if (valid && c2 + j <= c2hi) {
                        float d2 = d2_p, umax = umax_p, uabs = uabs_p; int base = base_p;
                        coop_axis<D>(dv, sh, o, D - 2, c2 + j, Ts, m, d2, umax, uabs, base);
                        have = coop_zrange<D>(dv, sh, o, Ts, rho2, m, d2, umax, uabs, base, pa, pb);
                    }
                }
                const int len = (have && pb > pa) ? pb - pa : 0;
                if (len) { ls.rows++; ls.cand32 += (u32)len; }
                const int nch = (len + 3) >> 2;
                int incl = nch;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) { const int t_ = __shfl_up_sync(FULL, incl, dd); if (lane >= dd) incl += t_; }
                int pos = total + incl - nch;
                total += __shfl_sync(FULL, incl, 31);
                for (int c = 0; c < nch; ++c, ++pos) {
                    const int p = pa + 4 * c;
                    const int cn = (pb - p < 4) ? pb - p : 4;
                    if (pos < MAXT && p < (1 << 25)) sh.chunk[pos] = ((unsigned)o << 27) | ((unsigned)(cn - 1) << 25) | (unsigned)p;
                    else pool_eval_chunk<D>(dv, sh, o, p, cn);            // no room in the round's array: scanned at once
                }
            }
        }
        __syncwarp();
        // ---- chunks: 32 per step, every lane on the same instruction ---------------------------------------------
        const int M = total < MAXT ? total : MAXT;
        for (int c = lane; c < M; c += 32) {
            const unsigned t_ = sh.chunk[c];
            pool_eval_chunk<D>(dv, sh, (int)(t_ >> 27), (int)(t_ & 0x1ffffffu), (int)((t_ >> 25) & 3u) + 1);
        }
        __syncwarp();
    }
}

// The min-t query of up to 32 rays of a warp (lane l owns ray q iff has_ray).  Every lane of the warp must call.
template <int D, bool POOL>
__device__ __forceinline__ Best coop_min_t(const Dev<D>& dv, CoopShared<D>& sh, const RayQ<D>& q, bool has_ray, LocalStats& ls) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int TA = CoopCfg<D>::TA, CAP = CoopCfg<D>::CAP;
    Best best;
    best.t = INFINITY; best.id = -1; best.t2 = INFINITY; best.tw = 0.0;
    bool active = has_ray, serial = false, tighten = true, redone = false;
    double R0 = 0, R0p = 0, perp2 = 0, scale = dv.probe_scale;
    int stage = 0;
    if (has_ray) {
        ls.raycasts++;
        plane_candidates<D>(dv, q, best);
        R0 = sqrt(q.R0sq);
        R0p = fmax(R0, 0.5 * dv.hmin);
        perp2 = ray_perp2<D>(q);                                  // squared distance of x0 to the ray's line
#pragma unroll
        for (int k = 0; k < D; ++k) {
            sh.uf[k][lane] = (float)q.u[k];
            sh.w2f[k][lane] = (float)(2.0 * (q.r[k] - q.x0[k]));
            sh.x0f[k][lane] = (float)(q.x0[k] - dv.lo[k]);
            sh.r32[k][lane] = (float)(q.r[k] - dv.lo[k]);
        }
#pragma unroll
        for (int e = 0; e < D + 1; ++e) sh.excl[e][lane] = (e < q.nexcl) ? q.excl[e] : -1;
        sh.a32[lane] = (float)q.a;
        sh.perp2[lane] = __double2float_ru(perp2 * (1.0 + 1e-6)) * 1.000001f;
    }
    for (int round = 0;; ++round) {
        // ---- stage setup by the owner lanes (min_t_query's stage head) ----------------------------------------
        int ntask = 0;
        double Ts0 = 0, Tst = 0;
        if (active && round >= 256) { serial = true; active = false; }
        if (active) {
            ls.stages++;
            const double rho_t = scale * R0p;
            Tst = (rho_t > 4.0 * dv.diag) ? INFINITY : q.a + sqrt(fmax(rho_t * rho_t - perp2, 0.0));
            const double Ts = fmin(Tst, best.t);
            if (!(Ts < INFINITY)) { serial = true; active = false; }      // half-space mode: the FP64 path
            else {
                double rho, rho2;
                ray_ball<D>(q, perp2, R0, Ts, rho, rho2);
                int clo[D], chi[D];
                long long nrows = 1; int nt = 1;
                bool empty = false;
                double cmax = 0;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double cen = q.r[k] + Ts * q.u[k];
                    const double gk = (double)dv.g[k];
                    const double vlo = fmin(fmax((cen - rho - dv.lo[k]) * dv.inv_h[k] - 1e-9, 0.0), gk - 1.0);
                    const double vhi = fmin(fmax((cen + rho - dv.lo[k]) * dv.inv_h[k] + 1e-9, -1.0), gk - 1.0);
                    clo[k] = (int)floor(vlo); chi[k] = (int)floor(vhi);
                    if (chi[k] < clo[k]) empty = true;
                    const int e = chi[k] - clo[k] + 1;
                    if (k < D - 1) nrows *= (long long)(e > 0 ? e : 1);
                    if (k < TA) nt *= (e > 0 ? e : 1);
                    cmax = fmax(cmax, fmax(fabs(cen - dv.lo[k]), fabs(q.r[k] - dv.lo[k])));
                }
                const bool use32 = rho < 32.0 * dv.diag && cmax < 32.0 * dv.diag && nrows < (1LL << 22);
                if (!use32) { serial = true; active = false; }
                else {
                    const Filt flt = make_filter<D>(Ts, rho, R0, dv.ext);
                    // FP32 upper bounds certify validity only while the filter's denominator margin dominates the half-space slack
                    if (!((float)(fabs(q.c) * 8e-12) < flt.ed)) tighten = false;
#pragma unroll
                    for (int k = 0; k < D - 1; ++k) {
                        const int e = chi[k] - clo[k] + 1;
                        sh.clo[k][lane] = clo[k]; sh.ext[k][lane] = e; sh.re[k][lane] = 1.0f / (float)(e > 0 ? e : 1);
                    }
                    sh.ts0[lane] = __double2float_ru(Ts) * 1.0000005f;
                    sh.m[lane] = (float)(4e-6 * (dv.ext + cmax + rho + Ts));
                    sh.en[lane] = flt.en; sh.ed[lane] = flt.ed;
                    sh.tb2[lane] = __float_as_uint(flt.tb2);
                    sh.tighten[lane] = tighten ? 1 : 0;
                    sh.cnt[lane] = 0;
                    Ts0 = Ts;
                    ntask = empty ? 0 : nt;
                }
            }
        }
        if (!__any_sync(FULL, active)) break;
        sh.ntask[lane] = ntask;
        sh.nexttask[lane] = 0;
        __syncwarp();
        if (__any_sync(FULL, ntask > 0)) { if (POOL) pool_scan<D>(dv, sh, ls); else coop_scan<D>(dv, sh, ls); }
        __syncwarp();
        // ---- settle: FP64 evaluation of everything that can be the winner or tie with it ------------------------
        if (active) {
            const int c = sh.cnt[lane];
            const float tb2f = __uint_as_float(sh.tb2[lane]);
            const int nlist = c < CAP ? c : CAP;
            bool rejected = false;
            for (int e = 0; e < nlist; ++e) {
                const u64 ent = sh.surv[lane][e];
                const bool bounded = ((ent >> 31) & 1ULL) != 0;
                const float lo = __uint_as_float((unsigned)(ent >> 32));
                if (bounded && !(lo <= tb2f)) continue;
                const bool ok = verify64<D>(dv, q, (int)(ent & 0x7fffffffULL), best, ls);
                if (bounded && !ok) rejected = true;
            }
            if (rejected) tighten = false;                         // FP64 refused a candidate that had tightened the bound: same stage, FP64 decides
            else if (c > CAP) {                                    // survivors were lost: same stage again, now bounded by the verified best
                if (redone) { serial = true; active = false; }
                redone = true;
            }
            else if (best.t <= Ts0 || !(Tst < INFINITY)) active = false;
            else { scale *= dv.probe_growth; redone = false; }
            if (++stage >= 96) active = false;
        }
    }
    if (serial) {
        // rays the pooled FP32 path does not handle: the exact one-lane query (its counters replace this one's)
        ls.raycasts--;
        TileDev<1> tile;
        best = min_t_query_call<D, TileDev<1> >(dv, tile, q, ls);
    }
    return best;
}

// ------------------------------------------------------------------------------------------------------------
// Commit of up to 32 new vertices of a warp (commit_vertex with one atomic per warp on the vertex counter and one on
// the queue tail instead of one per vertex / per opened edge: the single-address atomics of the lane-per-ray commit
// were 13 % of the stall samples at d = 3).  Every lane of the warp must call; `has`: this lane holds a vertex.
// ------------------------------------------------------------------------------------------------------------
// read-only probe of the vertex set: true if sig is stored already, else `slot` is the first empty slot met
template <int D>
__device__ __forceinline__ bool vertex_probe(const Dev<D>& dv, const int* sig, u64 h, u64 s, u64& slot) {
    u64 fp = (h >> 32) << 32;
    if (fp == 0) fp = 1ULL << 32;
    slot = h & dv.vmask;
    for (;;) {
        if (s == 0) return false;
        if ((s >> 32) == (fp >> 32) && sig_equal<D>(dv, (u32)(s & 0xffffffffu) - 1u, sig)) return true;
        slot = (slot + 1) & dv.vmask;
        s = ld_cg(dv.vtab + slot);
    }
}
// publishes record `mine` (already written) starting at `slot`; false: another walk stored the same vertex meanwhile
template <int D>
__device__ __forceinline__ bool vertex_publish(const Dev<D>& dv, const int* sig, u64 h, u64 slot, u32 mine, LocalStats& ls) {
    u64 fp = (h >> 32) << 32;
    if (fp == 0) fp = 1ULL << 32;
    u64 s = 0;
    for (;;) {
        if (s == 0) {
            s = atom_cas(dv.vtab + slot, 0ULL, fp | (u64)(mine + 1u));
            if (s == 0) return true;
        }
        if ((s >> 32) == (fp >> 32) && sig_equal<D>(dv, (u32)(s & 0xffffffffu) - 1u, sig)) {
            dv.vsig[(size_t)mine * (D + 1)] = -1; ls.dead++; ls.dup_hits++;      // lost a race: dead record
            return false;
        }
        slot = (slot + 1) & dv.vmask;
        s = ld_cg(dv.vtab + slot);
    }
}

template <int D>
__device__ __forceinline__ void commit_vertex_warp(const Dev<D>& dv, bool has, const int (&sig)[D + 1], const double (&r)[D],
                                                   const WalkQueue& wq, LocalStats& ls) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    // every independent load is issued up front: the vertex set's home slot and the home slots of the sub-facets
    u64 hv = 0, sv = 0, hs[D + 1], s0[D + 1];
    unsigned minemask = 0;
    if (has) {
        unsigned actmask = 0;
#pragma unroll
        for (int i = 0; i < D + 1; ++i)
            if ((sig[i] < dv.n) && (dv.active[sig[i]] != 0)) actmask |= 1u << i;
        hv = hash_ids<D>(sig, D + 1, -1);
        sv = ld_cg(dv.vtab + (hv & dv.vmask));
#pragma unroll
        for (int k = 0; k < D + 1; ++k) {
            hs[k] = 0; s0[k] = 0;
            if ((actmask & ~(1u << k)) != 0) {                   // the sub-facet keeps an explored real generator
                minemask |= 1u << k;
                hs[k] = hash_ids<D>(sig, D + 1, k);
                s0[k] = ld_cg(dv.etab + (hs[k] & dv.emask));
            }
        }
    }
    u64 vslot = 0;
    if (has && vertex_probe<D>(dv, sig, hv, sv, vslot)) { ls.dup_hits++; has = false; }
    // record indices: one atomic per warp
    u32 v = 0xffffffffu;
    {
        const unsigned am = __ballot_sync(FULL, has);
        if (am) {
            const int leader = __ffs(am) - 1;
            u32 base = 0;
            if (lane == leader) base = atomicAdd(dv.vcount, (u32)__popc(am));
            base = __shfl_sync(FULL, base, leader);
            if (has) {
                v = base + (u32)__popc(am & lt);
                if (v >= dv.vcap) { atom_or(&dv.ctr->flags, (u32)FLAG_VFULL); has = false; }
            }
        }
    }
    if (has) {
        int* ps = dv.vsig + (size_t)v * (D + 1);
        double* pr = dv.vr + (size_t)v * D;
#pragma unroll
        for (int k = 0; k < D + 1; ++k) ps[k] = sig[k];
#pragma unroll
        for (int k = 0; k < D; ++k) pr[k] = r[k];
        mem_fence();
        if (!vertex_publish<D>(dv, sig, hv, vslot, v, ls)) has = false;
    }
    // sub-facets: first endpoint -> open edge (queued below), second endpoint -> closed
    u32 pslot[D + 1];
    unsigned openmask = 0;
    if (has) {
#pragma unroll
        for (int k = 0; k < D + 1; ++k)
            if (sig[k] < dv.n) dv.has_vertex[sig[k]] = 1;
        int dummy[D + 1];
        // the first insertion attempt of ALL sub-facets is issued before any result is looked at: D + 1 independent
        // atomics in flight instead of D + 1 memory round trips in a row (a home slot seen empty is claimed here; the
        // sub-facets whose home slot is taken, or whose claim lost a race, continue in edge_register from what they saw)
        u64 s1[D + 1];
#pragma unroll
        for (int k = 0; k < D + 1; ++k) {
            s1[k] = s0[k];
            if (((minemask >> k) & 1u) && s0[k] == 0) s1[k] = atom_cas(dv.etab + (hs[k] & dv.emask), 0ULL, edge_slot(hs[k], v, k));
        }
#pragma unroll
        for (int k = 0; k < D + 1; ++k) {
            pslot[k] = 0;
            if (!((minemask >> k) & 1u)) continue;
            u64 slot;
            if (s0[k] == 0 && s1[k] == 0) slot = hs[k] & dv.emask;                     // claimed above: first endpoint, open edge
            else slot = edge_register<D>(dv, sig, v, k, hs[k], s1[k], false, dummy);
            if (slot != ~0ULL) { pslot[k] = (u32)slot; openmask |= 1u << k; }
        }
    }
    // queue tail: one atomic per warp
    {
        const int mycnt = __popc(openmask);
        int incl = mycnt;
#pragma unroll
        for (int d_ = 1; d_ < 32; d_ <<= 1) { const int o_ = __shfl_up_sync(FULL, incl, d_); if (lane >= d_) incl += o_; }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total) {
            u32 base = 0;
            if (lane == 31) base = atomicAdd(wq.tail, (u32)total);
            base = __shfl_sync(FULL, base, 31);
            u32 pos = base + (u32)(incl - mycnt);
#pragma unroll
            for (int k = 0; k < D + 1; ++k) {
                if (!((openmask >> k) & 1u)) continue;
                if (pos < wq.cap) wq.q[pos] = frontier_entry((u64)pslot[k], v, k);
                else atom_or(&dv.ctr->flags, (u32)FLAG_QFULL);
                ++pos;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// The persistent frontier walk (see k_walk) on the cooperative query.  Tickets of a warp are taken with one atomic.
// ------------------------------------------------------------------------------------------------------------
#ifndef HVB_COOP_MINB
#define HVB_COOP_MINB 4
#endif
#ifndef HVB_COOP_COMMIT
#define HVB_COOP_COMMIT 1         // 1: warp-aggregated commit (commit_vertex_warp), 0: the lane-per-ray commit_vertex
#endif
// QMODE 0: every lane runs the one-lane query of its own ray; 1: pooled query with dynamic row tickets (coop_scan);
// 2: pooled query with static scheduling (pool_scan)
template <int D, int QMODE>
static __global__ void __launch_bounds__(128, HVB_COOP_MINB) k_walk_coop(Dev<D> dv, WalkQueue wq) {
    extern __shared__ __align__(16) unsigned char hvb_smem_raw[];
    CoopShared<D>& sh = reinterpret_cast<CoopShared<D>*>(hvb_smem_raw)[threadIdx.x >> 5];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    TileDev<1> tile;
    LocalStats ls = {};
    u32 ticket = 0xffffffffu;          // a ticket whose entry is not published yet is kept for the next trip
    u32 my_done = 0;                   // processed entries not yet added to *done
    long long t_idle = clock64();
    for (u32 trip = 0;; ++trip) {
        // non-general position, a full table or the safety abort stop the walk at once (the host reports / regrows)
        // (looked at every 8th trip: the load is a full memory round trip in front of every trip otherwise, 6 % of the
        // stall samples of C2; an overflow / a non-general vertex only has to stop the walk soon, not at once)
        if ((trip & 7u) == 0u) {
            u32 stop = 0;
            if (lane == 0) {
                const u32 fl = __ldcg(&dv.ctr->flags);
                stop = ((fl & FLAG_OVERFLOW_MASK) || (wq.stop_on_degenerate && (fl & FLAG_DEGEN)) || __ldcg(wq.abort)) ? 1u : 0u;
            }
            if (__shfl_sync(FULL, stop, 0)) break;
        }
        u64 item = 0;
        bool live = false;
        for (int tries = 0; tries < 4; ++tries) {
            const bool need = !live && ticket == 0xffffffffu;
            const unsigned nm = __ballot_sync(FULL, need);
            if (nm) {
                const int leader = __ffs(nm) - 1;
                u32 base = 0;
                if (lane == leader) base = atomicAdd(wq.head, (u32)__popc(nm));
                base = __shfl_sync(FULL, base, leader);
                if (need) ticket = base + (u32)__popc(nm & ((1u << lane) - 1u));
            }
            bool skipped = false;
            if (!live) {
                // a ticket beyond the queue's capacity can never be served (pushes beyond it are refused and flagged)
                const u64 it = (ticket < wq.cap) ? __ldcg(wq.q + ticket) : HVB_Q_EMPTY;
                if (it != HVB_Q_EMPTY) {
                    const u64 s = __ldcg(dv.etab + (u32)(it >> 32));
                    if (s >> 63) {                                       // else: the edge slot is not visible yet, retry next trip
                        ticket = 0xffffffffu;
                        if (s & EDGE_CLOSED) { ls.closed_skips++; ++my_done; skipped = true; }
                        else { item = it; live = true; }
                    }
                }
            }
            if (!__any_sync(FULL, skipped)) break;
        }
        if (__any_sync(FULL, live)) {
            int sig[D + 1];
            RayQ<D> q;
            u32 v = 0; int kd = 0;
            bool ok = false;
            if (live) {
                ok = ray_setup<D>(dv, item, q, sig, v, kd);
                if (!ok) ls.seed_fail++;
            }
            // COOPQ: the pooled query; else every lane runs the one-lane query for its own ray (only the acquisition
            // and the commit are warp-aggregated)
            Best best;
            if (QMODE == 2) best = coop_min_t<D, true>(dv, sh, q, ok, ls);
            else if (QMODE == 1) best = coop_min_t<D, false>(dv, sh, q, ok, ls);
            else { best.t = INFINITY; best.id = -1; best.t2 = INFINITY; best.tw = 0.0; if (ok) best = min_t_query<D, TileDev<1> >(dv, tile, q, ls); }
#if HVB_COOP_COMMIT
            int sig2[D + 1];
            double r2[D];
            bool has = false;
            if (ok) has = ray_result<D>(dv, 0, q, sig, v, kd, best, sig2, r2, ls);
            commit_vertex_warp<D>(dv, has, sig2, r2, wq, ls);
#else
            if (ok) ray_finish<D, TileDev<1> >(dv, tile, q, sig, v, kd, best, wq.q, wq.tail, wq.cap, ls);
#endif
            if (live) ++my_done;
            t_idle = clock64();
        } else {
            // the whole warp is idle: publish progress, test for global completion, back off
            u32 dsum = my_done;
#pragma unroll
            for (int m_ = 16; m_ >= 1; m_ >>= 1) dsum += __shfl_xor_sync(FULL, dsum, m_);
            my_done = 0;
            u32 dn = 0, tl = 0, ab = 0;
            if (lane == 0) {
                if (dsum) { __threadfence(); atomicAdd(wq.done, dsum); }
                dn = __ldcg(wq.done); tl = __ldcg(wq.tail); ab = __ldcg(wq.abort);
            }
            dn = __shfl_sync(FULL, dn, 0); tl = __shfl_sync(FULL, tl, 0); ab = __shfl_sync(FULL, ab, 0);
            // done == tail with every ticket of the warp beyond tail: nothing can ever be published again
            const bool fin = ab || (dn == tl && ticket >= tl);
            if (__all_sync(FULL, fin)) break;
            __nanosleep(200);
            // a bug must never hang the device: 20 s without work anywhere in the warp
            if ((trip & 1023u) == 1023u && clock64() - t_idle > 40000000000LL) atomicExch(wq.abort, 1u);
        }
        __syncwarp();
    }
    {
        u32 dsum = my_done;
#pragma unroll
        for (int m_ = 16; m_ >= 1; m_ >>= 1) dsum += __shfl_xor_sync(FULL, dsum, m_);
        if (lane == 0 && dsum) { __threadfence(); atomicAdd(wq.done, dsum); }
    }
    flush_stats(ls, dv.ctr);
}

}  // namespace hvb
