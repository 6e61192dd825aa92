// sphb_walk.cuh — group neighbour search: replaces BHTree::neighbor_search (src/bhtree.cpp:114-126,
// 234-270) and exhaustive_search (src/exhaustive_search.cpp:11-42) for a whole warp at once.
//
// A group = 32 consecutive particles of the tree order = one warp, lane = particle i.
//   1. Tree descent, breadth-first with ONE LANE PER NODE: up to 32 nodes are popped from a
//      shared-memory stack, every lane fetches its node's 64-byte record and tests the node's cube
//      against the group's bounding box grown by the search radius (for the symmetric search: by
//      max(radius, BHNode::kernel_size of the node), src/bhtree.cpp:237).  Children of hit nodes are
//      pushed, hit leaves are streamed.  No per-particle pointer chasing, 32 node loads in flight.
//   2. Candidate streaming: the particles of the hit leaves are staged 32 at a time into a
//      shared-memory tile (group-relative FP32 coordinates); every lane runs the SAME loop over the
//      tile (convergent, broadcast reads) with a conservative FP32 distance test that only produces
//      a hit mask; the FP32 pipe is otherwise idle in this code and runs at twice the FP64 rate.
//   3. Each lane hands the indices of its (few) hits to the visitor; the caller then applies the
//      reference's exact FP64 predicate (r2 < h2 with the reference's operation order) in a
//      convergent pass over the recorded indices — so the neighbour SET is exactly the
//      reference's; the tree and the FP32 test only ever discard pairs that cannot pass.
// The cull is conservative by construction (margins below), hence the set equals the
// EXHAUSTIVE_SEARCH set, which the reference's tree search reproduces as well.
#pragma once
#include "sphb_tree.cuh"

namespace sphb {

constexpr int NW_STACK = 640;      // node stack entries per warp: a batch leaves <= 2^DIM (8 - 1) * 4 = 28 entries per level behind, 28 * 21 + 32 = 620 for the
                                   // deepest 3-D tree (maxTreeLevel 21); deeper 1-D / 2-D trees that outgrow it raise "node stack overflow"

struct NWalkSmem {
    int     stack[NW_STACK];       // child0 | (nchild - 1) << 29: all children of a hit node
    int     expand[32];            // nodes of the batch being fetched
    int     raw[32];               // particle indices of the hit leaves, 32 at a time, before staging
    float4  tile32[64];            // staged candidates that can be within reach of some lane (two blocks of 32, used as a ring):
                                   //   {x - ref, y - ref, z - ref, symmetric threshold} in FP32, scaled
    int     tilej[64];             //   particle index
};

// error bits reported through d_err[2]
enum { WALK_ERR_STACK = 1, WALK_ERR_GRAV_STACK = 2 };

// Bounding box of the group's valid lanes: centre and half width per axis (warp-uniform).
template <int DIM>
__device__ __forceinline__ void group_box(const double (&ri)[DIM], bool valid, double (&bc)[DIM], double (&bh)[DIM])
{
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        const double lo = warp_min(valid ? ri[d] : 1.7976931348623157e308);
        const double hi = warp_max(valid ? ri[d] : -1.7976931348623157e308);
        bc[d] = 0.5 * lo + 0.5 * hi;
        bh[d] = fmax(hi - bc[d], bc[d] - lo) * (1.0 + 1e-12);
    }
}

// Per-lane state of the FP32 pre-filter.
template <int DIM> struct Filter32 {
    float fi[DIM];        // x_i - ref
    float thr;            // squared threshold of lane i (negative: lane inactive)
    float L[DIM];         // periodic range (FP32), only if periodic
    double ref[DIM];      // = group box centre
    double rlim;          // candidates farther than this from ref (max norm) cannot be gather neighbours
    double delta;         // absolute per-axis error bound of the FP32 coordinates inside rlim
    double sc;            // warp-uniform power of two ~ 1 / rlim: every FP32 quantity is staged as (value * sc), so the
                          // filter is independent of the unit system (h ~ 1e20 cm or 1e-20 must not leave the FP32 range)
};

// V:  void hit(int j)   — particle j passed lane's conservative test; the exact test is the caller's.
// reach = search radius of the group (max over lanes), h_i = lane's own radius (gather: h_search,
// symmetric: sml_i).  hj_src / hj_stride: where h_j lives for the symmetric search.
template <int DIM, bool SYM, class V>
__device__ __forceinline__ void group_stream(const TreeDev & t, const DevParams & P, const double4 * __restrict__ posm,
                                             const double * __restrict__ hj_src, int hj_stride,
                                             NWalkSmem & sm, int lane, const double (&ri)[DIM], double h_i, bool valid,
                                             V & v, unsigned long long * __restrict__ d_err)
{
    double bc[DIM], bh[DIM];
    group_box<DIM>(ri, valid, bc, bh);
    const double reach = warp_max(valid ? h_i : 0.0);
    double bhmax = bh[0];
#pragma unroll
    for (int d = 1; d < DIM; ++d) bhmax = fmax(bhmax, bh[d]);
    // slack for the rounding of node centres / box arithmetic (coordinates are O(root edge))
    double cmax = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) cmax = fmax(cmax, fabs(bc[d]) + bh[d]);
    const double slack = 1e-13 * (cmax + reach) + 1e-300;

    // ---- FP32 filter set-up
    Filter32<DIM> F;
    double lmax = 0.0;
    F.rlim = (bhmax + reach) * 1.001 + slack;
    F.sc = scalbn(1.0, -ilogb(F.rlim));                  // exact scaling: staged coordinates inside rlim are in (-2, 2)
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        F.ref[d] = bc[d];
        F.fi[d] = (float)((ri[d] - bc[d]) * F.sc);
        F.L[d] = (float)(P.range[d] * F.sc);
        if (P.periodic) lmax = fmax(lmax, P.range[d]);
    }
    F.delta = 1.1920929e-7 * (F.rlim + lmax);            // 2^-23 * magnitude bound
    {
        const double hh = (h_i + 4.0 * F.delta) * F.sc;
        F.thr = valid ? (float)(hh * hh * (1.0 + 4e-6)) : -1.0f;
    }

    int fill = 0;                                        // raw indices waiting to be staged
    int st0 = 0, nst = 0;                                // ring of staged candidates: block start (0 or 32), entries
    const unsigned lt_mask = (1u << lane) - 1u;

    // every lane tests the m staged candidates of the block at st0 (conservative FP32 test) and hands its hits on
    auto process = [&](int m) {
        const float4 * tile = sm.tile32 + st0;
        unsigned hits = 0;
        if (P.periodic) {
#pragma unroll 8
            for (int k = 0; k < m; ++k) {
                const float4 c = tile[k];
                float ax = fabsf(F.fi[0] - c.x);
                ax = fminf(ax, F.L[0] - ax);
                float r2 = ax * ax;
                if (DIM >= 2) { float ay = fabsf(F.fi[DIM >= 2 ? 1 : 0] - c.y); ay = fminf(ay, F.L[DIM >= 2 ? 1 : 0] - ay); r2 = fmaf(ay, ay, r2); }
                if (DIM >= 3) { float az = fabsf(F.fi[DIM >= 3 ? 2 : 0] - c.z); az = fminf(az, F.L[DIM >= 3 ? 2 : 0] - az); r2 = fmaf(az, az, r2); }
                const float th = SYM ? fmaxf(F.thr, c.w) : F.thr;
                if (r2 < th) hits |= 1u << k;
            }
        } else {
#pragma unroll 8
            for (int k = 0; k < m; ++k) {
                const float4 c = tile[k];
                const float dx = F.fi[0] - c.x;
                float r2 = dx * dx;
                if (DIM >= 2) { const float dy = F.fi[DIM >= 2 ? 1 : 0] - c.y; r2 = fmaf(dy, dy, r2); }
                if (DIM >= 3) { const float dz = F.fi[DIM >= 3 ? 2 : 0] - c.z; r2 = fmaf(dz, dz, r2); }
                const float th = SYM ? fmaxf(F.thr, c.w) : F.thr;
                if (r2 < th) hits |= 1u << k;
            }
        }
        if (!valid) hits = 0;
        while (hits) {
            const int kb = __ffs(hits) - 1;
            hits &= hits - 1;
            v.hit(sm.tilej[st0 + kb]);
        }
        __syncwarp();                                    // the block is refilled later: every lane is done reading it
    };

    // FP32 image of the m raw candidates; those that cannot be within reach of ANY lane of the group (the hit leaves
    // stick out of the group's search region by up to a leaf edge: about half of them) are dropped here, the rest is
    // appended to the ring, and a full block of 32 is tested
    auto stage = [&](int m) {
        __syncwarp();
        bool keep = false;
        float4 f = make_float4(0.f, 0.f, 0.f, -1.0f);
        int j = 0;
        if (lane < m) {
            j = sm.raw[lane];
            const double4 pj = ldg4(&posm[j]);
            double dj[DIM];
            dj[0] = pj.x - F.ref[0];
            if (DIM >= 2) dj[DIM >= 2 ? 1 : 0] = pj.y - F.ref[DIM >= 2 ? 1 : 0];
            if (DIM >= 3) dj[DIM >= 3 ? 2 : 0] = pj.z - F.ref[DIM >= 3 ? 2 : 0];
            double amax = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                if (P.periodic) dj[d] = min_image(dj[d], P.range[d]);
                amax = fmax(amax, fabs(dj[d]));
            }
            f = make_float4((float)(dj[0] * F.sc), DIM >= 2 ? (float)(dj[DIM >= 2 ? 1 : 0] * F.sc) : 0.f,
                            DIM >= 3 ? (float)(dj[DIM >= 3 ? 2 : 0] * F.sc) : 0.f, -1.0f);
            if (SYM) {
                // a hit needs r < max(h_i, h_j), hence |x_j - ref|_inf < bhmax + max(reach, h_j)
                const double hj = __ldg(hj_src + (size_t)j * hj_stride);
                keep = amax <= (bhmax + fmax(reach, hj)) * 1.001 + slack;
                const double dl = 1.1920929e-7 * (fmax(amax, F.rlim) + lmax);
                const double hh = (hj + 4.0 * (dl + F.delta)) * F.sc;
                f.w = (float)(hh * hh * (1.0 + 4e-6));   // +inf for an h_j beyond FP32 range: passes (conservative)
            } else {
                keep = amax <= F.rlim;                   // farther than rlim from ref: not within h_search of any lane
            }
        }
        const unsigned kb = __ballot_sync(SPHB_FULL_MASK, keep);
        if (keep) {
            const int pos = (st0 + nst + __popc(kb & lt_mask)) & 63;
            sm.tile32[pos] = f;
            sm.tilej[pos] = j;
        }
        nst += __popc(kb);
        __syncwarp();
        if (nst >= 32) {
            process(32);
            st0 ^= 32;
            nst -= 32;
        }
    };
    // queue the particles [first, first + count) of a hit leaf; 32 raw indices at a time are staged
    auto feed = [&](int first, int count) {
        int off = 0;
        while (off < count) {
            const int take = min(count - off, 32 - fill);
            if (lane < take) sm.raw[fill + lane] = first + off + lane;
            fill += take;
            off += take;
            if (fill == 32) {
                stage(32);
                fill = 0;
            }
        }
    };

    // ---- breadth-first descent, one lane per node.  A stack entry stands for all (contiguous)
    // children of a hit node: a batch of <= 32 nodes pushes <= 32 entries and pops >= 32 / 2^DIM.
    int top = 1;
    if (lane == 0) sm.stack[0] = 0;                      // the root alone: child0 = 0, nchild = 1
    __syncwarp();
    while (top > 0) {
        const int ne = min(top, 32);
        int ent = 0, nc = 0;
        if (lane < ne) { ent = sm.stack[top - 1 - lane]; nc = (int)((unsigned)ent >> 29) + 1; }
        int incl = nc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(SPHB_FULL_MASK, incl, o);
            if (lane >= o) incl += y;
        }
        const int m = __popc(__ballot_sync(SPHB_FULL_MASK, lane < ne && incl <= 32));   // entries taken (a prefix)
        const int k = __shfl_sync(SPHB_FULL_MASK, incl, m - 1);
        if (lane < m) {
            const int c0 = ent & 0x1fffffff;
            for (int ci = 0; ci < nc; ++ci) sm.expand[incl - nc + ci] = c0 + ci;
        }
        __syncwarp();
        int node = -1;
        if (lane < k) node = sm.expand[lane];
        top -= m;
        __syncwarp();
        bool hit = false;
        int child0 = 0, nchild = 0, first = 0, count = 0;
        if (node >= 0) {
            const double2 * q = t.nn + (size_t)node * 4;
            const double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            const double half = 0.5 * q1.y;
            const double rn = SYM ? fmax(reach, q2.x) : reach;
            double c[DIM];
            c[0] = q0.x;
            if (DIM >= 2) c[DIM >= 2 ? 1 : 0] = q0.y;
            if (DIM >= 3) c[DIM >= 3 ? 2 : 0] = q1.x;
            double g2 = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double dc = bc[d] - c[d];
                if (P.periodic) dc = min_image(dc, P.range[d]);
                const double gap = fabs(dc) - (bh[d] + half) - slack;
                if (gap > 0.0) g2 += gap * gap;
            }
            hit = g2 <= rn * rn * (1.0 + 1e-9);
            if (hit) {
                child0 = __double2loint(q2.y); nchild = __double2hiint(q2.y);
                if (nchild == 0) {
                    const double2 q3 = __ldg(q + 3);
                    first = __double2loint(q3.x); count = __double2hiint(q3.x);
                }
            }
        }
        const bool is_leaf = hit && nchild == 0;
        unsigned leaf_b = __ballot_sync(SPHB_FULL_MASK, is_leaf);
        const unsigned int_b = __ballot_sync(SPHB_FULL_MASK, hit && nchild > 0);
        if (int_b) {
            const int total = __popc(int_b);
            if (top + total > NW_STACK) {
                if (lane == 0) atomicOr(&d_err[2], (unsigned long long)WALK_ERR_STACK);
            } else {
                if ((int_b >> lane) & 1u) sm.stack[top + __popc(int_b & lt_mask)] = child0 | ((nchild - 1) << 29);
                top += total;
            }
            __syncwarp();
        }
        while (leaf_b) {
            const int src = __ffs(leaf_b) - 1;
            leaf_b &= leaf_b - 1;
            const int f0 = __shfl_sync(SPHB_FULL_MASK, first, src);
            const int c0 = __shfl_sync(SPHB_FULL_MASK, count, src);
            feed(f0, c0);
        }
    }
    if (fill > 0) stage(fill);
    if (nst > 0) process(nst);
    __syncwarp();
}

} // namespace sphb
