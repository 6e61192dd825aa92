// sphb_tree.cuh — device replacement of sph::BHTree (include/bhtree.hpp, src/bhtree.cpp).
//
// The reference builds a pointer octree by serial recursive insertion (src/bhtree.cpp:136-204).
// Here the SAME node set is produced without pointers:
//   1. every particle walks the reference's own descent (`pos[d] > center[d]`, child centre
//      = centre +- edge/4; src/bhtree.cpp:168-196) for key_levels levels and records the child
//      index of each level in a 64-bit key (most significant = first split);
//   2. (key, index) pairs are radix-sorted (cub) and the particle SoA is permuted, so every
//      tree node is a contiguous particle range;
//   3. nodes are emitted level by level from the sorted keys with the reference's split rule
//      (`num > leaf_particle_num && parent.level < max_level`, src/bhtree.cpp:154-158; the
//      root always splits), children in child-index order;
//   4. mass / mass centre go bottom-up; the nodes stay in this breadth-first order, in which the
//      children of a node are contiguous (child0 .. child0 + nchild - 1), so that a walk can
//      expand 32 nodes at a time with one lane per node (sphb_walk.cuh).
#pragma once
#include "sphb_math.cuh"
#include <limits.h>

namespace sphb {

// Particle state, structure of arrays, always in tree (sorted-key) order.
struct PSoA {
    double *pos[3], *vel[3], *vel_p[3], *acc[3];
    double *mass, *dens, *pres, *ene, *ene_p, *dene, *sml, *sound, *balsara, *alpha, *gradh, *phi;
    int    *pid, *neighbor;        // SPHParticle::id, SPHParticle::neighbor
    int    *orig;                  // index of this particle in the caller's AoS buffer
    // GSPH MUSCL gradients (src/solver.cpp:373-385): grad_density, grad_pressure, grad_velocity_k
    double *grad_d[3], *grad_p[3], *grad_v[3][3];
};
constexpr int PSOA_NDOUBLE = 12 + 12;   // permuted double arrays (without gradients)

// Nodes under construction (BFS order).
struct TreeBuild {
    int *first, *count, *level, *parent, *child0, *nchild;
    double *center[3];
    double4 *msum4;            // {sum m, sum m x, sum m y, sum m z}: one contiguous array (all-reduced in the multi-GPU mode)
};

// Finished tree, breadth-first order, one packed 64-byte record (4 x double2) per node and walk
// kind, so that a lane fetches "its" node with four 16-byte loads off one address:
//   nn (neighbour walks): [0] cx, cy   [1] cz, edge   [2] ksize,  {child0, nchild}   [3] {first, count}, -
//   ng (gravity walk):    [0] mx, my   [1] mz, mass   [2] edge^2, {child0, nchild}   [3] {first, count}, -
// (c = geometric centre, m = mass centre, ksize = BHNode::kernel_size set by set_kernel,
//  child0/nchild = contiguous children (nchild == 0: leaf), first/count = particle range)
struct TreeDev {
    int      n_nodes;
    double2 *nn;
    double2 *ng;
    int     *parent;    // index of the parent, -1 for the root
    double  *ksize;     // BHNode::kernel_size while it is being reduced (contiguous: all-reduced (max) in the multi-GPU
                        // mode), then copied into the nn / ng records by k_apply_ksize
};
__device__ __forceinline__ double pack_ints(int lo, int hi) { return __hiloint2double(hi, lo); }

// root[0..2] = centre, root[3] = edge.
// ---- bounding cube: BHTree::make, src/bhtree.cpp:59-95 -----------------------------------------
template <int DIM>
__global__ void k_bbox_partial(PSoA p, int n, double * __restrict__ part /* [grid][2*DIM] */)
{
    double lo[DIM], hi[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { lo[d] = 1.7976931348623157e308; hi[d] = -1.7976931348623157e308; }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            const double x = p.pos[d][i];
            lo[d] = fmin(lo[d], x);
            hi[d] = fmax(hi[d], x);
        }
    }
    __shared__ double s[32][2 * DIM];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < DIM; ++d) { lo[d] = warp_min(lo[d]); hi[d] = warp_max(hi[d]); }
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) { s[w][d] = lo[d]; s[w][DIM + d] = hi[d]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int k = 1; k < nw; ++k) {
#pragma unroll
            for (int d = 0; d < DIM; ++d) { lo[d] = fmin(lo[d], s[k][d]); hi[d] = fmax(hi[d], s[k][DIM + d]); }
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d) { part[blockIdx.x * 2 * DIM + d] = lo[d]; part[blockIdx.x * 2 * DIM + DIM + d] = hi[d]; }
    }
}

template <int DIM>
__global__ void k_bbox_final(const double * __restrict__ part, int nblocks, double * __restrict__ root)
{
    // one warp
    double lo[DIM], hi[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { lo[d] = 1.7976931348623157e308; hi[d] = -1.7976931348623157e308; }
    for (int b = threadIdx.x; b < nblocks; b += 32) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) { lo[d] = fmin(lo[d], part[b * 2 * DIM + d]); hi[d] = fmax(hi[d], part[b * 2 * DIM + DIM + d]); }
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) { lo[d] = warp_min(lo[d]); hi[d] = warp_max(hi[d]); }
    if (threadIdx.x == 0) {
        double l = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            root[d] = (hi[d] + lo[d]) * 0.5;           // src/bhtree.cpp:85
            const double range = hi[d] - lo[d];
            if (l < range) l = range;                  // src/bhtree.cpp:87-93
        }
        for (int d = DIM; d < 3; ++d) root[d] = 0.0;
        root[3] = l;
    }
}

// ---- keys: the reference's descent, src/bhtree.cpp:163-196 -------------------------------------
template <int DIM>
__device__ __forceinline__ unsigned long long key_of(const PSoA & p, int i, const double * __restrict__ root, int key_levels)
{
    double c[DIM], x[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { c[d] = root[d]; x[d] = p.pos[d][i]; }
    double edge = root[3];
    unsigned long long key = 0;
    for (int l = 0; l < key_levels; ++l) {
        unsigned int bits = 0;
        const double q = edge * 0.25;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            if (x[d] > c[d]) { bits |= 1u << d; c[d] = __dadd_rn(c[d], q); }
            else             {                  c[d] = __dadd_rn(c[d], -q); }
        }
        key = (key << DIM) | bits;
        edge *= 0.5;
    }
    return key;
}
template <int DIM>
__global__ void k_keys(PSoA p, int i_begin, int i_end, const double * __restrict__ root, int key_levels,
                       unsigned long long * __restrict__ keys, int * __restrict__ idx)
{
    const int i = i_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i_end) return;
    keys[i] = key_of<DIM>(p, i, root, key_levels);
    idx[i]  = i;
}

// ---- level-by-level node emission ---------------------------------------------------------------
__device__ __forceinline__ int lower_bound_child(const unsigned long long * __restrict__ keys, int lo, int hi,
                                                 int shift, unsigned int mask, unsigned int c)
{
    // first position in [lo, hi) whose child index at this level is >= c
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const unsigned int v = (unsigned int)(keys[mid] >> shift) & mask;
        if (v < c) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int DIM>
__device__ __forceinline__ bool node_splits(const TreeBuild & t, int i, int leaf_num, int max_level_eff)
{
    // root: BHTree::make always calls m_root.create_tree (src/bhtree.cpp:106);
    // child: `child->num > leaf_particle_num && level < max_level` with level = the parent's
    // (src/bhtree.cpp:154), i.e. own level <= max_level.
    if (i == 0) return true;
    return t.count[i] > leaf_num && t.level[i] <= max_level_eff;
}

// One lane per CHILD (2^DIM lanes per node): lane c finds the upper boundary of child c by its own
// binary search over the node's key range, so the 2^DIM searches of a node run concurrently — the top
// levels have a handful of nodes with millions of keys each and were pure dependent-load latency with
// one thread per node.
template <int DIM>
__device__ __forceinline__ void child_range(const TreeBuild & t, const unsigned long long * __restrict__ keys, int node, bool splits,
                                            int key_levels, int c, int & b, int & e)
{
    constexpr int NCHILD = 1 << DIM;
    b = 0; e = 0;
    if (splits) {
        const int first = t.first[node], last = first + t.count[node];
        const int shift = (key_levels - t.level[node]) * DIM;
        e = (c == NCHILD - 1) ? last : lower_bound_child(keys, first, last, shift, NCHILD - 1, c + 1);
        b = first;
    }
    const int prev = __shfl_up_sync(SPHB_FULL_MASK, e, 1);
    if (splits && c > 0) b = prev;
}

template <int DIM>
__global__ void k_level_count(TreeBuild t, const unsigned long long * __restrict__ keys, int lvl_begin, int lvl_end,
                              int leaf_num, int max_level_eff, int key_levels, int * __restrict__ tmp,
                              const int * __restrict__ lvl /* speculative build: {begin, end} on the device */, int w_cap,
                              int true_max_level, int * __restrict__ deeper)
{
    constexpr int NCHILD = 1 << DIM;
    if (lvl) { lvl_begin = lvl[0]; lvl_end = lvl[1]; }
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = lvl_begin + tid / NCHILD, c = tid % NCHILD;
    const bool in = i < lvl_end;
    const bool splits = in && node_splits<DIM>(t, i, leaf_num, max_level_eff);
    // the keys are only sorted down to level max_level_eff (partial radix sort): a node below that which the
    // reference would still split means the sort has to be redone with more levels
    if (in && c == 0 && i != 0 && !splits && t.count[i] > leaf_num && t.level[i] <= true_max_level) *deeper = 1;
    int b, e;
    child_range<DIM>(t, keys, i, splits, key_levels, c, b, e);
    const unsigned full = __ballot_sync(SPHB_FULL_MASK, e > b);
    const int lane = threadIdx.x & 31;
    const unsigned grp = (full >> (lane - c)) & ((1u << NCHILD) - 1u);
    if (in && c == 0) tmp[i - lvl_begin] = __popc(grp);
    else if (lvl && c == 0 && tid / NCHILD < w_cap) tmp[tid / NCHILD] = 0;       // the scan runs over w_cap entries
}

template <int DIM>
__global__ void k_level_emit(TreeBuild t, const unsigned long long * __restrict__ keys, int lvl_begin, int lvl_end,
                             int leaf_num, int max_level_eff, int key_levels, const int * __restrict__ offs,
                             const double * __restrict__ root,
                             const int * __restrict__ lvl /* speculative build: {begin, end, next end} */, const int * __restrict__ bad)
{
    constexpr int NCHILD = 1 << DIM;
    if (lvl) { lvl_begin = lvl[0]; lvl_end = lvl[1]; if (*bad) return; }
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = lvl_begin + tid / NCHILD, c = tid % NCHILD;
    const bool in = i < lvl_end;
    const bool splits = in && node_splits<DIM>(t, i, leaf_num, max_level_eff);
    int b, e;
    child_range<DIM>(t, keys, i, splits, key_levels, c, b, e);
    const unsigned full = __ballot_sync(SPHB_FULL_MASK, e > b);
    const int lane = threadIdx.x & 31;
    const unsigned grp = (full >> (lane - c)) & ((1u << NCHILD) - 1u);
    if (!in) return;
    if (!splits) {
        if (c == 0) { t.child0[i] = -1; t.nchild[i] = 0; }
        return;
    }
    const int level = t.level[i];
    const int child0 = lvl_end + offs[i - lvl_begin];
    if (e > b) {
        const double q = ldexp(root[3], -(level - 1)) * 0.25;     // edge of this node / 4 (exact)
        const int j = child0 + __popc(grp & ((1u << c) - 1u));
        t.first[j] = b;
        t.count[j] = e - b;
        t.level[j] = level + 1;
        t.parent[j] = i;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            t.center[d][j] = __dadd_rn(t.center[d][i], ((c >> d) & 1) ? q : -q);   // src/bhtree.cpp:190-196
    }
    if (c == 0) { t.child0[i] = child0; t.nchild[i] = __popc(grp); }
}

// speculative build: close level l on the device.  lvl = &bounds[l]: {begin, end} -> writes the end of the next
// level; flags the build as bad if the next level does not fit the grid the host sized for it (w_next_cap)
// or the node pool.
__global__ void k_level_advance(int * __restrict__ lvl, const int * __restrict__ tmp, const int * __restrict__ offs,
                                int w_next_cap, int node_cap, int * __restrict__ bad)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int w = lvl[1] - lvl[0];
    const int total = (w > 0 && !*bad) ? offs[w - 1] + tmp[w - 1] : 0;
    if (total > w_next_cap || lvl[1] + total > node_cap) { *bad = 1; lvl[2] = lvl[1]; }
    else lvl[2] = lvl[1] + total;
}

__global__ void k_root_init(TreeBuild t, int n, const double * __restrict__ root)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        t.first[0] = 0; t.count[0] = n; t.level[0] = 1; t.parent[0] = -1;   // src/bhtree.cpp:16
        for (int d = 0; d < 3; ++d) if (t.center[d]) t.center[d][0] = root[d];
    }
}

// mass, sum m*pos; one level per launch, deepest level first
// (BHNode::assign accumulations, src/bhtree.cpp:199-201).  posm = packed {x, y, z, m} records, global tree-order index.
// mode 0: leaves and internal nodes of the level (single GPU);
// mode 1: leaves only, restricted to the particles [own_lo, own_hi) this rank owns (multi-GPU: partial sums, the ranks'
//         arrays are then all-reduced; run once over all nodes); mode 2: internal nodes of the level only.
template <int DIM>
__global__ void k_level_up(TreeBuild t, const double4 * __restrict__ posm, int lvl_begin, int lvl_end, int own_lo, int own_hi, int mode)
{
    const int i = lvl_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_end) return;
    double m = 0.0, mp[3] = {0.0, 0.0, 0.0};
    const int nc = t.nchild[i];
    if (nc == 0) {
        if (mode == 2) return;
        const int first = max(t.first[i], own_lo), last = min(t.first[i] + t.count[i], own_hi);
        for (int j = first; j < last; ++j) {
            const double4 pj = posm[j];
            m += pj.w;
            mp[0] += pj.x * pj.w;
            if (DIM >= 2) mp[1] += pj.y * pj.w;
            if (DIM >= 3) mp[2] += pj.z * pj.w;
        }
    } else {
        if (mode == 1) { t.msum4[i] = make_double4(0.0, 0.0, 0.0, 0.0); return; }
        const int c0 = t.child0[i];
        for (int k = 0; k < nc; ++k) {
            const double4 q = t.msum4[c0 + k];
            m += q.x; mp[0] += q.y; mp[1] += q.z; mp[2] += q.w;
        }
    }
    t.msum4[i] = make_double4(m, mp[0], mp[1], mp[2]);
}

template <int DIM>
__global__ void k_tree_scatter(TreeBuild t, TreeDev o, int n_nodes, const double * __restrict__ root)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int level = t.level[i];
    double c[3] = {0.0, 0.0, 0.0}, mc[3] = {0.0, 0.0, 0.0};
    double m = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) c[d] = t.center[d][i];
    if (i != 0) {
        // the reference never accumulates the root's mass / mass centre (root_clear,
        // include/bhtree.hpp:46-53): keep 0 so that the walk reproduces its behaviour.
        const double4 q = t.msum4[i];
        m = q.x;
        mc[0] = q.y / m;                                              // src/bhtree.cpp:152
        if (DIM >= 2) mc[1] = q.z / m;
        if (DIM >= 3) mc[2] = q.w / m;
    }
    const double edge = ldexp(root[3], -(level - 1));
    const int nc = t.nchild[i];
    const double ch = pack_ints(nc ? t.child0[i] : 0, nc);            // {child0, nchild}
    const double fc = pack_ints(t.first[i], t.count[i]);              // {first, count}
    double2 * nn = o.nn + (size_t)i * 4;
    nn[0] = make_double2(c[0], c[1]); nn[1] = make_double2(c[2], edge);
    nn[2] = make_double2(0.0, ch);    nn[3] = make_double2(fc, 0.0);
    double2 * ng = o.ng + (size_t)i * 4;
    ng[0] = make_double2(mc[0], mc[1]); ng[1] = make_double2(mc[2], m);
    ng[2] = make_double2(__dmul_rn(edge, edge), ch); ng[3] = make_double2(fc, 0.0);
    o.parent[i] = t.parent[i];
}

// ---- particle groups ------------------------------------------------------------------------------
// A group is the unit of work of every walk kernel: <= 32 consecutive particles of the tree order, one
// per lane.  Plain 32-particle slices of the sorted order would now and then straddle the boundary of
// two large cells (consecutive in the key order, far apart in space) and get a huge bounding box, so
// groups are cut at the boundaries of "group cells": the tree nodes with <= cell_max particles
// whose parent has more (and leaves above that size at the maximum level).  A group cell is chopped
// into slices of 32; every group therefore lies inside one cube holding <= cell_max particles.
// The neighbour walks want the cube small (their cost grows with the box + search radius), the gravity walk
// wants full warps more than a tight box: two group tables are built per tree.
#ifndef SPHB_GROUP_CELL_SPH
#define SPHB_GROUP_CELL_SPH 1024
#endif
#ifndef SPHB_GROUP_CELL_GRAV
#define SPHB_GROUP_CELL_GRAV 4096
#endif
constexpr int GROUP_CELL_SPH = SPHB_GROUP_CELL_SPH, GROUP_CELL_GRAV = SPHB_GROUP_CELL_GRAV;

// ctl[0] = next group (work counter), ctl[1] = end group, for the particle range [p_begin, p_end)
__global__ void k_group_range(const int * __restrict__ gstart, const int * __restrict__ n_groups, int p_begin, int p_end, int * __restrict__ ctl)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int ng = *n_groups;
    auto lower = [&](int v) { int lo = 0, hi = ng; while (lo < hi) { const int mid = (lo + hi) >> 1; if (gstart[mid] < v) lo = mid + 1; else hi = mid; } return lo; };
    ctl[0] = lower(p_begin);
    ctl[1] = lower(p_end);
}

struct GroupTable { const int * start; const int * n_groups; int * ctl; int n; };

// next group of this warp: particles [first, first + cnt); false when the range is exhausted
__device__ __forceinline__ bool next_group(const GroupTable & gt, int lane, int & first, int & cnt)
{
    int g = 0;
    if (lane == 0) g = atomicAdd(&gt.ctl[0], 1);
    g = __shfl_sync(SPHB_FULL_MASK, g, 0);
    if (g >= gt.ctl[1]) return false;
    first = gt.start[g];
    const int next = (g + 1 < *gt.n_groups) ? gt.start[g + 1] : gt.n;
    cnt = next - first;
    return true;
}

// BHNode::set_kernel (src/bhtree.cpp:206-232): per node the largest sml beneath it.  Three kernels: clear, leaf maxima
// carried up the parent chain by atomic max (over the particles [own_lo, own_hi) this rank owns; sml is indexed by the
// global tree-order index), and — after the all-reduce (max) of the multi-GPU mode — the copy into the walk records:
// nn[2].x = kernel_size (symmetric neighbour search, src/bhtree.cpp:237), ng[3].y = (largest h of a LEAF)^2 with the
// margin of hsoft.y (softening threshold of the gravity particle-particle pass).
__global__ void k_clear_kernel(TreeDev t)
{
    const int D = blockIdx.x * blockDim.x + threadIdx.x;
    if (D < t.n_nodes) t.ksize[D] = 0.0;
}
__global__ void k_set_kernel(TreeDev t, const double * __restrict__ sml, int own_lo, int own_hi)
{
    const int D = blockIdx.x * blockDim.x + threadIdx.x;
    if (D >= t.n_nodes) return;
    const double2 q2 = t.nn[(size_t)D * 4 + 2], q3 = t.nn[(size_t)D * 4 + 3];
    if (__double2hiint(q2.y)) return;                  // internal node
    const int first = max(__double2loint(q3.x), own_lo), last = min(__double2loint(q3.x) + __double2hiint(q3.x), own_hi);
    if (first >= last) return;
    double h = 0.0;
    for (int j = first; j < last; ++j) { const double s = sml[j]; if (s > h) h = s; }
    const unsigned long long hb = (unsigned long long)__double_as_longlong(h);
    int q = D;
    while (q >= 0) {
        const unsigned long long old = atomic_max_pos(&t.ksize[q], h);
        if (old >= hb) break;          // whoever wrote `old` carries it (or more) upward
        q = t.parent[q];
    }
}
__global__ void k_apply_ksize(TreeDev t)
{
    const int D = blockIdx.x * blockDim.x + threadIdx.x;
    if (D >= t.n_nodes) return;
    const double h = t.ksize[D];
    double2 * nn = t.nn + (size_t)D * 4;
    nn[2].x = h;
    if (__double2hiint(nn[2].y) == 0) t.ng[(size_t)D * 4 + 3].y = h * h * (1.0 + 1e-12);
}

// r_ij = r_i - {x,y,z of a staged particle}, minimum image if periodic
template <int DIM>
__device__ __forceinline__ void rij_from4(const DevParams & P, const double (&ri)[DIM], const double4 & pj, double (&d)[DIM])
{
    d[0] = ri[0] - pj.x;
    if (DIM >= 2) d[DIM >= 2 ? 1 : 0] = ri[DIM >= 2 ? 1 : 0] - pj.y;
    if (DIM >= 3) d[DIM >= 3 ? 2 : 0] = ri[DIM >= 3 ? 2 : 0] - pj.z;
    if (P.periodic) {
#pragma unroll
        for (int a = 0; a < DIM; ++a) d[a] = min_image(d[a], P.range[a]);
    }
}

} // namespace sphb
