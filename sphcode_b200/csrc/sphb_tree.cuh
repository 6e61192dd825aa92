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
//   4. mass / mass centre / subtree size go bottom-up, and the nodes are laid out in DFS
//      pre-order with a skip count per node, which makes every tree walk stackless.
// Walks are warp-synchronous: a warp of 32 Morton-consecutive particles traverses the union of
// its lanes' reference walks; each lane applies the reference's per-particle criterion itself
// (so the set of nodes it opens / accepts is exactly the reference's) and sits out the subtrees
// it did not open.  Node and leaf-particle loads are warp-uniform (one broadcast transaction).
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
    int *first, *count, *level, *parent, *child0, *nchild, *size, *dfs;
    double *center[3];
    double *msum, *mpos[3];
};

// Finished tree, DFS pre-order, one packed 64-byte record (4 x double2) per node and walk kind so
// that a node visit is four 16-byte warp-uniform loads off one address:
//   nn (neighbour walks): [0] cx, cy   [1] cz, edge   [2] ksize, {skip, first}   [3] {count, leaf}, -
//   ng (gravity walk):    [0] mx, my   [1] mz, mass   [2] edge^2, {skip, first}  [3] {count, leaf}, -
// (c = geometric centre, m = mass centre, ksize = BHNode::kernel_size set by set_kernel,
//  skip = nodes in the subtree incl. this one, first/count = particle range, leaf = is_leaf)
struct TreeDev {
    int      n_nodes;
    double2 *nn;
    double2 *ng;
    int     *parent;    // DFS index of the parent, -1 for the root
};
struct NodeRec { double x, y, z, w, e; int skip, first, count, leaf; };
__device__ __forceinline__ NodeRec load_node(const double2 * __restrict__ base, int idx)
{
    const double2 * q = base + (size_t)idx * 4;
    const double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
    NodeRec r;
    r.x = q0.x; r.y = q0.y; r.z = q1.x; r.w = q1.y; r.e = q2.x;
    r.skip = __double2loint(q2.y); r.first = __double2hiint(q2.y);
    r.count = __double2loint(q3.x); r.leaf = __double2hiint(q3.x);
    return r;
}
__device__ __forceinline__ double * node_ksize(const TreeDev & t, int idx) { return &t.nn[(size_t)idx * 4 + 2].x; }

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
__global__ void k_keys(PSoA p, int n, const double * __restrict__ root, int key_levels,
                       unsigned long long * __restrict__ keys, int * __restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
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
    keys[i] = key;
    idx[i]  = i;
}

// Gather-permute of the particle arrays by the sorted index.
__global__ void k_permute(const double * const * __restrict__ src, double * const * __restrict__ dst, int narr,
                          const int * const * __restrict__ isrc, int * const * __restrict__ idst, int niarr,
                          const int * __restrict__ perm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = perm[i];
    for (int a = 0; a < narr; ++a) dst[a][i] = src[a][s];
    for (int a = 0; a < niarr; ++a) idst[a][i] = isrc[a][s];
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

template <int DIM>
__global__ void k_level_count(TreeBuild t, const unsigned long long * __restrict__ keys, int lvl_begin, int lvl_end,
                              int leaf_num, int max_level_eff, int key_levels, int * __restrict__ tmp)
{
    const int i = lvl_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_end) return;
    int nchild = 0;
    if (node_splits<DIM>(t, i, leaf_num, max_level_eff)) {
        constexpr int NCHILD = 1 << DIM;
        const int first = t.first[i], last = first + t.count[i];
        const int shift = (key_levels - t.level[i]) * DIM;
        int b = first;
        for (int c = 0; c < NCHILD; ++c) {
            const int e = (c == NCHILD - 1) ? last : lower_bound_child(keys, b, last, shift, NCHILD - 1, c + 1);
            if (e > b) ++nchild;
            b = e;
        }
    }
    tmp[i - lvl_begin] = nchild;
}

template <int DIM>
__global__ void k_level_emit(TreeBuild t, const unsigned long long * __restrict__ keys, int lvl_begin, int lvl_end,
                             int leaf_num, int max_level_eff, int key_levels, const int * __restrict__ offs,
                             const double * __restrict__ root)
{
    const int i = lvl_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_end) return;
    if (!node_splits<DIM>(t, i, leaf_num, max_level_eff)) {
        t.child0[i] = -1;
        t.nchild[i] = 0;
        return;
    }
    constexpr int NCHILD = 1 << DIM;
    const int first = t.first[i], last = first + t.count[i], level = t.level[i];
    const int shift = (key_levels - level) * DIM;
    const int child0 = lvl_end + offs[i - lvl_begin];
    const double q = ldexp(root[3], -(level - 1)) * 0.25;     // edge of this node / 4 (exact)
    int b = first, k = 0;
    for (int c = 0; c < NCHILD; ++c) {
        const int e = (c == NCHILD - 1) ? last : lower_bound_child(keys, b, last, shift, NCHILD - 1, c + 1);
        if (e > b) {
            const int j = child0 + k;
            t.first[j] = b;
            t.count[j] = e - b;
            t.level[j] = level + 1;
            t.parent[j] = i;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
                t.center[d][j] = __dadd_rn(t.center[d][i], ((c >> d) & 1) ? q : -q);   // src/bhtree.cpp:190-196
            ++k;
        }
        b = e;
    }
    t.child0[i] = child0;
    t.nchild[i] = k;
}

__global__ void k_root_init(TreeBuild t, int n, const double * __restrict__ root)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        t.first[0] = 0; t.count[0] = n; t.level[0] = 1; t.parent[0] = -1;   // src/bhtree.cpp:16
        for (int d = 0; d < 3; ++d) if (t.center[d]) t.center[d][0] = root[d];
    }
}

// mass, sum m*pos, subtree size; one level per launch, deepest level first
// (BHNode::assign accumulations, src/bhtree.cpp:199-201)
template <int DIM>
__global__ void k_level_up(TreeBuild t, PSoA p, int lvl_begin, int lvl_end)
{
    const int i = lvl_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_end) return;
    double m = 0.0, mp[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) mp[d] = 0.0;
    int size = 1;
    const int nc = t.nchild[i];
    if (nc == 0) {
        const int first = t.first[i], last = first + t.count[i];
        for (int j = first; j < last; ++j) {
            const double mj = p.mass[j];
            m += mj;
#pragma unroll
            for (int d = 0; d < DIM; ++d) mp[d] += p.pos[d][j] * mj;
        }
    } else {
        const int c0 = t.child0[i];
        for (int k = 0; k < nc; ++k) {
            m += t.msum[c0 + k];
#pragma unroll
            for (int d = 0; d < DIM; ++d) mp[d] += t.mpos[d][c0 + k];
            size += t.size[c0 + k];
        }
    }
    t.msum[i] = m;
#pragma unroll
    for (int d = 0; d < DIM; ++d) t.mpos[d][i] = mp[d];
    t.size[i] = size;
}

// DFS pre-order index of the children of every node of one level (top-down).
__global__ void k_level_dfs(TreeBuild t, int lvl_begin, int lvl_end)
{
    const int i = lvl_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_end) return;
    if (i == 0) t.dfs[0] = 0;
    const int nc = t.nchild[i];
    if (nc == 0) return;
    const int c0 = t.child0[i];
    int d = t.dfs[i] + 1;
    for (int k = 0; k < nc; ++k) { t.dfs[c0 + k] = d; d += t.size[c0 + k]; }
}

template <int DIM>
__global__ void k_tree_scatter(TreeBuild t, TreeDev o, int n_nodes, const double * __restrict__ root)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int D = t.dfs[i];
    const int level = t.level[i];
    double c[3] = {0.0, 0.0, 0.0}, mc[3] = {0.0, 0.0, 0.0};
    double m = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) c[d] = t.center[d][i];
    if (i != 0) {
        // the reference never accumulates the root's mass / mass centre (root_clear,
        // include/bhtree.hpp:46-53): keep 0 so that the walk reproduces its behaviour.
        m = t.msum[i];
#pragma unroll
        for (int d = 0; d < DIM; ++d) mc[d] = t.mpos[d][i] / m;       // src/bhtree.cpp:152
    }
    const double edge = ldexp(root[3], -(level - 1));
    const double sf = __hiloint2double(t.first[i], t.size[i]);                 // {skip, first}
    const double cl = __hiloint2double(t.nchild[i] == 0 ? 1 : 0, t.count[i]);  // {count, leaf}
    double2 * nn = o.nn + (size_t)D * 4;
    nn[0] = make_double2(c[0], c[1]); nn[1] = make_double2(c[2], edge);
    nn[2] = make_double2(0.0, sf);    nn[3] = make_double2(cl, 0.0);
    double2 * ng = o.ng + (size_t)D * 4;
    ng[0] = make_double2(mc[0], mc[1]); ng[1] = make_double2(mc[2], m);
    ng[2] = make_double2(__dmul_rn(edge, edge), sf); ng[3] = make_double2(cl, 0.0);
    o.parent[D] = (i == 0) ? -1 : t.dfs[t.parent[i]];
}

// packed {x, y, z, m} gather records of the particles (tree order), rebuilt by every make_tree
template <int DIM>
__global__ void k_pack_posm(PSoA p, double4 * __restrict__ posm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = p.pos[0][i];
    const double y = DIM >= 2 ? p.pos[DIM >= 2 ? 1 : 0][i] : 0.0;
    const double z = DIM >= 3 ? p.pos[DIM >= 3 ? 2 : 0][i] : 0.0;
    posm[i] = make_double4(x, y, z, p.mass[i]);
}

// BHNode::set_kernel (src/bhtree.cpp:206-232): per node the largest sml beneath it.
__global__ void k_clear_kernel(TreeDev t)
{
    const int D = blockIdx.x * blockDim.x + threadIdx.x;
    if (D < t.n_nodes) *node_ksize(t, D) = 0.0;
}
__global__ void k_set_kernel(TreeDev t, const double * __restrict__ sml)
{
    const int D = blockIdx.x * blockDim.x + threadIdx.x;
    if (D >= t.n_nodes) return;
    const double2 q2 = t.nn[(size_t)D * 4 + 2], q3 = t.nn[(size_t)D * 4 + 3];
    if (!__double2hiint(q3.x)) return;
    const int first = __double2hiint(q2.y), count = __double2loint(q3.x);
    double h = 0.0;
    for (int j = first; j < first + count; ++j) { const double s = sml[j]; if (s > h) h = s; }
    const unsigned long long hb = (unsigned long long)__double_as_longlong(h);
    int q = D;
    while (q >= 0) {
        const unsigned long long old = atomic_max_pos(node_ksize(t, q), h);
        if (old >= hb) break;          // whoever wrote `old` carries it (or more) upward
        q = t.parent[q];
    }
}

// ---- the stackless warp walk ----------------------------------------------------------------------
// A warp of 32 consecutive particles visits the union of its lanes' reference walks.
// V provides   bool open(const NodeRec &)                          the lane's own criterion
//              void leaf(const NodeRec &, int base, int m, const double4 * s)
//                   the lane opened this leaf; s[0..m) = {x,y,z,m} of particles base .. base+m-1,
//                   staged in shared memory by one coalesced load of the whole warp.
// PREFETCH: fetch both possible successors before the (dependent) test of the current node.  It takes
// the uniform-load latency off the critical path at the price of ~30 registers; measured on B200 it
// pays for the pre-interaction walks and costs occupancy in the force / gravity walks.
template <bool PREFETCH, class V>
__device__ __forceinline__ void warp_walk(const TreeDev & t, const double4 * __restrict__ posm, double4 * s_leaf,
                                          int lane, V & v, bool lane_valid)
{
    int idx = 0;
    int resume = lane_valid ? 0 : INT_MAX;       // lane takes part iff idx >= resume
    const int n_nodes = t.n_nodes;
    NodeRec nd;
    if (PREFETCH) nd = load_node(t.nn, 0);
    while (idx < n_nodes) {
        NodeRec n_down, n_skip;
        if (!PREFETCH) nd = load_node(t.nn, idx);
        if (PREFETCH) {
            n_down = load_node(t.nn, min(idx + 1, n_nodes - 1));
            n_skip = load_node(t.nn, min(idx + nd.skip, n_nodes - 1));
        }
        bool open = false;
        if (idx >= resume) {
            open = v.open(nd);
            if (!open) resume = idx + nd.skip;
        }
        if (__any_sync(SPHB_FULL_MASK, open)) {
            if (nd.leaf) {
                const int last = nd.first + nd.count;
                for (int base = nd.first; base < last; base += 32) {
                    const int m = min(32, last - base);
                    __syncwarp();
                    if (lane < m) s_leaf[lane] = ldg4(&posm[base + lane]);
                    __syncwarp();
                    if (open) v.leaf(nd, base, m, s_leaf);
                }
            }
            idx += 1;
            if (PREFETCH) nd = n_down;
        } else {
            idx += nd.skip;
            if (PREFETCH) nd = n_skip;
        }
    }
    __syncwarp();
}

// Per-lane neighbour criterion of BHNode::neighbor_search (src/bhtree.cpp:236-249):
// Chebyshev minimum-image distance to the geometric centre <= edge/2 + h.
template <int DIM>
__device__ __forceinline__ bool node_in_reach(const DevParams & P, const NodeRec & g, const double (&ri)[DIM], double h)
{
    const double l2 = (g.w * 0.5 + h) * (g.w * 0.5 + h);
    double c[DIM];
    c[0] = g.x;
    if (DIM >= 2) c[DIM >= 2 ? 1 : 0] = g.y;
    if (DIM >= 3) c[DIM >= 3 ? 2 : 0] = g.z;
    double d[DIM];
    calc_r_ij<DIM>(P, ri, c, d);
    double dx2_max = d[0] * d[0];
#pragma unroll
    for (int k = 1; k < DIM; ++k) {
        const double dx2 = d[k] * d[k];
        if (dx2 > dx2_max) dx2_max = dx2;
    }
    return dx2_max <= l2;
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
