// sphb_dist.cuh — device side of the multi-GPU mode: Morton domain decomposition of the particle set over the
// ranks of one NVLink / NVSwitch box, one process per GPU.  The reference is a single process (nothing to cite but
// the loops whose reads define the halo: src/pre_interaction.cpp:83-134, src/fluid_force.cpp:62,
// src/bhtree.cpp:301-331).
//
// Ownership: the global tree order (sorted octree keys) is cut into `world` contiguous ranges; rank r owns the
// particles [off_r, off_r + n_r) and holds the full state (PSoA) of those only.  What the walks read of OTHER
// particles are the packed gather records (posm, velc, thermo, av, hsoft), which every rank keeps in arrays indexed by
// the GLOBAL tree-order index — its own range is written by its own kernels, the rest stays unpopulated except for
// the halo.  All record arrays of a rank live in one cudaMalloc slab that is exported with CUDA IPC; a rank fills
// its halo by READING the owners' slabs over NVLink from a kernel (k_pull_halo) — no send/recv pairing, no message
// sizes on the host, one launch per halo.  Which remote leaves are needed is decided against the replicated tree
// topology by k_mark_halo (one warp per group cell of the own range).
#pragma once
#include "sphb_gravity2.cuh"

namespace sphb {

constexpr int MAX_WORLD = 16;

// key bit that marks a particle leaving this rank: sorts behind every real key
__host__ __device__ inline unsigned long long leaver_bit(int key_bits) { return 1ull << key_bits; }

struct PeerTab {
    const char * slab[MAX_WORLD];   // base of every rank's record slab (own entry = own slab), IPC-mapped
    int off[MAX_WORLD + 1];         // global tree-order index of every rank's first particle; off[world] = N
    int world, rank;
};
// layout of a record slab: the five record arrays, each n_pad entries long, then the rank's sorted keys
struct SlabLayout {
    size_t posm, velc, thermo, av, hsoft, keys, mig;     // byte offsets
};
// migration pull: the block of records rank s packed for this rank sits at src_off[s] (records) of s's mig region
struct MigPull { int src_off[MAX_WORLD]; int dst_off[MAX_WORLD + 1]; };

__device__ __forceinline__ int owner_of(const PeerTab & pt, int g)
{
    int r = 0;
#pragma unroll 1
    for (int k = 1; k < pt.world; ++k) r += (g >= pt.off[k]) ? 1 : 0;
    return r;
}

// 32- / 16-byte record loads from a peer's slab that bypass stale cache lines (the owner rewrites them every step)
__device__ __forceinline__ double4 peer_ld4(const double4 * p)
{
    const double2 a = __ldcv(reinterpret_cast<const double2 *>(p)), b = __ldcv(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// ---- bounding box across ranks: part -> {-lo, hi} (all-reduce max) -> root --------------------------------------
template <int DIM>
__global__ void k_bbox_neg(const double * __restrict__ part, int nblocks, double * __restrict__ out /* [2*DIM]: -lo, hi */)
{
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
#pragma unroll
        for (int d = 0; d < DIM; ++d) { out[d] = -lo[d]; out[DIM + d] = hi[d]; }
    }
}
template <int DIM>
__global__ void k_root_from_bbox(const double * __restrict__ nb, double * __restrict__ root)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double l = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        const double lo = -nb[d], hi = nb[DIM + d];
        root[d] = (hi + lo) * 0.5;                     // src/bhtree.cpp:85
        const double range = hi - lo;
        if (l < range) l = range;                      // src/bhtree.cpp:87-93
    }
    for (int d = DIM; d < 3; ++d) root[d] = 0.0;
    root[3] = l;
}

// ---- migration ------------------------------------------------------------------------------------------------------
// split[0] = 0 <= split[1] <= ... <= split[world - 1]; rank d owns the keys in [split[d], split[d + 1])
__global__ void k_mig_mark(unsigned long long * __restrict__ keys, int n, const unsigned long long * __restrict__ split, int world, int rank,
                           int key_bits, int * __restrict__ cnt /* [world] + total at [world] */, int * __restrict__ list_idx, int * __restrict__ list_dest)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    int d = 0;
    for (int q = 1; q < world; ++q) d += (k >= split[q]) ? 1 : 0;
    if (d != rank) {
        const int t = atomicAdd(&cnt[world], 1);
        list_idx[t] = i;
        list_dest[t] = d;
        atomicAdd(&cnt[d], 1);
        keys[i] = k | leaver_bit(key_bits);
    }
}

// one migrating particle = MIG_REC(DIM) doubles: the 4 DIM + 12 state members, {pid, neighbor} and {orig, -} as bit patterns
__host__ __device__ constexpr int mig_rec(int dim) { return 4 * dim + 14; }

template <int DIM>
__global__ void k_mig_pack(PSoA s, const int * __restrict__ list_idx, const int * __restrict__ list_dest, int total,
                           const int * __restrict__ soff, int * __restrict__ cursor, double * __restrict__ buf)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int i = list_idx[t], d = list_dest[t];
    double * r = buf + (size_t)(soff[d] + atomicAdd(&cursor[d], 1)) * mig_rec(DIM);
#pragma unroll
    for (int a = 0; a < DIM; ++a) { r[a] = s.pos[a][i]; r[DIM + a] = s.vel[a][i]; r[2 * DIM + a] = s.vel_p[a][i]; r[3 * DIM + a] = s.acc[a][i]; }
    double * q = r + 4 * DIM;
    q[0] = s.mass[i]; q[1] = s.dens[i]; q[2] = s.pres[i]; q[3] = s.ene[i]; q[4] = s.ene_p[i]; q[5] = s.dene[i];
    q[6] = s.sml[i]; q[7] = s.sound[i]; q[8] = s.balsara[i]; q[9] = s.alpha[i]; q[10] = s.gradh[i]; q[11] = s.phi[i];
    q[12] = pack_ints(s.pid[i], s.neighbor[i]);
    q[13] = pack_ints(s.orig[i], 0);
}
// arrivals: read the blocks the other ranks packed for this rank out of their slabs (NVLink peer reads)
__global__ void k_mig_pull(PeerTab pt, size_t mig_off, MigPull mp, int rec, double * __restrict__ recv)
{
    const long long total = (long long)mp.dst_off[pt.world] * rec;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(e / rec);
        int s = 0;
#pragma unroll 1
        for (int k = 1; k < pt.world; ++k) s += (t >= mp.dst_off[k]) ? 1 : 0;
        const double * src = reinterpret_cast<const double *>(pt.slab[s] + mig_off) + (size_t)mp.src_off[s] * rec;
        recv[e] = __ldcv(src + (e - (long long)mp.dst_off[s] * rec));
    }
}

// arrival t goes into the slot of this rank's t-th leaver, the rest behind the own particles (index n_own, n_own + 1, ...)
template <int DIM>
__global__ void k_mig_unpack(PSoA s, const double * __restrict__ buf, int count, const int * __restrict__ list_idx, int n_leave, int n_own,
                             const double * __restrict__ root, int key_levels, unsigned long long * __restrict__ keys, int * __restrict__ idx)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int i = t < n_leave ? list_idx[t] : n_own + (t - n_leave);
    const double * r = buf + (size_t)t * mig_rec(DIM);
#pragma unroll
    for (int a = 0; a < DIM; ++a) { s.pos[a][i] = r[a]; s.vel[a][i] = r[DIM + a]; s.vel_p[a][i] = r[2 * DIM + a]; s.acc[a][i] = r[3 * DIM + a]; }
    const double * q = r + 4 * DIM;
    s.mass[i] = q[0]; s.dens[i] = q[1]; s.pres[i] = q[2]; s.ene[i] = q[3]; s.ene_p[i] = q[4]; s.dene[i] = q[5];
    s.sml[i] = q[6]; s.sound[i] = q[7]; s.balsara[i] = q[8]; s.alpha[i] = q[9]; s.gradh[i] = q[10]; s.phi[i] = q[11];
    s.pid[i] = __double2loint(q[12]); s.neighbor[i] = __double2hiint(q[12]);
    s.orig[i] = __double2loint(q[13]);
    keys[i] = key_of<DIM>(s, i, root, key_levels);
    idx[i] = i;
}

// splitters of the NEXT build: the key at the global position q N / world of the (replicated) sorted keys — every rank
// computes the same values, no communication
__global__ void k_next_splitters(const unsigned long long * __restrict__ keys_global, long long n, int world, unsigned long long * __restrict__ split)
{
    const int q = threadIdx.x;
    if (q == 0) split[0] = 0;
    else if (q < world) split[q] = keys_global[(n * q) / world];
}

// ---- replicated topology from the ranks' sorted keys: pull every rank's run over NVLink ----------------------------------
__global__ void k_gather_keys(PeerTab pt, size_t keys_off, unsigned long long * __restrict__ keys_global, int n)
{
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        const int r = owner_of(pt, g);
        const unsigned long long * src = reinterpret_cast<const unsigned long long *>(pt.slab[r] + keys_off);
        keys_global[g] = __ldcv(src + (g - pt.off[r]));      // peer memory: never from a stale cache line
    }
}

// ---- group tables of the own range -----------------------------------------------------------------------------------------
// As k_group_flags, for the particles [own_lo, own_hi) only: flags are indexed by (global index - own_lo); the first own
// particle always starts a group.  Also lists the group cells that overlap the own range (k_mark_halo's work units).
__global__ void k_group_flags_own(TreeBuild t, int n_nodes, unsigned char * __restrict__ flags, int own_lo, int own_hi, int cell_max,
                                  int * __restrict__ cells, int * __restrict__ n_cells, int chunk /* particles per group: 32, or 64 for k_gravity2 */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int cnt = t.count[i];
    const bool small = cnt <= cell_max;
    const bool parent_big = (i == 0) || t.count[t.parent[i]] > cell_max;
    if (!((small && parent_big) || (!small && t.nchild[i] == 0))) return;
    const int first = t.first[i];
    if (first + cnt <= own_lo || first >= own_hi) return;
    if (cells) cells[atomicAdd(n_cells, 1)] = i;
    for (int k = 0; k < cnt; k += chunk) {
        const int g = first + k;
        if (g >= own_lo && g < own_hi) flags[g - own_lo] = 1;
    }
    if (first < own_lo) flags[0] = 1;
}

// ---- halo marking -----------------------------------------------------------------------------------------------------------
// One warp per group cell C of the own range (a tree node with <= cell_max particles; its cube bounds every own
// particle in it).  Breadth-first descent of the replicated tree with one lane per node, as in the walks:
//   bit 1 (SPH)     node cube within `reach` of C's cube, reach = the largest search radius of C's own particles
//                   (gather pass: h_guess * factor, src/pre_interaction.cpp:61-64; symmetric pass: max(kernel_size(C),
//                   kernel_size(node)), src/bhtree.cpp:237) — a superset of what any group inside C can hit;
//   bit 2 (gravity) node not accepted by the opening criterion from the nearest point of C's cube (src/bhtree.cpp:308):
//                   every ancestor of a leaf some own particle opens is opened from C's cube as well.
// Leaves that reach beyond the own range get their bits set in `flags` (one byte per node).
constexpr int MK_STACK = 1024;
struct MarkSmem { int2 stack[MK_STACK]; int2 expand[32]; };

template <int DIM>
__global__ void __launch_bounds__(128)
k_mark_halo(TreeDev t, DevParams P, const int * __restrict__ cells, const int * __restrict__ n_cells, int symmetric, int bits0,
            double factor, const double * __restrict__ mass_g, const double * __restrict__ dens_g /* global-view arrays */,
            int own_lo, int own_hi, unsigned char * __restrict__ flags, unsigned long long * __restrict__ d_err)
{
    __shared__ MarkSmem s_mk[4];
    MarkSmem & sm = s_mk[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int nc = *n_cells;
    for (int ci = blockIdx.x * 4 + (threadIdx.x >> 5); ci < nc; ci += gridDim.x * 4) {
        const int cell = cells[ci];
        const double2 * cq = t.nn + (size_t)cell * 4;
        const double2 c0 = cq[0], c1 = cq[1], c2 = cq[2], c3 = cq[3];
        double bc[DIM], bh;
        bc[0] = c0.x;
        if (DIM >= 2) bc[DIM >= 2 ? 1 : 0] = c0.y;
        if (DIM >= 3) bc[DIM >= 3 ? 2 : 0] = c1.x;
        bh = 0.5 * c1.y;
        double reach;
        if (symmetric) reach = c2.x;
        else {
            const int first = max(__double2loint(c3.x), own_lo), last = min(__double2loint(c3.x) + __double2hiint(c3.x), own_hi);
            double hmax = 0.0;
            for (int j = first + lane; j < last; j += 32) hmax = fmax(hmax, h_guess<DIM>(P.ngb, mass_g[j], dens_g[j]));
            reach = warp_max(hmax) * factor * (1.0 + 1e-12);
        }
        double cmax = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) cmax = fmax(cmax, fabs(bc[d]) + bh);
        const double slack = 1e-12 * (cmax + reach) + 1e-300;
        bh += slack;                                                  // rounding of the node centres

        int top = 1;
        if (lane == 0) sm.stack[0] = make_int2(0, bits0);
        __syncwarp();
        while (top > 0) {
            const int ne = min(top, 32);
            int2 ent = make_int2(0, 0);
            int nch = 0;
            if (lane < ne) { ent = sm.stack[top - 1 - lane]; nch = (int)((unsigned)ent.x >> 29) + 1; }
            int incl = nch;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(SPHB_FULL_MASK, incl, o);
                if (lane >= o) incl += y;
            }
            const int m = __popc(__ballot_sync(SPHB_FULL_MASK, lane < ne && incl <= 32));
            const int k = __shfl_sync(SPHB_FULL_MASK, incl, m - 1);
            if (lane < m) {
                const int ch0 = ent.x & 0x1fffffff;
                for (int q = 0; q < nch; ++q) sm.expand[incl - nch + q] = make_int2(ch0 + q, ent.y);
            }
            __syncwarp();
            int node = -1, bits = 0;
            if (lane < k) { node = sm.expand[lane].x; bits = sm.expand[lane].y; }
            top -= m;
            __syncwarp();
            int nb = 0, child0 = 0, nchild = 0;
            if (node >= 0) {
                const double2 * q = t.nn + (size_t)node * 4;
                const double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
                child0 = __double2loint(q2.y); nchild = __double2hiint(q2.y);
                if (bits & 1) {
                    const double half = 0.5 * q1.y;
                    const double rn = symmetric ? fmax(reach, q2.x) : reach;
                    double cc[DIM];
                    cc[0] = q0.x;
                    if (DIM >= 2) cc[DIM >= 2 ? 1 : 0] = q0.y;
                    if (DIM >= 3) cc[DIM >= 3 ? 2 : 0] = q1.x;
                    double g2 = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) {
                        double dc = bc[d] - cc[d];
                        if (P.periodic) dc = min_image(dc, P.range[d]);
                        const double gap = fabs(dc) - (bh + half) - slack;
                        if (gap > 0.0) g2 += gap * gap;
                    }
                    if (g2 <= rn * rn * (1.0 + 1e-9)) nb |= 1;
                }
                if (bits & 2) {
                    const double2 * g = t.ng + (size_t)node * 4;
                    const double2 g0 = __ldg(g), g1 = __ldg(g + 1), g2q = __ldg(g + 2);
                    double mc[DIM];
                    mc[0] = g0.x;
                    if (DIM >= 2) mc[DIM >= 2 ? 1 : 0] = g0.y;
                    if (DIM >= 3) mc[DIM >= 3 ? 2 : 0] = g1.x;
                    double dmin2 = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) {
                        double dc = bc[d] - mc[d];
                        if (P.periodic) dc = min_image(dc, P.range[d]);
                        const double lo = fmax(fabs(dc) - bh - slack, 0.0);
                        dmin2 += lo * lo;
                    }
                    if (node == 0 || !(g2q.x <= P.theta2 * dmin2 * (1.0 - 1e-9))) nb |= 2;
                }
                if (nb && nchild == 0) {
                    const double2 q3 = __ldg(q + 3);
                    const int first = __double2loint(q3.x), count = __double2hiint(q3.x);
                    if ((first < own_lo || first + count > own_hi) && (flags[node] & nb) != nb)
                        atomicOr(reinterpret_cast<unsigned *>(flags + (node & ~3)), (unsigned)nb << (8 * (node & 3)));
                    nb = 0;
                }
            }
            const unsigned pb = __ballot_sync(SPHB_FULL_MASK, nb != 0);
            if (pb) {
                const int total = __popc(pb);
                if (top + total > MK_STACK) {
                    if (lane == 0) atomicOr(&d_err[2], (unsigned long long)WALK_ERR_STACK);
                } else {
                    if (nb) sm.stack[top + __popc(pb & lt_mask)] = make_int2(child0 | ((nchild - 1) << 29), nb);
                    top += total;
                }
            }
            __syncwarp();
        }
    }
}

// ---- halo pull ---------------------------------------------------------------------------------------------------------------
// record sets: 1 posm, 2 velc, 4 thermo (before PreInteraction: u), 8 thermo + av (after it), 16 hsoft
enum { PULL_POSM = 1, PULL_VELC = 2, PULL_THERMO_A = 4, PULL_THERMO_B = 8, PULL_HSOFT = 16 };

// One warp per 32 nodes: a flagged leaf's particles that other ranks own are copied from the owners' slabs into the
// same (global) index of this rank's record arrays.  need_sph / need_grav: record sets wanted for leaves flagged with
// bit 1 / bit 2; `have` remembers what was fetched since the tree was built (positions and velocities do not change in
// between; the thermo sets are fetched whenever asked).
__global__ void __launch_bounds__(128)
k_pull_halo(TreeDev t, PeerTab pt, SlabLayout L, char * __restrict__ my_slab, const unsigned char * __restrict__ flags,
            unsigned char * __restrict__ have, int need_sph, int need_grav, unsigned long long * __restrict__ pulled)
{
    const int lane = threadIdx.x & 31;
    const int base = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 32;
    if (base >= t.n_nodes) return;
    const int node = base + lane;
    int need = 0, first = 0, count = 0;
    if (node < t.n_nodes) {
        const int f = flags[node];
        if (f) {
            need = ((f & 1) ? need_sph : 0) | ((f & 2) ? need_grav : 0);
            const int hv = have[node];
            need &= ~(hv & (PULL_POSM | PULL_VELC | PULL_HSOFT));
            if (need) {
                have[node] = (unsigned char)(hv | need);
                const double2 q3 = t.nn[(size_t)node * 4 + 3];
                first = __double2loint(q3.x); count = __double2hiint(q3.x);
            }
        }
    }
    unsigned todo = __ballot_sync(SPHB_FULL_MASK, need != 0);
    const int own_lo = pt.off[pt.rank], own_hi = pt.off[pt.rank + 1];
    unsigned long long moved = 0;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int nd = __shfl_sync(SPHB_FULL_MASK, need, src);
        const int f0 = __shfl_sync(SPHB_FULL_MASK, first, src);
        const int c0 = __shfl_sync(SPHB_FULL_MASK, count, src);
        for (int j = f0 + lane; j < f0 + c0; j += 32) {
            if (j >= own_lo && j < own_hi) continue;
            const char * ps = pt.slab[owner_of(pt, j)];
            if (nd & PULL_POSM) reinterpret_cast<double4 *>(my_slab + L.posm)[j] = peer_ld4(reinterpret_cast<const double4 *>(ps + L.posm) + j);
            if (nd & PULL_VELC) reinterpret_cast<double4 *>(my_slab + L.velc)[j] = peer_ld4(reinterpret_cast<const double4 *>(ps + L.velc) + j);
            if (nd & (PULL_THERMO_A | PULL_THERMO_B)) reinterpret_cast<double4 *>(my_slab + L.thermo)[j] = peer_ld4(reinterpret_cast<const double4 *>(ps + L.thermo) + j);
            if (nd & PULL_THERMO_B) reinterpret_cast<double4 *>(my_slab + L.av)[j] = peer_ld4(reinterpret_cast<const double4 *>(ps + L.av) + j);
            if (nd & PULL_HSOFT) reinterpret_cast<double2 *>(my_slab + L.hsoft)[j] = __ldcv(reinterpret_cast<const double2 *>(ps + L.hsoft) + j);
            ++moved;
        }
    }
    if (pulled) {
        moved = warp_sum_u64(moved);
        if (lane == 0 && moved) atomicAdd(pulled, moved);
    }
}

// download in the multi-GPU mode renumbers: the caller's record k is the rank's k-th particle in tree order
__global__ void k_renumber_orig(int * __restrict__ orig, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) orig[i] = i;
}

} // namespace sphb
