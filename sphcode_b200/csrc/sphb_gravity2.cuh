// sphb_gravity2.cuh — GravityForce::calculation -> BHNode::calc_force (src/gravity_force.cpp:52-89, src/bhtree.cpp:301-331)
// with TWO particles per lane: a warp walks the tree for a group of up to 64 consecutive particles of the tree order.
//
// Same algorithm and the same per-particle decisions as k_gravity (sphb_stages.cuh): every stack entry carries the set of
// particles that opened all its ancestors (now two 32-bit words: particle s * 32 + lane is bit `lane` of word s), a
// batch of <= 32 nodes is classified against the group's bounding box, mixed nodes are decided particle by particle
// with the reference's own expression, and the interactions are deferred into the three lists (cells accepted by the whole
// group / by some particles / opened leaves).  What changes is the cost per particle:
//   * the walk (classification, stack, mask transposes) is paid once per 64 particles instead of once per 32;
//   * every record a lane loads — a leaf particle from global memory, a cell from shared memory — interacts with BOTH of
//     the lane's particles (they are Morton neighbours and mostly want the same leaves and cells; where only one does,
//     the other's contribution is multiplied by zero): half the L1 wavefronts and half the loop overhead per interaction
//     in the particle-particle pass.
// MEASURED (16 M Evrard, one B200, r02k): correct — interaction counters equal the reference algorithm's, parity 1e-10 — but
// SLOWER than k_gravity: 98.9 ms against 86.1 ms at 3 blocks/SM (166 registers), 108 ms at 4 blocks/SM (128 registers, 64 B of
// spills), 116 ms at 2; one / two / three leaf particles in flight 108 / 98.9 / 99.2 ms.  Fewer cells are accepted by a whole
// 64-particle group, the lane runs over the UNION of its two particles' cells and leaves (the non-accepting particle's
// FP64 work is wasted), and 12 warps per SM hide less latency than 16.  Not the default (environment SPHB_GRAVITY=2 selects it);
// kept because it is the evidence for "gathers are not what bounds the particle-particle pass".
#pragma once
#include "sphb_stages.cuh"

namespace sphb {

#ifndef SPHB_GV2_BLOCKS
#define SPHB_GV2_BLOCKS 3
#endif
#ifndef SPHB_GV2_STACK
#define SPHB_GV2_STACK 608
#endif
#ifndef SPHB_PP2_ILP
#define SPHB_PP2_ILP 2
#endif
constexpr int GV2_BLOCKS = SPHB_GV2_BLOCKS;   // resident blocks per SM of k_gravity2 (168 registers)
constexpr int GV2_STACK = SPHB_GV2_STACK;     // node stack entries per warp (<= 28 stay behind per tree level)
constexpr int PP2_ILP = SPHB_PP2_ILP;         // leaf particles in flight per lane (each meets two particles)
constexpr int GRAV2_GROUP = 64;               // particles per group

struct Grav2Smem {
    int4     stack[GV2_STACK];          // {child0 | (nchild - 1) << 29, particle mask word 0, word 1, -}: the children of an opened node
    int4     expand[32];                // {node, mask 0, mask 1, -} of the batch being fetched
    double4  pc[GV_PC];                 // accepted cells of the current chunk: mass centre, G * mass
    double   box[8];                    // the group's bounding box: centre[3], half width[3], cmax
    double4  gcell[GV_GC + 1];          // cells accepted by EVERY particle of the group (+ pad)
    double4  mx[32];                    // mixed nodes of the current batch: mass centre + mass
    double   me2[32];                   //   edge^2
    int4     minfo[32];                 //   {child0, nchild, first, count}
    double   mh2[32];                   //   leaves: largest h^2 among the leaf's particles
    unsigned mmask[2][32];              //   particle masks; afterwards the accept rows of the batch, compacted
    int      msl[32];                   //   chunk slot (within the batch) of the mixed node
};

// one monopole / unsoftened pair interaction: phi -= w / r, acc -= d w / r^3
template <int DIM>
__device__ __forceinline__ void mono(const double (&d)[DIM], double w, double & phi, double (&acc)[DIM])
{
    const double y = fast_rsqrt(dot<DIM>(d, d));
    phi -= w * y;
    const double s = w * y * (y * y);
#pragma unroll
    for (int a = 0; a < DIM; ++a) acc[a] -= d[a] * s;
}

// the same with a weight that may be 0 for a cell the particle did NOT accept: such a cell can be a single-particle leaf
// sitting exactly on the particle (d = 0), so the distance is replaced before the reciprocal square root
template <int DIM>
__device__ __forceinline__ void mono_masked(const double (&d)[DIM], double w, double & phi, double (&acc)[DIM])
{
    const double r2 = dot<DIM>(d, d);
    const double y = fast_rsqrt(w == 0.0 ? 1.0 : r2);
    phi -= w * y;
    const double s = w * y * (y * y);
#pragma unroll
    for (int a = 0; a < DIM; ++a) acc[a] -= d[a] * s;
}

template <int DIM, bool PERIODIC, bool COUNT>
__global__ void __launch_bounds__(128, GV2_BLOCKS)
k_gravity2(PSoA p, TreeDev t, DevParams P, GroupTable gt, const double4 * __restrict__ posm,
           const double2 * __restrict__ hsoft /* {2/h_j, h_j^2} */, double2 * __restrict__ scratch_lq,
           int * __restrict__ scratch_near, Counters * __restrict__ cnt, unsigned long long * __restrict__ d_err)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    Grav2Smem & sm = reinterpret_cast<Grav2Smem *>(s_dyn)[w];
    // per-lane queue of opened leaves, entry = {{first, count | which of the lane's particles << 28}, h_leaf^2}, and per-lane list
    // of possibly softened pairs (j | particle << 31); both in this warp's global scratch slot, [entry][lane]
    double2 * const lq = scratch_lq + ((size_t)(blockIdx.x * (blockDim.x >> 5) + w) * GRAV_LQ) * 32 + lane;
    int * const nearq = scratch_near + ((size_t)(blockIdx.x * (blockDim.x >> 5) + w) * GRAV_NEAR) * 32 + lane;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long tot_pp = 0, tot_pc = 0, tot_visit = 0, tot_pcg = 0, tot_ppg = 0;
    int g_first, g_cnt;
    while (next_group(gt, lane, g_first, g_cnt)) {
    bool valid[2];
    double ri[2][DIM], acc[2][DIM], phi[2], einv_i[2], h_i2[2];
    unsigned vmask[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int i = g_first + s * 32 + lane;
        valid[s] = s * 32 + lane < g_cnt;
        double h_i = 1.0;
        phi[s] = 0.0;                                        // phi = 0: src/bhtree.cpp:130
#pragma unroll
        for (int a = 0; a < DIM; ++a) { ri[s][a] = 0.0; acc[s][a] = 0.0; }
        if (valid[s]) {
            load_vec<DIM>(p.pos, i, ri[s]);
            load_vec<DIM>(p.acc, i, acc[s]);                  // gravity adds onto the fluid acceleration
            h_i = p.sml[i];
        }
        einv_i[s] = 2.0 / h_i;
        h_i2[s] = h_i * h_i * (1.0 + 1e-12);                  // softening test: r2 < max(h_i, h_j)^2 with a margin
        vmask[s] = __ballot_sync(SPHB_FULL_MASK, valid[s]);
    }
    unsigned int n_pp = 0, n_pc = 0, n_visit = 0, n_pcg = 0, n_ppg = 0;   // per lane (both particles): fit 32 bits
    unsigned pcw[2][2] = {{0u, 0u}, {0u, 0u}};               // [particle][word]: accept bits over the 64 chunk slots
    int npb = 0, nlq = 0, ngc = 0;                           // chunk slots, leaf queue entries, group cells in use
    {
        // bounding box of the group's (<= 64) particles: warp-uniform, kept in shared memory
        double lo[DIM], hi[DIM];
        double cmax = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double l = 1.7976931348623157e308, h = -1.7976931348623157e308;
#pragma unroll
            for (int s = 0; s < 2; ++s) if (valid[s]) { l = fmin(l, ri[s][d]); h = fmax(h, ri[s][d]); }
            lo[d] = warp_min(l);
            hi[d] = warp_max(h);
            const double bc = 0.5 * lo[d] + 0.5 * hi[d];
            const double bh = fmax(hi[d] - bc, bc - lo[d]) * (1.0 + 1e-12);
            cmax = fmax(cmax, fabs(bc) + bh);
            if (lane == 0) { sm.box[d] = bc; sm.box[3 + d] = bh; }
        }
        if (lane == 0) sm.box[6] = cmax;
    }

    int top = 1;
    if (lane == 0) sm.stack[0] = make_int4(0, (int)vmask[0], (int)vmask[1], 0);     // the root alone: child0 = 0, nchild = 1
    __syncwarp();
    int k = 0, node = -1;
    unsigned mask[2] = {0u, 0u};
    double2 q0 = make_double2(0.0, 0.0), q1 = q0, q2 = q0;

    for (;;) {
        const bool last = (k == 0 && top == 0);
        // ================= interaction loops (each exists once; all lanes arrive together) =================
        // (2) accepted cells of the chunk (monopole, src/bhtree.cpp:326-330): every lane runs over the union of its two
        // particles' accept bits; a cell is read once and meets both particles (weight 0 where not accepted)
        if (last || npb > GV_PC - 32) {
            __syncwarp();
            if (COUNT) n_pc += __popc(pcw[0][0]) + __popc(pcw[0][1]) + __popc(pcw[1][0]) + __popc(pcw[1][1]);
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                const unsigned m0 = pcw[0][blk], m1 = pcw[1][blk];
                unsigned mm = m0 | m1;
                const double4 * pcs = sm.pc + blk * 32;
                while (mm) {
                    const int e0 = __ffs(mm) - 1;
                    mm &= mm - 1;
                    const bool two = mm != 0;
                    const int e1 = two ? __ffs(mm) - 1 : e0;
                    mm &= mm - 1;                                  // stays 0 when !two
                    const double4 c0 = pcs[e0], c1 = pcs[e1];
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const unsigned ms = s == 0 ? m0 : m1;
                        double d0[DIM], d1[DIM];
                        grav_rij<DIM, PERIODIC>(P, ri[s], c0, d0);
                        grav_rij<DIM, PERIODIC>(P, ri[s], c1, d1);
                        mono_masked<DIM>(d0, ((ms >> e0) & 1u) ? c0.w : 0.0, phi[s], acc[s]);
                        mono_masked<DIM>(d1, (two && ((ms >> e1) & 1u)) ? c1.w : 0.0, phi[s], acc[s]);
                    }
                }
            }
            pcw[0][0] = pcw[0][1] = pcw[1][0] = pcw[1][1] = 0u;
            npb = 0;
            __syncwarp();
        }
        // (3) cells accepted by every particle of the group: all lanes run the same loop over the list, broadcast reads
        if (last || ngc > GV_GC - 32) {
            __syncwarp();
            if (COUNT) n_pc += ngc * ((valid[0] ? 1 : 0) + (valid[1] ? 1 : 0));
            if (ngc & 1) {     // pad to even: the last cell again (far from every particle by construction), massless
                if (lane == 0) { double4 pad = sm.gcell[ngc - 1]; pad.w = 0.0; sm.gcell[ngc] = pad; }
                ++ngc;
            }
            __syncwarp();
#pragma unroll 1
            for (int kk = 0; kk < ngc; kk += 2) {
                const double4 c0 = sm.gcell[kk], c1 = sm.gcell[kk + 1];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    double d0[DIM], d1[DIM];
                    grav_rij<DIM, PERIODIC>(P, ri[s], c0, d0);
                    grav_rij<DIM, PERIODIC>(P, ri[s], c1, d1);
                    mono<DIM>(d0, c0.w, phi[s], acc[s]);
                    mono<DIM>(d1, c1.w, phi[s], acc[s]);
                }
            }
            ngc = 0;
            __syncwarp();
        }
        // ================= the walk: classify the batch against the group's bounding box =================
        int cls = 0, child0 = 0, nchild = 0, first = 0, count = 0;
        double c[DIM], e2 = 0.0, mass = 0.0, hl2 = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) c[d] = 0.0;
        if (node >= 0) {
            c[0] = q0.x;
            if (DIM >= 2) c[DIM >= 2 ? 1 : 0] = q0.y;
            if (DIM >= 3) c[DIM >= 3 ? 2 : 0] = q1.x;
            mass = q1.y;
            e2 = q2.x;
            child0 = __double2loint(q2.y); nchild = __double2hiint(q2.y);
            double dmin2 = 0.0, dmax2 = 0.0;
            const double cmax = sm.box[6];
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                const double slack = 1e-13 * (cmax + fabs(c[d])) + 1e-300;
                const double bhd = sm.box[3 + d];
                double dc = sm.box[d] - c[d];
                if (PERIODIC) dc = min_image(dc, P.range[d]);
                dc = fabs(dc);
                const double lo = fmax(dc - bhd - slack, 0.0), hi = dc + bhd + slack;
                dmin2 += lo * lo;
                dmax2 += hi * hi;
            }
            if (e2 <= P.theta2 * dmin2 * (1.0 - 1e-9)) cls = 1;               // every particle accepts
            else if (e2 > P.theta2 * dmax2 * (1.0 + 1e-9)) cls = 2;           // every particle opens
            else cls = 3;
            if (nchild == 0 && cls != 1) {
                const double2 q3 = __ldg(t.ng + (size_t)node * 4 + 3);
                first = __double2loint(q3.x); count = __double2hiint(q3.x);
                hl2 = q3.y;
            }
        }
        const bool whole = mask[0] == vmask[0] && mask[1] == vmask[1];
        if (COUNT) {
            for (int s = 0; s < k; ++s) {
                n_visit += (__shfl_sync(SPHB_FULL_MASK, mask[0], s) >> lane) & 1u;
                n_visit += (__shfl_sync(SPHB_FULL_MASK, mask[1], s) >> lane) & 1u;
            }
            const int nv = (valid[0] ? 1 : 0) + (valid[1] ? 1 : 0);
            const unsigned bf = __ballot_sync(SPHB_FULL_MASK, cls == 1 && whole);
            n_pcg += __popc(bf) * nv;
            unsigned lf = __ballot_sync(SPHB_FULL_MASK, cls == 2 && nchild == 0 && whole);
            while (lf) { const int src = __ffs(lf) - 1; lf &= lf - 1; const int c0 = __shfl_sync(SPHB_FULL_MASK, count, src); n_ppg += c0 * nv; }
        }
        const bool grp = cls == 1 && whole;                                    // accepted by the whole group
        const unsigned b_grp = __ballot_sync(SPHB_FULL_MASK, grp);
        const unsigned b_acc = __ballot_sync(SPHB_FULL_MASK, cls == 1 && !grp);
        const unsigned b_mix = __ballot_sync(SPHB_FULL_MASK, cls == 3);
        const unsigned b_oleaf = __ballot_sync(SPHB_FULL_MASK, cls == 2 && nchild == 0);
        const unsigned b_oint = __ballot_sync(SPHB_FULL_MASK, cls == 2 && nchild > 0);

        // (a) opened by every particle of the mask, internal: push the children with the same mask
        if (b_oint) {
            const int total = __popc(b_oint);
            if (top + total > GV2_STACK) {
                if (lane == 0) atomicOr(&d_err[2], (unsigned long long)WALK_ERR_GRAV_STACK);
            } else {
                if ((b_oint >> lane) & 1u)
                    sm.stack[top + __popc(b_oint & lt_mask)] = make_int4(child0 | ((nchild - 1) << 29), (int)mask[0], (int)mask[1], 0);
                top += total;
            }
        }
        // (b) mixed nodes -> list in shared memory (tested particle by particle below)
        const int nmix = __popc(b_mix);
        const int mslot = __popc(b_mix & lt_mask);         // class 3: position in the mixed list
        const unsigned b_sel = b_acc | b_mix;              // nodes that get a slot of the masked chunk
        const int sslot = __popc(b_sel & lt_mask);         // position among the batch's chunk cells
        if (cls == 3) {
            sm.msl[mslot] = sslot;
            sm.mx[mslot] = make_double4(c[0], DIM >= 2 ? c[DIM >= 2 ? 1 : 0] : 0.0, DIM >= 3 ? c[DIM >= 3 ? 2 : 0] : 0.0, mass);
            sm.me2[mslot] = e2;
            sm.minfo[mslot] = make_int4(child0, nchild, first, count);
            sm.mh2[mslot] = hl2;
            sm.mmask[0][mslot] = mask[0];
            sm.mmask[1][mslot] = mask[1];
        }
        // (c) cells the whole group accepts -> group list
        if (grp) sm.gcell[ngc + __popc(b_grp & lt_mask)] = make_double4(c[0], DIM >= 2 ? c[DIM >= 2 ? 1 : 0] : 0.0, DIM >= 3 ? c[DIM >= 3 ? 2 : 0] : 0.0, P.G * mass);
        ngc += __popc(b_grp);
        // (d) cells some particles accept get the next free chunk slots; racc = particles that accept this lane's node
        const bool part = cls == 1 && !grp;
        const unsigned racc0 = part ? mask[0] : 0u, racc1 = part ? mask[1] : 0u;
        if (part || cls == 3)
            sm.pc[npb + sslot] = make_double4(c[0], DIM >= 2 ? c[DIM >= 2 ? 1 : 0] : 0.0, DIM >= 3 ? c[DIM >= 3 ? 2 : 0] : 0.0, P.G * mass);
        // (e) opened by every particle of the mask, leaf -> per-lane queues (one entry for both of the lane's particles)
        {
            unsigned bl = b_oleaf;
            while (bl) {
                const int src = __ffs(bl) - 1;
                bl &= bl - 1;
                const int f0 = __shfl_sync(SPHB_FULL_MASK, first, src);
                const int c0 = __shfl_sync(SPHB_FULL_MASK, count, src);
                const unsigned m0 = __shfl_sync(SPHB_FULL_MASK, mask[0], src), m1 = __shfl_sync(SPHB_FULL_MASK, mask[1], src);
                const double l2 = __shfl_sync(SPHB_FULL_MASK, hl2, src);
                const int fl = (int)((m0 >> lane) & 1u) | ((int)((m1 >> lane) & 1u) << 1);
                if (fl) { lq[nlq * 32] = make_double2(pack_ints(f0, c0 | (fl << 28)), l2); ++nlq; }
            }
        }
        // ================= particle-particle sums of the queued leaves (src/bhtree.cpp:309-317) =================
        // (between the classification of a batch and the fetch of the next one: the batch's registers are dead)
        // Pass 1: one flattened loop over all particles of the lane's queued leaves, every loaded record meets both of
        // the lane's particles with the unsoftened form; pairs that may be softened (r2 < max(h_i, h_leaf)^2) are only
        // LISTED.  Pass 2: the full Hernquist-Katz form (src/bhtree.cpp:273-299) over the listed pairs.
        if (last || __any_sync(SPHB_FULL_MASK, nlq > GRAV_LQ - 64)) {
            int q = 0, j = 0, jend = 0, fl = 0;
            double hl = 0.0;
            double2 en = make_double2(0.0, 0.0);           // the entry after the current one, already loaded
            if (nlq > 0) {
                const double2 e = lq[0];
                const int cf = __double2hiint(e.x);
                j = __double2loint(e.x); jend = j + (cf & 0x0fffffff); fl = cf >> 28; hl = e.y;
                q = 1;
                if (nlq > 1) en = lq[32];
            }
            int nav = min(PP2_ILP, jend - j);                // records of the coming trip, 0 = done
            double4 pr[PP2_ILP];
#pragma unroll
            for (int u = 0; u < PP2_ILP; ++u) pr[u] = make_double4(0.0, 0.0, 0.0, 0.0);
            if (nav > 0) {
#pragma unroll
                for (int u = 0; u < PP2_ILP; ++u) pr[u] = ldg4(&posm[j + min(u, nav - 1)]);
            }
            do {
                int nnear = 0;
                while (nav > 0 && nnear <= GRAV_NEAR - 2 * PP2_ILP) {
                    double4 cr[PP2_ILP];
#pragma unroll
                    for (int u = 0; u < PP2_ILP; ++u) cr[u] = pr[u];
                    const int cj = j, cn = nav, cfl = fl;
                    const double thr0 = fmax(h_i2[0], hl), thr1 = fmax(h_i2[1], hl);
                    j += cn;
                    if (j >= jend) {                            // next leaf: its entry is in registers already
                        const bool more = q < nlq;
                        const int cf = more ? __double2hiint(en.x) : 0;
                        j = more ? __double2loint(en.x) : 0;
                        jend = j + (cf & 0x0fffffff);
                        fl = cf >> 28;
                        hl = en.y;
                        ++q;
                        if (q < nlq) en = lq[q * 32];
                    }
                    nav = min(PP2_ILP, jend - j);
                    if (nav > 0) {
#pragma unroll
                        for (int u = 0; u < PP2_ILP; ++u) pr[u] = ldg4(&posm[j + min(u, nav - 1)]);
                    }
#pragma unroll
                    for (int u = 0; u < PP2_ILP; ++u) {
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            double d[DIM];
                            grav_rij<DIM, PERIODIC>(P, ri[s], cr[u], d);
                            const double r2 = dot<DIM>(d, d);
                            const bool on = u < cn && ((cfl >> s) & 1);
                            const bool nx = r2 < (s == 0 ? thr0 : thr1);
                            const double y = fast_rsqrt(nx ? 1.0 : r2);
                            const double gm = (nx || !on) ? 0.0 : P.G * cr[u].w;
                            phi[s] -= gm * y;
                            const double sc = gm * y * (y * y);
#pragma unroll
                            for (int a = 0; a < DIM; ++a) acc[s][a] -= d[a] * sc;
                            if (nx && on) { nearq[nnear * 32] = (int)((unsigned)(cj + u) | ((unsigned)s << 31)); ++nnear; }
                            if (COUNT) n_pp += on ? 1 : 0;
                        }
                    }
                }
                for (int kk = 0; kk < nnear; ++kk) {
                    const int jn = nearq[kk * 32];
                    const int jj = jn & 0x7fffffff;
                    const bool s1 = jn < 0;
                    const double4 pj = ldg4(&posm[jj]);
                    const double einv_j = __ldg(&hsoft[jj]).x;
                    double rs[DIM];
#pragma unroll
                    for (int a = 0; a < DIM; ++a) rs[a] = s1 ? ri[1][a] : ri[0][a];
                    double d[DIM];
                    grav_rij<DIM, PERIODIC>(P, rs, pj, d);
                    const double r2 = dot<DIM>(d, d);
                    const double rinv = rsqrt(r2);              // inf at r == 0, unused there (u < 1 branch)
                    const double r = r2 > 0.0 ? r2 * rinv : 0.0;
                    double fi, gi, fj, gj;
                    soft_fg_fast(r, rinv, s1 ? einv_i[1] : einv_i[0], fi, gi);
                    soft_fg_fast(r, rinv, einv_j, fj, gj);
                    const double gm = P.G * pj.w;
                    const double dphi = gm * (fi + fj) * 0.5;   // src/bhtree.cpp:314-315
                    const double sg = gm * (gi + gj) * 0.5;
                    if (s1) {
                        phi[1] -= dphi;
#pragma unroll
                        for (int a = 0; a < DIM; ++a) acc[1][a] -= d[a] * sg;
                    } else {
                        phi[0] -= dphi;
#pragma unroll
                        for (int a = 0; a < DIM; ++a) acc[0][a] -= d[a] * sg;
                    }
                }
            } while (nav > 0);                                  // only if the softened-pair list ran full
            nlq = 0;
        }
        if (last) break;

        // ---- fetch the next batch now: its loads are in flight during the per-particle tests
        __syncwarp();
        {
            const int ne = min(top, 32);
            int4 ent = make_int4(0, 0, 0, 0);
            int nc = 0;
            if (lane < ne) { ent = sm.stack[top - 1 - lane]; nc = (int)((unsigned)ent.x >> 29) + 1; }
            int incl = nc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(SPHB_FULL_MASK, incl, o);
                if (lane >= o) incl += y;
            }
            const int m = __popc(__ballot_sync(SPHB_FULL_MASK, lane < ne && incl <= 32));   // entries taken (a prefix)
            k = m > 0 ? __shfl_sync(SPHB_FULL_MASK, incl, m - 1) : 0;
            if (lane < m) {
                const int c0 = ent.x & 0x1fffffff;
                for (int ci = 0; ci < nc; ++ci) sm.expand[incl - nc + ci] = make_int4(c0 + ci, ent.y, ent.z, 0);
            }
            __syncwarp();
            node = -1;
            mask[0] = mask[1] = 0u;
            if (lane < k) {
                const int4 e = sm.expand[lane];
                node = e.x;
                mask[0] = (unsigned)e.y; mask[1] = (unsigned)e.z;
                const double2 * q = t.ng + (size_t)node * 4;
                q0 = __ldg(q); q1 = __ldg(q + 1); q2 = __ldg(q + 2);
            }
            top -= m;
            __syncwarp();
        }
        // (f) mixed nodes: the reference's own per-particle test (src/bhtree.cpp:303-308), straight-line over the mixed list
        // for both of the lane's particles; bit-matrix transposes then give lane q the particles that accept / open mixed node q
        unsigned my_open[2] = {0u, 0u}, my_acc[2] = {0u, 0u};
#pragma unroll 2
        for (int q = 0; q < nmix; ++q) {
            const double4 c4 = sm.mx[q];
            const double me2 = sm.me2[q];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const unsigned mm = sm.mmask[s][q];
                double d[DIM];
                grav_rij<DIM, PERIODIC>(P, ri[s], c4, d);
                const double d2 = abs2_exact<DIM>(d);
                const bool in = (mm >> lane) & 1u;
                const bool op = in && me2 > __dmul_rn(P.theta2, d2);
                my_open[s] |= (op ? 1u : 0u) << q;
                my_acc[s] |= ((in && !op) ? 1u : 0u) << q;
            }
        }
        if (nmix) {
            const unsigned acc_row0 = warp_transpose32(my_acc[0], lane), acc_row1 = warp_transpose32(my_acc[1], lane);
            const unsigned open_row0 = warp_transpose32(my_open[0], lane), open_row1 = warp_transpose32(my_open[1], lane);
            int4 info = make_int4(0, 0, 0, 0);
            int msl = 0;
            if (lane < nmix) { info = sm.minfo[lane]; msl = sm.msl[lane]; }
            // opened internal nodes: push the children with the mask of the particles that opened
            const bool psh = lane < nmix && (open_row0 | open_row1) != 0u && info.y != 0;
            const unsigned pb = __ballot_sync(SPHB_FULL_MASK, psh);
            if (pb) {
                const int total = __popc(pb);
                if (top + total > GV2_STACK) {
                    if (lane == 0) atomicOr(&d_err[2], (unsigned long long)WALK_ERR_GRAV_STACK);
                } else {
                    if (psh) sm.stack[top + __popc(pb & lt_mask)] = make_int4(info.x | ((info.y - 1) << 29), (int)open_row0, (int)open_row1, 0);
                    top += total;
                }
            }
            // opened leaves: every lane appends the leaves its particles opened to its queue
            const unsigned leaf_bits = __ballot_sync(SPHB_FULL_MASK, lane < nmix && info.y == 0);
            unsigned lb = (my_open[0] | my_open[1]) & leaf_bits;
            while (lb) {
                const int q = __ffs(lb) - 1;
                lb &= lb - 1;
                const int4 inf = sm.minfo[q];
                const int fl = (int)((my_open[0] >> q) & 1u) | ((int)((my_open[1] >> q) & 1u) << 1);
                lq[nlq * 32] = make_double2(pack_ints(inf.z, inf.w | (fl << 28)), sm.mh2[q]);
                ++nlq;
            }
            __syncwarp();                                  // the particle masks in sm.mmask are dead now
            if (lane < nmix) { sm.mmask[0][msl] = acc_row0; sm.mmask[1][msl] = acc_row1; }   // accept rows of mixed node `lane` at its chunk slot
        }
        // accept masks of the batch: rows compacted to the chunk slots, then node-major -> particle-major
        if (b_sel) {
            if (part) { sm.mmask[0][sslot] = racc0; sm.mmask[1][sslot] = racc1; }
            __syncwarp();
            const int nsel = __popc(b_sel);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const unsigned row = lane < nsel ? sm.mmask[s][lane] : 0u;
                const unsigned long long tw = (unsigned long long)warp_transpose32(row, lane) << npb;
                pcw[s][0] |= (unsigned)tw;
                pcw[s][1] |= (unsigned)(tw >> 32);
            }
            npb += nsel;
        }
        __syncwarp();
    }

#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (valid[s]) {
            const int i = g_first + s * 32 + lane;
#pragma unroll
            for (int a = 0; a < DIM; ++a) p.acc[a][i] = acc[s][a];
            p.phi[i] = phi[s];
        }
    }
    if (COUNT) { tot_pp += n_pp; tot_pc += n_pc; tot_visit += n_visit; tot_pcg += n_pcg; tot_ppg += n_ppg; }
    __syncwarp();
    }
    if (COUNT) {
        const unsigned long long a = warp_sum_u64(tot_pp), b = warp_sum_u64(tot_pc), cc = warp_sum_u64(tot_visit);
        const unsigned long long pg = warp_sum_u64(tot_pcg), qg = warp_sum_u64(tot_ppg);
        if (lane == 0) { atomicAdd(&cnt->grav_pp, a); atomicAdd(&cnt->grav_pc, b); atomicAdd(&cnt->grav_node_visits, cc);
                         atomicAdd(&cnt->grav_pc_group, pg); atomicAdd(&cnt->grav_pp_group, qg); }
    }
}

} // namespace sphb
