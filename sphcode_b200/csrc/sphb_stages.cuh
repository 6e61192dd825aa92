// sphb_stages.cuh — the per-step stage kernels (sm_100a, FP64):
//   PreInteraction   src/pre_interaction.cpp:39-283, src/disph/d_pre_interaction.cpp:21-230,
//                    src/gsph/g_pre_interaction.cpp:27-144
//   FluidForce       src/fluid_force.cpp:26-116, src/disph/d_fluid_force.cpp:26-86,
//                    src/gsph/g_fluid_force.cpp:28-199
//   GravityForce     src/gravity_force.cpp:52-89 -> src/bhtree.cpp:128-132,301-331
//   TimeStep         src/timestep.cpp:18-38
//   predict/correct  src/solver.cpp:431-474
// One warp owns 32 consecutive particles of the sorted order; lane = particle i.  Every sum over
// neighbours j is a warp_walk (sphb_tree.cuh) whose leaf handler streams the leaf's particles from a
// shared-memory tile, so there are no per-particle neighbour lists in memory at all; the only
// per-lane list is the r (and m) column of the Newton iteration, which must see one fixed
// candidate set several times (src/pre_interaction.cpp:227-283).
#pragma once
#include "sphb_walk.cuh"

namespace sphb {

#ifndef SPHB_PF_BLOCKS
#define SPHB_PF_BLOCKS 5
#endif
constexpr int PF_BLOCKS = SPHB_PF_BLOCKS;   // resident blocks per SM of k_pre_interaction (96 registers)
#ifndef SPHB_FF_BLOCKS
#define SPHB_FF_BLOCKS 4
#endif
constexpr int FF_BLOCKS = SPHB_FF_BLOCKS;   // ... of k_fluid_force: 4 x 128 registers (16 bytes of spills) beat 5 x 96 with 88 bytes of spills (measured)

struct Counters {   // device mirror of sphb_counters (include/sphb.h), all summed over particles
    unsigned long long newton_evals, newton_iters, pre_candidates, pre_neighbors, force_pairs,
                       grav_pp, grav_pc, grav_node_visits, nonconverged, list_overflow,
                       grav_pc_group, grav_pp_group;
};

template <int DIM> __device__ __forceinline__ void load_vec(double * const (&a)[3], int i, double (&o)[DIM])
{
#pragma unroll
    for (int d = 0; d < DIM; ++d) o[d] = a[d][i];
}
template <int DIM> __device__ __forceinline__ double dot(const double (&a)[DIM], const double (&b)[DIM])
{
    double s = a[0] * b[0];
#pragma unroll
    for (int d = 1; d < DIM; ++d) s += a[d] * b[d];
    return s;
}

// Smoothing-length guess, src/pre_interaction.cpp:61-64.
template <int DIM> __device__ __forceinline__ double h_guess(int ngb, double mass, double dens)
{
    const double x = ngb * mass / (dens * unit_ball<DIM>());
    if (DIM == 1) return x;
    if (DIM == 2) return sqrt(x);
    return cbrt(x);
}

// Packed gather records (tree order), 32 bytes each, so that a pair body fetches a neighbour with a few
// 16-byte loads instead of one scattered 8-byte load per field:
//   posm   {x, y, z, m}            rebuilt by make_tree
//   velc   {vx, vy, vz, c}         c = sound speed
//   thermo {u, h, dens, pres}      u before PreInteraction; h, dens, pres written by PreInteraction
//   av     {gradh, alpha, balsara, -}   written by PreInteraction
//   hsoft  {2/h, h^2}              gravity softening record
struct Recs { double4 *posm, *velc, *thermo, *av; double2 *hsoft; };

// p: the rank's own particles (local index i); the records are indexed by the global tree-order index off + i
template <int DIM>
__global__ void k_pack_recs(PSoA p, Recs r, int n, int what, int off)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = off + i;
    constexpr int Y = DIM >= 2 ? 1 : 0, Z = DIM >= 3 ? 2 : 0;
    if (what & 1) r.posm[g] = make_double4(p.pos[0][i], DIM >= 2 ? p.pos[Y][i] : 0.0, DIM >= 3 ? p.pos[Z][i] : 0.0, p.mass[i]);
    if (what & 2) r.velc[g] = make_double4(p.vel[0][i], DIM >= 2 ? p.vel[Y][i] : 0.0, DIM >= 3 ? p.vel[Z][i] : 0.0, p.sound[i]);
    if (what & 4) {
        r.thermo[g] = make_double4(p.ene[i], p.sml[i], p.dens[i], p.pres[i]);
        r.av[g] = make_double4(p.gradh[i], p.alpha[i], p.balsara[i], 0.0);
    }
}

// Gather-permute of the whole particle state by the sorted index, fused with the packing of the gather
// records (the values are in registers anyway): replaces k_permute + k_pack_recs(7) in the tree build.
template <int DIM>
__global__ void k_permute_pack(PSoA s, PSoA d, Recs r, const int * __restrict__ perm, int n, int gsph, int off)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int q = perm[i];
    double pos[3] = {0.0, 0.0, 0.0}, vel[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        pos[a] = s.pos[a][q]; d.pos[a][i] = pos[a];
        vel[a] = s.vel[a][q]; d.vel[a][i] = vel[a];
        d.vel_p[a][i] = s.vel_p[a][q];
        d.acc[a][i] = s.acc[a][q];
    }
    const double mass = s.mass[q], dens = s.dens[q], pres = s.pres[q], ene = s.ene[q], sml = s.sml[q], sound = s.sound[q],
                 balsara = s.balsara[q], alpha = s.alpha[q], gradh = s.gradh[q];
    d.mass[i] = mass; d.dens[i] = dens; d.pres[i] = pres; d.ene[i] = ene; d.sml[i] = sml; d.sound[i] = sound;
    d.balsara[i] = balsara; d.alpha[i] = alpha; d.gradh[i] = gradh;
    d.ene_p[i] = s.ene_p[q]; d.dene[i] = s.dene[q]; d.phi[i] = s.phi[q];
    d.pid[i] = s.pid[q]; d.neighbor[i] = s.neighbor[q]; d.orig[i] = s.orig[q];
    if (gsph) {
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            d.grad_d[a][i] = s.grad_d[a][q]; d.grad_p[a][i] = s.grad_p[a][q];
#pragma unroll
            for (int v = 0; v < DIM; ++v) d.grad_v[v][a][i] = s.grad_v[v][a][q];
        }
    }
    const int g = off + i;                  // records: global tree-order index
    r.posm[g] = make_double4(pos[0], pos[1], pos[2], mass);
    r.velc[g] = make_double4(vel[0], vel[1], vel[2], sound);
    r.thermo[g] = make_double4(ene, sml, dens, pres);
    r.av[g] = make_double4(gradh, alpha, balsara, 0.0);
}

template <int DIM> __device__ __forceinline__ void vec_from4(const double4 & q, double (&o)[DIM])
{
    o[0] = q.x;
    if (DIM >= 2) o[DIM >= 2 ? 1 : 0] = q.y;
    if (DIM >= 3) o[DIM >= 3 ? 2 : 0] = q.z;
}

// =================================================================================================
// initial_smoothing, src/pre_interaction.cpp:171-215: dens_i = sum_{r < h} m_j W(r, h), h = guess
// =================================================================================================
template <int DIM, int KT>
struct InitSmoothV {
    const DevParams & P; const double4 * posm;
    double ri[DIM], h, h2, dens;
    KernelCoef<DIM, KT> kc;
    __device__ __forceinline__ void hit(int j)
    {
        const double4 pj = ldg4(&posm[j]);
        double d[DIM];
        rij_from4<DIM>(P, ri, pj, d);
        const double r2 = abs2_exact<DIM>(d);
        if (r2 < h2) {
            const double r = sqrt(r2);
            if (r < h) dens += pj.w * kc.w(r);
        }
    }
};

template <int DIM, int KT>
__global__ void __launch_bounds__(128) k_initial_smoothing(PSoA p, Recs rc, TreeDev t, DevParams P, GroupTable gt, unsigned long long * __restrict__ d_err)
{
    __shared__ NWalkSmem s_walk[4];
    const int lane = threadIdx.x & 31;
    int g_first, g_cnt;
    while (next_group(gt, lane, g_first, g_cnt)) {
    const int i = g_first + lane;
    const bool valid = lane < g_cnt;
    InitSmoothV<DIM, KT> v{P, rc.posm};
    v.dens = 0.0;
    v.h = 1.0;
    if (valid) {
        load_vec<DIM>(p.pos, i, v.ri);
        v.h = h_guess<DIM>(P.ngb, p.mass[i], p.dens[i]);
    } else {
#pragma unroll
        for (int d = 0; d < DIM; ++d) v.ri[d] = 0.0;
    }
    v.h2 = __dmul_rn(v.h, v.h);
    v.kc.init(v.h);
    group_stream<DIM, false>(t, P, rc.posm, nullptr, 0, s_walk[threadIdx.x >> 5], lane, v.ri, v.h, valid, v, d_err);
    if (valid) { p.sml[i] = v.h; p.dens[i] = v.dens; }
    }
}

// =================================================================================================
// PreInteraction::calculation
// =================================================================================================
// candidate set {j : r2 < h_search^2} (src/bhtree.cpp:251-261) -> per-lane columns r, j (and m) in
// the warp's scratch slot (column layout [k][lane]: every later pass reads them coalesced)
// per-lane index column in the warp's scratch slot (layout [k][lane]: later passes read it coalesced)
struct IndexListV {
    int * lj;
    int cap, cnt, lane;
    __device__ __forceinline__ void hit(int j)
    {
        if (cnt < cap) lj[cnt * 32 + lane] = j;
        ++cnt;
    }
};

// sums over {j : r2 < h_search^2 and r < h_i}: density pass + Balsara / MUSCL-gradient pass
// (the reference runs them as two loops over the same sorted list prefix,
//  src/pre_interaction.cpp:83-103 and 116-161; the second only needs dens_i at the very end)
template <int DIM, int KT, int SPH>
struct DensityAcc {
    const DevParams & P; const PSoA & p; const Recs & rc;
    int i;
    double ri[DIM], vi[DIM], ci, ui;
    KernelCoef<DIM, KT> kc;
    bool need_div;
    // accumulators
    double dens, dh_dens, n_i, dh_n, pres, dh_pres, v_sig_max;
    int n_neighbor;
    double div_v, rot_v[3];
    double dd[DIM], du[DIM], dv[DIM][DIM];   // GSPH

    __device__ __forceinline__ void pair(int j, double r)
    {
        const double4 pj = ldg4(&rc.posm[j]);
        const double4 vc = ldg4(&rc.velc[j]);
        double d[DIM];
        rij_from4<DIM>(P, ri, pj, d);
        ++n_neighbor;
        const double mj = pj.w;
        const double w = kc.w(r);
        dens += mj * w;
        double uj = 0.0;
        if (SPH == T_SSPH) {
            dh_dens += mj * kc.dhw(r);
        } else if (SPH == T_DISPH) {
            const double dhw = kc.dhw(r);
            uj = __ldg(&rc.thermo[j].x);
            n_i += w;
            pres += mj * uj * w;
            dh_pres += mj * uj * dhw;
            dh_n += dhw;
        }
        double vij[DIM];
        {
            double vj[DIM];
            vec_from4<DIM>(vc, vj);
#pragma unroll
            for (int k = 0; k < DIM; ++k) vij[k] = vi[k] - vj[k];
        }
        if (j != i) {
            const double v_sig = ci + vc.w - 3.0 * dot<DIM>(d, vij) / r;
            if (v_sig > v_sig_max) v_sig_max = v_sig;
        }
        if (SPH == T_GSPH) {
            if (P.gsph2) {
                const double c = kc.dwc(r);
                const double uji = __ldg(&rc.thermo[j].x) - ui;
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const double dwa = d[a] * c;
                    dd[a] += dwa * mj;
                    du[a] += dwa * (mj * uji);
#pragma unroll
                    for (int k = 0; k < DIM; ++k) dv[k][a] += dwa * (mj * (-vij[k]));
                }
            }
        } else if (need_div) {
            const double c = kc.dwc(r);
            double dw[DIM];
#pragma unroll
            for (int a = 0; a < DIM; ++a) dw[a] = d[a] * c;
            const double wgt = (SPH == T_DISPH) ? mj * uj : mj;
            div_v -= wgt * dot<DIM>(vij, dw);
            if (DIM == 2) {
                rot_v[0] += (vij[0] * dw[DIM > 1 ? 1 : 0] - vij[DIM > 1 ? 1 : 0] * dw[0]) * wgt;
            } else if (DIM == 3) {
                constexpr int Y = DIM > 1 ? 1 : 0, Z = DIM > 2 ? 2 : 0;
                rot_v[0] += (vij[Y] * dw[Z] - vij[Z] * dw[Y]) * wgt;
                rot_v[1] += (vij[Z] * dw[0] - vij[0] * dw[Z]) * wgt;
                rot_v[2] += (vij[0] * dw[Y] - vij[Y] * dw[0]) * wgt;
            }
        }
    }
};

template <int DIM, int KT, int SPH>
__global__ void __launch_bounds__(128, PF_BLOCKS)
k_pre_interaction(PSoA p, Recs rc, TreeDev t, DevParams P, GroupTable gt,
                  double * __restrict__ scratch_r, double * __restrict__ scratch_m, int * __restrict__ scratch_j,
                  const double * __restrict__ d_dt, double * __restrict__ d_hpvs,
                  unsigned long long * __restrict__ d_err, Counters * __restrict__ cnt)
{
    __shared__ NWalkSmem s_walk[4];
    NWalkSmem & sm = s_walk[threadIdx.x >> 5];
    constexpr bool NEED_M = (SPH != T_DISPH);     // DISPH Newton uses unit weights (d_pre_interaction.cpp:208-209)
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    double * lr = scratch_r + (size_t)slot * P.list_cap * 32;
    int * lj = scratch_j + (size_t)slot * P.list_cap * 32;
    double * lm = NEED_M ? scratch_m + (size_t)slot * P.list_cap * 32 : nullptr;
    const double dt = *d_dt;
    double hpvs_min = 1.7976931348623157e308;
    unsigned int c_evals = 0, c_iters = 0, c_cand = 0, c_ngb = 0, c_nonconv = 0, c_over = 0;   // per lane: fits 32 bits

    int g_first, g_cnt;
    while (next_group(gt, lane, g_first, g_cnt)) {
        const int i = g_first + lane;
        const bool valid = lane < g_cnt;

        double ri[DIM], vi[DIM];
        double mass_i = 1.0, dens_old = 1.0, ene_i = 0.0, c_i = 0.0, alpha_i = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { ri[d] = 0.0; vi[d] = 0.0; }
        if (valid) {
            load_vec<DIM>(p.pos, i, ri);
            load_vec<DIM>(p.vel, i, vi);
            mass_i = p.mass[i]; dens_old = p.dens[i]; ene_i = p.ene[i]; c_i = p.sound[i]; alpha_i = p.alpha[i];
        }
        // guess smoothing length (src/pre_interaction.cpp:61-64)
        const double hs = h_guess<DIM>(P.ngb, mass_i, dens_old) * P.kernel_ratio;
        const double hs2 = __dmul_rn(hs, hs);
        double h = hs;

        // ---- candidates (src/pre_interaction.cpp:66-71): conservative hits of the group search ...
        int nraw;
        {
            IndexListV cv{lj, P.list_cap, 0, lane};
            group_stream<DIM, false>(t, P, rc.posm, nullptr, 0, sm, lane, ri, hs, valid, cv, d_err);
            nraw = valid ? cv.cnt : 0;
        }
        if (nraw > P.list_cap) { ++c_over; nraw = P.list_cap; }
        __syncwarp();
        // ... reduced to the candidate set {j : r2 < h_search^2} (src/bhtree.cpp:251-261) by the exact test,
        // all lanes together; columns r, j (and m) are compacted in place
        int ncand = 0;
        // the sums of the FIRST Newton iteration (h = h0) are taken here, while r is in a register: most
        // particles converge in one iteration, which then never re-reads the r column
        const double h0 = hs / P.kernel_ratio;
        double s_first = 0.0, sd_first = 0.0;
        unsigned int evals_first = 0;
        KernelCoef<DIM, KT> kc0;
        kc0.init(h0);
        {
            const int nraw_max = __reduce_max_sync(SPHB_FULL_MASK, nraw);
            // two raw hits per trip: both gathers are in flight before the first is used
            for (int k0 = 0; k0 < nraw_max; k0 += 2) {
                const bool in0 = k0 < nraw, in1 = k0 + 1 < nraw;
                const int j0 = in0 ? lj[k0 * 32 + lane] : i, j1 = in1 ? lj[(k0 + 1) * 32 + lane] : i;
                const double4 p0 = ldg4(&rc.posm[valid ? j0 : 0]), p1 = ldg4(&rc.posm[valid ? j1 : 0]);
                double d0[DIM], d1[DIM];
                rij_from4<DIM>(P, ri, p0, d0);
                rij_from4<DIM>(P, ri, p1, d1);
                const double r20 = abs2_exact<DIM>(d0), r21 = abs2_exact<DIM>(d1);
                if (in0 && r20 < hs2) {
                    const double r = sqrt(r20);
                    lr[ncand * 32 + lane] = r;
                    lj[ncand * 32 + lane] = j0;
                    if (NEED_M) lm[ncand * 32 + lane] = p0.w;
                    ++ncand;
                    if (P.iterative && r < h0) {
                        s_first += NEED_M ? p0.w * kc0.w(r) : kc0.w(r);
                        sd_first += NEED_M ? p0.w * kc0.dhw(r) : kc0.dhw(r);
                        ++evals_first;
                    }
                }
                if (in1 && r21 < hs2) {
                    const double r = sqrt(r21);
                    lr[ncand * 32 + lane] = r;
                    lj[ncand * 32 + lane] = j1;
                    if (NEED_M) lm[ncand * 32 + lane] = p1.w;
                    ++ncand;
                    if (P.iterative && r < h0) {
                        s_first += NEED_M ? p1.w * kc0.w(r) : kc0.w(r);
                        sd_first += NEED_M ? p1.w * kc0.dhw(r) : kc0.dhw(r);
                        ++evals_first;
                    }
                }
            }
        }
        c_cand += ncand;
        const int ncand_max = __reduce_max_sync(SPHB_FULL_MASK, ncand);

        if (P.iterative) {
            // ---- Newton-Raphson (src/pre_interaction.cpp:227-283)
            const double b = NEED_M ? mass_i * P.ngb / unit_ball<DIM>() : P.ngb / unit_ball<DIM>();
            h = h0;
            bool done = !valid, conv = false;
            for (int it = 0; it < 10; ++it) {
                if (!__any_sync(SPHB_FULL_MASK, !done)) break;
                double s = s_first, sd = sd_first;
                if (it == 0) {
                    if (!done) c_evals += evals_first;
                } else {
                    KernelCoef<DIM, KT> kc;
                    kc.init(h);
                    s = 0.0; sd = 0.0;
                    const int nk = done ? 0 : ncand;
#pragma unroll 4
                    for (int k = 0; k < ncand_max; ++k) {
                        if (k < nk) {
                            const double r = lr[k * 32 + lane];
                            if (r < h) {
                                if (NEED_M) {
                                    const double m = lm[k * 32 + lane];
                                    s += m * kc.w(r);
                                    sd += m * kc.dhw(r);
                                } else {
                                    s += kc.w(r);
                                    sd += kc.dhw(r);
                                }
                                ++c_evals;
                            }
                        }
                    }
                }
                if (!done) {
                    ++c_iters;
                    const double f = s * powh<DIM>(h) - b;
                    const double df = sd * powh<DIM>(h) + DIM * s * powh_<DIM>(h);
                    const double hn = h - f / df;
                    if (fabs(hn - h) < (hn + h) * 1e-4) { done = true; conv = true; }
                    h = hn;
                }
            }
            if (valid && !conv) { h = h0; ++c_nonconv; }    // logged + fallback, pre_interaction.cpp:277-282
        }

        // ---- density pass (+ Balsara / MUSCL gradients) over the candidates with r < h
        DensityAcc<DIM, KT, SPH> dv{P, p, rc};
        dv.i = i;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { dv.ri[d] = ri[d]; dv.vi[d] = vi[d]; }
        dv.ci = c_i; dv.ui = ene_i;
        dv.kc.init(h);
        dv.need_div = (SPH != T_GSPH) && ((P.use_balsara && DIM != 1) || P.use_tdav);
        dv.dens = dv.dh_dens = dv.n_i = dv.dh_n = dv.pres = dv.dh_pres = 0.0;
        dv.v_sig_max = c_i * 2.0;
        dv.n_neighbor = 0;
        dv.div_v = 0.0; dv.rot_v[0] = dv.rot_v[1] = dv.rot_v[2] = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            dv.dd[a] = 0.0; dv.du[a] = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) dv.dv[k][a] = 0.0;
        }
        {
            // the reference's sorted loop breaks at the first r >= h (src/pre_interaction.cpp:87-91): here the
            // entries with r < h are first compacted in place by all lanes together (independent, coalesced
            // column reads; the write cursor never passes the read cursor), then every lane runs the pair
            // body over its dense prefix
            const int nk = valid ? ncand : 0;
            int nn = 0;
            for (int k0 = 0; k0 < ncand_max; k0 += 4) {
                double r4[4]; int j4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool in = k0 + u < nk;
                    r4[u] = in ? lr[(k0 + u) * 32 + lane] : 1.7976931348623157e308;
                    j4[u] = in ? lj[(k0 + u) * 32 + lane] : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (r4[u] < h) { lr[nn * 32 + lane] = r4[u]; lj[nn * 32 + lane] = j4[u]; ++nn; }
                }
            }
            const int nn_max = __reduce_max_sync(SPHB_FULL_MASK, nn);
            for (int k = 0; k < nn_max; ++k) {
                if (k < nn) dv.pair(lj[k * 32 + lane], lr[k * 32 + lane]);
            }
        }

        if (valid) {
            const double dens_i = dv.dens;
            double pres_i, div_norm, gradh_i = 0.0, bal_i = 1.0, alpha_new = alpha_i;
            if (SPH == T_SSPH) {
                pres_i = (P.gamma - 1.0) * dens_i * ene_i;
                gradh_i = 1.0 / (1.0 + h / (DIM * dens_i) * dv.dh_dens);
                p.gradh[i] = gradh_i;
                div_norm = 1.0 / dens_i;
            } else if (SPH == T_DISPH) {
                pres_i = (P.gamma - 1.0) * dv.pres;
                gradh_i = h / (DIM * dv.n_i) * dv.dh_pres / (1.0 + h / (DIM * dv.n_i) * dv.dh_n);
                p.gradh[i] = gradh_i;
                div_norm = (P.gamma - 1.0) / pres_i;
            } else {
                pres_i = (P.gamma - 1.0) * dens_i * ene_i;
                div_norm = 0.0;
                gradh_i = p.gradh[i];
            }
            p.sml[i] = h;
            p.dens[i] = dens_i;
            p.pres[i] = pres_i;
            p.neighbor[i] = dv.n_neighbor;
            c_ngb += dv.n_neighbor;
            hpvs_min = fmin(hpvs_min, h / dv.v_sig_max);

            if (SPH == T_GSPH) {
                if (P.gsph2) {
                    const double rho_inv = 1.0 / dens_i;
#pragma unroll
                    for (int a = 0; a < DIM; ++a) {
                        p.grad_d[a][i] = dv.dd[a];
                        p.grad_p[a][i] = (dv.dd[a] * ene_i + dv.du[a]) * (P.gamma - 1.0);
#pragma unroll
                        for (int k = 0; k < DIM; ++k) p.grad_v[k][a][i] = dv.dv[k][a] * rho_inv;
                    }
                }
                bal_i = p.balsara[i];
            } else if (P.use_balsara && DIM != 1) {
                const double div_v = (SPH == T_SSPH) ? dv.div_v / dens_i : dv.div_v * div_norm;
                double rot_abs;
                if (DIM == 2) {
                    rot_abs = fabs((SPH == T_SSPH) ? dv.rot_v[0] / dens_i : dv.rot_v[0] * div_norm);
                } else {
                    double rr[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) rr[a] = (SPH == T_SSPH) ? dv.rot_v[a] / dens_i : dv.rot_v[a] * div_norm;
                    rot_abs = sqrt(rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2]);
                }
                bal_i = fabs(div_v) / (fabs(div_v) + rot_abs + 1e-4 * c_i / h);
                p.balsara[i] = bal_i;
                if (P.use_tdav) {
                    const double tau_inv = P.epsilon_av * c_i / h;
                    const double dalpha = (-(alpha_i - P.alpha_min) * tau_inv + fmax(-div_v, 0.0) * (P.alpha_max - alpha_i)) * dt;
                    alpha_new = alpha_i + dalpha;
                    p.alpha[i] = alpha_new;
                }
            } else {
                bal_i = p.balsara[i];
                if (P.use_tdav) {
                    const double div_v = (SPH == T_SSPH) ? dv.div_v / dens_i : dv.div_v * div_norm;
                    const double tau_inv = P.epsilon_av * c_i / h;
                    const double s_i = fmax(-div_v, 0.0);
                    alpha_new = (alpha_i + dt * tau_inv * P.alpha_min + s_i * dt * P.alpha_max) / (1.0 + dt * tau_inv + s_i * dt);
                    p.alpha[i] = alpha_new;
                }
            }
            // gather records of the force pass (u stays as packed: other warps may be reading it)
            double * th = reinterpret_cast<double *>(&rc.thermo[i]);
            th[1] = h;
            *reinterpret_cast<double2 *>(th + 2) = make_double2(dens_i, pres_i);
            rc.av[i] = make_double4(gradh_i, alpha_new, bal_i, 0.0);
        }
        __syncwarp();
    }

    hpvs_min = warp_min(hpvs_min);
    if (lane == 0) atomic_min_pos(d_hpvs, hpvs_min);
    if (cnt) {
        const unsigned long long e = warp_sum_u64(c_evals), it = warp_sum_u64(c_iters), ca = warp_sum_u64(c_cand),
                                 ng = warp_sum_u64(c_ngb);
        if (lane == 0) {
            atomicAdd(&cnt->newton_evals, e); atomicAdd(&cnt->newton_iters, it);
            atomicAdd(&cnt->pre_candidates, ca); atomicAdd(&cnt->pre_neighbors, ng);
        }
    }
    const unsigned long long nc = warp_sum_u64(c_nonconv), ov = warp_sum_u64(c_over);
    if (lane == 0 && (nc | ov)) {      // always-on error counters
        atomicAdd(&d_err[0], nc);
        atomicAdd(&d_err[1], ov);
    }
}

// =================================================================================================
// FluidForce::calculation — symmetric pair set {j : r2 < max(h_i, ksize(leaf))^2, 0 < r < max(h_i, h_j)}
// =================================================================================================
// Monaghan (1997) signal-velocity viscosity, src/fluid_force.cpp:89-106
__device__ __forceinline__ double art_visc(double vr, double r, double ci, double cj, double ai, double aj,
                                           double bi, double bj, double di, double dj)
{
    if (vr < 0) {
        const double alpha = 0.5 * (ai + aj);
        const double balsara = 0.5 * (bi + bj);
        const double w_ij = vr / r;
        const double v_sig = ci + cj - 3.0 * w_ij;
        const double rho_ij_inv = 2.0 / (di + dj);
        return -0.5 * balsara * alpha * v_sig * w_ij * rho_ij_inv;
    }
    return 0.0;
}

// van Leer (1979) limiter, src/gsph/g_fluid_force.cpp:28-36
__device__ __forceinline__ double van_leer(double dq1, double dq2)
{
    const double dq1dq2 = dq1 * dq2;
    if (dq1dq2 <= 0) return 0.0;
    return 2.0 * dq1dq2 / (dq1 + dq2);
}

// HLL solver, src/gsph/g_fluid_force.cpp:168-199.  state = {u, rho, p, c}
__device__ __forceinline__ void hll(const double (&left)[4], const double (&right)[4], double & pstar, double & vstar)
{
    const double u_l = left[0], rho_l = left[1], p_l = left[2], c_l = left[3];
    const double u_r = right[0], rho_r = right[1], p_r = right[2], c_r = right[3];
    const double roe_l = sqrt(rho_l);
    const double roe_r = sqrt(rho_r);
    const double roe_inv = 1.0 / (roe_l + roe_r);
    const double u_t = (roe_l * u_l + roe_r * u_r) * roe_inv;
    const double c_t = (roe_l * c_l + roe_r * c_r) * roe_inv;
    const double s_l = fmin(u_l - c_l, u_t - c_t);
    const double s_r = fmax(u_r + c_r, u_t + c_t);
    const double c1 = rho_l * (s_l - u_l);
    const double c2 = rho_r * (s_r - u_r);
    const double c3 = 1.0 / (c1 - c2);
    const double c4 = p_l - u_l * c1;
    const double c5 = p_r - u_r * c2;
    vstar = (c5 - c4) * c3;
    pstar = (c1 * c5 - c2 * c4) * c3;
}

// pair set of lane i: {j : 0 < r < max(h_i, h_j)} (src/fluid_force.cpp:62; the reference's candidate
// test r2 < max(h_i, kernel_size(leaf))^2, src/bhtree.cpp:255-256, is implied by it).  The group search
// records the conservative hits (IndexListV); ForceAcc::pair applies this exact filter first.
template <int DIM, int KT, int SPH>
struct ForceAcc {
    const DevParams & P; const PSoA & p; const Recs & rc;
    int i;
    double dt;
    double ri[DIM], vi[DIM], h_i, m_i, dens_i, pres_i, gradh_i, alpha_i, bal_i, c_i, u_i;
    double gdi[DIM], gpi[DIM], gvi[DIM][DIM];          // GSPH gradients of i
    double pp_i;            // SSPH: P_i/rho_i^2 * gradh_i ; DISPH: (g-1)^2 u_i / P_i ; GSPH: 1/rho_i^2
    double g2u_i;           // DISPH: (g-1)^2 u_i
    double m_u_inv;         // DISPH: 1/(m_i u_i)
    KernelCoef<DIM, KT> ki;
    double acc[DIM], dene;
    unsigned int pairs;

    __device__ __forceinline__ void pair(int j)
    {
        const double4 pj = ldg4(&rc.posm[j]);
        const double4 th = ldg4(&rc.thermo[j]);          // {u, h, dens, pres}
        double d[DIM];
        rij_from4<DIM>(P, ri, pj, d);
        const double r = sqrt(abs2_exact<DIM>(d));
        const double h_j = th.y;
        if (r >= fmax(h_i, h_j) || r == 0.0) return;     // src/fluid_force.cpp:62
        ++pairs;
        const double4 vc = ldg4(&rc.velc[j]);
        KernelCoef<DIM, KT> kj;
        kj.init(h_j);
        const double cwi = ki.dwc(r), cwj = kj.dwc(r);
        double dw_i[DIM], dw_j[DIM], vij[DIM], vj[DIM];
        vec_from4<DIM>(vc, vj);
#pragma unroll
        for (int a = 0; a < DIM; ++a) { dw_i[a] = d[a] * cwi; dw_j[a] = d[a] * cwj; vij[a] = vi[a] - vj[a]; }
        const double m_j = pj.w;
        const double dens_j = th.z, pres_j = th.w;
        const double c_j = vc.w;

        if (SPH == T_GSPH) {
            const double r_inv = 1.0 / r;
            double e[DIM];
#pragma unroll
            for (int a = 0; a < DIM; ++a) e[a] = d[a] * r_inv;
            const double ve_i = dot<DIM>(vi, e);
            const double ve_j = dot<DIM>(vj, e);
            double vstar, pstar;
            if (P.gsph2) {
                // Murante et al. (2011), src/gsph/g_fluid_force.cpp:96-134
                double right[4], left[4];
                const double delta_i = 0.5 * (1.0 - c_i * dt * r_inv);
                const double delta_j = 0.5 * (1.0 - c_j * dt * r_inv);
                const double dv_ij = ve_i - ve_j;
                double dvi[DIM], dvj[DIM], gdj[DIM], gpj[DIM];
#pragma unroll
                for (int k = 0; k < DIM; ++k) {
                    double gvj[DIM];
#pragma unroll
                    for (int a = 0; a < DIM; ++a) gvj[a] = p.grad_v[k][a][j];
                    dvi[k] = dot<DIM>(gvi[k], e);
                    dvj[k] = dot<DIM>(gvj, e);
                }
#pragma unroll
                for (int a = 0; a < DIM; ++a) { gdj[a] = p.grad_d[a][j]; gpj[a] = p.grad_p[a][j]; }
                const double dve_i = dot<DIM>(dvi, e) * r;
                const double dve_j = dot<DIM>(dvj, e) * r;
                right[0] = ve_i - van_leer(dv_ij, dve_i) * delta_i;
                left[0] = ve_j + van_leer(dv_ij, dve_j) * delta_j;
                const double dd_ij = dens_i - dens_j;
                const double dd_i = dot<DIM>(gdi, e) * r;
                const double dd_j = dot<DIM>(gdj, e) * r;
                right[1] = dens_i - van_leer(dd_ij, dd_i) * delta_i;
                left[1] = dens_j + van_leer(dd_ij, dd_j) * delta_j;
                const double dp_ij = pres_i - pres_j;
                const double dp_i = dot<DIM>(gpi, e) * r;
                const double dp_j = dot<DIM>(gpj, e) * r;
                right[2] = pres_i - van_leer(dp_ij, dp_i) * delta_i;
                left[2] = pres_j + van_leer(dp_ij, dp_j) * delta_j;
                right[3] = sqrt(P.gamma * right[2] / right[1]);
                left[3] = sqrt(P.gamma * left[2] / left[1]);
                hll(left, right, pstar, vstar);
            } else {
                const double right[4] = {ve_i, dens_i, pres_i, c_i};
                const double left[4] = {ve_j, dens_j, pres_j, c_j};
                hll(left, right, pstar, vstar);
            }
            const double rho2_inv_j = 1.0 / (dens_j * dens_j);
            double fdotv = 0.0;
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                const double f = dw_i[a] * (m_j * pstar * pp_i) + dw_j[a] * (m_j * pstar * rho2_inv_j);
                acc[a] -= f;
                fdotv += f * (e[a] * vstar - vi[a]);
            }
            dene -= fdotv;
        } else {
            const double4 avj = ldg4(&rc.av[j]);          // {gradh, alpha, balsara, -}
            const double vr = dot<DIM>(vij, d);
            const double pi_ij = art_visc(vr, r, c_i, c_j, alpha_i, avj.y, bal_i, avj.z, dens_i, dens_j);
            double dw_ij[DIM];
#pragma unroll
            for (int a = 0; a < DIM; ++a) dw_ij[a] = (dw_i[a] + dw_j[a]) * 0.5;
            const double u_j = th.x;
            double dene_ac = 0.0;
            if (P.use_ac) {
                // src/fluid_force.cpp:108-116
                const double v_sig = P.use_gravity ? fabs(vr / r) : sqrt(2.0 * fabs(pres_i - pres_j) / (dens_i + dens_j));
                dene_ac = P.alpha_ac * m_j * v_sig * (u_i - u_j) * dot<DIM>(dw_ij, d) / r;
            }
            double ti, tj, ei;
            if (SPH == T_SSPH) {
                // src/fluid_force.cpp:78-79
                ti = m_j * (pp_i + 0.5 * pi_ij);
                tj = m_j * (pres_j / (dens_j * dens_j) * avj.x + 0.5 * pi_ij);
                ei = m_j * pp_i;
            } else {
                // src/disph/d_fluid_force.cpp:70-79
                const double f_ij = 1.0 - gradh_i / (m_j * u_j);
                const double f_ji = 1.0 - avj.x * m_u_inv;
                const double u_per_pres_j = u_j / pres_j;
                ti = m_j * (pp_i * u_j * f_ij + 0.5 * pi_ij);
                tj = m_j * (g2u_i * u_per_pres_j * f_ji + 0.5 * pi_ij);
                ei = m_j * pp_i * u_j * f_ij;
            }
#pragma unroll
            for (int a = 0; a < DIM; ++a) acc[a] -= dw_i[a] * ti + dw_j[a] * tj;
            dene += ei * dot<DIM>(vij, dw_i) + 0.5 * m_j * pi_ij * dot<DIM>(vij, dw_ij) + dene_ac;
        }
    }
};

template <int DIM, int KT, int SPH>
__global__ void __launch_bounds__(128, FF_BLOCKS)
k_fluid_force(PSoA p, Recs rc, TreeDev t, DevParams P, GroupTable gt,
              int * __restrict__ scratch_j, const double * __restrict__ d_dt,
              unsigned long long * __restrict__ d_err, Counters * __restrict__ cnt)
{
    __shared__ NWalkSmem s_walk[4];
    NWalkSmem & sm = s_walk[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int * lj = scratch_j + (size_t)slot * P.list_cap * 32;
    const double dt = *d_dt;
    unsigned int c_pairs = 0, c_over = 0;

    int g_first, g_cnt;
    while (next_group(gt, lane, g_first, g_cnt)) {
        const int i = g_first + lane;
        const bool valid = lane < g_cnt;

        ForceAcc<DIM, KT, SPH> v{P, p, rc};
        v.i = i;
        v.dt = dt;
        v.dene = 0.0;
        v.pairs = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) { v.acc[a] = 0.0; v.ri[a] = 0.0; v.vi[a] = 0.0; }
        v.h_i = 1.0; v.m_i = 1.0; v.dens_i = 1.0; v.pres_i = 1.0; v.gradh_i = 0.0; v.alpha_i = 0.0; v.bal_i = 0.0; v.c_i = 0.0; v.u_i = 1.0;
        if (valid) {
            load_vec<DIM>(p.pos, i, v.ri);
            load_vec<DIM>(p.vel, i, v.vi);
            v.h_i = p.sml[i]; v.m_i = p.mass[i]; v.dens_i = p.dens[i]; v.pres_i = p.pres[i];
            v.c_i = p.sound[i]; v.u_i = p.ene[i];
            if (SPH != T_GSPH) { v.gradh_i = p.gradh[i]; v.alpha_i = p.alpha[i]; v.bal_i = p.balsara[i]; }
            if (SPH == T_GSPH && P.gsph2) {
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    v.gdi[a] = p.grad_d[a][i]; v.gpi[a] = p.grad_p[a][i];
#pragma unroll
                    for (int k = 0; k < DIM; ++k) v.gvi[k][a] = p.grad_v[k][a][i];
                }
            }
        }
        if (SPH == T_SSPH) {
            v.pp_i = v.pres_i / (v.dens_i * v.dens_i) * v.gradh_i;
        } else if (SPH == T_DISPH) {
            v.g2u_i = (P.gamma - 1.0) * (P.gamma - 1.0) * v.u_i;
            v.pp_i = v.g2u_i / v.pres_i;
            v.m_u_inv = 1.0 / (v.m_i * v.u_i);
        } else {
            v.pp_i = 1.0 / (v.dens_i * v.dens_i);
        }
        v.ki.init(v.h_i);

        // ---- symmetric pair search
        int npair;
        {
            IndexListV cv{lj, P.list_cap, 0, lane};
            group_stream<DIM, true>(t, P, rc.posm, reinterpret_cast<const double *>(rc.thermo) + 1, 4, sm, lane, v.ri, v.h_i, valid, cv, d_err);
            npair = cv.cnt;
        }
        if (npair > P.list_cap) { ++c_over; npair = P.list_cap; }
        if (!valid) npair = 0;
        __syncwarp();
        const int npair_max = __reduce_max_sync(SPHB_FULL_MASK, npair);
        // ---- pair bodies, all lanes together
        for (int k = 0; k < npair_max; ++k) {
            if (k < npair) v.pair(lj[k * 32 + lane]);
        }
        if (valid) {
#pragma unroll
            for (int a = 0; a < DIM; ++a) p.acc[a][i] = v.acc[a];
            p.dene[i] = v.dene;
            c_pairs += v.pairs;
        }
        __syncwarp();
    }
    if (cnt) {
        const unsigned long long s = warp_sum_u64(c_pairs);
        if (lane == 0) atomicAdd(&cnt->force_pairs, s);
    }
    const unsigned long long ov = warp_sum_u64(c_over);
    if (lane == 0 && ov) atomicAdd(&d_err[1], ov);
}

// =================================================================================================
// GravityForce::calculation -> BHNode::calc_force, src/bhtree.cpp:301-331
// =================================================================================================
// Group walk, breadth-first with one lane per node (as in sphb_walk.cuh), that reproduces the
// reference's PER-PARTICLE opening decisions exactly:
//   * every stack entry carries the mask of lanes (particles) that opened all its ancestors;
//   * a node is first classified against the group's bounding box: if even the nearest point of the
//     box accepts it (edge^2 <= theta^2 d_min^2) every lane of the mask accepts it; if even the
//     farthest point opens it, every lane opens it; both with a safety margin, so these decisions
//     coincide with the per-particle test of src/bhtree.cpp:308;
//   * the remaining ("mixed") nodes are tested lane by lane with the reference's own expression,
//     which splits the mask into the lanes that accept and the lanes that descend.
// The set of cells a particle accepts and of leaves it opens is therefore exactly
// BHNode::calc_force's.  The interactions themselves are deferred so that lanes do them together:
//   * accepted cells go to a 32-entry chunk staged in shared memory (mass centre + mass) with a
//     per-lane accept mask; when the chunk is full every lane runs over ITS mask;
//   * opened leaves go to a per-lane queue of (first, count); when any lane's queue is full, every
//     lane runs one flattened loop over all particles of its queued leaves (packed x,y,z,m + 2/h
//     read through L1; lanes of a warp are Morton neighbours and share these lines).
#ifndef SPHB_GRAV_LQ
#define SPHB_GRAV_LQ 128
#endif
#ifndef SPHB_PP_ILP
#define SPHB_PP_ILP 3
#endif
#ifndef SPHB_GV_BLOCKS
#define SPHB_GV_BLOCKS 4
#endif
#ifndef SPHB_GV_GC
#define SPHB_GV_GC 64
#endif
#ifndef SPHB_GV_GROUPCELLS
#define SPHB_GV_GROUPCELLS 1
#endif
#ifndef SPHB_GC_UNROLL
#define SPHB_GC_UNROLL 1
#endif
constexpr int GC_UNROLL = SPHB_GC_UNROLL;   // pairs of group cells per loop trip
constexpr int GV_BLOCKS = SPHB_GV_BLOCKS;   // resident blocks per SM of k_gravity
constexpr int GRAV_LQ = SPHB_GRAV_LQ;       // leaf queue depth per lane (global scratch, [entry][lane]); flushed above GRAV_LQ - 64
constexpr int PP_ILP = SPHB_PP_ILP;         // particle-particle pairs in flight per lane
constexpr int GRAV_NEAR = 128;    // softened-pair list depth per lane (global scratch, [entry][lane])
#ifndef SPHB_GV_STACK
#define SPHB_GV_STACK 704
#endif
constexpr int GV_STACK = SPHB_GV_STACK;     // node stack entries per warp (<= 28 stay behind per tree level, see pop_load)
constexpr int GV_NB = 2;          // 32-bit words of a lane's accept mask over the chunk
constexpr int GV_PC = 32 * GV_NB; // chunk slots, filled densely; flushed when another batch might not fit
constexpr int GV_GC = SPHB_GV_GC; // list of cells accepted by the whole group, flushed likewise

struct GravSmem {
    int2     stack[GV_STACK];           // {child0 | (nchild - 1) << 29, lane mask}: the children of an opened node
    int2     expand[32];                // {node, lane mask} of the batch being fetched
    double4  pc[GV_PC];                 // accepted cells of the current chunk: mass centre, G * mass (one 32-byte slot per cell)
    double   box[8];                    // the group's bounding box: centre[3], half width[3], cmax (warp-uniform, read on use)
    double4  gcell[GV_GC + 1];          // cells accepted by EVERY particle of the group: mass centre, G * mass (+ pad)
    double4  mx[32];                    // mixed nodes of the current batch: mass centre + mass
    double   me2[32];                   //   edge^2
    int4     minfo[32];                 //   {child0, nchild, first, count}
    double   mh2[32];                   //   leaves: largest h^2 among the leaf's particles (k_apply_ksize)
    unsigned mmask[32];                 //   lane mask; afterwards the accept rows of the batch, compacted
    int      msl[32];                   //   chunk slot (within the batch) of the mixed node
};

// Softening functions with the divisions by constants turned into products (soft_fg in
// sphb_math.cuh keeps the literal form for the direct-sum checker kernel).
__device__ __forceinline__ void soft_fg_fast(double r, double rinv, double einv, double & f, double & g)
{
    const double u = r * einv;
    if (u < 1.0) {
        const double u2 = u * u;
        f = (-0.5 * u2 * (1.0 / 3.0 - (3.0 / 20) * u2 + (1.0 / 20) * (u2 * u)) + 1.4) * einv;
        g = (4.0 / 3.0 - 1.2 * u2 + 0.5 * (u2 * u)) * (einv * einv * einv);
    } else if (u < 2.0) {
        const double u2 = u * u, u3 = u2 * u;
        f = (-1.0 / 15) * rinv + (-u2 * (4.0 / 3.0 - u + 0.3 * u2 - (1.0 / 30) * u3) + 1.6) * einv;
        g = (-1.0 / 15 + (8.0 / 3) * u3 - 3 * (u3 * u) + 1.2 * (u3 * u2) - (1.0 / 6.0) * (u3 * u3)) * (rinv * rinv * rinv);
    } else {
        f = rinv;
        g = rinv * rinv * rinv;
    }
}

// r_i - c with the minimum image folded in at compile time
template <int DIM, bool PERIODIC>
__device__ __forceinline__ void grav_rij(const DevParams & P, const double (&ri)[DIM], const double4 & pj, double (&d)[DIM])
{
    d[0] = ri[0] - pj.x;
    if (DIM >= 2) d[DIM >= 2 ? 1 : 0] = ri[DIM >= 2 ? 1 : 0] - pj.y;
    if (DIM >= 3) d[DIM >= 3 ? 2 : 0] = ri[DIM >= 3 ? 2 : 0] - pj.z;
    if (PERIODIC) {
#pragma unroll
        for (int a = 0; a < DIM; ++a) d[a] = min_image(d[a], P.range[a]);
    }
}

// Code size matters here: the SM's instruction cache holds 32 KB, the warps of an SM sit in different
// phases of this kernel, and instruction fetch from L2 showed up as the top stall as soon as the
// interaction loops were inlined at several call sites.  Hence every interaction loop exists ONCE
// (all flushes happen at the head of the walk loop), PERIODIC and COUNT are compile-time, and the
// counters are a separate instantiation.
template <int DIM, bool PERIODIC, bool COUNT>
__global__ void __launch_bounds__(128, GV_BLOCKS)
k_gravity(PSoA p, TreeDev t, DevParams P, GroupTable gt, const double4 * __restrict__ posm,
          const double2 * __restrict__ hsoft /* {2/h_j, h_j^2} */, double2 * __restrict__ scratch_lq,
          int * __restrict__ scratch_near, Counters * __restrict__ cnt, unsigned long long * __restrict__ d_err)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    GravSmem & sm = reinterpret_cast<GravSmem *>(s_dyn)[w];
    // per-lane queue of opened leaves, entry = {{first, count}, softening threshold max(h_i, h_leaf)^2},
    // deep enough that the lanes' particle-particle work evens out before a flush, and per-lane list of
    // possibly softened pairs; both live in this warp's global scratch slot, [entry][lane]
    double2 * const lq = scratch_lq + ((size_t)(blockIdx.x * (blockDim.x >> 5) + w) * GRAV_LQ) * 32 + lane;
    int * const nearq = scratch_near + ((size_t)(blockIdx.x * (blockDim.x >> 5) + w) * GRAV_NEAR) * 32 + lane;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long tot_pp = 0, tot_pc = 0, tot_visit = 0, tot_pcg = 0, tot_ppg = 0;
    int g_first, g_cnt;
    while (next_group(gt, lane, g_first, g_cnt)) {
    const int i = g_first + lane;
    const bool valid = lane < g_cnt;
    double ri[DIM], acc[DIM], phi = 0.0, h_i = 1.0;          // phi = 0: src/bhtree.cpp:130
#pragma unroll
    for (int a = 0; a < DIM; ++a) { ri[a] = 0.0; acc[a] = 0.0; }
    if (valid) {
        load_vec<DIM>(p.pos, i, ri);
        load_vec<DIM>(p.acc, i, acc);                         // gravity adds onto the fluid acceleration
        h_i = p.sml[i];
    }
    const double einv_i = 2.0 / h_i;
    const double h_i2 = h_i * h_i * (1.0 + 1e-12);           // softening test: r2 < max(h_i, h_j)^2 with a margin
    unsigned int n_pp = 0, n_pc = 0, n_visit = 0, n_pcg = 0, n_ppg = 0;   // per lane: fit 32 bits
    unsigned pcw0 = 0, pcw1 = 0;                             // lane's accept bits over the chunk slots
    int npb = 0, nlq = 0, ngc = 0;                           // chunk slots, leaf queue entries, group cells in use

    const unsigned vmask = __ballot_sync(SPHB_FULL_MASK, valid);
    {
        // bounding box of the group: warp-uniform, kept in shared memory (14 registers less across the interaction loops)
        double bc[DIM], bh[DIM];
        group_box<DIM>(ri, valid, bc, bh);
        double cmax = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) cmax = fmax(cmax, fabs(bc[d]) + bh[d]);
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < DIM; ++d) { sm.box[d] = bc[d]; sm.box[3 + d] = bh[d]; }
            sm.box[6] = cmax;
        }
    }

    // ---- node stack and the batch held in registers.  A stack entry stands for ALL children of an
    // opened node (they are contiguous), so a batch of <= 32 nodes pushes <= 32 entries and pops >= 32 / NCH:
    // at most 28 entries stay behind per tree level, GV_STACK covers the deepest tree (21 levels).
    int top = 1;
    if (lane == 0) sm.stack[0] = make_int2(0, (int)vmask);          // the root alone: child0 = 0, nchild = 1
    __syncwarp();
    int k = 0, node = -1;
    unsigned mask = 0;
    double2 q0 = make_double2(0.0, 0.0), q1 = q0, q2 = q0;

    for (;;) {
        const bool last = (k == 0 && top == 0);
        // ================= interaction loops (each exists once; all lanes arrive together) =================
        // (2) accepted cells of the chunk (monopole, src/bhtree.cpp:326-330): every lane runs over ITS
        // accept bits, two cells in flight
        if (last || npb > GV_PC - 32) {
            __syncwarp();
            if (COUNT) n_pc += __popc(pcw0) + __popc(pcw1);
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                unsigned mm = blk == 0 ? pcw0 : pcw1;
                const double4 * pcs = sm.pc + blk * 32;
                while (mm) {
                    const int e0 = __ffs(mm) - 1;
                    mm &= mm - 1;
                    const bool two = mm != 0;
                    const int e1 = two ? __ffs(mm) - 1 : e0;
                    mm &= mm - 1;                                  // stays 0 when !two
                    const double4 c0 = pcs[e0];
                    double4 c1 = pcs[e1];
                    if (!two) c1.w = 0.0;
                    double d0[DIM], d1[DIM];
                    grav_rij<DIM, PERIODIC>(P, ri, c0, d0);
                    grav_rij<DIM, PERIODIC>(P, ri, c1, d1);
                    const double y0 = fast_rsqrt(dot<DIM>(d0, d0)), y1 = fast_rsqrt(dot<DIM>(d1, d1));
                    phi -= c0.w * y0;
                    phi -= c1.w * y1;
                    const double s0 = c0.w * y0 * (y0 * y0), s1 = c1.w * y1 * (y1 * y1);
#pragma unroll
                    for (int a = 0; a < DIM; ++a) { acc[a] -= d0[a] * s0; acc[a] -= d1[a] * s1; }
                }
            }
            pcw0 = 0; pcw1 = 0;
            npb = 0;
            __syncwarp();
        }
        // (3) cells accepted by every particle of the group: all lanes run the same loop over the list,
        // broadcast reads, two cells in flight
        if (last || ngc > GV_GC - 32) {
            __syncwarp();
            if (COUNT) n_pc += ngc;
            if (ngc & 1) {     // pad to even: the last cell again (far from every lane by construction), massless
                if (lane == 0) { double4 pad = sm.gcell[ngc - 1]; pad.w = 0.0; sm.gcell[ngc] = pad; }
                ++ngc;
            }
            __syncwarp();
#pragma unroll (GC_UNROLL)
            for (int kk = 0; kk < ngc; kk += 2) {
                const double4 c0 = sm.gcell[kk], c1 = sm.gcell[kk + 1];
                double d0[DIM], d1[DIM];
                grav_rij<DIM, PERIODIC>(P, ri, c0, d0);
                grav_rij<DIM, PERIODIC>(P, ri, c1, d1);
                const double y0 = fast_rsqrt(dot<DIM>(d0, d0)), y1 = fast_rsqrt(dot<DIM>(d1, d1));
                phi -= c0.w * y0;
                phi -= c1.w * y1;
                const double s0 = c0.w * y0 * (y0 * y0), s1 = c1.w * y1 * (y1 * y1);
#pragma unroll
                for (int a = 0; a < DIM; ++a) { acc[a] -= d0[a] * s0; acc[a] -= d1[a] * s1; }
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
            if (e2 <= P.theta2 * dmin2 * (1.0 - 1e-9)) cls = 1;               // every lane accepts
            else if (e2 > P.theta2 * dmax2 * (1.0 + 1e-9)) cls = 2;           // every lane opens
            else cls = 3;
            if (nchild == 0 && cls != 1) {
                const double2 q3 = __ldg(t.ng + (size_t)node * 4 + 3);
                first = __double2loint(q3.x); count = __double2hiint(q3.x);
                hl2 = q3.y;
            }
        }
        if (COUNT) {
            for (int s = 0; s < k; ++s) n_visit += (__shfl_sync(SPHB_FULL_MASK, mask, s) >> lane) & 1u;
            const unsigned bf = __ballot_sync(SPHB_FULL_MASK, cls == 1 && mask == vmask);
            if (valid) n_pcg += __popc(bf);
            unsigned lf = __ballot_sync(SPHB_FULL_MASK, cls == 2 && nchild == 0 && mask == vmask);
            while (lf) { const int src = __ffs(lf) - 1; lf &= lf - 1; const int c0 = __shfl_sync(SPHB_FULL_MASK, count, src); if (valid) n_ppg += c0; }
        }
        const bool grp = SPHB_GV_GROUPCELLS && cls == 1 && mask == vmask;      // accepted by the whole group
        const unsigned b_grp = __ballot_sync(SPHB_FULL_MASK, grp);
        const unsigned b_acc = __ballot_sync(SPHB_FULL_MASK, cls == 1 && !grp);
        const unsigned b_mix = __ballot_sync(SPHB_FULL_MASK, cls == 3);
        const unsigned b_oleaf = __ballot_sync(SPHB_FULL_MASK, cls == 2 && nchild == 0);
        const unsigned b_oint = __ballot_sync(SPHB_FULL_MASK, cls == 2 && nchild > 0);

        // (a) opened by every lane of the mask, internal: push the children with the same mask
        if (b_oint) {
            const int total = __popc(b_oint);
            if (top + total > GV_STACK) {
                if (lane == 0) atomicOr(&d_err[2], (unsigned long long)WALK_ERR_GRAV_STACK);
            } else {
                if ((b_oint >> lane) & 1u)
                    sm.stack[top + __popc(b_oint & lt_mask)] = make_int2(child0 | ((nchild - 1) << 29), (int)mask);
                top += total;
            }
        }
        // (b) mixed nodes -> list in shared memory (tested lane by lane below)
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
            sm.mmask[mslot] = mask;
        }
        // (c) cells the whole group accepts -> group list
        if (grp) sm.gcell[ngc + __popc(b_grp & lt_mask)] = make_double4(c[0], DIM >= 2 ? c[DIM >= 2 ? 1 : 0] : 0.0, DIM >= 3 ? c[DIM >= 3 ? 2 : 0] : 0.0, P.G * mass);
        ngc += __popc(b_grp);
        // (d) cells some lanes accept (class 1 with a partial mask: every lane of the mask; class 3: decided
        // below) get the next free chunk slots; racc = lanes that accept this lane's node
        const unsigned racc = (cls == 1 && !grp) ? mask : 0u;
        if ((cls == 1 && !grp) || cls == 3)
            sm.pc[npb + sslot] = make_double4(c[0], DIM >= 2 ? c[DIM >= 2 ? 1 : 0] : 0.0, DIM >= 3 ? c[DIM >= 3 ? 2 : 0] : 0.0, P.G * mass);
        // (e) opened by every lane of the mask, leaf -> per-lane queues (at most one entry per node of the
        // batch and lane: the head of the loop leaves room for 32)
        {
            unsigned bl = b_oleaf;
            while (bl) {
                const int src = __ffs(bl) - 1;
                bl &= bl - 1;
                const int f0 = __shfl_sync(SPHB_FULL_MASK, first, src);
                const int c0 = __shfl_sync(SPHB_FULL_MASK, count, src);
                const unsigned m = __shfl_sync(SPHB_FULL_MASK, mask, src);
                const double l2 = __shfl_sync(SPHB_FULL_MASK, hl2, src);
                if ((m >> lane) & 1u) { lq[nlq * 32] = make_double2(pack_ints(f0, c0), fmax(h_i2, l2)); ++nlq; }
            }
        }
        // ================= particle-particle sums of the queued leaves (src/bhtree.cpp:309-317) =================
        // Placed here, where the batch of the walk is consumed and the next one is not fetched yet: the registers of
        // the batch are dead, which is what lets PP_ILP pairs be in flight per lane without spills.  Pass 1 runs one
        // flattened loop over all particles of the lane's queued leaves with the unsoftened form and only LISTS the
        // pairs that may be softened (r2 < max(h_i, h_leaf)^2 >= max(h_i, h_j)^2: no per-pair load of h_j); pass 2
        // runs the full Hernquist-Katz form over the listed pairs (it reduces to the unsoftened form for u >= 2, so
        // listing too many is harmless).  A lane gets at most 64 entries between two visits of this point (the
        // mixed leaves of this batch and the opened leaves of the next).
        if (last || __any_sync(SPHB_FULL_MASK, nlq > GRAV_LQ - 64)) {
            int q = 0, j = 0, jend = 0;
            double thr2 = 0.0;
            double2 en = make_double2(0.0, 0.0);           // the entry after the current one, already loaded
            if (nlq > 0) {
                const double2 e = lq[0];
                j = __double2loint(e.x); jend = j + __double2hiint(e.x); thr2 = e.y;
                q = 1;
                if (nlq > 1) en = lq[32];
            }
            // software pipeline: the records of the NEXT PP_ILP pairs (all of one leaf) are loaded before the current ones are used
            int nav = min(PP_ILP, jend - j);                // pairs of the coming trip, 0 = done
            double4 pr[PP_ILP];
#pragma unroll
            for (int u = 0; u < PP_ILP; ++u) pr[u] = make_double4(0.0, 0.0, 0.0, 0.0);
            if (nav > 0) {
#pragma unroll
                for (int u = 0; u < PP_ILP; ++u) pr[u] = ldg4(&posm[j + min(u, nav - 1)]);
            }
            do {
                int nnear = 0;
                while (nav > 0 && nnear <= GRAV_NEAR - PP_ILP) {
                    double4 cr[PP_ILP];
#pragma unroll
                    for (int u = 0; u < PP_ILP; ++u) cr[u] = pr[u];
                    const int cj = j, cn = nav;
                    const double cthr2 = thr2;
                    j += cn;
                    if (j >= jend) {                            // next leaf: its entry is in registers already
                        const bool more = q < nlq;
                        j = more ? __double2loint(en.x) : 0;
                        jend = more ? j + __double2hiint(en.x) : 0;
                        thr2 = en.y;
                        ++q;
                        if (q < nlq) en = lq[q * 32];
                    }
                    nav = min(PP_ILP, jend - j);
                    if (nav > 0) {
#pragma unroll
                        for (int u = 0; u < PP_ILP; ++u) pr[u] = ldg4(&posm[j + min(u, nav - 1)]);
                    }
                    // branch-free: a possibly softened pair contributes 0 here and is listed for pass 2
                    bool nr[PP_ILP];
#pragma unroll
                    for (int u = 0; u < PP_ILP; ++u) {
                        double d[DIM];
                        grav_rij<DIM, PERIODIC>(P, ri, cr[u], d);
                        const double r2 = dot<DIM>(d, d);
                        const bool nx = r2 < cthr2;
                        nr[u] = nx && u < cn;
                        const double y = fast_rsqrt(nx ? 1.0 : r2);
                        const double gm = (nx || u >= cn) ? 0.0 : P.G * cr[u].w;
                        phi -= gm * y;
                        const double sc = gm * y * (y * y);
#pragma unroll
                        for (int a = 0; a < DIM; ++a) acc[a] -= d[a] * sc;
                    }
#pragma unroll
                    for (int u = 0; u < PP_ILP; ++u) { if (nr[u]) { nearq[nnear * 32] = cj + u; ++nnear; } }
                    if (COUNT) n_pp += cn;
                }
                for (int kk = 0; kk < nnear; ++kk) {
                    const int jn = nearq[kk * 32];
                    const double4 pj = ldg4(&posm[jn]);
                    const double einv_j = __ldg(&hsoft[jn]).x;
                    double d[DIM];
                    grav_rij<DIM, PERIODIC>(P, ri, pj, d);
                    const double r2 = dot<DIM>(d, d);
                    const double rinv = rsqrt(r2);              // inf at r == 0, unused there (u < 1 branch)
                    const double r = r2 > 0.0 ? r2 * rinv : 0.0;
                    double fi, gi, fj, gj;
                    soft_fg_fast(r, rinv, einv_i, fi, gi);
                    soft_fg_fast(r, rinv, einv_j, fj, gj);
                    const double gm = P.G * pj.w;
                    phi -= gm * (fi + fj) * 0.5;                // src/bhtree.cpp:314-315
                    const double s = gm * (gi + gj) * 0.5;
#pragma unroll
                    for (int a = 0; a < DIM; ++a) acc[a] -= d[a] * s;
                }
            } while (nav > 0);                                  // only if the softened-pair list ran full
            nlq = 0;
        }
        if (last) break;

        // ---- fetch the next batch now: its loads are in flight during the per-lane tests
        __syncwarp();
        {
            const int ne = min(top, 32);
            int2 ent = make_int2(0, 0);
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
                for (int ci = 0; ci < nc; ++ci) sm.expand[incl - nc + ci] = make_int2(c0 + ci, ent.y);
            }
            __syncwarp();
            node = -1;
            mask = 0;
            if (lane < k) {
                const int2 e = sm.expand[lane];
                node = e.x;
                mask = (unsigned)e.y;
                const double2 * q = t.ng + (size_t)node * 4;
                q0 = __ldg(q); q1 = __ldg(q + 1); q2 = __ldg(q + 2);
            }
            top -= m;
            __syncwarp();
        }
        // (f) mixed nodes: the reference's own per-particle test (src/bhtree.cpp:303-308).  Phase 1 is
        // straight-line (no votes, no branches: the loads and distance chains of successive nodes overlap)
        // and leaves each lane with its open / accept bits over the mixed list; two bit-matrix transposes
        // then give lane q the lanes that accept / open mixed node q.
        unsigned my_open = 0, my_acc = 0;
#pragma unroll 4
        for (int q = 0; q < nmix; ++q) {
            const double4 c4 = sm.mx[q];
            const double me2 = sm.me2[q];
            const unsigned mm = sm.mmask[q];
            double d[DIM];
            grav_rij<DIM, PERIODIC>(P, ri, c4, d);
            const double d2 = abs2_exact<DIM>(d);
            const bool in = (mm >> lane) & 1u;
            const bool op = in && me2 > __dmul_rn(P.theta2, d2);
            my_open |= (op ? 1u : 0u) << q;
            my_acc |= ((in && !op) ? 1u : 0u) << q;
        }
        if (nmix) {
            const unsigned acc_row = warp_transpose32(my_acc, lane);
            const unsigned open_row = warp_transpose32(my_open, lane);
            int4 info = make_int4(0, 0, 0, 0);
            int msl = 0;
            if (lane < nmix) { info = sm.minfo[lane]; msl = sm.msl[lane]; }
            // opened internal nodes: push the children with the mask of the lanes that opened
            const bool psh = lane < nmix && open_row != 0u && info.y != 0;
            const unsigned pb = __ballot_sync(SPHB_FULL_MASK, psh);
            if (pb) {
                const int total = __popc(pb);
                if (top + total > GV_STACK) {
                    if (lane == 0) atomicOr(&d_err[2], (unsigned long long)WALK_ERR_GRAV_STACK);
                } else {
                    if (psh) sm.stack[top + __popc(pb & lt_mask)] = make_int2(info.x | ((info.y - 1) << 29), (int)open_row);
                    top += total;
                }
            }
            // opened leaves: every lane appends the leaves it opened to its queue
            const unsigned leaf_bits = __ballot_sync(SPHB_FULL_MASK, lane < nmix && info.y == 0);
            unsigned lb = my_open & leaf_bits;
            while (lb) {
                const int q = __ffs(lb) - 1;
                lb &= lb - 1;
                const int4 inf = sm.minfo[q];
                lq[nlq * 32] = make_double2(pack_ints(inf.z, inf.w), fmax(h_i2, sm.mh2[q]));
                ++nlq;
            }
            __syncwarp();                                  // the lane masks in sm.mmask are dead now
            if (lane < nmix) sm.mmask[msl] = acc_row;      // accept row of mixed node `lane` at its chunk slot
        }
        // accept masks of the batch: rows compacted to the chunk slots, then node-major -> particle-major
        if (b_sel) {
            if (cls == 1 && !grp) sm.mmask[sslot] = racc;
            __syncwarp();
            const int nsel = __popc(b_sel);
            const unsigned row = lane < nsel ? sm.mmask[lane] : 0u;
            const unsigned long long tw = (unsigned long long)warp_transpose32(row, lane) << npb;
            pcw0 |= (unsigned)tw;
            pcw1 |= (unsigned)(tw >> 32);
            npb += nsel;
        }
        __syncwarp();
    }

    if (valid) {
#pragma unroll
        for (int a = 0; a < DIM; ++a) p.acc[a][i] = acc[a];
        p.phi[i] = phi;
        if (COUNT) { tot_pp += n_pp; tot_pc += n_pc; tot_visit += n_visit; tot_pcg += n_pcg; tot_ppg += n_ppg; }
    }
    __syncwarp();
    }
    if (COUNT) {
        const unsigned long long a = warp_sum_u64(tot_pp), b = warp_sum_u64(tot_pc), cc = warp_sum_u64(tot_visit);
        const unsigned long long pg = warp_sum_u64(tot_pcg), qg = warp_sum_u64(tot_ppg);
        if (lane == 0) { atomicAdd(&cnt->grav_pp, a); atomicAdd(&cnt->grav_pc, b); atomicAdd(&cnt->grav_node_visits, cc);
                         atomicAdd(&cnt->grav_pc_group, pg); atomicAdd(&cnt->grav_pp_group, qg); }
    }
}

// per-particle softening record of the gravity particle-particle loop: {2 / h, h^2}
__global__ void k_grav_pack(const double * __restrict__ sml, double2 * __restrict__ hsoft, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double h = sml[i];
    hsoft[i] = make_double2(2.0 / h, h * h * (1.0 + 1e-12));
}

// Direct sum, the EXHAUSTIVE_SEARCH flavour of GravityForce (src/gravity_force.cpp:70-84).  targets == nullptr: every
// particle; else the n_targets particles (sorted-order indices) listed there, against all n sources.
__global__ void k_select_targets(const int * __restrict__ orig, int n, int k, int * __restrict__ list, int * __restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && orig[i] < k) list[atomicAdd(count, 1)] = i;
}
template <int DIM>
__global__ void __launch_bounds__(128) k_gravity_direct(PSoA p, DevParams P, int n, const int * __restrict__ targets, int n_targets)
{
    __shared__ double sx[DIM][128], sm[128], sh[128];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < n_targets;
    const int i = valid ? (targets ? targets[t] : t) : 0;
    double ri[DIM], f[DIM], phi = 0.0, h_i = 1.0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) { ri[a] = 0.0; f[a] = 0.0; }
    if (valid) { load_vec<DIM>(p.pos, i, ri); h_i = p.sml[i]; }
    const double einv_i = 2.0 / h_i;
    for (int base = 0; base < n; base += 128) {
        const int j = base + threadIdx.x;
        __syncthreads();
        if (j < n) {
#pragma unroll
            for (int a = 0; a < DIM; ++a) sx[a][threadIdx.x] = p.pos[a][j];
            sm[threadIdx.x] = p.mass[j];
            sh[threadIdx.x] = p.sml[j];
        }
        __syncthreads();
        const int m = min(128, n - base);
        for (int k = 0; k < m; ++k) {
            double rj[DIM], d[DIM];
#pragma unroll
            for (int a = 0; a < DIM; ++a) rj[a] = sx[a][k];
            calc_r_ij<DIM>(P, ri, rj, d);
            const double r = sqrt(abs2_exact<DIM>(d));
            const double rinv = 1.0 / r;
            double fi, gi, fj, gj;
            soft_fg(r, rinv, einv_i, fi, gi);
            soft_fg(r, rinv, 2.0 / sh[k], fj, gj);
            const double gm = P.G * sm[k];
            phi -= gm * (fi + fj) * 0.5;
            const double s = gm * (gi + gj) * 0.5;
#pragma unroll
            for (int a = 0; a < DIM; ++a) f[a] -= d[a] * s;
        }
    }
    if (valid) {
#pragma unroll
        for (int a = 0; a < DIM; ++a) p.acc[a][i] += f[a];
        p.phi[i] = phi;
    }
}

// =================================================================================================
// TimeStep::calculation (src/timestep.cpp:18-38): dt[0] = min(cflSound * h_per_v_sig, cflForce * min sqrt(h/|a|))
// =================================================================================================
template <int DIM>
__global__ void k_timestep_partial(PSoA p, int i_begin, int i_end, double c_force, double * __restrict__ d_min)
{
    double m = 1.7976931348623157e308;
    for (int i = i_begin + blockIdx.x * blockDim.x + threadIdx.x; i < i_end; i += gridDim.x * blockDim.x) {
        double a[DIM];
        load_vec<DIM>(p.acc, i, a);
        const double acc_abs = sqrt(dot<DIM>(a, a));
        if (acc_abs > 0.0) {
            const double dt_force_i = c_force * sqrt(p.sml[i] / acc_abs);
            if (dt_force_i < m) m = dt_force_i;
        }
    }
    m = warp_min(m);
    if ((threadIdx.x & 31) == 0) atomic_min_pos(d_min, m);
}
__global__ void k_timestep_final(const double * __restrict__ d_min, const double * __restrict__ d_hpvs, double c_sound, double * __restrict__ d_dt)
{
    const double dt_sound = c_sound * d_hpvs[0];
    d_dt[0] = fmin(dt_sound, d_min[0]);
}

// =================================================================================================
// Solver::predict / correct (src/solver.cpp:431-474) and the post-IC state (392-404)
// =================================================================================================
template <int DIM>
__global__ void k_predict(PSoA p, DevParams P, int n, const double * __restrict__ d_dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dt = *d_dt;
    const double c_sound = P.gamma * (P.gamma - 1.0);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        const double v = p.vel[d][i], a = p.acc[d][i];
        const double vp = v + a * (0.5 * dt);
        p.vel_p[d][i] = vp;
        double x = p.pos[d][i] + vp * dt;
        p.vel[d][i] = v + a * dt;
        if (P.periodic) {                              // Periodic::apply, include/periodic.hpp:61-72
            if (x < P.rmin[d]) x += P.range[d];
            else if (x > P.rmax[d]) x -= P.range[d];
        }
        p.pos[d][i] = x;
    }
    const double u = p.ene[i], du = p.dene[i];
    p.ene_p[i] = u + du * (0.5 * dt);
    const double un = u + du * dt;
    p.ene[i] = un;
    p.sound[i] = sqrt(c_sound * un);
}

template <int DIM>
__global__ void k_correct(PSoA p, DevParams P, int n, const double * __restrict__ d_dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dt = *d_dt;
    const double c_sound = P.gamma * (P.gamma - 1.0);
#pragma unroll
    for (int d = 0; d < DIM; ++d) p.vel[d][i] = p.vel_p[d][i] + p.acc[d][i] * (0.5 * dt);
    const double un = p.ene_p[i] + p.dene[i] * (0.5 * dt);
    p.ene[i] = un;
    p.sound[i] = sqrt(c_sound * un);
}

__global__ void k_init_state(PSoA p, DevParams P, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p.alpha[i] = P.av_alpha;
    p.balsara[i] = 1.0;
    p.sound[i] = sqrt(P.gamma * (P.gamma - 1.0) * p.ene[i]);
}

// Output::output_energy sums, src/output.cpp:72-83: out = {kinetic, thermal, potential}
template <int DIM>
__global__ void k_energy(PSoA p, int n, double * __restrict__ out)
{
    double ek = 0.0, et = 0.0, ep = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double v[DIM];
        load_vec<DIM>(p.vel, i, v);
        const double m = p.mass[i];
        ek += 0.5 * m * dot<DIM>(v, v);
        et += m * p.ene[i];
        ep += 0.5 * m * p.phi[i];
    }
    ek = warp_sum(ek); et = warp_sum(et); ep = warp_sum(ep);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], ek); atomicAdd(&out[1], et); atomicAdd(&out[2], ep); }
}

// =================================================================================================
// test hook: explicit neighbour lists (sphb_neighbor_lists)
// =================================================================================================
template <int DIM>
struct ListV {
    const DevParams & P; const PSoA & p;
    double ri[DIM], h_i2;
    bool symmetric, fill;
    int cnt;
    int * out;           // fill: write p.orig[j] at out[cnt]
    long long cap_left;
    const double4 * posm;
    __device__ __forceinline__ void hit(int j)
    {
        const double4 pj = ldg4(&posm[j]);
        double d[DIM];
        rij_from4<DIM>(P, ri, pj, d);
        const double r2 = abs2_exact<DIM>(d);
        double k2 = h_i2;
        if (symmetric) { const double hj = p.sml[j]; k2 = fmax(h_i2, __dmul_rn(hj, hj)); }     // exhaustive_search.cpp:28
        if (r2 < k2) {
            if (fill && cnt < cap_left) out[cnt] = p.orig[j];
            ++cnt;
        }
    }
};

template <int DIM>
__global__ void __launch_bounds__(128)
k_neighbor_lists(PSoA p, Recs rc, TreeDev t, DevParams P, GroupTable gt, const double * __restrict__ h_override /* sorted order or null */,
                 int symmetric, int fill, int * __restrict__ counts, const long long * __restrict__ offsets,
                 int * __restrict__ ids, long long cap_total, unsigned long long * __restrict__ d_err)
{
    __shared__ NWalkSmem s_walk[4];
    const int lane = threadIdx.x & 31;
    int g_first, g_cnt;
    while (next_group(gt, lane, g_first, g_cnt)) {
    const int i = g_first + lane;
    const bool valid = lane < g_cnt;
    ListV<DIM> v{P, p};
    v.posm = rc.posm;
    v.symmetric = symmetric != 0; v.fill = fill != 0; v.cnt = 0; v.out = nullptr; v.cap_left = 0;
    double h_i = 1.0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) v.ri[a] = 0.0;
    if (valid) {
        load_vec<DIM>(p.pos, i, v.ri);
        h_i = h_override ? h_override[i] : p.sml[i];
        if (fill) {
            const long long o = offsets[i];
            v.out = ids + (o < cap_total ? o : 0);
            v.cap_left = o < cap_total ? cap_total - o : 0;
        }
    }
    v.h_i2 = __dmul_rn(h_i, h_i);
    if (symmetric) group_stream<DIM, true>(t, P, rc.posm, p.sml, 1, s_walk[threadIdx.x >> 5], lane, v.ri, h_i, valid, v, d_err);
    else group_stream<DIM, false>(t, P, rc.posm, nullptr, 0, s_walk[threadIdx.x >> 5], lane, v.ri, h_i, valid, v, d_err);
    if (valid && !fill) counts[i] = v.cnt;
    }
}

// FP64 FMA micro-benchmark (roofline denominator)
__global__ void k_fp64_peak(double * out, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int k = 0; k < iters; ++k) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

} // namespace sphb
