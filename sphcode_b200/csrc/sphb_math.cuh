// sphb_math.cuh — device numerics primitives of the hot path (sm_100a, FP64 CUDA cores).
//
// Restates, for the device, the semantics of
//   include/kernel/cubic_spline.hpp:21-52, include/kernel/wendland_kernel.hpp:23-49  (W, grad W, dW/dh)
//   include/periodic.hpp:34-72                                                         (minimum image, wrap)
//   src/bhtree.cpp:273-299 == src/gravity_force.cpp:16-42                             (softening f, g)
// of mitchiinaga/sphcode.  Nothing here is a dense contraction, so no tensor cores: the binding
// resource is the FP64 pipe (DFMA) and every per-pair division that does not depend on the pair
// is hoisted into a per-particle coefficient set (KernelCoef).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SPHB_FULL_MASK 0xffffffffu
#define SPHB_PI 3.14159265358979323846264338327950288

namespace sphb {

enum { K_CUBIC = 0, K_WENDLAND = 1 };
enum { T_SSPH = 0, T_DISPH = 1, T_GSPH = 2 };

// Device copy of sph::SPHParameters (include/parameters.hpp:20-79) + derived constants.
struct DevParams {
    int    sph_type, kernel;
    double cfl_sound, cfl_force;
    double av_alpha;
    int    use_balsara, use_tdav;
    double alpha_max, alpha_min, epsilon_av;
    int    use_ac;
    double alpha_ac;
    int    max_level, leaf_num, ngb, iterative;
    double gamma;
    int    periodic, use_gravity;
    double rmax[3], rmin[3], range[3];
    double G, theta, theta2;
    int    gsph2;
    double kernel_ratio;     // 1.2 if iterative else 1.0, src/pre_interaction.cpp:31-35
    int    key_levels;       // split levels encoded in the 64-bit key
    int    list_cap;         // neighbor_number * neighbor_list_size, include/defines.hpp:29
};

template <int DIM> struct Vec { double v[DIM]; };

template <int DIM> __device__ __forceinline__ double powh(double h)
{
    if (DIM == 1) return h;
    if (DIM == 2) return h * h;
    return h * h * h;
}
template <int DIM> __device__ __forceinline__ double powh_(double h)   // h^(DIM-1)
{
    if (DIM == 1) return 1.0;
    if (DIM == 2) return h;
    return h * h;
}
template <int DIM> __device__ __forceinline__ constexpr double unit_ball()   // A_d, src/pre_interaction.cpp:61-63
{
    return DIM == 1 ? 2.0 : DIM == 2 ? SPHB_PI : 4.0 * SPHB_PI / 3.0;
}

// |a|^2 with the reference's evaluation order and NO fma contraction (x86-64 build of the
// reference has no FMA): these values feed set-membership predicates (r2 < h2) that must be
// bit-exact (include/vector_type.hpp:227-230, include/defines.hpp:12-21).
template <int DIM> __device__ __forceinline__ double abs2_exact(const double (&a)[DIM])
{
    double s = __dmul_rn(a[0], a[0]);
    if (DIM >= 2) s = __dadd_rn(s, __dmul_rn(a[DIM >= 2 ? 1 : 0], a[DIM >= 2 ? 1 : 0]));
    if (DIM >= 3) s = __dadd_rn(s, __dmul_rn(a[DIM >= 3 ? 2 : 0], a[DIM >= 3 ? 2 : 0]));
    return s;
}

// Periodic::calc_r_ij for one axis (include/periodic.hpp:34-59): pick the smallest |.| among
// d, d+L, d-L, ties resolved in that order by <=.
__device__ __forceinline__ double min_image(double d1, double L)
{
    const double d2 = d1 + L;
    const double d3 = d1 - L;
    const double a1 = fabs(d1), a2 = fabs(d2), a3 = fabs(d3);
    if (a1 <= a2 && a1 <= a3) return d1;
    if (a2 <= a3 && a2 <= a1) return d2;
    return d3;
}

template <int DIM>
__device__ __forceinline__ void calc_r_ij(const DevParams & P, const double (&ri)[DIM], const double (&rj)[DIM], double (&out)[DIM])
{
#pragma unroll
    for (int d = 0; d < DIM; ++d) out[d] = ri[d] - rj[d];
    if (P.periodic) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) out[d] = min_image(out[d], P.range[d]);
    }
}

// ---- SPH kernel functions ----------------------------------------------------------------
// KernelCoef holds everything that depends on h only; w/dhw/dwc are then division-free.
// dwc is the scalar c of grad W = r_ij * c.
template <int DIM, int KT> struct KernelCoef;

template <int DIM> struct KernelCoef<DIM, K_CUBIC> {
    double qinv, cw, cdw, cdhw;
    __device__ __forceinline__ void init(double h)
    {
        // include/kernel/cubic_spline.hpp:13-19; one reciprocal, the rest are products
        const double sigma = DIM == 1 ? 2.0 / 3.0 : DIM == 2 ? 10.0 / (7.0 * SPHB_PI) : 1.0 / SPHB_PI;
        qinv = 2.0 / h;                                 // 1 / h_,  h_ = h / 2
        cw   = sigma * powh<DIM>(qinv);                 // sigma / h_^d
        cdw  = -sigma * (powh<DIM>(qinv) * qinv);       // -sigma / (h_^d h_), times 1/r per pair
        cdhw = 0.5 * sigma * (powh<DIM>(qinv) * qinv);
    }
    __device__ __forceinline__ double w(double r) const        // cubic_spline.hpp:27-32
    {
        const double q = r * qinv;
        const double a = 0.5 * (2.0 - q + fabs(2.0 - q));
        const double b = 0.5 * (1.0 - q + fabs(1.0 - q));
        return cw * (0.25 * (a * a * a) - b * b * b);
    }
    __device__ __forceinline__ double dwc(double r) const      // cubic_spline.hpp:34-43
    {
        if (r == 0.0) return 0.0;
        const double q = r * qinv;
        const double a = 0.5 * (2.0 - q + fabs(2.0 - q));
        const double b = 0.5 * (1.0 - q + fabs(1.0 - q));
        return cdw / r * (0.75 * (a * a) - 3.0 * (b * b));
    }
    __device__ __forceinline__ double dhw(double r) const      // cubic_spline.hpp:45-51
    {
        const double q = r * qinv;
        const double a = (fabs(2.0 - q) + 2.0 - q) * 0.5;
        const double b = (fabs(1.0 - q) + 1.0 - q) * 0.5;
        return cdhw * (a * a * ((3.0 + DIM) * 0.25 * q - 0.5 * DIM) + b * b * ((-3.0 - DIM) * q + DIM));
    }
};

template <int DIM> struct KernelCoef<DIM, K_WENDLAND> {
    double qinv, cw, cdw, cdhw;
    __device__ __forceinline__ void init(double h)
    {
        // include/kernel/wendland_kernel.hpp:14-20 (DIM == 1 is asserted out there; sigma 0 here)
        const double sigma = DIM == 1 ? 0.0 : DIM == 2 ? 9.0 / SPHB_PI : 495.0 / (32.0 * SPHB_PI);
        qinv = 1.0 / h;
        cw   = sigma * powh<DIM>(qinv);                                  // sigma / h^d
        cdw  = -56.0 / 3.0 * sigma * (powh<DIM>(qinv) * (qinv * qinv));  // / (h^d h^2)
        cdhw = -sigma / 3.0 * (powh<DIM>(qinv) * qinv);                  // / (h^d h 3)
    }
    __device__ __forceinline__ double w(double r) const        // wendland_kernel.hpp:30-34
    {
        const double q = r * qinv;
        const double a = 0.5 * (1.0 - q + fabs(1.0 - q));
        const double a2 = a * a;
        return cw * (a2 * a2 * a2) * (1.0 + 6.0 * q + 35.0 / 3.0 * q * q);
    }
    __device__ __forceinline__ double dwc(double r) const      // wendland_kernel.hpp:36-41
    {
        const double q = r * qinv;
        const double a = 0.5 * (1.0 - q + fabs(1.0 - q));
        const double a2 = a * a;
        return cdw * (a2 * a2 * a) * (1.0 + 5.0 * q);
    }
    __device__ __forceinline__ double dhw(double r) const      // wendland_kernel.hpp:43-48
    {
        const double q = r * qinv;
        const double a = 0.5 * (1.0 - q + fabs(1.0 - q));
        const double a2 = a * a;
        return cdhw * (a2 * a2 * a)
             * (3.0 * DIM + 15.0 * DIM * q + (-56.0 + 17.0 * DIM) * q * q - 35.0 * (8.0 + DIM) * (q * q * q));
    }
};

// ---- gravitational softening (Hernquist & Katz 1989), src/bhtree.cpp:273-299 -----------------
// rinv = 1/r (unused when u < 1, so r == 0 is safe), einv = 2/h.
__device__ __forceinline__ void soft_fg(double r, double rinv, double einv, double & f, double & g)
{
    const double u = r * einv;
    if (u < 1.0) {
        const double u2 = u * u;
        f = (-0.5 * u2 * (1.0 / 3.0 - 3.0 / 20 * u2 + u2 * u / 20) + 1.4) * einv;
        g = (4.0 / 3.0 - 1.2 * u2 + 0.5 * u2 * u) * (einv * einv * einv);
    } else if (u < 2.0) {
        const double u2 = u * u, u3 = u2 * u;
        f = -rinv / 15 + (-u2 * (4.0 / 3.0 - u + 0.3 * u2 - u3 / 30) + 1.6) * einv;
        g = (-1.0 / 15 + 8.0 / 3 * u3 - 3 * u3 * u + 1.2 * u3 * u2 - u3 * u3 / 6.0) * (rinv * rinv * rinv);
    } else {
        f = rinv;
        g = rinv * rinv * rinv;
    }
}

// 1/sqrt(x) for normal x > 0: hardware seed (MUFU.RSQ64H, relative error < 2^-22) + one polynomial step in e = 1 - x y^2.
// No special cases: x = 0 / inf / nan give nan — callers use it only on squared distances known to be positive.
//   SPHB_RSQRT_ORDER 3: y (1 + e/2 + 3 e^2/8), error ~ e^3: below 1 ulp;
//   SPHB_RSQRT_ORDER 2 (default): y (1 + e/2) (Newton), relative error -1.5 d^2 >= -8.6e-14 for a seed error d <= 2^-22 (2.6e-13 on a
//   force term, three orders below the 1e-10 parity bar): one FP64 instruction less per interaction, k_gravity 86.0 -> 83.7 ms at 16 M.
#ifndef SPHB_RSQRT_ORDER
#define SPHB_RSQRT_ORDER 2
#endif
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y * y, 1.0);
#if SPHB_RSQRT_ORDER == 2
    return fma(0.5 * y, e, y);
#else
    const double p = fma(0.375, e, 0.5);
    return fma(y, e * p, y);
#endif
}

// read-only 32-byte load: ONE 256-bit non-coherent load (sm_100a: LDG.E.ENL2.256.CONSTANT), i.e. one
// L1 wavefront per distinct record instead of the two of a pair of 16-byte loads.  p must be 32-byte
// aligned (the packed record arrays are).
__device__ __forceinline__ double4 ldg4(const double4 * p)
{
    double4 v;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

// ---- warp / atomic helpers -------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(SPHB_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(SPHB_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SPHB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SPHB_FULL_MASK, v, o);
    return v;
}
// 32 x 32 bit-matrix transpose across a warp: lane k passes row k, lane j receives column j
// (bit k of the result = bit j of lane k's input).  Five butterfly steps.
__device__ __forceinline__ unsigned warp_transpose32(unsigned x, int lane)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const unsigned m = s == 16 ? 0x0000ffffu : s == 8 ? 0x00ff00ffu : s == 4 ? 0x0f0f0f0fu : s == 2 ? 0x33333333u : 0x55555555u;
        const unsigned y = __shfl_xor_sync(SPHB_FULL_MASK, x, s);
        x = (lane & s) ? ((x & ~m) | ((y >> s) & m)) : ((x & m) | ((y & m) << s));
    }
    return x;
}
// Non-negative doubles order like their bit patterns: min / max through 64-bit integer atomics.
__device__ __forceinline__ void atomic_min_pos(double * addr, double v)
{
    atomicMin(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ unsigned long long atomic_max_pos(double * addr, double v)
{
    return atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

} // namespace sphb
