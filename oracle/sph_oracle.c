/*
 * TEST INFRASTRUCTURE — not part of the product path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * sph_oracle.c — plain-C restatement ("port") of the per-step hot path of mitchiinaga/sphcode, the
 * CPU checker that exists wherever gcc exists (the unmodified reference itself, oracle/_ref/*.so,
 * can only be built where /root/reference is mounted).  It follows the reference's own algorithm
 * and data structures — pointer Barnes-Hut tree with intrusive particle lists, recursive range
 * search + sort by r^2, recursive gravity walk — NOT the device design, and every function cites the
 * reference lines it restates.  DIM is a run-time value here (the reference compiles one DIM at a
 * time, include/defines.hpp:5).
 *
 * Parity pinned: tests/test_oracle_cpu.py checks this port against the unmodified reference
 * (oracle/_ref) on every sample / SPH type / kernel combination and against the golden vectors in
 * tests/golden/ (generated from the unmodified reference by tests/golden/make_golden.py).
 *
 * Exports the same entry points as oracle/ref_driver.cpp with the prefix spho_.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795028841971693993751
#endif

typedef double real;

/* mirror of ref_params in oracle/ref_driver.cpp (= SPHParameters, include/parameters.hpp:20-79) */
typedef struct {
    int    sph_type, kernel;
    double cfl_sound, cfl_force;
    double av_alpha;
    int    use_balsara, use_tdav;
    double alpha_max, alpha_min, epsilon_av;
    int    use_ac;
    double alpha_ac;
    int    max_tree_level, leaf_particle_num;
    int    neighbor_number;
    double gamma;
    int    iterative_sml;
    int    periodic;
    double range_max[3], range_min[3];
    int    use_gravity;
    double G, theta;
    int    gsph_2nd_order;
} spho_params;

/* include/particle.hpp:8-33, vectors padded to 3 */
typedef struct Particle {
    real pos[3], vel[3], vel_p[3], acc[3];
    real mass, dens, pres, ene, ene_p, dene, sml, sound, balsara, alpha, gradh, phi;
    int id, neighbor;
    struct Particle * next;
} Particle;

/* include/bhtree.hpp:17-31 */
typedef struct Node {
    Particle * first;
    real mass;
    int num;
    struct Node * childs[8];
    real center[3], m_center[3];
    real edge;
    int level;
    real kernel_size;
    int is_leaf;
} Node;

typedef struct {
    spho_params P;
    int dim, nchild, n;
    Particle * p;
    Node root;
    Node * nodes;
    int node_size;
    real range[3];
    real dt, time, h_per_v_sig;
    int first_pre;
    real * grad_d, * grad_p, * grad_v[3];    /* GSPH arrays, n*3 each */
    /* interaction counts of the reference algorithm (the numerators of the algorithmic-FLOP model):
     * [0] Newton kernel evaluations [1] Newton iterations [2] candidates r2 < h_search^2 [3] neighbours r < h_i
     * [4] force pairs [5] gravity particle-particle [6] gravity particle-cell [7] gravity node visits */
    unsigned long long cnt[8];
    char err[256];
} Ctx;

/* ---- include/periodic.hpp:34-72 ------------------------------------------------------------ */
static void calc_r_ij(const Ctx * c, const real * ri, const real * rj, real * out)
{
    for (int i = 0; i < c->dim; ++i) {
        const real d1 = ri[i] - rj[i];
        if (!c->P.periodic) { out[i] = d1; continue; }
        const real d2 = d1 + c->range[i], d3 = d1 - c->range[i];
        if (fabs(d1) <= fabs(d2) && fabs(d1) <= fabs(d3)) out[i] = d1;
        else if (fabs(d2) <= fabs(d3) && fabs(d2) <= fabs(d1)) out[i] = d2;
        else out[i] = d3;
    }
}
static void periodic_apply(const Ctx * c, real * r)
{
    if (!c->P.periodic) return;
    for (int i = 0; i < c->dim; ++i) {
        if (r[i] < c->P.range_min[i]) r[i] += c->range[i];
        else if (r[i] > c->P.range_max[i]) r[i] -= c->range[i];
    }
}
static real abs2v(const Ctx * c, const real * a)
{
    real s = a[0] * a[0];
    for (int i = 1; i < c->dim; ++i) s += a[i] * a[i];
    return s;
}
static real dotv(const Ctx * c, const real * a, const real * b)
{
    real s = a[0] * b[0];
    for (int i = 1; i < c->dim; ++i) s += a[i] * b[i];
    return s;
}

/* ---- kernels: include/kernel/cubic_spline.hpp:21-52, wendland_kernel.hpp:23-49 ------------- */
static real powh(int dim, real h) { return dim == 1 ? h : dim == 2 ? h * h : h * h * h; }
static real powh_(int dim, real h) { return dim == 1 ? 1 : dim == 2 ? h : h * h; }
static real pos_part(real x) { return 0.5 * (x + fabs(x)); }

static real sigma_cubic(int dim) { return dim == 1 ? 2.0 / 3.0 : dim == 2 ? 10.0 / (7.0 * M_PI) : 1.0 / M_PI; }
static real sigma_c4(int dim) { return dim == 1 ? 0.0 : dim == 2 ? 9.0 / M_PI : 495. / (32 * M_PI); }

static real kernel_w(const Ctx * c, real r, real h)
{
    const int dim = c->dim;
    if (c->P.kernel == 0) {
        const real h_ = h * 0.5, q = r / h_;
        const real a = pos_part(2.0 - q), b = pos_part(1.0 - q);
        return sigma_cubic(dim) / powh(dim, h_) * (0.25 * a * a * a - b * b * b);
    } else {
        const real q = r / h, a = pos_part(1.0 - q);
        return sigma_c4(dim) / powh(dim, h) * (a * a * a * a * a * a) * (1.0 + 6.0 * q + 35.0 / 3.0 * q * q);
    }
}
/* scalar c of grad W = r_ij * c */
static real kernel_dwc(const Ctx * c, real r, real h)
{
    const int dim = c->dim;
    if (c->P.kernel == 0) {
        if (r == 0.0) return 0.0;
        const real h_ = h * 0.5, q = r / h_;
        const real a = pos_part(2.0 - q), b = pos_part(1.0 - q);
        return -sigma_cubic(dim) / (powh(dim, h_) * h_ * r) * (0.75 * a * a - 3.0 * b * b);
    } else {
        const real q = r / h, a = pos_part(1.0 - q);
        return -56.0 / 3.0 * sigma_c4(dim) / (powh(dim, h) * h * h) * (a * a * a * a * a) * (1.0 + 5.0 * q);
    }
}
static real kernel_dhw(const Ctx * c, real r, real h)
{
    const int dim = c->dim;
    if (c->P.kernel == 0) {
        const real h_ = h * 0.5, q = r / h_;
        const real a = pos_part(2.0 - q), b = pos_part(1.0 - q);
        return 0.5 * sigma_cubic(dim) / (powh(dim, h_) * h_) * (a * a * ((3. + dim) * 0.25 * q - 0.5 * dim) + b * b * ((-3.0 - dim) * q + dim));
    } else {
        const real q = r / h, a = pos_part(1.0 - q);
        return -sigma_c4(dim) / (powh(dim, h) * h * 3.0) * (a * a * a * a * a)
             * (3.0 * dim + 15.0 * dim * q + (-56.0 + 17.0 * dim) * q * q - 35.0 * (8.0 + dim) * q * q * q);
    }
}

/* ---- BHTree: src/bhtree.cpp ---------------------------------------------------------------- */
static void node_clear(Node * nd)
{
    memset(nd, 0, sizeof(*nd));
}

/* BHNode::assign, src/bhtree.cpp:163-204 */
static int node_assign(Ctx * c, Node * self, Particle * pt, Node ** pool, int * remaind)
{
    int index = 0, mask = 1;
    for (int i = 0; i < c->dim; ++i) {
        if (pt->pos[i] > self->center[i]) index |= mask;
        mask <<= 1;
    }
    Node * child = self->childs[index];
    if (!child) {
        if (*remaind < 0) { snprintf(c->err, sizeof(c->err), "There is no free node."); return 1; }
        self->childs[index] = *pool;
        child = *pool;
        ++*pool;
        --*remaind;
        node_clear(child);
        child->level = self->level + 1;
        child->edge = self->edge * 0.5;
        int a = 1;
        real b = 2.0;
        for (int i = 0; i < c->dim; ++i) {
            child->center[i] = self->center[i] + ((index & a) * b - 1.0) * self->edge * 0.25;
            a <<= 1;
            b *= 0.5;
        }
    }
    child->num++;
    child->mass += pt->mass;
    for (int i = 0; i < c->dim; ++i) child->m_center[i] += pt->pos[i] * pt->mass;
    pt->next = child->first;
    child->first = pt;
    return 0;
}

/* BHNode::create_tree, src/bhtree.cpp:136-161 */
static int node_create_tree(Ctx * c, Node * self, Node ** pool, int * remaind)
{
    for (int i = 0; i < 8; ++i) self->childs[i] = NULL;
    Particle * pp = self->first;
    do {
        Particle * pnext = pp->next;
        if (node_assign(c, self, pp, pool, remaind)) return 1;
        pp = pnext;
    } while (pp != NULL);
    for (int i = 0; i < c->nchild; ++i) {
        Node * child = self->childs[i];
        if (child) {
            for (int k = 0; k < c->dim; ++k) child->m_center[k] /= child->mass;
            if (child->num > c->P.leaf_particle_num && self->level < c->P.max_tree_level) {
                if (node_create_tree(c, child, pool, remaind)) return 1;
            } else {
                child->is_leaf = 1;
            }
        }
    }
    return 0;
}

/* BHTree::make, src/bhtree.cpp:55-107 (+ initialize 12-40 for the periodic root) */
static int tree_make(Ctx * c)
{
    Node * root = &c->root;
    root->first = NULL; root->mass = 0.0; root->num = 0;                 /* root_clear, include/bhtree.hpp:46-53 */
    for (int i = 0; i < 3; ++i) root->m_center[i] = 0.0;
    root->kernel_size = 0.0; root->is_leaf = 0;
    root->level = 1;
    if (c->P.periodic) {
        real l = 0.0;
        for (int i = 0; i < c->dim; ++i) {
            root->center[i] = (c->P.range_max[i] + c->P.range_min[i]) * 0.5;
            const real range = c->P.range_max[i] - c->P.range_min[i];
            if (l < range) l = range;
        }
        root->edge = l;
    } else {
        real rmin[3], rmax[3];
        for (int i = 0; i < c->dim; ++i) { rmin[i] = DBL_MAX; rmax[i] = -DBL_MAX; }
        for (int k = 0; k < c->n; ++k)
            for (int i = 0; i < c->dim; ++i) {
                if (rmin[i] > c->p[k].pos[i]) rmin[i] = c->p[k].pos[i];
                if (rmax[i] < c->p[k].pos[i]) rmax[i] = c->p[k].pos[i];
            }
        real l = 0.0;
        for (int i = 0; i < c->dim; ++i) {
            root->center[i] = (rmax[i] + rmin[i]) * 0.5;
            const real range = rmax[i] - rmin[i];
            if (l < range) l = range;
        }
        root->edge = l;
    }
    for (int k = 0; k < c->n - 1; ++k) c->p[k].next = &c->p[k + 1];
    c->p[c->n - 1].next = NULL;
    root->first = &c->p[0];
    int remaind = c->node_size;
    Node * pool = c->nodes;
    return node_create_tree(c, root, &pool, &remaind);
}

/* BHNode::set_kernel, src/bhtree.cpp:206-232 */
static real node_set_kernel(Ctx * c, Node * nd)
{
    real kernel = 0.0;
    if (nd->is_leaf) {
        for (Particle * p = nd->first; p; p = p->next) if (p->sml > kernel) kernel = p->sml;
    } else {
        for (int i = 0; i < c->nchild; ++i)
            if (nd->childs[i]) {
                const real h = node_set_kernel(c, nd->childs[i]);
                if (h > kernel) kernel = h;
            }
    }
    nd->kernel_size = kernel;
    return kernel;
}

/* BHNode::neighbor_search, src/bhtree.cpp:234-270 */
static void node_neighbor_search(const Ctx * c, const Node * nd, const real * r_i, real sml_i, int * list, int * n, int cap, int is_ij)
{
    const real h = is_ij ? (sml_i > nd->kernel_size ? sml_i : nd->kernel_size) : sml_i;
    const real h2 = h * h;
    const real l2 = (nd->edge * 0.5 + h) * (nd->edge * 0.5 + h);
    real d[3];
    calc_r_ij(c, r_i, nd->center, d);
    real dx2_max = d[0] * d[0];
    for (int i = 1; i < c->dim; ++i) {
        const real dx2 = d[i] * d[i];
        if (dx2 > dx2_max) dx2_max = dx2;
    }
    if (dx2_max <= l2) {
        if (nd->is_leaf) {
            for (const Particle * p = nd->first; p; p = p->next) {
                real r_ij[3];
                calc_r_ij(c, r_i, p->pos, r_ij);
                const real r2 = abs2v(c, r_ij);
                if (r2 < h2) {
                    if (*n < cap) list[*n] = (int)(p - c->p);     /* the reference stores p->id == index */
                    ++*n;
                }
            }
        } else {
            for (int i = 0; i < c->nchild; ++i)
                if (nd->childs[i]) node_neighbor_search(c, nd->childs[i], r_i, sml_i, list, n, cap, is_ij);
        }
    }
}

typedef struct { real r2; int j; } SortItem;
static int cmp_item(const void * a, const void * b)
{
    const real x = ((const SortItem *)a)->r2, y = ((const SortItem *)b)->r2;
    return x < y ? -1 : x > y ? 1 : 0;
}

/* BHTree::neighbor_search, src/bhtree.cpp:114-126 (search + sort by r^2), or exhaustive_search
 * (src/exhaustive_search.cpp:11-42) when exhaustive != 0.  Returns the count (may exceed cap). */
static int neighbor_search(const Ctx * c, const real * r_i, real sml_i, int * list, SortItem * tmp, int cap, int is_ij, int exhaustive)
{
    int n = 0;
    if (exhaustive) {
        const real k2i = sml_i * sml_i;
        for (int j = 0; j < c->n; ++j) {
            real r_ij[3];
            calc_r_ij(c, r_i, c->p[j].pos, r_ij);
            const real r2 = abs2v(c, r_ij);
            const real hj2 = c->p[j].sml * c->p[j].sml;
            const real k2 = is_ij ? (k2i > hj2 ? k2i : hj2) : k2i;
            if (r2 < k2) { if (n < cap) list[n] = j; ++n; }
        }
    } else {
        node_neighbor_search(c, &c->root, r_i, sml_i, list, &n, cap, is_ij);
    }
    const int m = n < cap ? n : cap;
    for (int k = 0; k < m; ++k) {
        real r_ij[3];
        calc_r_ij(c, r_i, c->p[list[k]].pos, r_ij);
        tmp[k].r2 = abs2v(c, r_ij);
        tmp[k].j = list[k];
    }
    qsort(tmp, m, sizeof(SortItem), cmp_item);
    for (int k = 0; k < m; ++k) list[k] = tmp[k].j;
    return n;
}

/* Hernquist & Katz (1989), src/bhtree.cpp:273-299 */
static real soft_f(real r, real h)
{
    const real e = h * 0.5, u = r / e;
    if (u < 1.0) return (-0.5 * u * u * (1.0 / 3.0 - 3.0 / 20 * u * u + u * u * u / 20) + 1.4) / e;
    else if (u < 2.0) return -1.0 / (15 * r) + (-u * u * (4.0 / 3.0 - u + 0.3 * u * u - u * u * u / 30) + 1.6) / e;
    else return 1 / r;
}
static real soft_g(real r, real h)
{
    const real e = h * 0.5, u = r / e;
    if (u < 1.0) return (4.0 / 3.0 - 1.2 * u * u + 0.5 * u * u * u) / (e * e * e);
    else if (u < 2.0) return (-1.0 / 15 + 8.0 / 3 * u * u * u - 3 * u * u * u * u + 1.2 * u * u * u * u * u - u * u * u * u * u * u / 6.0) / (r * r * r);
    else return 1 / (r * r * r);
}

/* BHNode::calc_force, src/bhtree.cpp:301-331 */
static void node_calc_force(const Ctx * c, const Node * nd, Particle * p_i, real theta2, unsigned long long * k)
{
    ++k[2];                                             /* node visit = one opening test */
    const real l2 = nd->edge * nd->edge;
    real d[3];
    calc_r_ij(c, p_i->pos, nd->m_center, d);
    const real d2 = abs2v(c, d);
    if (l2 > theta2 * d2) {
        if (nd->is_leaf) {
            for (const Particle * p = nd->first; p; p = p->next) {
                real r_ij[3];
                ++k[0];
                calc_r_ij(c, p_i->pos, p->pos, r_ij);
                const real r = sqrt(abs2v(c, r_ij));
                p_i->phi -= c->P.G * p->mass * (soft_f(r, p_i->sml) + soft_f(r, p->sml)) * 0.5;
                const real s = c->P.G * p->mass * (soft_g(r, p_i->sml) + soft_g(r, p->sml)) * 0.5;
                for (int k = 0; k < c->dim; ++k) p_i->acc[k] -= r_ij[k] * s;
            }
        } else {
            for (int i = 0; i < c->nchild; ++i)
                if (nd->childs[i]) node_calc_force(c, nd->childs[i], p_i, theta2, k);
        }
    } else {
        ++k[1];
        const real r_inv = 1.0 / sqrt(d2);
        p_i->phi -= c->P.G * nd->mass * r_inv;
        const real s = c->P.G * nd->mass * r_inv * r_inv * r_inv;
        for (int k = 0; k < c->dim; ++k) p_i->acc[k] -= d[k] * s;
    }
}

/* ---- PreInteraction ------------------------------------------------------------------------ */
static real unit_ball(int dim) { return dim == 1 ? 2.0 : dim == 2 ? M_PI : 4.0 * M_PI / 3.0; }

/* newton_raphson, src/pre_interaction.cpp:227-283 and src/disph/d_pre_interaction.cpp:174-230 */
static real newton_raphson(const Ctx * c, const Particle * p_i, const int * list, int n_neighbor, real kernel_ratio, int * nonconv,
                           unsigned long long * evals, unsigned long long * iters)
{
    const int dim = c->dim, disph = c->P.sph_type == 1;
    real h_i = p_i->sml / kernel_ratio;
    const real A = unit_ball(dim);
    const real b = disph ? c->P.neighbor_number / A : p_i->mass * c->P.neighbor_number / A;
    for (int it = 0; it < 10; ++it) {
        const real h_b = h_i;
        real dens = 0.0, ddens = 0.0;
        for (int n = 0; n < n_neighbor; ++n) {
            const Particle * p_j = &c->p[list[n]];
            real r_ij[3];
            calc_r_ij(c, p_i->pos, p_j->pos, r_ij);
            const real r = sqrt(abs2v(c, r_ij));
            if (r >= h_i) break;
            ++*evals;
            if (disph) { dens += kernel_w(c, r, h_i); ddens += kernel_dhw(c, r, h_i); }
            else { dens += p_j->mass * kernel_w(c, r, h_i); ddens += p_j->mass * kernel_dhw(c, r, h_i); }
        }
        ++*iters;
        const real f = dens * powh(dim, h_i) - b;
        const real df = ddens * powh(dim, h_i) + dim * dens * powh_(dim, h_i);
        h_i -= f / df;
        if (fabs(h_i - h_b) < (h_i + h_b) * 1e-4) return h_i;
    }
    ++*nonconv;
    return p_i->sml / kernel_ratio;
}

/* initial_smoothing, src/pre_interaction.cpp:171-215 */
static void initial_smoothing(Ctx * c, int exhaustive)
{
    const int cap = c->P.neighbor_number * 20 > c->n + 16 ? c->P.neighbor_number * 20 : c->n + 16;
#pragma omp parallel
    {
        int * list = (int *)malloc(sizeof(int) * cap);
        SortItem * tmp = (SortItem *)malloc(sizeof(SortItem) * cap);
#pragma omp for
        for (int i = 0; i < c->n; ++i) {
            Particle * p_i = &c->p[i];
            p_i->sml = pow(c->P.neighbor_number * p_i->mass / (p_i->dens * unit_ball(c->dim)), 1.0 / c->dim);
            const int nn = neighbor_search(c, p_i->pos, p_i->sml, list, tmp, cap, 0, exhaustive);
            real dens_i = 0.0;
            for (int n = 0; n < nn; ++n) {
                const Particle * p_j = &c->p[list[n]];
                real r_ij[3];
                calc_r_ij(c, p_i->pos, p_j->pos, r_ij);
                const real r = sqrt(abs2v(c, r_ij));
                if (r >= p_i->sml) break;
                dens_i += p_j->mass * kernel_w(c, r, p_i->sml);
            }
            p_i->dens = dens_i;      /* reads of dens by other threads: only p_i's own dens is read */
        }
        free(list); free(tmp);
    }
}

/* PreInteraction::calculation for SSPH (src/pre_interaction.cpp:39-169), DISPH
 * (src/disph/d_pre_interaction.cpp:21-162) and GSPH (src/gsph/g_pre_interaction.cpp:27-144) */
static int pre_interaction(Ctx * c, int exhaustive)
{
    const int dim = c->dim, type = c->P.sph_type;
    if (c->first_pre) { initial_smoothing(c, exhaustive); c->first_pre = 0; }
    const real kernel_ratio = c->P.iterative_sml ? 1.2 : 1.0;
    const real dt = c->dt, gamma = c->P.gamma;
    const int cap = c->P.neighbor_number * 20 > c->n + 16 ? c->P.neighbor_number * 20 : c->n + 16;
    real hpvs_min = DBL_MAX;
    int nonconv_total = 0;
#pragma omp parallel
    {
        int * list = (int *)malloc(sizeof(int) * cap);
        SortItem * tmp = (SortItem *)malloc(sizeof(SortItem) * cap);
        real hpvs_local = DBL_MAX;
        int nonconv = 0;
        unsigned long long k_ev = 0, k_it = 0, k_cand = 0, k_ngb = 0;
#pragma omp for
        for (int i = 0; i < c->n; ++i) {
            Particle * p_i = &c->p[i];
            p_i->sml = pow(c->P.neighbor_number * p_i->mass / (p_i->dens * unit_ball(dim)), 1.0 / dim) * kernel_ratio;
            const int n_tmp = neighbor_search(c, p_i->pos, p_i->sml, list, tmp, cap, 0, exhaustive);
            k_cand += (unsigned long long)(n_tmp < cap ? n_tmp : cap);
            if (c->P.iterative_sml) p_i->sml = newton_raphson(c, p_i, list, n_tmp, kernel_ratio, &nonconv, &k_ev, &k_it);

            real dens_i = 0.0, dh_dens_i = 0.0, pres_i = 0.0, dh_pres_i = 0.0, n_i = 0.0, dh_n_i = 0.0;
            real v_sig_max = p_i->sound * 2.0;
            int n_neighbor = 0;
            for (int n = 0; n < n_tmp; ++n) {
                const int j = list[n];
                const Particle * p_j = &c->p[j];
                real r_ij[3], v_ij[3];
                calc_r_ij(c, p_i->pos, p_j->pos, r_ij);
                const real r = sqrt(abs2v(c, r_ij));
                if (r >= p_i->sml) break;
                ++n_neighbor;
                ++k_ngb;
                const real w_ij = kernel_w(c, r, p_i->sml);
                dens_i += p_j->mass * w_ij;
                if (type == 0) {
                    dh_dens_i += p_j->mass * kernel_dhw(c, r, p_i->sml);
                } else if (type == 1) {
                    const real dhw_ij = kernel_dhw(c, r, p_i->sml);
                    n_i += w_ij;
                    pres_i += p_j->mass * p_j->ene * w_ij;
                    dh_pres_i += p_j->mass * p_j->ene * dhw_ij;
                    dh_n_i += dhw_ij;
                }
                if (i != j) {
                    for (int k = 0; k < dim; ++k) v_ij[k] = p_i->vel[k] - p_j->vel[k];
                    const real v_sig = p_i->sound + p_j->sound - 3.0 * dotv(c, r_ij, v_ij) / r;
                    if (v_sig > v_sig_max) v_sig_max = v_sig;
                }
            }
            p_i->dens = dens_i;
            if (type == 0) {
                p_i->pres = (gamma - 1.0) * dens_i * p_i->ene;
                p_i->gradh = 1.0 / (1.0 + p_i->sml / (dim * dens_i) * dh_dens_i);
            } else if (type == 1) {
                p_i->pres = (gamma - 1.0) * pres_i;
                p_i->gradh = p_i->sml / (dim * n_i) * dh_pres_i / (1.0 + p_i->sml / (dim * n_i) * dh_n_i);
            } else {
                p_i->pres = (gamma - 1.0) * dens_i * p_i->ene;
            }
            p_i->neighbor = n_neighbor;
            const real hpvs_i = p_i->sml / v_sig_max;
            if (hpvs_local > hpvs_i) hpvs_local = hpvs_i;

            if (type == 2) {
                /* MUSCL gradients, src/gsph/g_pre_interaction.cpp:112-136 */
                if (!c->P.gsph_2nd_order) continue;
                real dd[3] = {0, 0, 0}, du[3] = {0, 0, 0}, dv[3][3] = {{0}};
                for (int n = 0; n < n_neighbor; ++n) {
                    const Particle * p_j = &c->p[list[n]];
                    real r_ij[3];
                    calc_r_ij(c, p_i->pos, p_j->pos, r_ij);
                    const real r = sqrt(abs2v(c, r_ij));
                    const real cw = kernel_dwc(c, r, p_i->sml);
                    for (int a = 0; a < dim; ++a) {
                        const real dw = r_ij[a] * cw;
                        dd[a] += dw * p_j->mass;
                        du[a] += dw * (p_j->mass * (p_j->ene - p_i->ene));
                        for (int k = 0; k < dim; ++k) dv[k][a] += dw * (p_j->mass * (p_j->vel[k] - p_i->vel[k]));
                    }
                }
                const real rho_inv = 1.0 / p_i->dens;
                for (int a = 0; a < dim; ++a) {
                    c->grad_d[i * 3 + a] = dd[a];
                    c->grad_p[i * 3 + a] = (dd[a] * p_i->ene + du[a]) * (gamma - 1.0);
                    for (int k = 0; k < dim; ++k) c->grad_v[k][i * 3 + a] = dv[k][a] * rho_inv;
                }
                continue;
            }

            /* Balsara switch and time dependent alpha, src/pre_interaction.cpp:116-161 */
            const int use_balsara = c->P.use_balsara && dim != 1;
            if (use_balsara || c->P.use_tdav) {
                real div_v = 0.0, rot_v[3] = {0, 0, 0};
                for (int n = 0; n < n_neighbor; ++n) {
                    const Particle * p_j = &c->p[list[n]];
                    real r_ij[3], dw[3] = {0, 0, 0}, v_ij[3] = {0, 0, 0};
                    calc_r_ij(c, p_i->pos, p_j->pos, r_ij);
                    const real r = sqrt(abs2v(c, r_ij));
                    const real cw = kernel_dwc(c, r, p_i->sml);
                    for (int k = 0; k < dim; ++k) { dw[k] = r_ij[k] * cw; v_ij[k] = p_i->vel[k] - p_j->vel[k]; }
                    const real wgt = type == 1 ? p_j->mass * p_j->ene : p_j->mass;
                    div_v -= wgt * dotv(c, v_ij, dw);
                    if (use_balsara) {
                        if (dim == 2) rot_v[0] += (v_ij[0] * dw[1] - v_ij[1] * dw[0]) * wgt;
                        else {
                            rot_v[0] += (v_ij[1] * dw[2] - v_ij[2] * dw[1]) * wgt;
                            rot_v[1] += (v_ij[2] * dw[0] - v_ij[0] * dw[2]) * wgt;
                            rot_v[2] += (v_ij[0] * dw[1] - v_ij[1] * dw[0]) * wgt;
                        }
                    }
                }
                if (type == 0) { div_v /= p_i->dens; for (int k = 0; k < 3; ++k) rot_v[k] /= p_i->dens; }
                else { const real p_inv = (gamma - 1.0) / p_i->pres; div_v *= p_inv; for (int k = 0; k < 3; ++k) rot_v[k] *= p_inv; }
                if (use_balsara) {
                    const real rot_abs = dim == 2 ? fabs(rot_v[0]) : sqrt(rot_v[0] * rot_v[0] + rot_v[1] * rot_v[1] + rot_v[2] * rot_v[2]);
                    p_i->balsara = fabs(div_v) / (fabs(div_v) + rot_abs + 1e-4 * p_i->sound / p_i->sml);
                    if (c->P.use_tdav) {
                        const real tau_inv = c->P.epsilon_av * p_i->sound / p_i->sml;
                        const real dalpha = (-(p_i->alpha - c->P.alpha_min) * tau_inv + fmax(-div_v, 0.0) * (c->P.alpha_max - p_i->alpha)) * dt;
                        p_i->alpha += dalpha;
                    }
                } else {
                    const real tau_inv = c->P.epsilon_av * p_i->sound / p_i->sml;
                    const real s_i = fmax(-div_v, 0.0);
                    p_i->alpha = (p_i->alpha + dt * tau_inv * c->P.alpha_min + s_i * dt * c->P.alpha_max) / (1.0 + dt * tau_inv + s_i * dt);
                }
            }
        }
#pragma omp critical
        {
            if (hpvs_local < hpvs_min) hpvs_min = hpvs_local;
            nonconv_total += nonconv;
            c->cnt[0] += k_ev; c->cnt[1] += k_it; c->cnt[2] += k_cand; c->cnt[3] += k_ngb;
        }
        free(list); free(tmp);
    }
    c->h_per_v_sig = hpvs_min;
    if (!exhaustive) node_set_kernel(c, &c->root);
    (void)nonconv_total;
    return 0;
}

/* ---- FluidForce ---------------------------------------------------------------------------- */
/* src/fluid_force.cpp:89-106 */
static real artificial_viscosity(const Ctx * c, const Particle * p_i, const Particle * p_j, const real * r_ij)
{
    real v_ij[3];
    for (int k = 0; k < c->dim; ++k) v_ij[k] = p_i->vel[k] - p_j->vel[k];
    const real vr = dotv(c, v_ij, r_ij);
    if (vr < 0) {
        const real alpha = 0.5 * (p_i->alpha + p_j->alpha);
        const real balsara = 0.5 * (p_i->balsara + p_j->balsara);
        const real w_ij = vr / sqrt(abs2v(c, r_ij));
        const real v_sig = p_i->sound + p_j->sound - 3.0 * w_ij;
        const real rho_ij_inv = 2.0 / (p_i->dens + p_j->dens);
        return -0.5 * balsara * alpha * v_sig * w_ij * rho_ij_inv;
    }
    return 0;
}
/* src/fluid_force.cpp:108-116 */
static real artificial_conductivity(const Ctx * c, const Particle * p_i, const Particle * p_j, const real * r_ij, const real * dw_ij)
{
    real v_ij[3];
    for (int k = 0; k < c->dim; ++k) v_ij[k] = p_i->vel[k] - p_j->vel[k];
    const real r = sqrt(abs2v(c, r_ij));
    const real v_sig = c->P.use_gravity ? fabs(dotv(c, v_ij, r_ij) / r) : sqrt(2.0 * fabs(p_i->pres - p_j->pres) / (p_i->dens + p_j->dens));
    return c->P.alpha_ac * p_j->mass * v_sig * (p_i->ene - p_j->ene) * dotv(c, dw_ij, r_ij) / r;
}
/* van Leer (1979), src/gsph/g_fluid_force.cpp:28-36 */
static real limiter(real dq1, real dq2)
{
    const real q = dq1 * dq2;
    return q <= 0 ? 0.0 : 2.0 * q / (dq1 + dq2);
}
/* HLL, src/gsph/g_fluid_force.cpp:168-199 */
static void hll_solver(const real * left, const real * right, real * pstar, real * vstar)
{
    const real u_l = left[0], rho_l = left[1], p_l = left[2], c_l = left[3];
    const real u_r = right[0], rho_r = right[1], p_r = right[2], c_r = right[3];
    const real roe_l = sqrt(rho_l), roe_r = sqrt(rho_r), roe_inv = 1.0 / (roe_l + roe_r);
    const real u_t = (roe_l * u_l + roe_r * u_r) * roe_inv;
    const real c_t = (roe_l * c_l + roe_r * c_r) * roe_inv;
    const real s_l = fmin(u_l - c_l, u_t - c_t), s_r = fmax(u_r + c_r, u_t + c_t);
    const real c1 = rho_l * (s_l - u_l), c2 = rho_r * (s_r - u_r), c3 = 1.0 / (c1 - c2);
    const real c4 = p_l - u_l * c1, c5 = p_r - u_r * c2;
    *vstar = (c5 - c4) * c3;
    *pstar = (c1 * c5 - c2 * c4) * c3;
}

/* FluidForce::calculation: src/fluid_force.cpp:26-86, src/disph/d_fluid_force.cpp:26-86,
 * src/gsph/g_fluid_force.cpp:39-166 */
static int fluid_force(Ctx * c, int exhaustive)
{
    const int dim = c->dim, type = c->P.sph_type;
    const real gamma = c->P.gamma, dt = c->dt;
    const int cap = c->P.neighbor_number * 20 > c->n + 16 ? c->P.neighbor_number * 20 : c->n + 16;
#pragma omp parallel
    {
        int * list = (int *)malloc(sizeof(int) * cap);
        SortItem * tmp = (SortItem *)malloc(sizeof(SortItem) * cap);
        unsigned long long k_pairs = 0;
#pragma omp for
        for (int i = 0; i < c->n; ++i) {
            Particle * p_i = &c->p[i];
            const int nn = neighbor_search(c, p_i->pos, p_i->sml, list, tmp, cap, 1, exhaustive);
            const real h_i = p_i->sml, gradh_i = p_i->gradh;
            const real p_per_rho2_i = p_i->pres / (p_i->dens * p_i->dens);
            const real gamma2_u_i = (gamma - 1.0) * (gamma - 1.0) * p_i->ene;
            const real gamma2_u_per_pres_i = gamma2_u_i / p_i->pres;
            const real m_u_inv = 1.0 / (p_i->mass * p_i->ene);
            const real rho2_inv_i = 1.0 / (p_i->dens * p_i->dens);
            real acc[3] = {0, 0, 0}, dene = 0.0;
            for (int n = 0; n < nn; ++n) {
                const int j = list[n];
                const Particle * p_j = &c->p[j];
                real r_ij[3];
                calc_r_ij(c, p_i->pos, p_j->pos, r_ij);
                const real r = sqrt(abs2v(c, r_ij));
                if (r >= (h_i > p_j->sml ? h_i : p_j->sml) || r == 0.0) continue;
                ++k_pairs;
                const real cwi = kernel_dwc(c, r, h_i), cwj = kernel_dwc(c, r, p_j->sml);
                real dw_i[3], dw_j[3], dw_ij[3], v_ij[3];
                for (int k = 0; k < dim; ++k) {
                    dw_i[k] = r_ij[k] * cwi; dw_j[k] = r_ij[k] * cwj;
                    dw_ij[k] = (dw_i[k] + dw_j[k]) * 0.5;
                    v_ij[k] = p_i->vel[k] - p_j->vel[k];
                }
                if (type == 2) {
                    const real r_inv = 1.0 / r;
                    real e_ij[3];
                    for (int k = 0; k < dim; ++k) e_ij[k] = r_ij[k] * r_inv;
                    const real ve_i = dotv(c, p_i->vel, e_ij), ve_j = dotv(c, p_j->vel, e_ij);
                    real vstar, pstar;
                    if (c->P.gsph_2nd_order) {
                        real right[4], left[4];
                        const real delta_i = 0.5 * (1.0 - p_i->sound * dt * r_inv);
                        const real delta_j = 0.5 * (1.0 - p_j->sound * dt * r_inv);
                        const real dv_ij = ve_i - ve_j;
                        real dv_i[3], dv_j[3];
                        for (int k = 0; k < dim; ++k) {
                            dv_i[k] = dotv(c, &c->grad_v[k][i * 3], e_ij);
                            dv_j[k] = dotv(c, &c->grad_v[k][j * 3], e_ij);
                        }
                        const real dve_i = dotv(c, dv_i, e_ij) * r, dve_j = dotv(c, dv_j, e_ij) * r;
                        right[0] = ve_i - limiter(dv_ij, dve_i) * delta_i;
                        left[0] = ve_j + limiter(dv_ij, dve_j) * delta_j;
                        const real dd_ij = p_i->dens - p_j->dens;
                        const real dd_i = dotv(c, &c->grad_d[i * 3], e_ij) * r, dd_j = dotv(c, &c->grad_d[j * 3], e_ij) * r;
                        right[1] = p_i->dens - limiter(dd_ij, dd_i) * delta_i;
                        left[1] = p_j->dens + limiter(dd_ij, dd_j) * delta_j;
                        const real dp_ij = p_i->pres - p_j->pres;
                        const real dp_i = dotv(c, &c->grad_p[i * 3], e_ij) * r, dp_j = dotv(c, &c->grad_p[j * 3], e_ij) * r;
                        right[2] = p_i->pres - limiter(dp_ij, dp_i) * delta_i;
                        left[2] = p_j->pres + limiter(dp_ij, dp_j) * delta_j;
                        right[3] = sqrt(gamma * right[2] / right[1]);
                        left[3] = sqrt(gamma * left[2] / left[1]);
                        hll_solver(left, right, &pstar, &vstar);
                    } else {
                        const real right[4] = {ve_i, p_i->dens, p_i->pres, p_i->sound};
                        const real left[4] = {ve_j, p_j->dens, p_j->pres, p_j->sound};
                        hll_solver(left, right, &pstar, &vstar);
                    }
                    const real rho2_inv_j = 1.0 / (p_j->dens * p_j->dens);
                    real fdotv = 0.0;
                    for (int k = 0; k < dim; ++k) {
                        const real f = dw_i[k] * (p_j->mass * pstar * rho2_inv_i) + dw_j[k] * (p_j->mass * pstar * rho2_inv_j);
                        acc[k] -= f;
                        fdotv += f * (e_ij[k] * vstar - p_i->vel[k]);
                    }
                    dene -= fdotv;
                    continue;
                }
                const real pi_ij = artificial_viscosity(c, p_i, p_j, r_ij);
                const real dene_ac = c->P.use_ac ? artificial_conductivity(c, p_i, p_j, r_ij, dw_ij) : 0.0;
                if (type == 0) {
                    for (int k = 0; k < dim; ++k)
                        acc[k] -= dw_i[k] * (p_j->mass * (p_per_rho2_i * gradh_i + 0.5 * pi_ij))
                                + dw_j[k] * (p_j->mass * (p_j->pres / (p_j->dens * p_j->dens) * p_j->gradh + 0.5 * pi_ij));
                    dene += p_j->mass * p_per_rho2_i * gradh_i * dotv(c, v_ij, dw_i) + 0.5 * p_j->mass * pi_ij * dotv(c, v_ij, dw_ij) + dene_ac;
                } else {
                    const real f_ij = 1.0 - gradh_i / (p_j->mass * p_j->ene);
                    const real f_ji = 1.0 - p_j->gradh * m_u_inv;
                    const real u_per_pres_j = p_j->ene / p_j->pres;
                    for (int k = 0; k < dim; ++k)
                        acc[k] -= dw_i[k] * (p_j->mass * (gamma2_u_per_pres_i * p_j->ene * f_ij + 0.5 * pi_ij))
                                + dw_j[k] * (p_j->mass * (gamma2_u_i * u_per_pres_j * f_ji + 0.5 * pi_ij));
                    dene += p_j->mass * gamma2_u_per_pres_i * p_j->ene * f_ij * dotv(c, v_ij, dw_i)
                          + 0.5 * p_j->mass * pi_ij * dotv(c, v_ij, dw_ij) + dene_ac;
                }
            }
            for (int k = 0; k < dim; ++k) p_i->acc[k] = acc[k];
            p_i->dene = dene;
        }
#pragma omp atomic
        c->cnt[4] += k_pairs;
        free(list); free(tmp);
    }
    return 0;
}

/* GravityForce::calculation, src/gravity_force.cpp:52-89 */
static int gravity_force(Ctx * c, int exhaustive)
{
    if (!c->P.use_gravity) return 0;
    const real theta2 = c->P.theta * c->P.theta;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < c->n; ++i) {
        Particle * p_i = &c->p[i];
        if (exhaustive) {
            real phi = 0.0, force[3] = {0, 0, 0};
            for (int j = 0; j < c->n; ++j) {
                const Particle * p_j = &c->p[j];
                real r_ij[3];
                calc_r_ij(c, p_i->pos, p_j->pos, r_ij);
                const real r = sqrt(abs2v(c, r_ij));
                phi -= c->P.G * p_j->mass * (soft_f(r, p_i->sml) + soft_f(r, p_j->sml)) * 0.5;
                const real s = c->P.G * p_j->mass * (soft_g(r, p_i->sml) + soft_g(r, p_j->sml)) * 0.5;
                for (int k = 0; k < c->dim; ++k) force[k] -= r_ij[k] * s;
            }
            for (int k = 0; k < c->dim; ++k) p_i->acc[k] += force[k];
            p_i->phi = phi;
        } else {
            unsigned long long k[3] = {0, 0, 0};
            p_i->phi = 0.0;                                            /* BHTree::tree_force, src/bhtree.cpp:128-132 */
            node_calc_force(c, &c->root, p_i, theta2, k);
#pragma omp atomic
            c->cnt[5] += k[0];
#pragma omp atomic
            c->cnt[6] += k[1];
#pragma omp atomic
            c->cnt[7] += k[2];
        }
    }
    return 0;
}

/* TimeStep::calculation, src/timestep.cpp:18-38 */
static void timestep(Ctx * c)
{
    real dt_min = DBL_MAX;
    for (int i = 0; i < c->n; ++i) {
        const real acc_abs = sqrt(abs2v(c, c->p[i].acc));
        if (acc_abs > 0.0) {
            const real dt_force_i = c->P.cfl_force * sqrt(c->p[i].sml / acc_abs);
            if (dt_force_i < dt_min) dt_min = dt_force_i;
        }
    }
    const real dt_sound = c->P.cfl_sound * c->h_per_v_sig;
    c->dt = dt_sound < dt_min ? dt_sound : dt_min;
}

/* ============================================================================================ */
/* C ABI (same surface as oracle/ref_driver.cpp)                                                 */
static size_t rec_size(int dim) { return (size_t)(4 * dim + 12) * 8 + 16; }

static void unpack(const Ctx * c, const void * aos)
{
    const int D = c->dim;
    const size_t rec = rec_size(D);
    for (int i = 0; i < c->n; ++i) {
        const double * r = (const double *)((const char *)aos + rec * i);
        Particle * p = &c->p[i];
        memset(p, 0, sizeof(*p));
        for (int d = 0; d < D; ++d) { p->pos[d] = r[d]; p->vel[d] = r[D + d]; p->vel_p[d] = r[2 * D + d]; p->acc[d] = r[3 * D + d]; }
        const double * q = r + 4 * D;
        p->mass = q[0]; p->dens = q[1]; p->pres = q[2]; p->ene = q[3]; p->ene_p = q[4]; p->dene = q[5];
        p->sml = q[6]; p->sound = q[7]; p->balsara = q[8]; p->alpha = q[9]; p->gradh = q[10]; p->phi = q[11];
        const int * iq = (const int *)(q + 12);
        p->id = iq[0]; p->neighbor = iq[1];
    }
}
static void pack(const Ctx * c, void * aos)
{
    const int D = c->dim;
    const size_t rec = rec_size(D);
    for (int i = 0; i < c->n; ++i) {
        double * r = (double *)((char *)aos + rec * i);
        const Particle * p = &c->p[i];
        for (int d = 0; d < D; ++d) { r[d] = p->pos[d]; r[D + d] = p->vel[d]; r[2 * D + d] = p->vel_p[d]; r[3 * D + d] = p->acc[d]; }
        double * q = r + 4 * D;
        q[0] = p->mass; q[1] = p->dens; q[2] = p->pres; q[3] = p->ene; q[4] = p->ene_p; q[5] = p->dene;
        q[6] = p->sml; q[7] = p->sound; q[8] = p->balsara; q[9] = p->alpha; q[10] = p->gradh; q[11] = p->phi;
        int * iq = (int *)(q + 12);
        iq[0] = p->id; iq[1] = p->neighbor;
        q[13] = 0.0;
    }
}

int spho_dim(void) { return 0; }                 /* run-time DIM */
int spho_sizeof_particle(void) { return 0; }
void spho_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
    (void)n;
}
int spho_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void * spho_create(const spho_params * P, int dim, int n, const void * particles)
{
    Ctx * c = (Ctx *)calloc(1, sizeof(Ctx));
    c->P = *P; c->dim = dim; c->nchild = 1 << dim; c->n = n;
    c->p = (Particle *)calloc((size_t)n, sizeof(Particle));
    c->node_size = n * 5;                                           /* BHTree::resize, src/bhtree.cpp:42-53 */
    c->nodes = (Node *)calloc((size_t)c->node_size + 1, sizeof(Node));
    for (int i = 0; i < 3; ++i) c->range[i] = i < dim ? P->range_max[i] - P->range_min[i] : 0.0;
    c->first_pre = 1;
    if (P->sph_type == 2) {
        c->grad_d = (real *)calloc((size_t)n * 3, sizeof(real));
        c->grad_p = (real *)calloc((size_t)n * 3, sizeof(real));
        for (int k = 0; k < 3; ++k) c->grad_v[k] = (real *)calloc((size_t)n * 3, sizeof(real));
    }
    if (P->kernel == 1 && dim == 1) snprintf(c->err, sizeof(c->err), "Wendland C4 is not defined for DIM == 1");
    unpack(c, particles);
    return c;
}
void spho_destroy(void * v)
{
    Ctx * c = (Ctx *)v;
    if (!c) return;
    free(c->p); free(c->nodes); free(c->grad_d); free(c->grad_p);
    for (int k = 0; k < 3; ++k) free(c->grad_v[k]);
    free(c);
}
const char * spho_error(void * v) { return ((Ctx *)v)->err; }
void spho_get_particles(void * v, void * out) { pack((Ctx *)v, out); }
void spho_set_particles(void * v, const void * in) { unpack((Ctx *)v, in); }

/* Solver::initialize, src/solver.cpp:392-404 */
void spho_init_state(void * v)
{
    Ctx * c = (Ctx *)v;
    const real c_sound = c->P.gamma * (c->P.gamma - 1.0);
    for (int i = 0; i < c->n; ++i) {
        c->p[i].alpha = c->P.av_alpha;
        c->p[i].balsara = 1.0;
        c->p[i].sound = sqrt(c_sound * c->p[i].ene);
    }
}
int spho_make_tree(void * v) { return tree_make((Ctx *)v); }
int spho_pre(void * v) { return pre_interaction((Ctx *)v, 0); }
int spho_fluid(void * v) { return fluid_force((Ctx *)v, 0); }
int spho_gravity(void * v) { return gravity_force((Ctx *)v, 0); }
int spho_timestep(void * v) { timestep((Ctx *)v); return 0; }
double spho_get_dt(void * v) { return ((Ctx *)v)->dt; }
void spho_set_dt(void * v, double dt) { ((Ctx *)v)->dt = dt; }
double spho_get_h_per_v_sig(void * v) { return ((Ctx *)v)->h_per_v_sig; }
void spho_set_h_per_v_sig(void * v, double x) { ((Ctx *)v)->h_per_v_sig = x; }

/* Solver::predict, src/solver.cpp:431-456 */
void spho_predict(void * v)
{
    Ctx * c = (Ctx *)v;
    const real dt = c->dt, c_sound = c->P.gamma * (c->P.gamma - 1.0);
    for (int i = 0; i < c->n; ++i) {
        Particle * p = &c->p[i];
        for (int k = 0; k < c->dim; ++k) {
            p->vel_p[k] = p->vel[k] + p->acc[k] * (0.5 * dt);
            p->pos[k] += p->vel_p[k] * dt;
            p->vel[k] += p->acc[k] * dt;
        }
        p->ene_p = p->ene + p->dene * (0.5 * dt);
        p->ene += p->dene * dt;
        p->sound = sqrt(c_sound * p->ene);
        periodic_apply(c, p->pos);
    }
}
/* Solver::correct, src/solver.cpp:458-474 */
void spho_correct(void * v)
{
    Ctx * c = (Ctx *)v;
    const real dt = c->dt, c_sound = c->P.gamma * (c->P.gamma - 1.0);
    for (int i = 0; i < c->n; ++i) {
        Particle * p = &c->p[i];
        for (int k = 0; k < c->dim; ++k) p->vel[k] = p->vel_p[k] + p->acc[k] * (0.5 * dt);
        p->ene = p->ene_p + p->dene * (0.5 * dt);
        p->sound = sqrt(c_sound * p->ene);
    }
}
/* Solver::initialize after the IC, src/solver.cpp:392-414 */
int spho_initialize(void * v)
{
    spho_init_state(v);
    return spho_make_tree(v) || spho_pre(v) || spho_fluid(v) || spho_gravity(v);
}
/* Solver::integrate, src/solver.cpp:417-429 */
int spho_integrate(void * v)
{
    Ctx * c = (Ctx *)v;
    timestep(c);
    spho_predict(v);
    if (spho_make_tree(v) || spho_pre(v) || spho_fluid(v) || spho_gravity(v)) return 1;
    spho_correct(v);
    c->time += c->dt;
    return 0;
}
/* Output::output_energy sums, src/output.cpp:72-83 */
void spho_energy(void * v, double * out)
{
    Ctx * c = (Ctx *)v;
    real ek = 0, et = 0, ep = 0;
    for (int i = 0; i < c->n; ++i) {
        const Particle * p = &c->p[i];
        ek += 0.5 * p->mass * abs2v(c, p->vel);
        et += p->mass * p->ene;
        ep += 0.5 * p->mass * p->phi;
    }
    out[0] = ek; out[1] = et; out[2] = ep;
}

/* all neighbour lists (CSR); is_ij & 2 selects exhaustive_search semantics
 * (src/exhaustive_search.cpp:11-42), else the tree search (src/bhtree.cpp:114-126) */
long long spho_neighbor_search_all(void * v, const double * h, int is_ij, long long * offsets, int * ids, long long cap_total)
{
    Ctx * c = (Ctx *)v;
    const int exhaustive = (is_ij & 2) != 0, sym = (is_ij & 1) != 0;
    const int cap = c->n + 16;
    int * list = (int *)malloc(sizeof(int) * cap);
    SortItem * tmp = (SortItem *)malloc(sizeof(SortItem) * cap);
    long long tot = 0;
    for (int i = 0; i < c->n; ++i) {
        offsets[i] = tot;
        const real hi = h ? h[i] : c->p[i].sml;
        const int n = neighbor_search(c, c->p[i].pos, hi, list, tmp, cap, sym, exhaustive);
        for (int k = 0; k < n; ++k) { if (tot < cap_total) ids[tot] = list[k]; ++tot; }
    }
    offsets[c->n] = tot;
    free(list); free(tmp);
    return tot;
}

/* interaction counts since the last reset (see Ctx::cnt) */
int spho_counters(void * v, unsigned long long * out, int reset)
{
    Ctx * c = (Ctx *)v;
    for (int k = 0; k < 8; ++k) { out[k] = c->cnt[k]; if (reset) c->cnt[k] = 0; }
    return 0;
}

int spho_get_vector_array(void * v, const char * name, double * out)
{
    Ctx * c = (Ctx *)v;
    const real * src = NULL;
    if (c->P.sph_type == 2) {
        if (!strcmp(name, "grad_density")) src = c->grad_d;
        else if (!strcmp(name, "grad_pressure")) src = c->grad_p;
        else if (!strncmp(name, "grad_velocity_", 14) && name[14] >= '0' && name[14] < '0' + c->dim && !name[15]) src = c->grad_v[name[14] - '0'];
    }
    if (!src) { snprintf(c->err, sizeof(c->err), "additional_vector_array does not have %s", name); return 1; }
    for (int i = 0; i < c->n; ++i) for (int k = 0; k < c->dim; ++k) out[i * c->dim + k] = src[i * 3 + k];
    return 0;
}

/* out = {w, dhw, dw[0..dim)} */
void spho_kernel_eval(void * v, const double * rij, double h, double * out)
{
    Ctx * c = (Ctx *)v;
    const real r = sqrt(abs2v(c, rij));
    out[0] = kernel_w(c, r, h);
    out[1] = kernel_dhw(c, r, h);
    const real cw = kernel_dwc(c, r, h);
    for (int k = 0; k < c->dim; ++k) out[2 + k] = rij[k] * cw;
}
