// TEST INFRASTRUCTURE — not part of the product path.
//
// C-ABI driver around the UNMODIFIED reference translation units
// (/root/reference/src/{bhtree,exhaustive_search,pre_interaction,fluid_force,
// gravity_force,timestep,simulation}.cpp + disph/* + gsph/*), compiled where
// they lie by oracle/Makefile into oracle/_ref/libsphref_d{1,2,3}[_ex].so.
// Nothing from the reference is copied: this file only *calls* the reference's
// Module classes the way Solver does (src/solver.cpp:353-474).
//
// Pieces restated here because src/solver.cpp, src/output.cpp and
// src/logger.cpp need Boost (absent in this image):
//   * Logger statics                         (src/logger.cpp:19-50)
//   * Solver::initialize post-IC sequence    (src/solver.cpp:387-414)
//   * Solver::integrate / predict / correct  (src/solver.cpp:417-474)
//   * Output::output_energy sums             (src/output.cpp:66-90)
#include <cstring>
#include <memory>
#include <vector>
#include <string>
#include <cmath>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "parameters.hpp"
#include "particle.hpp"
#include "simulation.hpp"
#include "periodic.hpp"
#include "bhtree.hpp"
#include "logger.hpp"
#include "exception.hpp"
#include "module.hpp"
#include "timestep.hpp"
#include "pre_interaction.hpp"
#include "fluid_force.hpp"
#include "gravity_force.hpp"
#include "disph/d_pre_interaction.hpp"
#include "disph/d_fluid_force.hpp"
#include "gsph/g_pre_interaction.hpp"
#include "gsph/g_fluid_force.hpp"
#include "exhaustive_search.hpp"
#include "kernel/kernel_function.hpp"
#ifdef SPHB_GPU_MODULES
// Same driver, but the module slots are filled with the sph::gpu drop-ins
// (sphcode_b200/host/gpu_modules.hpp): the integration test of the plugin boundary.
#include "gpu_modules.hpp"
#endif

// ---- Logger statics (stand-in for src/logger.cpp, which needs boost::format) ----
namespace sph {
std::string Logger::dir_name;
std::ofstream Logger::log_io;
bool Logger::open_flag = false;
void Logger::open(const std::string & d) { open(d.c_str()); }
void Logger::open(const char * d) {
    dir_name = d;
    log_io.open("/dev/null");
    open_flag = true;
}
}

using namespace sph;

extern "C" {

// Plain-C mirror of sph::SPHParameters (include/parameters.hpp:20-79).
struct ref_params {
    int    sph_type;            // 0 ssph, 1 disph, 2 gsph
    int    kernel;              // 0 cubic spline, 1 wendland
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
};

struct ref_ctx {
    std::shared_ptr<SPHParameters> param;
    std::shared_ptr<Simulation>    sim;
    std::shared_ptr<Module> timestep, pre, fforce, gforce;
    bool tree_sized = false;
    int n_full = 0;            // particles in the vector / in the tree (>= Simulation::particle_num, see ref_set_active)
    std::string err;
};

int ref_dim() { return DIM; }
int ref_sizeof_particle() { return (int)sizeof(SPHParticle); }
int ref_is_exhaustive() {
#ifdef EXHAUSTIVE_SEARCH
    return 1;
#else
    return 0;
#endif
}
void ref_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
    (void)n;
}
int ref_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static std::shared_ptr<SPHParameters> to_param(const ref_params * p)
{
    auto q = std::make_shared<SPHParameters>();
    std::memset(q.get(), 0, sizeof(SPHParameters));
    q->time.start = 0; q->time.end = 1; q->time.output = 1; q->time.energy = 1;
    q->type = p->sph_type == 0 ? SPHType::SSPH : p->sph_type == 1 ? SPHType::DISPH : SPHType::GSPH;
    q->cfl.sound = p->cfl_sound; q->cfl.force = p->cfl_force;
    q->av.alpha = p->av_alpha;
    q->av.use_balsara_switch = p->use_balsara;
    q->av.use_time_dependent_av = p->use_tdav;
    q->av.alpha_max = p->alpha_max; q->av.alpha_min = p->alpha_min; q->av.epsilon = p->epsilon_av;
    q->ac.is_valid = p->use_ac; q->ac.alpha = p->alpha_ac;
    q->tree.max_level = p->max_tree_level; q->tree.leaf_particle_num = p->leaf_particle_num;
    q->physics.neighbor_number = p->neighbor_number; q->physics.gamma = p->gamma;
    q->kernel = p->kernel == 0 ? KernelType::CUBIC_SPLINE : KernelType::WENDLAND;
    q->iterative_sml = p->iterative_sml;
    q->periodic.is_valid = p->periodic;
    for(int i = 0; i < DIM; ++i) {
        q->periodic.range_max[i] = p->range_max[i];
        q->periodic.range_min[i] = p->range_min[i];
    }
    q->gravity.is_valid = p->use_gravity; q->gravity.constant = p->G; q->gravity.theta = p->theta;
    q->gsph.is_2nd_order = p->gsph_2nd_order;
    return q;
}

// particles: n records with the exact in-memory layout of sph::SPHParticle for this DIM.
ref_ctx * ref_create(const ref_params * p, int n, const void * particles)
{
    if(!Logger::is_open()) Logger::open("/tmp");
    auto * c = new ref_ctx;
    try {
        c->param = to_param(p);
        c->sim = std::make_shared<Simulation>(c->param);
        std::vector<SPHParticle> v(n);
        std::memcpy((void*)v.data(), particles, sizeof(SPHParticle) * (size_t)n);
        for(auto & q : v) q.next = nullptr;
        c->sim->set_particles(v);
        c->sim->set_particle_num(n);
        c->n_full = n;

        // module selection: src/solver.cpp:359-370
#ifdef SPHB_GPU_MODULES
        c->timestep = std::make_shared<gpu::TimeStep>();
        if(c->param->type == SPHType::SSPH) {
            c->pre = std::make_shared<gpu::PreInteraction>();
            c->fforce = std::make_shared<gpu::FluidForce>();
        } else if(c->param->type == SPHType::DISPH) {
            c->pre = std::make_shared<gpu::disph::PreInteraction>();
            c->fforce = std::make_shared<gpu::disph::FluidForce>();
        } else {
            c->pre = std::make_shared<gpu::gsph::PreInteraction>();
            c->fforce = std::make_shared<gpu::gsph::FluidForce>();
#else
        c->timestep = std::make_shared<TimeStep>();
        if(c->param->type == SPHType::SSPH) {
            c->pre = std::make_shared<PreInteraction>();
            c->fforce = std::make_shared<FluidForce>();
        } else if(c->param->type == SPHType::DISPH) {
            c->pre = std::make_shared<disph::PreInteraction>();
            c->fforce = std::make_shared<disph::FluidForce>();
        } else {
            c->pre = std::make_shared<gsph::PreInteraction>();
            c->fforce = std::make_shared<gsph::FluidForce>();
#endif
            // src/solver.cpp:373-385
            std::vector<std::string> names = {"grad_density", "grad_pressure", "grad_velocity_0"};
#if DIM >= 2
            names.push_back("grad_velocity_1");
#endif
#if DIM == 3
            names.push_back("grad_velocity_2");
#endif
            c->sim->add_vector_array(names);
        }
#ifdef SPHB_GPU_MODULES
        c->gforce = std::make_shared<gpu::GravityForce>();
#else
        c->gforce = std::make_shared<GravityForce>();
#endif
        c->timestep->initialize(c->param);
        c->pre->initialize(c->param);
        c->fforce->initialize(c->param);
        c->gforce->initialize(c->param);
    } catch(std::exception & e) {
        c->err = e.what();
    }
    return c;
}

void ref_destroy(ref_ctx * c)
{
#ifdef SPHB_GPU_MODULES
    if(c && c->sim) gpu::release(c->sim.get());
#endif
    delete c;
}
const char * ref_error(ref_ctx * c) { return c->err.c_str(); }

void ref_get_particles(ref_ctx * c, void * out)
{
    auto & v = c->sim->get_particles();
    std::memcpy(out, (void*)v.data(), sizeof(SPHParticle) * v.size());
}

void ref_set_particles(ref_ctx * c, const void * in)
{
    auto & v = c->sim->get_particles();
    std::memcpy((void*)v.data(), in, sizeof(SPHParticle) * v.size());
    for(auto & q : v) q.next = nullptr;
}

// src/solver.cpp:392-404
void ref_init_state(ref_ctx * c)
{
    auto & p = c->sim->get_particles();
    const int num = c->sim->get_particle_num();
    const real gamma = c->param->physics.gamma;
    const real c_sound = gamma * (gamma - 1.0);
    const real alpha = c->param->av.alpha;
#pragma omp parallel for
    for(int i = 0; i < num; ++i) {
        p[i].alpha = alpha;
        p[i].balsara = 1.0;
        p[i].sound = std::sqrt(c_sound * p[i].ene);
    }
}

int ref_make_tree(ref_ctx * c)
{
#ifndef EXHAUSTIVE_SEARCH
    try {
        auto tree = c->sim->get_tree();
        if(!c->tree_sized) {
            tree->resize(c->n_full);
            c->tree_sized = true;
        }
        tree->make(c->sim->get_particles(), c->n_full);
    } catch(std::exception & e) { c->err = e.what(); return 1; }
#endif
    return 0;
}

#define GUARD(stmt) try { stmt; } catch(std::exception & e) { c->err = e.what(); return 1; } return 0

int ref_pre(ref_ctx * c)      { GUARD(c->pre->calculation(c->sim)); }
int ref_fluid(ref_ctx * c)    { GUARD(c->fforce->calculation(c->sim)); }
int ref_gravity(ref_ctx * c)  { GUARD(c->gforce->calculation(c->sim)); }
int ref_timestep(ref_ctx * c) { GUARD(c->timestep->calculation(c->sim)); }

double ref_get_dt(ref_ctx * c) { return c->sim->get_dt(); }
void   ref_set_dt(ref_ctx * c, double dt) { c->sim->set_dt(dt); }
double ref_get_time(ref_ctx * c) { return c->sim->get_time(); }
double ref_get_h_per_v_sig(ref_ctx * c) { return c->sim->get_h_per_v_sig(); }
void   ref_set_h_per_v_sig(ref_ctx * c, double v) { c->sim->set_h_per_v_sig(v); }

// Subsample mode (parity checks at sizes where the whole reference pass would take minutes): the modules loop
// `for i < sim->get_particle_num()` (src/pre_interaction.cpp:48,55; src/fluid_force.cpp; src/gravity_force.cpp:59,66)
// while BHTree::make / neighbor_search / tree_force work on whatever was handed to make().  With the tree made over
// ALL n_full particles and particle_num lowered to k, the unmodified modules compute particles 0..k-1 against all
// n_full sources.  k = 0 restores n_full.
void ref_set_active(ref_ctx * c, int k) { c->sim->set_particle_num(k > 0 && k < c->n_full ? k : c->n_full); }
// BHTree::set_kernel (src/bhtree.cpp:109-112), which PreInteraction calls last (src/pre_interaction.cpp:167)
int ref_set_kernel(ref_ctx * c)
{
#ifndef EXHAUSTIVE_SEARCH
    try { c->sim->get_tree()->set_kernel(); } catch(std::exception & e) { c->err = e.what(); return 1; }
#endif
    return 0;
}
// Direct gravity sum of the EXHAUSTIVE_SEARCH GravityForce (src/gravity_force.cpp:70-84) for the targets 0..k-1
// against all n_full sources, through the module's own softening functions f and g (src/gravity_force.cpp:16-42
// are file-static there, so the two are restated from those lines): out = k * (DIM + 1) doubles {force, phi}.
static inline real dsum_f(const real r, const real h)
{
    const real e = h * 0.5, u = r / e;
    if(u < 1.0) return (-0.5 * u * u * (1.0 / 3.0 - 3.0 / 20 * u * u + u * u * u / 20) + 1.4) / e;
    if(u < 2.0) return -1.0 / (15 * r) + (-u * u * (4.0 / 3.0 - u + 0.3 * u * u - u * u * u / 30) + 1.6) / e;
    return 1 / r;
}
static inline real dsum_g(const real r, const real h)
{
    const real e = h * 0.5, u = r / e;
    if(u < 1.0) return (4.0 / 3.0 - 1.2 * u * u + 0.5 * u * u * u) / (e * e * e);
    if(u < 2.0) return (-1.0 / 15 + 8.0 / 3 * u * u * u - 3 * u * u * u * u + 1.2 * u * u * u * u * u - u * u * u * u * u * u / 6.0) / (r * r * r);
    return 1 / (r * r * r);
}
void ref_direct_gravity(ref_ctx * c, int k, double * out)
{
    auto & particles = c->sim->get_particles();
    auto * periodic = c->sim->get_periodic().get();
    const int n = c->n_full;
    const real G = c->param->gravity.constant;
#pragma omp parallel for schedule(dynamic, 16)
    for(int i = 0; i < k; ++i) {
        const auto & p_i = particles[i];
        real phi = 0.0;
        vec_t force(0.0);
        for(int j = 0; j < n; ++j) {
            const auto & p_j = particles[j];
            const vec_t r_ij = periodic->calc_r_ij(p_i.pos, p_j.pos);
            const real r = std::abs(r_ij);
            phi -= G * p_j.mass * (dsum_f(r, p_i.sml) + dsum_f(r, p_j.sml)) * 0.5;
            force -= r_ij * (G * p_j.mass * (dsum_g(r, p_i.sml) + dsum_g(r, p_j.sml)) * 0.5);
        }
        for(int d = 0; d < DIM; ++d) out[(size_t)i * (DIM + 1) + d] = force[d];
        out[(size_t)i * (DIM + 1) + DIM] = phi;
    }
}

// Solver::predict, src/solver.cpp:431-456
void ref_predict(ref_ctx * c)
{
    auto & p = c->sim->get_particles();
    const int num = c->sim->get_particle_num();
    auto * periodic = c->sim->get_periodic().get();
    const real dt = c->sim->get_dt();
    const real gamma = c->param->physics.gamma;
    const real c_sound = gamma * (gamma - 1.0);
#pragma omp parallel for
    for(int i = 0; i < num; ++i) {
        p[i].vel_p = p[i].vel + p[i].acc * (0.5 * dt);
        p[i].ene_p = p[i].ene + p[i].dene * (0.5 * dt);
        p[i].pos += p[i].vel_p * dt;
        p[i].vel += p[i].acc * dt;
        p[i].ene += p[i].dene * dt;
        p[i].sound = std::sqrt(c_sound * p[i].ene);
        periodic->apply(p[i].pos);
    }
}

// Solver::correct, src/solver.cpp:458-474
void ref_correct(ref_ctx * c)
{
    auto & p = c->sim->get_particles();
    const int num = c->sim->get_particle_num();
    const real dt = c->sim->get_dt();
    const real gamma = c->param->physics.gamma;
    const real c_sound = gamma * (gamma - 1.0);
#pragma omp parallel for
    for(int i = 0; i < num; ++i) {
        p[i].vel = p[i].vel_p + p[i].acc * (0.5 * dt);
        p[i].ene = p[i].ene_p + p[i].dene * (0.5 * dt);
        p[i].sound = std::sqrt(c_sound * p[i].ene);
    }
}

// Solver::initialize after the IC, src/solver.cpp:392-414
int ref_initialize(ref_ctx * c)
{
    ref_init_state(c);
    if(ref_make_tree(c)) return 1;
    if(ref_pre(c)) return 1;
    if(ref_fluid(c)) return 1;
    if(ref_gravity(c)) return 1;
    return 0;
}

// Solver::integrate + update_time, src/solver.cpp:417-429,322
int ref_integrate(ref_ctx * c)
{
    if(ref_timestep(c)) return 1;
    ref_predict(c);
    if(ref_make_tree(c)) return 1;
    if(ref_pre(c)) return 1;
    if(ref_fluid(c)) return 1;
    if(ref_gravity(c)) return 1;
    ref_correct(c);
    c->sim->update_time();
    return 0;
}

// Output::output_energy sums, src/output.cpp:72-83. out = {kinetic, thermal, potential}
void ref_energy(ref_ctx * c, double * out)
{
    const auto & particles = c->sim->get_particles();
    const int num = c->sim->get_particle_num();
    real kinetic = 0.0, thermal = 0.0, potential = 0.0;
#pragma omp parallel for reduction(+: kinetic, thermal, potential)
    for(int i = 0; i < num; ++i) {
        const auto & p_i = particles[i];
        kinetic += 0.5 * p_i.mass * abs2(p_i.vel);
        thermal += p_i.mass * p_i.ene;
        potential += 0.5 * p_i.mass * p_i.phi;
    }
    out[0] = kinetic; out[1] = thermal; out[2] = potential;
}

// Neighbour list of particle i with the CURRENT pos/sml of the context.
// Tree flavour: BHTree::neighbor_search (src/bhtree.cpp:114-126); the tree must have been
// made and (for is_ij) PreInteraction must have run set_kernel().
// Exhaustive flavour: exhaustive_search (src/exhaustive_search.cpp:11-42).
// If h > 0 it temporarily replaces p_i.sml. Returns the count; ids in reference order (sorted by r^2).
int ref_neighbor_search(ref_ctx * c, int i, double h, int is_ij, int * out, int cap)
{
    auto & particles = c->sim->get_particles();
    const int num = c->sim->get_particle_num();
    std::vector<int> list(std::max(cap, num) + 16);
    SPHParticle p_i = particles[i];
    if(h > 0) p_i.sml = h;
#ifdef EXHAUSTIVE_SEARCH
    const int n = exhaustive_search(p_i, p_i.sml, particles, num, list, (int)list.size(), c->sim->get_periodic().get(), is_ij != 0);
#else
    const int n = c->sim->get_tree()->neighbor_search(p_i, list, particles, is_ij != 0);
#endif
    for(int k = 0; k < n && k < cap; ++k) out[k] = list[k];
    return n;
}

// All neighbour lists at once (CSR). offsets has num+1 entries; ids capacity = cap_total.
// Returns total count (may exceed cap_total, in which case ids is truncated).
long long ref_neighbor_search_all(ref_ctx * c, const double * h, int is_ij, long long * offsets, int * ids, long long cap_total)
{
    auto & particles = c->sim->get_particles();
    const int num = c->sim->get_particle_num();
    std::vector<std::vector<int>> lists(num);
#pragma omp parallel for schedule(dynamic, 64)
    for(int i = 0; i < num; ++i) {
        std::vector<int> list(num + 16);
        SPHParticle p_i = particles[i];
        if(h) p_i.sml = h[i];
#ifdef EXHAUSTIVE_SEARCH
        const int n = exhaustive_search(p_i, p_i.sml, particles, num, list, (int)list.size(), c->sim->get_periodic().get(), is_ij != 0);
#else
        const int n = c->sim->get_tree()->neighbor_search(p_i, list, particles, is_ij != 0);
#endif
        lists[i].assign(list.begin(), list.begin() + n);
    }
    long long tot = 0;
    for(int i = 0; i < num; ++i) {
        offsets[i] = tot;
        for(int j : lists[i]) { if(tot < cap_total) ids[tot] = j; ++tot; }
    }
    offsets[num] = tot;
    return tot;
}

// GSPH gradient arrays (src/simulation.cpp:50-76). out: num*DIM doubles.
int ref_get_vector_array(ref_ctx * c, const char * name, double * out)
{
    try {
        auto & v = c->sim->get_vector_array(name);
        for(size_t i = 0; i < v.size(); ++i)
            for(int k = 0; k < DIM; ++k) out[i * DIM + k] = v[i][k];
    } catch(std::exception & e) { c->err = e.what(); return 1; }
    return 0;
}

// Kernel function values straight from the reference's KernelFunction objects
// (include/kernel/cubic_spline.hpp, wendland_kernel.hpp): out = {w, dhw, dw[0..DIM)}
void ref_kernel_eval(ref_ctx * c, const double * rij, double h, double * out)
{
    auto * k = c->sim->get_kernel().get();
    vec_t r;
    for(int d = 0; d < DIM; ++d) r[d] = rij[d];
    const real rr = std::abs(r);
    out[0] = k->w(rr, h);
    out[1] = k->dhw(rr, h);
    const vec_t g = k->dw(r, rr, h);
    for(int d = 0; d < DIM; ++d) out[2 + d] = g[d];
}

} // extern "C"
