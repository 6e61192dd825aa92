// TEST INFRASTRUCTURE — not part of the product path.
//
// C-ABI access to the UNMODIFIED reference Solver (src/solver.cpp, src/sample/*.cpp, src/output.cpp, src/logger.cpp and
// the module translation units, compiled where they lie by oracle/Makefile with the Boost stand-in of
// oracle/boost_standin/): what Solver::read_parameterfile (src/solver.cpp:155-299) makes of a parameter file and what
// Solver::make_initial_condition + src/sample/*.cpp generate — the ground truth the Boost-free JSON reader / sample
// generators of this repository (sphcode_b200/host/sph_gpu.cpp, sphcode_b200/params.py, samples.py) are pinned to.
// Solver keeps everything private (class default access) and main() is its only caller; this one translation unit reads
// the members by including include/solver.hpp with `class` spelled `struct` (everything solver.hpp itself includes is
// included first, so the macro only touches the Solver declaration; the reference's own translation units are compiled
// unchanged, and class / struct does not change the object layout).
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <unordered_map>
#include <sstream>
#include <fstream>
#include <iostream>
#include <typeinfo>
#include <stdexcept>
#include <boost/property_tree/ptree.hpp>
#include <boost/property_tree/json_parser.hpp>
#include <boost/any.hpp>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "defines.hpp"
#define class struct
#include "solver.hpp"
#undef class
#include "parameters.hpp"
#include "particle.hpp"
#include "simulation.hpp"
#include "exception.hpp"

using namespace sph;

extern "C" {

struct refsolver_params {        // sph::SPHParameters (include/parameters.hpp:20-79) flattened
    double t_start, t_end, t_output, t_energy;
    int    sph_type, kernel;
    double cfl_sound, cfl_force, av_alpha;
    int    use_balsara, use_tdav;
    double alpha_max, alpha_min, epsilon_av;
    int    use_ac;
    double alpha_ac;
    int    max_tree_level, leaf_particle_num, neighbor_number;
    double gamma;
    int    iterative_sml, periodic;
    double range_max[3], range_min[3];
    int    use_gravity;
    double G, theta;
    int    gsph_2nd_order;
    int    n_side;               // m_sample_parameters["N"] (-1: not a sample)
    int    sample;               // enum Sample
    char   output_dir[512];
};

struct refsolver { std::unique_ptr<Solver> s; std::string err; };

int refsolver_dim() { return DIM; }
int refsolver_sizeof_particle() { return (int)sizeof(SPHParticle); }

// arg = sample name or path of a parameter file, relative to the CURRENT directory (the reference reads
// "sample/<name>/<name>.json" and creates its outputDirectory there, src/solver.cpp:164-189, src/logger.cpp:24-37)
refsolver * refsolver_create(const char * arg)
{
    auto * r = new refsolver;
    std::string a0 = "sph", a1 = arg;
    char * argv[] = {&a0[0], &a1[0], nullptr};
    try {
        r->s.reset(new Solver(2, argv));
    } catch (std::exception & e) {
        r->err = e.what();
    }
    return r;
}
void refsolver_destroy(refsolver * r) { delete r; }
const char * refsolver_error(refsolver * r) { return r->err.c_str(); }

int refsolver_get_params(refsolver * r, refsolver_params * o)
{
    if (!r->s) return 1;
    std::memset(o, 0, sizeof(*o));
    const SPHParameters & p = *r->s->m_param;
    o->t_start = p.time.start; o->t_end = p.time.end; o->t_output = p.time.output; o->t_energy = p.time.energy;
    o->sph_type = p.type == SPHType::SSPH ? 0 : p.type == SPHType::DISPH ? 1 : 2;
    o->kernel = p.kernel == KernelType::CUBIC_SPLINE ? 0 : 1;
    o->cfl_sound = p.cfl.sound; o->cfl_force = p.cfl.force; o->av_alpha = p.av.alpha;
    o->use_balsara = p.av.use_balsara_switch; o->use_tdav = p.av.use_time_dependent_av;
    o->alpha_max = p.av.alpha_max; o->alpha_min = p.av.alpha_min; o->epsilon_av = p.av.epsilon;
    o->use_ac = p.ac.is_valid; o->alpha_ac = p.ac.alpha;
    o->max_tree_level = p.tree.max_level; o->leaf_particle_num = p.tree.leaf_particle_num;
    o->neighbor_number = p.physics.neighbor_number; o->gamma = p.physics.gamma;
    o->iterative_sml = p.iterative_sml; o->periodic = p.periodic.is_valid;
    for (int d = 0; d < DIM; ++d) { o->range_max[d] = p.periodic.range_max[d]; o->range_min[d] = p.periodic.range_min[d]; }
    o->use_gravity = p.gravity.is_valid; o->G = p.gravity.constant; o->theta = p.gravity.theta;
    o->gsph_2nd_order = p.gsph.is_2nd_order;
    o->sample = (int)r->s->m_sample;
    o->n_side = -1;
    auto it = r->s->m_sample_parameters.find("N");
    if (it != r->s->m_sample_parameters.end()) o->n_side = boost::any_cast<int>(it->second);
    std::strncpy(o->output_dir, r->s->m_output_dir.c_str(), sizeof(o->output_dir) - 1);
    return 0;
}

// Solver::initialize's first two lines (src/solver.cpp:355-357): the Simulation, then the sample's generator.
// Returns the particle count, -1 on error.
int refsolver_make_ic(refsolver * r)
{
    if (!r->s) return -1;
    try {
        r->s->m_sim = std::make_shared<Simulation>(r->s->m_param);
        r->s->make_initial_condition();
    } catch (std::exception & e) { r->err = e.what(); return -1; }
    return r->s->m_sim->get_particle_num();
}
void refsolver_get_particles(refsolver * r, void * out)
{
    auto & v = r->s->m_sim->get_particles();
    std::memcpy(out, (void *)v.data(), sizeof(SPHParticle) * v.size());
}

} // extern "C"
