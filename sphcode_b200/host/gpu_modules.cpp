// gpu_modules.cpp — see gpu_modules.hpp.  Compiled against the reference's headers.
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "parameters.hpp"
#include "particle.hpp"
#include "simulation.hpp"
#include "exception.hpp"

#include "gpu_modules.hpp"
#include "sphb.h"

namespace sph
{
namespace gpu
{
namespace
{

// SPHParameters (include/parameters.hpp:20-79) -> sphb_params, field for field
sphb_params to_c(const SPHParameters & p)
{
    sphb_params q{};
    q.sph_type = p.type == SPHType::SSPH ? SPHB_SSPH : p.type == SPHType::DISPH ? SPHB_DISPH : SPHB_GSPH;
    q.kernel = p.kernel == KernelType::CUBIC_SPLINE ? SPHB_CUBIC_SPLINE : SPHB_WENDLAND;
    q.cfl_sound = p.cfl.sound; q.cfl_force = p.cfl.force;
    q.av_alpha = p.av.alpha;
    q.use_balsara_switch = p.av.use_balsara_switch; q.use_time_dependent_av = p.av.use_time_dependent_av;
    q.alpha_max = p.av.alpha_max; q.alpha_min = p.av.alpha_min; q.epsilon_av = p.av.epsilon;
    q.use_ac = p.ac.is_valid; q.alpha_ac = p.ac.alpha;
    q.max_tree_level = p.tree.max_level; q.leaf_particle_num = p.tree.leaf_particle_num;
    q.neighbor_number = p.physics.neighbor_number; q.iterative_sml = p.iterative_sml;
    q.gamma = p.physics.gamma;
    q.periodic = p.periodic.is_valid; q.use_gravity = p.gravity.is_valid;
    for(int i = 0; i < DIM; ++i) {
        q.range_max[i] = p.periodic.is_valid ? p.periodic.range_max[i] : 0.0;
        q.range_min[i] = p.periodic.is_valid ? p.periodic.range_min[i] : 0.0;
    }
    q.G = p.gravity.is_valid ? p.gravity.constant : 1.0;
    q.theta = p.gravity.is_valid ? p.gravity.theta : 0.5;
    q.gsph_2nd_order = p.gsph.is_2nd_order;
    return q;
}

struct Session {
    sphb_ctx * ctx = nullptr;
    bool resident = false;       // the device holds a full copy of the particle set
    std::weak_ptr<Simulation> owner;   // the Simulation this context mirrors: a session dies with it
    ~Session() { if(ctx) sphb_destroy(ctx); }
};

std::mutex g_mutex;
std::map<Simulation *, std::shared_ptr<Session>> g_sessions;

// CUDA ordinal of the device the modules run on: SPHB_DEVICE (default 0)
int device_ordinal()
{
    const char * e = std::getenv("SPHB_DEVICE");
    return e ? std::atoi(e) : 0;
}

#define SPHB_CALL(s, call)\
    do {\
        if((call) != 0) {\
            THROW_ERROR("libsphb: ", sphb_last_error((s).ctx));\
        }\
    } while(0)

// The device context of a Simulation.  Sessions are keyed by address but OWNED through a weak_ptr: when a Simulation
// was destroyed (and another one may have been allocated at the same address) its context is dropped — never reused
// with stale resident state — and the contexts of all expired Simulations are released on the way.
Session & session(const std::shared_ptr<Simulation> & sim_sp, const std::shared_ptr<SPHParameters> & param)
{
    Simulation * sim = sim_sp.get();
    std::lock_guard<std::mutex> lock(g_mutex);
    for(auto it = g_sessions.begin(); it != g_sessions.end();) {
        if(it->second && it->second->owner.expired()) it = g_sessions.erase(it);
        else ++it;
    }
    auto & s = g_sessions[sim];
    if(s && s->owner.lock() != sim_sp) s.reset();
    if(!s) {
        s = std::make_shared<Session>();
        s->owner = sim_sp;
        static_assert(sizeof(SPHParticle) == (4 * DIM + 12) * 8 + 16, "SPHParticle layout (include/particle.hpp:8-33)");
        const sphb_params q = to_c(*param);
        if(sphb_create(&q, DIM, device_ordinal(), &s->ctx) != 0) {
            const std::string msg = sphb_last_error(nullptr);
            g_sessions.erase(sim);
            THROW_ERROR("libsphb: ", msg);
        }
    }
    return *s;
}

void upload(Session & s, Simulation & sim, uint32_t mask)
{
    auto & p = sim.get_particles();
    SPHB_CALL(s, sphb_upload_aos(s.ctx, p.data(), sim.get_particle_num(), sizeof(SPHParticle), mask));
}
void download(Session & s, Simulation & sim, uint32_t mask)
{
    auto & p = sim.get_particles();
    SPHB_CALL(s, sphb_download_aos(s.ctx, p.data(), sim.get_particle_num(), sizeof(SPHParticle), mask));
}

// members Solver::predict / Solver::correct write on the host between two stage calls
// (src/solver.cpp:442-455, 468-473)
constexpr uint32_t HOST_WRITES = SPHB_F_POS | SPHB_F_VEL | SPHB_F_VEL_P | SPHB_F_ENE | SPHB_F_ENE_P | SPHB_F_SOUND;

} // namespace

void release(Simulation * sim)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_sessions.erase(sim);
}

// ---- PreInteraction -----------------------------------------------------------------------------
void PreInteraction::initialize(std::shared_ptr<SPHParameters> param) { m_param = param; }

void PreInteraction::calculation(std::shared_ptr<Simulation> sim)
{
    Session & s = session(sim, m_param);
    if(!s.resident) {
        upload(s, *sim, SPHB_F_ALL);      // includes alpha / balsara / sound set by Solver::initialize (src/solver.cpp:399-404)
        s.resident = true;
    } else {
        upload(s, *sim, HOST_WRITES);
    }
    SPHB_CALL(s, sphb_set_dt(s.ctx, sim->get_dt()));
    SPHB_CALL(s, sphb_make_tree(s.ctx));              // the device tree replaces BHTree::make for this stage
    SPHB_CALL(s, sphb_pre_interaction(s.ctx));        // + set_kernel (src/pre_interaction.cpp:164-168)
    download(s, *sim, SPHB_F_SML | SPHB_F_DENS | SPHB_F_PRES | SPHB_F_GRADH | SPHB_F_NEIGHBOR | SPHB_F_BALSARA | SPHB_F_ALPHA);
    double hpvs = 0.0;
    SPHB_CALL(s, sphb_get_h_per_v_sig(s.ctx, &hpvs));
    sim->set_h_per_v_sig(hpvs);
    if(m_param->type == SPHType::GSPH && m_param->gsph.is_2nd_order) {
        // MUSCL gradients, src/gsph/g_pre_interaction.cpp:48-58
        const char * names[] = {"grad_density", "grad_pressure", "grad_velocity_0", "grad_velocity_1", "grad_velocity_2"};
        std::vector<double> buf((size_t)sim->get_particle_num() * DIM);
        for(int a = 0; a < 2 + DIM; ++a) {
            SPHB_CALL(s, sphb_get_vector_array(s.ctx, names[a], buf.data()));
            auto & dst = sim->get_vector_array(names[a]);
            for(size_t i = 0; i < dst.size(); ++i)
                for(int k = 0; k < DIM; ++k) dst[i][k] = buf[i * DIM + k];
        }
    }
}

// ---- FluidForce -----------------------------------------------------------------------------------
void FluidForce::initialize(std::shared_ptr<SPHParameters> param) { m_param = param; }

void FluidForce::calculation(std::shared_ptr<Simulation> sim)
{
    Session & s = session(sim, m_param);
    if(!s.resident) { upload(s, *sim, SPHB_F_ALL); s.resident = true; SPHB_CALL(s, sphb_make_tree(s.ctx)); }
    SPHB_CALL(s, sphb_set_dt(s.ctx, sim->get_dt()));
    SPHB_CALL(s, sphb_fluid_force(s.ctx));
    download(s, *sim, SPHB_F_ACC | SPHB_F_DENE);
}

// ---- GravityForce ---------------------------------------------------------------------------------
void GravityForce::initialize(std::shared_ptr<SPHParameters> param) { m_param = param; }

void GravityForce::calculation(std::shared_ptr<Simulation> sim)
{
    if(!m_param->gravity.is_valid) {
        return;                                        // src/gravity_force.cpp:54-56
    }
    Session & s = session(sim, m_param);
    if(!s.resident) { upload(s, *sim, SPHB_F_ALL); s.resident = true; SPHB_CALL(s, sphb_make_tree(s.ctx)); }
    SPHB_CALL(s, sphb_gravity_force(s.ctx));
    download(s, *sim, SPHB_F_ACC | SPHB_F_PHI);
}

// ---- TimeStep -------------------------------------------------------------------------------------
void TimeStep::initialize(std::shared_ptr<SPHParameters> param) { m_param = param; }

void TimeStep::calculation(std::shared_ptr<Simulation> sim)
{
    Session & s = session(sim, m_param);
    if(!s.resident) { upload(s, *sim, SPHB_F_ALL); s.resident = true; }
    SPHB_CALL(s, sphb_set_h_per_v_sig(s.ctx, sim->get_h_per_v_sig()));
    double dt = 0.0;
    SPHB_CALL(s, sphb_timestep(s.ctx, &dt));           // acc and sml are the device's own (last force stages)
    sim->set_dt(dt);
}

// ---- whole-step fast path ---------------------------------------------------------------------------
void DeviceSolver::initialize(std::shared_ptr<Simulation> sim)
{
    Session & s = session(sim, m_param);
    upload(s, *sim, SPHB_F_ALL);
    s.resident = true;
    SPHB_CALL(s, sphb_initialize(s.ctx));
}

void DeviceSolver::integrate(std::shared_ptr<Simulation> sim)
{
    Session & s = session(sim, m_param);
    double dt = 0.0;
    SPHB_CALL(s, sphb_integrate(s.ctx, &dt));
    sim->set_dt(dt);
}

void DeviceSolver::download(std::shared_ptr<Simulation> sim)
{
    Session & s = session(sim, m_param);
    gpu::download(s, *sim, SPHB_F_ALL);
    double hpvs = 0.0;
    SPHB_CALL(s, sphb_get_h_per_v_sig(s.ctx, &hpvs));
    sim->set_h_per_v_sig(hpvs);
}

}
}
