// gpu_modules.hpp — the C++ host of the B200 hot path: drop-in replacements of the reference's
// Module subclasses, one per stage, with the reference's own plugin surface
//     void initialize(std::shared_ptr<SPHParameters>)   /   void calculation(std::shared_ptr<Simulation>)
// (include/module.hpp:10-14 of mitchiinaga/sphcode).  This file is compiled AGAINST THE REFERENCE'S
// HEADERS (module.hpp, parameters.hpp, simulation.hpp, particle.hpp, exception.hpp): it is the code
// a reference maintainer adds to the tree, next to a one-line change per module in
// Solver::initialize (src/solver.cpp:359-370), e.g.
//     m_pre = std::make_shared<gpu::PreInteraction>();      // was PreInteraction / disph:: / gsph::
// The classes hold no physics: every call goes through the extern "C" layer of libsphb.so
// (include/sphb.h).  All modules of one Simulation share one device context, found through the
// Simulation pointer, so that the particle state stays resident in HBM between the stage calls and
// only the members a stage reads / writes cross PCIe.
#pragma once

#include <memory>

#include "module.hpp"

namespace sph
{
struct SPHParameters;
class Simulation;

namespace gpu
{

// replaces sph::PreInteraction, sph::disph::PreInteraction and sph::gsph::PreInteraction
// (the SPH type is read from SPHParameters::type)
class PreInteraction : public Module {
    std::shared_ptr<SPHParameters> m_param;
public:
    void initialize(std::shared_ptr<SPHParameters> param) override;
    void calculation(std::shared_ptr<Simulation> sim) override;
};

// replaces sph::FluidForce, sph::disph::FluidForce and sph::gsph::FluidForce
class FluidForce : public Module {
    std::shared_ptr<SPHParameters> m_param;
public:
    void initialize(std::shared_ptr<SPHParameters> param) override;
    void calculation(std::shared_ptr<Simulation> sim) override;
};

// replaces sph::GravityForce
class GravityForce : public Module {
    std::shared_ptr<SPHParameters> m_param;
public:
    void initialize(std::shared_ptr<SPHParameters> param) override;
    void calculation(std::shared_ptr<Simulation> sim) override;
};

// replaces sph::TimeStep
class TimeStep : public Module {
    std::shared_ptr<SPHParameters> m_param;
public:
    void initialize(std::shared_ptr<SPHParameters> param) override;
    void calculation(std::shared_ptr<Simulation> sim) override;
};

// the reference's namespaces for the variants, so that the three branches of Solver::initialize read
// the same as before
namespace disph { using PreInteraction = gpu::PreInteraction; using FluidForce = gpu::FluidForce; }
namespace gsph  { using PreInteraction = gpu::PreInteraction; using FluidForce = gpu::FluidForce; }

// Whole-step fast path (device-resident Solver::initialize / Solver::integrate, src/solver.cpp:353-429):
// no per-stage transfers; the host copy is refreshed only by download().
class DeviceSolver {
    std::shared_ptr<SPHParameters> m_param;
public:
    explicit DeviceSolver(std::shared_ptr<SPHParameters> param) : m_param(param) {}
    void initialize(std::shared_ptr<Simulation> sim);     // upload + init_state + tree + pre + fluid + gravity
    void integrate(std::shared_ptr<Simulation> sim);      // one step; sets sim dt; particles stay on the device
    void download(std::shared_ptr<Simulation> sim);       // device -> sim->get_particles()
};

// drop the device context of a Simulation (optional; contexts die with the process otherwise)
void release(Simulation * sim);

}
}
