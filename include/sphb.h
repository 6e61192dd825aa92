/*
 * sphb.h — C ABI of libsphb.so, the B200 (sm_100a) implementation of sphcode's per-step
 * particle hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * mitchiinaga/sphcode tree).  The C++ Module subclasses in sphcode_b200/host/ (same class names
 * and initialize()/calculation() entry points as include/module.hpp:10-14) are thin callers of
 * these functions; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; sphb_last_error(ctx) gives the
 *     message (the C++ wrappers turn it into THROW_ERROR, include/exception.hpp:9-14);
 *   - all calls are synchronous with respect to the host when they return host-visible results
 *     (dt, h_per_v_sig, downloads); kernels run on the context's stream (sphb_set_stream);
 *   - "AoS" buffers have exactly the in-memory layout of sph::SPHParticle for the context's
 *     DIM (include/particle.hpp:8-33): pos, vel, vel_p, acc (DIM doubles each), mass, dens,
 *     pres, ene, ene_p, dene, sml, sound, balsara, alpha, gradh, phi (doubles), id, neighbor
 *     (int32), next (pointer, ignored) = 144 / 176 / 208 bytes for DIM 1 / 2 / 3;
 *   - particle k of an AoS buffer is "particle k" everywhere in this API (neighbour ids,
 *     vector arrays), no matter how the device orders its copy internally.
 *   - there is no CPU fallback: without a CUDA device sphb_create fails.
 */
#ifndef SPHB_H
#define SPHB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB_VERSION 1

/* SPHType, include/parameters.hpp:9-13 */
enum { SPHB_SSPH = 0, SPHB_DISPH = 1, SPHB_GSPH = 2 };
/* KernelType, include/parameters.hpp:15-19 */
enum { SPHB_CUBIC_SPLINE = 0, SPHB_WENDLAND = 1 };

/* Field-for-field mirror of sph::SPHParameters (include/parameters.hpp:20-79); the time block
 * (start/end/output/energy) stays on the host because no device code needs it. */
typedef struct sphb_params {
    int32_t sph_type;             /* SPHParameters::type                      */
    int32_t kernel;               /* SPHParameters::kernel                    */
    double  cfl_sound;            /* cfl.sound                                */
    double  cfl_force;            /* cfl.force                                */
    double  av_alpha;             /* av.alpha                                 */
    int32_t use_balsara_switch;   /* av.use_balsara_switch                    */
    int32_t use_time_dependent_av;/* av.use_time_dependent_av                 */
    double  alpha_max;            /* av.alpha_max                             */
    double  alpha_min;            /* av.alpha_min                             */
    double  epsilon_av;           /* av.epsilon                               */
    int32_t use_ac;               /* ac.is_valid                              */
    int32_t _pad0;
    double  alpha_ac;             /* ac.alpha                                 */
    int32_t max_tree_level;       /* tree.max_level                           */
    int32_t leaf_particle_num;    /* tree.leaf_particle_num                   */
    int32_t neighbor_number;      /* physics.neighbor_number                  */
    int32_t iterative_sml;        /* iterative_sml                            */
    double  gamma;                /* physics.gamma                            */
    int32_t periodic;             /* periodic.is_valid                        */
    int32_t use_gravity;          /* gravity.is_valid                         */
    double  range_max[3];         /* periodic.range_max (first DIM used)      */
    double  range_min[3];         /* periodic.range_min                       */
    double  G;                    /* gravity.constant                         */
    double  theta;                /* gravity.theta                            */
    int32_t gsph_2nd_order;       /* gsph.is_2nd_order                        */
    int32_t _pad1;
} sphb_params;

typedef struct sphb_ctx sphb_ctx;

/* ---- life cycle ------------------------------------------------------------------------- */

/* Replaces Simulation::Simulation + BHTree::initialize + Module::initialize x4
 * (src/simulation.cpp:12-30, src/bhtree.cpp:12-40, src/solver.cpp:387-390).
 * dim in {1,2,3}; device = CUDA ordinal. */
int sphb_create(const sphb_params *params, int dim, int device, sphb_ctx **out);
void sphb_destroy(sphb_ctx *ctx);
const char *sphb_last_error(const sphb_ctx *ctx);   /* ctx may be NULL: last create error */
/* Run all later kernels/copies of this context on `cuda_stream` (a cudaStream_t). */
int sphb_set_stream(sphb_ctx *ctx, void *cuda_stream);
int sphb_synchronize(sphb_ctx *ctx);
int sphb_dim(const sphb_ctx *ctx);
int sphb_particle_num(const sphb_ctx *ctx);           /* particles THIS rank holds (all of them on one GPU) */
long long sphb_global_particle_num(const sphb_ctx *ctx);  /* particles of the whole job */
long long sphb_first_global_index(const sphb_ctx *ctx);   /* tree-order index of this rank's first particle */
size_t sphb_sizeof_particle(int dim);

/* Multi-GPU (one process per GPU, all GPUs on one NVLink box): make this context rank `rank` of `world` contexts that
 * together hold ONE particle set, split by Morton-curve domain decomposition.  `nccl_comm` is an ncclComm_t created
 * by the caller (one per rank).  After this call
 *   - sphb_upload_aos takes the rank's SHARE of the particles (any split of the global set; ids are the caller's);
 *     the first tree build ships every particle to the rank that owns its key range and re-balances every step;
 *   - a rank keeps the full state of its own particles only; what it needs of others (ghost particles within reach of
 *     its h, leaves its gravity walk opens) it reads from the owners' memory over NVLink (CUDA IPC peer mappings);
 *     tree topology, node masses / mass centres and kernel sizes are global (NCCL all-reduce), dt is all-reduced (min);
 *   - every stage call below is collective; sphb_particle_num() is the rank's current count, sphb_download_aos
 *     returns the rank's current particles in its tree order (and an upload with that count updates them in place);
 *   - GSPH, sphb_neighbor_lists, sphb_gravity_direct and the vector-array calls are single-GPU only. */
int sphb_set_distributed(sphb_ctx *ctx, int rank, int world, void *nccl_comm);
/* Same, but the library creates (and owns) the communicator: rank 0 calls sphb_nccl_unique_id,
 * ships the 128 bytes to the other ranks by any means (MPI, torch.distributed, a file), and every
 * rank calls sphb_set_distributed_id with them (ncclCommInitRank inside; libnccl is dlopen'ed). */
int sphb_nccl_unique_id(void *out128);
int sphb_set_distributed_id(sphb_ctx *ctx, int rank, int world, const void *unique_id128);

/* ---- state transfer (Simulation::get_particles(), include/simulation.hpp:25) ------------- */

/* Field groups for partial transfers. */
#define SPHB_F_POS      (1u << 0)
#define SPHB_F_VEL      (1u << 1)
#define SPHB_F_VEL_P    (1u << 2)
#define SPHB_F_ACC      (1u << 3)
#define SPHB_F_MASS     (1u << 4)
#define SPHB_F_DENS     (1u << 5)
#define SPHB_F_PRES     (1u << 6)
#define SPHB_F_ENE      (1u << 7)
#define SPHB_F_ENE_P    (1u << 8)
#define SPHB_F_DENE     (1u << 9)
#define SPHB_F_SML      (1u << 10)
#define SPHB_F_SOUND    (1u << 11)
#define SPHB_F_BALSARA  (1u << 12)
#define SPHB_F_ALPHA    (1u << 13)
#define SPHB_F_GRADH    (1u << 14)
#define SPHB_F_PHI      (1u << 15)
#define SPHB_F_ID       (1u << 16)
#define SPHB_F_NEIGHBOR (1u << 17)
#define SPHB_F_ALL      0x3FFFFu

/* Host AoS -> device.  First call (or a different n) sizes the context (BHTree::resize,
 * src/bhtree.cpp:42-53).  `stride` = bytes between records (>= sphb_sizeof_particle(dim)).
 * field_mask selects which members are taken from the host copy; the first upload must
 * use SPHB_F_ALL.  With a partial mask, a pinned / registered host buffer (sphb_host_alloc,
 * cudaHostRegister) and fewer than 2^18 records the selected members are read in place over PCIe;
 * larger sets go through a whole-record DMA and a masked unpack on the device. */
int sphb_upload_aos(sphb_ctx *ctx, const void *particles, int n, size_t stride, uint32_t field_mask);
/* Device -> host AoS; only members in field_mask are written (partial mask: small sets in pinned /
 * registered memory in place over PCIe by the pack kernel, larger ones by a staged DMA + host threads). */
int sphb_download_aos(sphb_ctx *ctx, void *particles, int n, size_t stride, uint32_t field_mask);

/* GSPH MUSCL gradient arrays, Simulation::get_vector_array(name) (src/simulation.cpp:68-76):
 * "grad_density", "grad_pressure", "grad_velocity_0..DIM-1"; out/in = n*DIM doubles. */
int sphb_get_vector_array(sphb_ctx *ctx, const char *name, double *out);
int sphb_set_vector_array(sphb_ctx *ctx, const char *name, const double *in);

/* Simulation scalars (include/simulation.hpp:27-29). */
int sphb_set_dt(sphb_ctx *ctx, double dt);
int sphb_get_dt(sphb_ctx *ctx, double *dt);
int sphb_set_h_per_v_sig(sphb_ctx *ctx, double v);
int sphb_get_h_per_v_sig(sphb_ctx *ctx, double *v);

/* ---- the hot path ------------------------------------------------------------------------ */

/* alpha = avAlpha, balsara = 1, sound = sqrt(gamma (gamma-1) u): src/solver.cpp:392-404. */
int sphb_init_state(sphb_ctx *ctx);

/* Simulation::make_tree -> BHTree::make (src/simulation.cpp:37-40, src/bhtree.cpp:55-107):
 * bounding cube, Morton keys by the reference's own `pos > center` descent, device radix
 * sort, linear octree with the reference's node set, per-node mass / centre of mass. */
int sphb_make_tree(sphb_ctx *ctx);

/* PreInteraction::calculation for the context's SPHType (src/pre_interaction.cpp:39-169,
 * src/disph/d_pre_interaction.cpp:21-162, src/gsph/g_pre_interaction.cpp:27-144), including
 * initial_smoothing on the first call (src/pre_interaction.cpp:171-215), the Newton-Raphson
 * smoothing length (227-283) and BHTree::set_kernel (src/bhtree.cpp:206-232).
 * Uses the context's dt.  Sets h_per_v_sig. */
int sphb_pre_interaction(sphb_ctx *ctx);

/* FluidForce::calculation (src/fluid_force.cpp:26-116, src/disph/d_fluid_force.cpp:26-86,
 * src/gsph/g_fluid_force.cpp:39-199): overwrites acc and dene. */
int sphb_fluid_force(sphb_ctx *ctx);

/* GravityForce::calculation -> BHTree::tree_force (src/gravity_force.cpp:52-89,
 * src/bhtree.cpp:128-132,301-331): adds to acc, overwrites phi.  No-op when gravity is off. */
int sphb_gravity_force(sphb_ctx *ctx);
/* The EXHAUSTIVE_SEARCH flavour of the same module (src/gravity_force.cpp:70-84): direct sum. */
int sphb_gravity_direct(sphb_ctx *ctx);
/* The same direct sum for the first k particles of the caller's buffer only (targets), over ALL particles as
 * sources: the "64k-particle subsample" gate of the tree-gravity error distribution at sizes where N^2 is out
 * of reach.  Other particles keep acc / phi. */
int sphb_gravity_direct_targets(sphb_ctx *ctx, int k);

/* TimeStep::calculation (src/timestep.cpp:18-38): sets and returns dt. */
int sphb_timestep(sphb_ctx *ctx, double *dt);

/* Solver::predict / Solver::correct (src/solver.cpp:431-474) with the context's dt. */
int sphb_predict(sphb_ctx *ctx);
int sphb_correct(sphb_ctx *ctx);

/* Solver::initialize after the IC (src/solver.cpp:392-414): init_state, make_tree, pre,
 * fluid, gravity.  Solver::integrate (417-429): timestep, predict, make_tree, pre, fluid,
 * gravity, correct; returns the dt used. */
int sphb_initialize(sphb_ctx *ctx);
int sphb_integrate(sphb_ctx *ctx, double *dt);

/* Output::output_energy sums (src/output.cpp:72-83): out = {kinetic, thermal, potential}. */
int sphb_energy(sphb_ctx *ctx, double out[3]);

/* ---- test / measurement hooks ------------------------------------------------------------ */

/* Neighbour lists for the current positions (tree must be made).
 *   h == NULL : use the device's sml; else h[k] is the search radius of particle k.
 *   symmetric == 0 : BHTree::neighbor_search(is_ij=false) sets {j : r2 < h_i^2}
 *                    (src/bhtree.cpp:251-261, src/exhaustive_search.cpp:22-33);
 *   symmetric == 1 : exhaustive_search(is_ij=true) sets {j : r2 < max(h_i^2, h_j^2)}
 *                    (src/exhaustive_search.cpp:28); h_j is always the device's sml.
 * offsets: n+1 int64; ids: capacity cap_total int32, each list sorted by id.
 * *total receives the full count (ids is truncated if it exceeds cap_total). */
int sphb_neighbor_lists(sphb_ctx *ctx, const double *h, int symmetric,
                        int64_t *offsets, int32_t *ids, int64_t cap_total, int64_t *total);

/* Interaction counters of the last stage calls (summed over particles), the inputs of the
 * algorithmic-FLOP model in DESIGN.md. */
typedef struct sphb_counters {
    uint64_t n_particles;
    uint64_t newton_evals;      /* kernel evaluations inside Newton iterations          */
    uint64_t newton_iters;
    uint64_t pre_candidates;    /* candidates r2 < h_search^2                           */
    uint64_t pre_neighbors;     /* neighbours r < h_i (density loop; Balsara loop same) */
    uint64_t force_pairs;       /* pairs 0 < r < max(h_i,h_j)                           */
    uint64_t grav_pp;           /* particle-particle gravity interactions               */
    uint64_t grav_pc;           /* accepted particle-cell (monopole) interactions       */
    uint64_t grav_node_visits;  /* node opening tests, reference semantics per particle */
    uint64_t tree_nodes;
    uint64_t tree_leaves;
    uint64_t grav_pc_group;     /* of grav_pc: cells accepted by every particle of the walking group */
    uint64_t grav_pp_group;     /* of grav_pp: leaves opened by every particle of the walking group  */
    uint64_t n_groups;          /* particle groups (walk work units) of the current tree             */
} sphb_counters;
/* enable != 0 makes the stage kernels count (slower); read with sphb_get_counters. */
int sphb_enable_counters(sphb_ctx *ctx, int enable);
int sphb_get_counters(sphb_ctx *ctx, sphb_counters *out);

/* Device time (ms, CUDA events on the context's stream) of the most recent call of each
 * stage: index by SPHB_T_*. */
enum { SPHB_T_TREE = 0, SPHB_T_PRE = 1, SPHB_T_FLUID = 2, SPHB_T_GRAVITY = 3,
       SPHB_T_TIMESTEP = 4, SPHB_T_PREDICT = 5, SPHB_T_CORRECT = 6,
       SPHB_T_EXCHANGE = 7,   /* multi-GPU: all communication phases of the step (sum of the four below) */
       SPHB_T_MIGRATE = 8,    /*   particle migration (counts + ncclSend/Recv)                           */
       SPHB_T_KEYS = 9,       /*   pull of the ranks' sorted keys                                        */
       SPHB_T_REDUCE = 10,    /*   all-reduces of node sums, kernel sizes, h/v_sig                       */
       SPHB_T_HALO = 11,      /*   halo marking + record pulls over NVLink                               */
       SPHB_T_COUNT = 12 };
/* These times include waiting for the slowest rank at the phase's first collective. */
int sphb_enable_timers(sphb_ctx *ctx, int enable);
int sphb_get_timers(sphb_ctx *ctx, float ms[SPHB_T_COUNT]);

/* Number of kernel launches issued by this context since creation. */
uint64_t sphb_launch_count(const sphb_ctx *ctx);
/* Multi-GPU bookkeeping since creation: ghost records read from peers / particles shipped to other ranks. */
uint64_t sphb_halo_records(const sphb_ctx *ctx);
uint64_t sphb_migrated(const sphb_ctx *ctx);
/* Particles whose Newton-Raphson smoothing-length iteration did not converge since creation (the
 * reference only logs "Particle id N is not convergence", src/pre_interaction.cpp:277-280, and
 * falls back to the guess; so does the device). */
uint64_t sphb_nonconverged(const sphb_ctx *ctx);

/* Pinned host memory for e2e transfers. */
void *sphb_host_alloc(size_t bytes);
void sphb_host_free(void *p);

/* FP64 FMA micro-benchmark (roofline denominator): returns achieved TFLOP/s. */
int sphb_bench_fp64(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* SPHB_H */
