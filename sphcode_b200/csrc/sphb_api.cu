// sphb_api.cu — the extern "C" layer of libsphb.so (declared in include/sphb.h).
//
// Host-side orchestration only: memory, the level loop of the tree build, kernel launches on the
// context's stream, and the NCCL exchange of the multi-GPU mode.  All arithmetic lives in the
// kernels of sphb_tree.cuh / sphb_stages.cuh.  There is no CPU fallback anywhere in this file.
#include "../../include/sphb.h"
#include "sphb_dist.cuh"

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <dlfcn.h>
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

using namespace sphb;

namespace {

constexpr int SPHB_MAX_LEVELS = 72;      // tree levels (maxTreeLevel * DIM <= 63 -> at most 64)
std::string g_create_error;

// ---- minimal NCCL binding (dlopen: single-GPU use must not need libnccl) ----------------------
typedef struct ncclComm * ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclInt8 = 0, ncclInt32 = 2, ncclInt = 2, ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };
struct NcclApi {
    void * lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char * (*GetErrorString)(int) = nullptr;
    bool load(std::string & err)
    {
        if (lib) return true;
        const char * names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char * nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define NSYM(f) f = reinterpret_cast<decltype(f)>(dlsym(lib, "nccl" #f)); if (!f) { err = "libnccl lacks nccl" #f; return false; }
        NSYM(GetUniqueId) NSYM(CommInitRank) NSYM(CommDestroy) NSYM(AllGather) NSYM(AllReduce) NSYM(Send) NSYM(Recv) NSYM(GroupStart) NSYM(GroupEnd)
        NSYM(GetErrorString)
#undef NSYM
        return true;
    }
} g_nccl;

} // namespace

constexpr int SPHB_TSEG = 8;             // event pairs per timer kind between two reads (a stage may run several phases)

struct sphb_ctx {
    int dim = 0, device = 0;
    sphb_params hp{};
    DevParams P{};
    cudaStream_t stream = nullptr;       // stream all work is issued on
    cudaStream_t own_stream = nullptr;   // created by sphb_create, destroyed by sphb_destroy
    int sm_count = 148;

    int n = 0;                 // particles this rank owns (= all particles on a single GPU)
    int n_glob = 0;            // particles of the whole job (tree, keys and gather records are global)
    int off = 0;               // global tree-order index of the first own particle
    int cap = 0;               // length of the state arrays (own particles + room for arrivals)
    int n_rec = 0;             // length of the gather-record arrays (n_glob, padded)
    PSoA cur{}, alt{};         // state of the own particles, local index
    PSoA gv{};                 // "global view" of cur: pointers shifted by -off, valid for indices [off, off + n)
    std::vector<void *> allocs;            // everything freed at destroy / resize

    unsigned long long * keys = nullptr, * keys_alt = nullptr;   // local keys / sorted local keys (keys_alt lives in the slab)
    unsigned long long * keys_glob = nullptr;                    // multi-GPU: all ranks' sorted keys (single GPU: = keys_alt)
    int * idx = nullptr, * idx_alt = nullptr;
    void * cub_tmp = nullptr; size_t cub_tmp_bytes = 0;

    TreeBuild tb{}; TreeDev td{};
    std::vector<void *> node_allocs;
    int node_cap = 0;
    std::vector<std::pair<int, int>> levels;
    int * lvl_tmp = nullptr, * lvl_offs = nullptr;
    bool force_full_sort = false;          // redo of a tree build whose partial key sort was too shallow
    int * d_lvl = nullptr, * d_lvl_bad = nullptr;   // speculative tree build: level bounds / failure flag on the device
    bool tree_valid = false;
    bool ksize_valid = false;              // node kernel sizes (nn[2].x, ng[3].y) match the current sml
    bool hsoft_valid = false;              // gravity softening records match the current sml and tree order

    double * d_root = nullptr;             // centre[3], edge
    double * d_bbox_part = nullptr; int bbox_blocks = 0;
    double * d_scal = nullptr;             // [0] dt, [1] h_per_v_sig, [2] dt_force_min, [3..5] energy, [8..13] -lo / hi of the bounding box
    unsigned long long * d_err = nullptr;  // [0] newton non-converged, [1] list overflow, [2] walk error bits, [3] halo records pulled
    Counters * d_cnt = nullptr;
    double dt = 0.0, hpvs = 0.0;
    bool first_pre = true;
    unsigned long long nonconverged_total = 0;

    double * scratch_r = nullptr, * scratch_m = nullptr; int * scratch_j = nullptr; int pre_grid = 0, grav_grid = 0;
    unsigned char * grp_flags = nullptr;   // 1 where a group starts (own particles, local index)
    int * grp_start = nullptr;             // first particle (global index) of every group, ascending (neighbour walks)
    int * d_ngroups = nullptr;             // number of groups (device)
    int * grp_start_g = nullptr;           // the same for the gravity walk (larger group cells)
    int * d_ngroups_g = nullptr;
    int * d_grp_ctl = nullptr;             // [0] work counter, [1] end group of the current kernel
    double2 * grav_lq = nullptr; int * grav_near = nullptr;   // per-warp leaf queues / softened-pair lists of the gravity walk
    bool grav_attr_set[2][2] = {{false, false}, {false, false}};   // dynamic-smem attribute set for k_gravity<DIM, PER, CNT>
    int grav_mode = 1;                     // 1: k_gravity (one particle per lane); 2: k_gravity2 (two per lane: correct, 15 % slower, kept as
                                           // the measured alternative — DESIGN.md section 3); environment SPHB_GRAVITY selects
    Recs rc{};                             // packed gather records, GLOBAL tree-order index (inside the slab)
    bool recs_dirty = true;                // SoA fields changed since the records were packed

    void * d_aos = nullptr; size_t d_aos_bytes = 0;
    void * h_stage = nullptr; size_t h_stage_bytes = 0;

    bool counters_on = false, timers_on = false;
    sphb_counters last_counters{};
    int last_ngroups = 0;
    cudaEvent_t ev[SPHB_T_COUNT][SPHB_TSEG][2] = {};
    int nseg[SPHB_T_COUNT] = {};
    float ms[SPHB_T_COUNT] = {};
    uint64_t launches = 0;

    // ---- multi-GPU: Morton domain decomposition (sphb_dist.cuh)
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr; bool own_comm = false;
    std::vector<int> n_all, off_all;       // every rank's particle count / first global index (off_all has world + 1 entries)
    char * slab = nullptr; size_t slab_bytes = 0; SlabLayout lay{};
    std::vector<void *> peer_open;         // IPC mappings to close
    PeerTab pt{};
    unsigned long long * d_split = nullptr; bool split_valid = false;
    int * d_mig = nullptr;                 // [world + 1] leavers per destination + total, [world] cursors, [world] send offsets, [world * world] all ranks' counts
    int * mig_idx = nullptr, * mig_dest = nullptr;
    double * mig_send = nullptr, * mig_recv = nullptr;
    int * cells_s = nullptr, * d_ncells = nullptr;   // group cells (<= 1024 particles) overlapping the own range: work units of the halo marking
    unsigned char * halo_flags = nullptr, * halo_have = nullptr;
    int * d_bar = nullptr;
    bool orig_valid = true;                // `orig` is a permutation of the caller's buffer indices (false once particles migrated)
    uint64_t halo_pulled = 0, migrated = 0;

    std::string err;
};

namespace {

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    c->err = std::string(#call) + ": " + cudaGetErrorString(e_); return 1; } } while (0)
#define CKN(call) do { int e_ = (call); if (e_ != 0) { \
    c->err = std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e_) : "nccl error"); return 1; } } while (0)
#define LAUNCH_CHECK() do { ++c->launches; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    c->err = std::string("kernel launch: ") + cudaGetErrorString(e_); return 1; } } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

size_t rec_size(int dim) { return (size_t)(4 * dim + 12) * 8 + 16; }

// One timed phase of kind k: every phase between two reads gets its own event pair and the read sums them (a stage
// can have several phases of a kind, e.g. the exchange steps of one time step).
struct Timer {
    sphb_ctx * c; int k, seg;
    Timer(sphb_ctx * c_, int k_) : c(c_), k(k_), seg(-1)
    {
        if (c->timers_on && c->nseg[k] < SPHB_TSEG) { seg = c->nseg[k]++; cudaEventRecord(c->ev[k][seg][0], c->stream); }
    }
    ~Timer() { if (seg >= 0) cudaEventRecord(c->ev[k][seg][1], c->stream); }
};
void read_timers(sphb_ctx * c)
{
    for (int k = 0; k < SPHB_T_COUNT; ++k) {
        if (!c->nseg[k]) continue;
        float tot = 0.f;
        for (int s = 0; s < c->nseg[k]; ++s) { float m = 0.f; cudaEventElapsedTime(&m, c->ev[k][s][0], c->ev[k][s][1]); tot += m; }
        c->ms[k] = tot;
        c->nseg[k] = 0;
    }
}

template <class T> int dev_alloc(sphb_ctx * c, T ** p, size_t count, std::vector<void *> & bag)
{
    void * q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    bag.push_back(q);
    *p = static_cast<T *>(q);
    return 0;
}

void free_bag(std::vector<void *> & bag) { for (void * q : bag) cudaFree(q); bag.clear(); }

// the permuted double / int arrays of a PSoA, in a fixed order
void list_arrays(sphb_ctx * c, PSoA & s, std::vector<double **> & d, std::vector<int **> & i)
{
    const int D = c->dim;
    for (int k = 0; k < D; ++k) d.push_back(&s.pos[k]);
    for (int k = 0; k < D; ++k) d.push_back(&s.vel[k]);
    for (int k = 0; k < D; ++k) d.push_back(&s.vel_p[k]);
    for (int k = 0; k < D; ++k) d.push_back(&s.acc[k]);
    double ** sc[] = {&s.mass, &s.dens, &s.pres, &s.ene, &s.ene_p, &s.dene, &s.sml, &s.sound, &s.balsara, &s.alpha, &s.gradh, &s.phi};
    for (auto q : sc) d.push_back(q);
    if (c->P.sph_type == T_GSPH) {
        for (int k = 0; k < D; ++k) d.push_back(&s.grad_d[k]);
        for (int k = 0; k < D; ++k) d.push_back(&s.grad_p[k]);
        for (int v = 0; v < D; ++v) for (int k = 0; k < D; ++k) d.push_back(&s.grad_v[v][k]);
    }
    i.push_back(&s.pid); i.push_back(&s.neighbor); i.push_back(&s.orig);
}

// gv = cur with every pointer shifted by -off: the walk kernels index particles by their global tree-order index
void make_global_view(sphb_ctx * c)
{
    c->gv = c->cur;
    std::vector<double **> d; std::vector<int **> iv;
    list_arrays(c, c->gv, d, iv);
    for (double ** q : d) *q -= c->off;
    for (int ** q : iv) *q -= c->off;
}

int nccl_barrier(sphb_ctx * c)
{
    if (c->world == 1) return 0;
    CKN(g_nccl.AllReduce(c->d_bar, c->d_bar, 1, ncclInt32, ncclSum, c->comm, c->stream));
    return 0;
}

void close_peers(sphb_ctx * c)
{
    for (void * q : c->peer_open) cudaIpcCloseMemHandle(q);
    c->peer_open.clear();
}

// exchange the IPC handle of the record slab: every rank maps every other rank's slab
int open_peers(sphb_ctx * c)
{
    close_peers(c);
    c->pt = PeerTab{};
    c->pt.world = c->world; c->pt.rank = c->rank;
    c->pt.slab[c->rank] = c->slab;
    if (c->world == 1) return 0;
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, c->slab));
    std::vector<cudaIpcMemHandle_t> all(c->world);
    char * d_h = nullptr;
    CK(cudaMalloc(&d_h, sizeof(mine) * c->world));
    CK(cudaMemcpyAsync(d_h + sizeof(mine) * c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    CKN(g_nccl.AllGather(d_h + sizeof(mine) * c->rank, d_h, sizeof(mine), ncclInt8, c->comm, c->stream));
    CK(cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * c->world, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_h);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        void * q = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            c->err = std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e) +
                     " — the multi-GPU mode needs all ranks on one box with peer access (every GPU visible to every process)";
            return 1;
        }
        c->peer_open.push_back(q);
        c->pt.slab[r] = static_cast<const char *>(q);
    }
    return 0;
}

// n_up = particles this rank was handed; sizes everything for the job
int alloc_particles(sphb_ctx * c, int n_up)
{
    close_peers(c);
    free_bag(c->allocs);
    c->cur = PSoA{}; c->alt = PSoA{};
    c->n = n_up;
    c->n_all.assign(c->world, n_up);
    c->off_all.assign(c->world + 1, 0);
    long long ng = n_up;
    if (c->world > 1) {
        if (c->world > MAX_WORLD) { c->err = "world size above MAX_WORLD"; return 1; }
        if (c->P.sph_type == T_GSPH) { c->err = "GSPH is single-GPU only (the MUSCL gradient arrays are not part of the halo records)"; return 1; }
        // every rank's count
        int * d_n = nullptr;
        CK(cudaMalloc(&d_n, sizeof(int) * c->world));
        CK(cudaMemcpyAsync(d_n + c->rank, &n_up, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CKN(g_nccl.AllGather(d_n + c->rank, d_n, 1, ncclInt32, c->comm, c->stream));
        CK(cudaMemcpyAsync(c->n_all.data(), d_n, sizeof(int) * c->world, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(d_n);
        ng = 0;
        for (int r = 0; r < c->world; ++r) { c->off_all[r] = (int)ng; ng += c->n_all[r]; }
        if (ng > 2000000000LL) { c->err = "more than 2e9 particles"; return 1; }
    }
    c->off_all[c->world] = (int)ng;
    c->off = c->off_all[c->rank];
    c->n_glob = (int)ng;
    const int n_max = *std::max_element(c->n_all.begin(), c->n_all.end());
    c->cap = c->world == 1 ? n_up : std::max<long long>(n_max, (long long)(ng / c->world * 1.3) + 4096) + 32;
    c->n_rec = c->n_glob + 64;
    const size_t np = (size_t)c->cap;
    const int n = c->cap;

    for (int side = 0; side < 2; ++side) {
        PSoA & s = side == 0 ? c->cur : c->alt;
        std::vector<double **> d; std::vector<int **> iv;
        list_arrays(c, s, d, iv);
        double * pool = nullptr; int * ipool = nullptr;
        if (dev_alloc(c, &pool, np * d.size(), c->allocs)) return 1;
        if (dev_alloc(c, &ipool, np * iv.size(), c->allocs)) return 1;
        CK(cudaMemsetAsync(pool, 0, np * d.size() * sizeof(double), c->stream));
        CK(cudaMemsetAsync(ipool, 0, np * iv.size() * sizeof(int), c->stream));
        for (size_t k = 0; k < d.size(); ++k) *d[k] = pool + k * np;
        for (size_t k = 0; k < iv.size(); ++k) *iv[k] = ipool + k * np;
    }
    // record slab (exported to the peers): posm | velc | thermo | av | hsoft | sorted local keys
    {
        const size_t nr = (size_t)c->n_rec;
        SlabLayout & L = c->lay;
        L.posm = 0; L.velc = nr * 32; L.thermo = nr * 64; L.av = nr * 96; L.hsoft = nr * 128; L.keys = nr * 144;
        L.mig = L.keys + (np + 32) * sizeof(unsigned long long);
        c->slab_bytes = L.mig + (c->world > 1 ? np * mig_rec(c->dim) * sizeof(double) : 0);
        if (dev_alloc(c, &c->slab, c->slab_bytes, c->allocs)) return 1;
        CK(cudaMemsetAsync(c->slab, 0, L.mig, c->stream));
        c->rc.posm = reinterpret_cast<double4 *>(c->slab + L.posm); c->rc.velc = reinterpret_cast<double4 *>(c->slab + L.velc);
        c->rc.thermo = reinterpret_cast<double4 *>(c->slab + L.thermo); c->rc.av = reinterpret_cast<double4 *>(c->slab + L.av);
        c->rc.hsoft = reinterpret_cast<double2 *>(c->slab + L.hsoft);
        c->keys_alt = reinterpret_cast<unsigned long long *>(c->slab + L.keys);
    }
    if (dev_alloc(c, &c->keys, np, c->allocs) || dev_alloc(c, &c->idx, np, c->allocs) || dev_alloc(c, &c->idx_alt, np, c->allocs)) return 1;
    if (c->world > 1) { if (dev_alloc(c, &c->keys_glob, (size_t)c->n_glob + 32, c->allocs)) return 1; }
    else c->keys_glob = c->keys_alt;
    c->cub_tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, c->cub_tmp_bytes, c->keys, c->keys_alt, c->idx, c->idx_alt, n, 0, 64, c->stream);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int *)nullptr, (int *)nullptr, (int)std::min<long long>(5LL * c->n_glob + 2, 2000000000LL), c->stream);
    size_t sel_bytes = 0;
    cub::DeviceSelect::Flagged(nullptr, sel_bytes, thrust::counting_iterator<int>(0), (unsigned char *)nullptr, (int *)nullptr, (int *)nullptr, n, c->stream);
    size_t keysort_bytes = 0;
    if (c->world > 1) cub::DeviceRadixSort::SortKeys(nullptr, keysort_bytes, c->keys_glob, c->keys_glob, c->n_glob, 0, 64, c->stream);
    c->cub_tmp_bytes = std::max(std::max(std::max(c->cub_tmp_bytes, scan_bytes), sel_bytes), keysort_bytes) + 256;
    { char * t = nullptr; if (dev_alloc(c, &t, c->cub_tmp_bytes, c->allocs)) return 1; c->cub_tmp = t; }
    c->bbox_blocks = std::min(cdiv(n, 256), 4 * c->sm_count);
    if (dev_alloc(c, &c->d_bbox_part, (size_t)c->bbox_blocks * 6, c->allocs)) return 1;

    // list scratch: one r, j (and m) column set per resident warp of the persistent pre / force kernels
    const int groups = cdiv(n, 32);
    c->pre_grid = std::min(cdiv(groups, 4), c->sm_count * PF_BLOCKS);
    c->grav_grid = std::min(cdiv(groups, 4), c->sm_count * GV_BLOCKS);
    const size_t slots = (size_t)c->pre_grid * 4;
    if (dev_alloc(c, &c->scratch_r, slots * c->P.list_cap * 32, c->allocs)) return 1;
    if (dev_alloc(c, &c->scratch_j, slots * c->P.list_cap * 32, c->allocs)) return 1;
    if (c->P.sph_type != T_DISPH) { if (dev_alloc(c, &c->scratch_m, slots * c->P.list_cap * 32, c->allocs)) return 1; }
    else c->scratch_m = nullptr;
    if (dev_alloc(c, &c->grp_flags, np + 32, c->allocs) || dev_alloc(c, &c->grp_start, np + 32, c->allocs) ||
        dev_alloc(c, &c->grp_start_g, np + 32, c->allocs)) return 1;
    if (c->P.use_gravity) {
        if (dev_alloc(c, &c->grav_lq, slots * GRAV_LQ * 32, c->allocs) || dev_alloc(c, &c->grav_near, slots * GRAV_NEAR * 32, c->allocs)) return 1;
    }
    c->recs_dirty = true;

    c->d_aos_bytes = (size_t)n * rec_size(c->dim);
    { char * t = nullptr; if (dev_alloc(c, &t, c->d_aos_bytes, c->allocs)) return 1; c->d_aos = t; }
    if (c->world > 1) {
        const int W = c->world;
        if (dev_alloc(c, &c->d_split, W + 1, c->allocs) || dev_alloc(c, &c->d_mig, 3 * W + 1 + W * W, c->allocs) ||
            dev_alloc(c, &c->mig_idx, np, c->allocs) || dev_alloc(c, &c->mig_dest, np, c->allocs) ||
            dev_alloc(c, &c->mig_recv, np * mig_rec(c->dim), c->allocs) ||
            dev_alloc(c, &c->cells_s, np + 32, c->allocs) ||
            dev_alloc(c, &c->d_ncells, 2, c->allocs) || dev_alloc(c, &c->d_bar, 1, c->allocs)) return 1;
        CK(cudaMemsetAsync(c->d_bar, 0, sizeof(int), c->stream));
        c->mig_send = reinterpret_cast<double *>(c->slab + c->lay.mig);
        if (open_peers(c)) return 1;
    } else {
        c->pt = PeerTab{}; c->pt.world = 1; c->pt.slab[0] = c->slab;
    }
    for (int r = 0; r <= c->world; ++r) c->pt.off[r] = c->off_all[r];
    c->split_valid = false;
    c->orig_valid = true;
    c->tree_valid = false; c->ksize_valid = false; c->hsoft_valid = false;
    c->first_pre = true;
    c->levels.clear();
    make_global_view(c);
    return 0;
}

int alloc_nodes(sphb_ctx * c, int cap, int keep, int keep_offs = 0)
{
    // (re)allocate node arrays with capacity `cap`, preserving the first `keep` BFS nodes and the
    // first `keep_offs` child offsets of the level being emitted
    TreeBuild old = c->tb;
    int * old_offs = c->lvl_offs;
    std::vector<void *> old_bag;
    old_bag.swap(c->node_allocs);
    TreeBuild & t = c->tb;
    t = TreeBuild{};
    int ** ia[] = {&t.first, &t.count, &t.level, &t.parent, &t.child0, &t.nchild};
    int * const oi[] = {old.first, old.count, old.level, old.parent, old.child0, old.nchild};
    for (int k = 0; k < 6; ++k) {
        if (dev_alloc(c, ia[k], cap, c->node_allocs)) return 1;
        if (keep) CK(cudaMemcpyAsync(*ia[k], oi[k], (size_t)keep * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    }
    for (int d = 0; d < c->dim; ++d) {
        if (dev_alloc(c, &t.center[d], cap, c->node_allocs)) return 1;
        if (keep) CK(cudaMemcpyAsync(t.center[d], old.center[d], (size_t)keep * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    if (dev_alloc(c, &t.msum4, cap, c->node_allocs)) return 1;
    if (dev_alloc(c, &c->lvl_tmp, cap, c->node_allocs) || dev_alloc(c, &c->lvl_offs, cap, c->node_allocs)) return 1;
    if (keep_offs) CK(cudaMemcpyAsync(c->lvl_offs, old_offs, (size_t)keep_offs * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    TreeDev & o = c->td;
    if (dev_alloc(c, &o.nn, (size_t)cap * 4, c->node_allocs) || dev_alloc(c, &o.ng, (size_t)cap * 4, c->node_allocs) ||
        dev_alloc(c, &o.parent, cap, c->node_allocs) || dev_alloc(c, &o.ksize, cap, c->node_allocs)) return 1;
    if (c->world > 1) {
        if (dev_alloc(c, &c->halo_flags, (size_t)cap + 8, c->node_allocs) || dev_alloc(c, &c->halo_have, (size_t)cap + 8, c->node_allocs)) return 1;
    }
    CK(cudaStreamSynchronize(c->stream));
    free_bag(old_bag);
    c->node_cap = cap;
    return 0;
}

// ---- AoS <-> SoA --------------------------------------------------------------------------------
template <int DIM>
__global__ void k_unpack(const char * __restrict__ aos, size_t stride, PSoA s, int n, uint32_t mask, int first)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = first ? i : s.orig[i];
    const double * r = reinterpret_cast<const double *>(aos + (size_t)k * stride);
    if (first) s.orig[i] = i;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        if (mask & SPHB_F_POS)   s.pos[d][i]   = r[d];
        if (mask & SPHB_F_VEL)   s.vel[d][i]   = r[DIM + d];
        if (mask & SPHB_F_VEL_P) s.vel_p[d][i] = r[2 * DIM + d];
        if (mask & SPHB_F_ACC)   s.acc[d][i]   = r[3 * DIM + d];
    }
    const double * q = r + 4 * DIM;
    if (mask & SPHB_F_MASS)    s.mass[i]    = q[0];
    if (mask & SPHB_F_DENS)    s.dens[i]    = q[1];
    if (mask & SPHB_F_PRES)    s.pres[i]    = q[2];
    if (mask & SPHB_F_ENE)     s.ene[i]     = q[3];
    if (mask & SPHB_F_ENE_P)   s.ene_p[i]   = q[4];
    if (mask & SPHB_F_DENE)    s.dene[i]    = q[5];
    if (mask & SPHB_F_SML)     s.sml[i]     = q[6];
    if (mask & SPHB_F_SOUND)   s.sound[i]   = q[7];
    if (mask & SPHB_F_BALSARA) s.balsara[i] = q[8];
    if (mask & SPHB_F_ALPHA)   s.alpha[i]   = q[9];
    if (mask & SPHB_F_GRADH)   s.gradh[i]   = q[10];
    if (mask & SPHB_F_PHI)     s.phi[i]     = q[11];
    const int * iq = reinterpret_cast<const int *>(q + 12);
    if (mask & SPHB_F_ID)       s.pid[i]      = iq[0];
    if (mask & SPHB_F_NEIGHBOR) s.neighbor[i] = iq[1];
}

// SoA -> the caller's AoS order.  Only the members in `mask` are written (a full mask also nulls SPHParticle::next), so
// the same kernel serves the staged full download and a field-masked download straight into mapped host memory.
template <int DIM>
__global__ void k_pack(char * __restrict__ aos, size_t stride, PSoA s, int n, uint32_t mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = s.orig[i];
    double * r = reinterpret_cast<double *>(aos + (size_t)k * stride);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        if (mask & SPHB_F_POS)   r[d] = s.pos[d][i];
        if (mask & SPHB_F_VEL)   r[DIM + d] = s.vel[d][i];
        if (mask & SPHB_F_VEL_P) r[2 * DIM + d] = s.vel_p[d][i];
        if (mask & SPHB_F_ACC)   r[3 * DIM + d] = s.acc[d][i];
    }
    double * q = r + 4 * DIM;
    if (mask & SPHB_F_MASS)    q[0] = s.mass[i];
    if (mask & SPHB_F_DENS)    q[1] = s.dens[i];
    if (mask & SPHB_F_PRES)    q[2] = s.pres[i];
    if (mask & SPHB_F_ENE)     q[3] = s.ene[i];
    if (mask & SPHB_F_ENE_P)   q[4] = s.ene_p[i];
    if (mask & SPHB_F_DENE)    q[5] = s.dene[i];
    if (mask & SPHB_F_SML)     q[6] = s.sml[i];
    if (mask & SPHB_F_SOUND)   q[7] = s.sound[i];
    if (mask & SPHB_F_BALSARA) q[8] = s.balsara[i];
    if (mask & SPHB_F_ALPHA)   q[9] = s.alpha[i];
    if (mask & SPHB_F_GRADH)   q[10] = s.gradh[i];
    if (mask & SPHB_F_PHI)     q[11] = s.phi[i];
    int * iq = reinterpret_cast<int *>(q + 12);
    if (mask & SPHB_F_ID)       iq[0] = s.pid[i];
    if (mask & SPHB_F_NEIGHBOR) iq[1] = s.neighbor[i];
    if ((mask & SPHB_F_ALL) == SPHB_F_ALL) q[13] = 0.0;       // SPHParticle::next: tree scratch in the reference, null here
}

__global__ void k_scatter_by_orig(const double * __restrict__ src, const int * __restrict__ orig, double * __restrict__ dst, int n, int ncomp, int comp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[(size_t)orig[i] * ncomp + comp] = src[i];
}
__global__ void k_gather_by_orig(const double * __restrict__ src, const int * __restrict__ orig, double * __restrict__ dst, int n, int ncomp, int comp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[(size_t)orig[i] * ncomp + comp];
}
__global__ void k_set_scalars(double * s, double dt, double hpvs, int which)
{
    if (which & 1) s[0] = dt;
    if (which & 2) s[1] = hpvs;
}

// (re)pack the gather records of the own particles from the SoA state: what = 1 posm | 2 velc | 4 thermo + av
int pack_recs(sphb_ctx * c, int what)
{
    const int n = c->n, B = 256;
    if (n > 0) {
        switch (c->dim) {
        case 1: k_pack_recs<1><<<cdiv(n, B), B, 0, c->stream>>>(c->cur, c->rc, n, what, c->off); break;
        case 2: k_pack_recs<2><<<cdiv(n, B), B, 0, c->stream>>>(c->cur, c->rc, n, what, c->off); break;
        default: k_pack_recs<3><<<cdiv(n, B), B, 0, c->stream>>>(c->cur, c->rc, n, what, c->off); break;
        }
        LAUNCH_CHECK();
    }
    if (what == 7) c->recs_dirty = false;
    return 0;
}
int ensure_recs(sphb_ctx * c) { return c->recs_dirty ? pack_recs(c, 7) : 0; }

// group table view for a kernel that works on the own particles
int group_table(sphb_ctx * c, GroupTable & gt, bool gravity = false)
{
    const int * start = gravity ? c->grp_start_g : c->grp_start, * ng = gravity ? c->d_ngroups_g : c->d_ngroups;
    k_group_range<<<1, 32, 0, c->stream>>>(start, ng, c->off, c->off + c->n, c->d_grp_ctl); LAUNCH_CHECK();
    gt.start = start; gt.n_groups = ng; gt.ctl = c->d_grp_ctl; gt.n = c->off + c->n;
    return 0;
}

// ---- multi-GPU: particle migration to the owners of their keys ------------------------------------------------------------
// On entry c->keys holds the keys of the n own particles (local order).  Particles whose key lies outside this rank's
// splitter range move to their owner (packed by destination into the slab, pulled by the receivers over NVLink), arrivals fill the leavers' slots / go behind the
// own particles, and the leavers' keys get the leaver bit, so that the sort moves them behind everything that stays.
// *n_sort = entries to sort, *n_new = particles this rank owns afterwards.
template <int DIM> int migrate_t(sphb_ctx * c, int * n_sort, int * n_new)
{
    Timer tx(c, SPHB_T_EXCHANGE), ty(c, SPHB_T_MIGRATE);
    const int W = c->world, n = c->n, B = 256, key_bits = c->P.key_levels * DIM;
    int * cnt = c->d_mig, * cursor = c->d_mig + (W + 1), * soff_d = c->d_mig + (2 * W + 1), * mat = c->d_mig + (3 * W + 1);
    CK(cudaMemsetAsync(c->d_mig, 0, sizeof(int) * (2 * W + 1), c->stream));
    if (n > 0) { k_mig_mark<<<cdiv(n, B), B, 0, c->stream>>>(c->keys, n, c->d_split, W, c->rank, key_bits, cnt, c->mig_idx, c->mig_dest); LAUNCH_CHECK(); }
    CKN(g_nccl.AllGather(cnt, mat, W, ncclInt32, c->comm, c->stream));
    std::vector<int> hm((size_t)W * W);
    CK(cudaMemcpyAsync(hm.data(), mat, sizeof(int) * W * W, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<int> soff(W + 1, 0), roff(W + 1, 0);
    long long moved_all = 0;
    for (int d = 0; d < W; ++d) { soff[d + 1] = soff[d] + hm[(size_t)c->rank * W + d]; roff[d + 1] = roff[d] + hm[(size_t)d * W + c->rank]; }
    for (int v : hm) moved_all += v;
    const int n_leave = soff[W], n_recv = roff[W];
    // every rank's new count (all ranks hold the same matrix); the capacity check is made for EVERY rank by every rank, so
    // that all of them leave with the error instead of one returning and the others waiting in the next collective
    bool fits = true;
    for (int r = 0; r < W; ++r) {
        int out = 0, in = 0;
        for (int d = 0; d < W; ++d) { out += hm[(size_t)r * W + d]; in += hm[(size_t)d * W + r]; }
        if ((long long)c->n_all[r] + std::max(0, in - out) > c->cap) fits = false;
        c->n_all[r] += in - out;
    }
    *n_sort = n + std::max(0, n_recv - n_leave);          // arrivals first fill the leavers' slots
    *n_new = n - n_leave + n_recv;
    if (!fits) { c->err = "domain decomposition: arrivals exceed the state capacity of a rank (load imbalance above 30 %)"; return 1; }
    if (moved_all == 0) return 0;
    c->migrated += (uint64_t)n_leave;
    c->orig_valid = false;
    if (n_leave > 0) {
        CK(cudaMemcpyAsync(soff_d, soff.data(), sizeof(int) * W, cudaMemcpyHostToDevice, c->stream));
        k_mig_pack<DIM><<<cdiv(n_leave, B), B, 0, c->stream>>>(c->cur, c->mig_idx, c->mig_dest, n_leave, soff_d, cursor, c->mig_send); LAUNCH_CHECK();
    }
    // arrivals are PULLED out of the senders' slabs (the blocks were packed by destination; every rank knows all counts)
    if (nccl_barrier(c)) return 1;                      // every rank's leavers are packed
    if (n_recv > 0) {
        MigPull mp{};
        for (int sr = 0; sr < W; ++sr) {
            int so = 0;
            for (int d = 0; d < c->rank; ++d) so += hm[(size_t)sr * W + d];
            mp.src_off[sr] = so;
            mp.dst_off[sr] = roff[sr];
        }
        mp.dst_off[W] = roff[W];
        k_mig_pull<<<std::min(cdiv((long long)n_recv * mig_rec(DIM), B), c->sm_count * 8), B, 0, c->stream>>>(c->pt, c->lay.mig, mp, mig_rec(DIM), c->mig_recv);
        LAUNCH_CHECK();
    }
    if (n_recv > 0) {
        k_mig_unpack<DIM><<<cdiv(n_recv, B), B, 0, c->stream>>>(c->cur, c->mig_recv, n_recv, c->mig_idx, n_leave, n,
                                                                c->d_root, c->P.key_levels, c->keys, c->idx);
        LAUNCH_CHECK();
    }
    return 0;
}

// pull every rank's sorted key run into keys_glob (replicated topology); pt.off must be current
int gather_keys(sphb_ctx * c)
{
    Timer tx(c, SPHB_T_EXCHANGE), ty(c, SPHB_T_KEYS);
    if (nccl_barrier(c)) return 1;                      // every rank's keys are sorted
    k_gather_keys<<<c->sm_count * 8, 256, 0, c->stream>>>(c->pt, c->lay.keys, c->keys_glob, c->n_glob); LAUNCH_CHECK();
    return 0;
}

void set_offsets(sphb_ctx * c)
{
    int o = 0;
    for (int r = 0; r < c->world; ++r) { c->off_all[r] = o; o += c->n_all[r]; }
    c->off_all[c->world] = o;
    c->off = c->off_all[c->rank];
    c->n = c->n_all[c->rank];
    for (int r = 0; r <= c->world; ++r) c->pt.off[r] = c->off_all[r];
}

// ---- tree ---------------------------------------------------------------------------------------------
template <int DIM> int make_tree_t(sphb_ctx * c)
{
    Timer tm(c, SPHB_T_TREE);
    const int B = 256, W = c->world;
    const bool dist = W > 1;
    if (!c->P.periodic) {
        k_bbox_partial<DIM><<<c->bbox_blocks, B, 0, c->stream>>>(c->cur, c->n, c->d_bbox_part); LAUNCH_CHECK();
        if (!dist) { k_bbox_final<DIM><<<1, 32, 0, c->stream>>>(c->d_bbox_part, c->bbox_blocks, c->d_root); LAUNCH_CHECK(); }
        else {
            k_bbox_neg<DIM><<<1, 32, 0, c->stream>>>(c->d_bbox_part, c->bbox_blocks, c->d_scal + 8); LAUNCH_CHECK();
            CKN(g_nccl.AllReduce(c->d_scal + 8, c->d_scal + 8, 2 * DIM, ncclFloat64, ncclMax, c->comm, c->stream));
            k_root_from_bbox<DIM><<<1, 32, 0, c->stream>>>(c->d_scal + 8, c->d_root); LAUNCH_CHECK();
        }
    }
    if (c->n > 0) { k_keys<DIM><<<cdiv(c->n, B), B, 0, c->stream>>>(c->cur, 0, c->n, c->d_root, c->P.key_levels, c->keys, c->idx); LAUNCH_CHECK(); }
    int n_sort = c->n, n_new = c->n;
    if (dist) {
        if (!c->split_valid) {
            // first build: splitters = exact quantiles of all keys (one-time replicated sort of the gathered, unsorted keys)
            CK(cudaMemcpyAsync(c->keys_alt, c->keys, sizeof(unsigned long long) * c->n, cudaMemcpyDeviceToDevice, c->stream));
            if (gather_keys(c)) return 1;
            unsigned long long * tmpk = nullptr;
            CK(cudaMalloc(&tmpk, sizeof(unsigned long long) * ((size_t)c->n_glob + 32)));
            size_t tb = c->cub_tmp_bytes;
            CK(cub::DeviceRadixSort::SortKeys(c->cub_tmp, tb, c->keys_glob, tmpk, c->n_glob, 0, std::min(64, c->P.key_levels * DIM), c->stream));
            ++c->launches;
            k_next_splitters<<<1, 32, 0, c->stream>>>(tmpk, c->n_glob, W, c->d_split); LAUNCH_CHECK();
            CK(cudaStreamSynchronize(c->stream));
            cudaFree(tmpk);
            if (nccl_barrier(c)) return 1;              // peers are done reading my unsorted keys
            c->split_valid = true;
        }
        if (migrate_t<DIM>(c, &n_sort, &n_new)) return 1;
    }
    // Partial radix sort: only the key levels the tree can use are sorted — one more than the depth of the
    // previous tree (the sort is stable and its input is the previous tree order).  If a node below the sorted
    // levels turns out to need a split, the build is redone with all levels (deeper flag).
    const int true_max_level = std::min(c->P.max_level, c->P.key_levels);
    int sort_levels = c->P.key_levels;
    {
        const int depth = (int)c->levels.size();
        if (!c->force_full_sort && depth > 0 && depth <= true_max_level) sort_levels = std::min(c->P.key_levels, depth + 1);
    }
    const int key_bits = c->P.key_levels * DIM;
    if (n_sort > 0) {
        size_t tmp = c->cub_tmp_bytes;
        CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, c->keys, c->keys_alt, c->idx, c->idx_alt, n_sort,
                                           (c->P.key_levels - sort_levels) * DIM, std::min(64, key_bits + (dist ? 1 : 0)), c->stream));
        ++c->launches;
    }
    if (dist) { set_offsets(c); }                       // n, off of every rank after the migration
    const int n = c->n, N = c->n_glob;
    if (n > 0) {
        k_permute_pack<DIM><<<cdiv(n, B), B, 0, c->stream>>>(c->cur, c->alt, c->rc, c->idx_alt, n, c->P.sph_type == T_GSPH ? 1 : 0, c->off);
        LAUNCH_CHECK();
    }
    c->recs_dirty = false;
    std::swap(c->cur, c->alt);
    make_global_view(c);
    if (dist) {
        if (gather_keys(c)) return 1;
        k_next_splitters<<<1, 32, 0, c->stream>>>(c->keys_glob, N, W, c->d_split); LAUNCH_CHECK();   // lag-one re-balancing
    }
    const unsigned long long * keys = c->keys_glob;

    if (c->node_cap == 0) { if (alloc_nodes(c, std::max(1024, N / 2 + 64), 0)) return 1; }
    const int max_level_eff = std::min(true_max_level, sort_levels);
    const long long node_limit = 5LL * N + 1;          // BHTree::resize: 5 N nodes + the root (src/bhtree.cpp:46)
    CK(cudaMemsetAsync(c->d_lvl_bad, 0, 2 * sizeof(int), c->stream));     // [0] speculation failed, [1] deeper sort needed
    k_root_init<<<1, 32, 0, c->stream>>>(c->tb, N, c->d_root); LAUNCH_CHECK();
    int lb = 0, le = 1;
    bool built = false;
    // ---- speculative build: the level widths of the previous tree (+25 %) size the grids, the level bounds
    // stay on the device, and the host synchronises ONCE at the end instead of once per level.  Any surprise
    // (a level wider than its grid, a deeper tree, a full node pool) falls back to the exact loop below.
    if (!c->levels.empty() && c->levels.size() + 3 <= SPHB_MAX_LEVELS) {
        const std::vector<std::pair<int, int>> prev = c->levels;
        const int nl = (int)prev.size() + 1;                  // one level more than last time, must come out empty
        auto cap_of = [&](int l) -> int {
            if (l == 0) return 1;
            const long long w = l < (int)prev.size() ? prev[l].second - prev[l].first : 0;
            return (int)std::min<long long>(c->node_cap, w + w / 4 + 1024);
        };
        int init[3] = {0, 1, 0};
        CK(cudaMemcpyAsync(c->d_lvl, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
        for (int l = 0; l < nl; ++l) {
            const int wc = cap_of(l);
            k_level_count<DIM><<<cdiv((long long)wc << DIM, B), B, 0, c->stream>>>(c->tb, keys, 0, 0, c->P.leaf_num, max_level_eff,
                c->P.key_levels, c->lvl_tmp, c->d_lvl + l, wc, true_max_level, c->d_lvl_bad + 1);
            LAUNCH_CHECK();
            size_t tb = c->cub_tmp_bytes;
            CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp, tb, c->lvl_tmp, c->lvl_offs, wc, c->stream));
            ++c->launches;
            k_level_advance<<<1, 32, 0, c->stream>>>(c->d_lvl + l, c->lvl_tmp, c->lvl_offs, cap_of(l + 1), c->node_cap, c->d_lvl_bad); LAUNCH_CHECK();
            k_level_emit<DIM><<<cdiv((long long)wc << DIM, B), B, 0, c->stream>>>(c->tb, keys, 0, 0, c->P.leaf_num, max_level_eff,
                c->P.key_levels, c->lvl_offs, c->d_root, c->d_lvl + l, c->d_lvl_bad);
            LAUNCH_CHECK();
        }
        int h_lvl[SPHB_MAX_LEVELS + 2], bad[2] = {0, 0};
        CK(cudaMemcpyAsync(h_lvl, c->d_lvl, (size_t)(nl + 2) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(bad, c->d_lvl_bad, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (bad[1]) { c->force_full_sort = true; const int r = make_tree_t<DIM>(c); c->force_full_sort = false; return r; }
        if (!bad[0] && h_lvl[nl + 1] == h_lvl[nl]) {             // the extra level is empty: the tree is complete
            c->levels.clear();
            for (int l = 0; l < nl && h_lvl[l + 1] > h_lvl[l]; ++l) c->levels.push_back({h_lvl[l], h_lvl[l + 1]});
            lb = h_lvl[nl];
            built = true;
        } else {
            k_root_init<<<1, 32, 0, c->stream>>>(c->tb, N, c->d_root); LAUNCH_CHECK();
        }
    }
    if (!built) {
    c->levels.clear();
    lb = 0; le = 1;
    while (le > lb) {
        c->levels.push_back({lb, le});
        const int w = le - lb;
        k_level_count<DIM><<<cdiv((long long)w << DIM, B), B, 0, c->stream>>>(c->tb, keys, lb, le, c->P.leaf_num, max_level_eff, c->P.key_levels, c->lvl_tmp, nullptr, 0, true_max_level, c->d_lvl_bad + 1);
        LAUNCH_CHECK();
        size_t tb = c->cub_tmp_bytes;
        CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp, tb, c->lvl_tmp, c->lvl_offs, w, c->stream));
        ++c->launches;
        int last_off = 0, last_cnt = 0;
        CK(cudaMemcpyAsync(&last_off, c->lvl_offs + (w - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(&last_cnt, c->lvl_tmp + (w - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        const int total = last_off + last_cnt;
        if ((long long)le + total > node_limit) { c->err = "There is no free node."; return 1; }    // src/bhtree.cpp:179-181
        if (le + total > c->node_cap) {
            const int ncap = (int)std::min<long long>(node_limit, std::max<long long>(2LL * c->node_cap, (long long)le + total + 1024));
            if (alloc_nodes(c, ncap, le, w)) return 1;
        }
        k_level_emit<DIM><<<cdiv((long long)w << DIM, B), B, 0, c->stream>>>(c->tb, keys, lb, le, c->P.leaf_num, max_level_eff, c->P.key_levels, c->lvl_offs, c->d_root, nullptr, nullptr);
        LAUNCH_CHECK();
        lb = le; le += total;
    }
    if (sort_levels < c->P.key_levels) {
        int deeper = 0;
        CK(cudaMemcpyAsync(&deeper, c->d_lvl_bad + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (deeper) { c->force_full_sort = true; const int r = make_tree_t<DIM>(c); c->force_full_sort = false; return r; }
    }
    }
    const int n_nodes = lb;
    if (!dist) {
        for (int l = (int)c->levels.size() - 1; l >= 0; --l) {
            const int a = c->levels[l].first, b = c->levels[l].second;
            k_level_up<DIM><<<cdiv(b - a, B), B, 0, c->stream>>>(c->tb, c->rc.posm, a, b, 0, N, 0); LAUNCH_CHECK();
        }
    } else {
        // partial sums of the leaves over the own particles, summed across ranks, then the internal nodes bottom-up
        k_level_up<DIM><<<cdiv(n_nodes, B), B, 0, c->stream>>>(c->tb, c->rc.posm, 0, n_nodes, c->off, c->off + n, 1); LAUNCH_CHECK();
        {
            Timer tx(c, SPHB_T_EXCHANGE), ty(c, SPHB_T_REDUCE);
            CKN(g_nccl.AllReduce(c->tb.msum4, c->tb.msum4, (size_t)n_nodes * 4, ncclFloat64, ncclSum, c->comm, c->stream));
        }
        for (int l = (int)c->levels.size() - 1; l >= 0; --l) {
            const int a = c->levels[l].first, b = c->levels[l].second;
            k_level_up<DIM><<<cdiv(b - a, B), B, 0, c->stream>>>(c->tb, c->rc.posm, a, b, 0, N, 2); LAUNCH_CHECK();
        }
    }
    c->td.n_nodes = n_nodes;
    k_tree_scatter<DIM><<<cdiv(n_nodes, B), B, 0, c->stream>>>(c->tb, c->td, n_nodes, c->d_root); LAUNCH_CHECK();
    // particle groups (sphb_tree.cuh) of the own range: flags at group starts -> ascending list of starts
    if (dist) {
        CK(cudaMemsetAsync(c->d_ncells, 0, 2 * sizeof(int), c->stream));
        CK(cudaMemsetAsync(c->halo_flags, 0, (size_t)n_nodes + 8, c->stream));
        CK(cudaMemsetAsync(c->halo_have, 0, (size_t)n_nodes + 8, c->stream));
    }
    for (int kind = 0; kind < (c->P.use_gravity ? 2 : 1); ++kind) {
        CK(cudaMemsetAsync(c->grp_flags, 0, (size_t)n + 1, c->stream));
        k_group_flags_own<<<cdiv(n_nodes, B), B, 0, c->stream>>>(c->tb, n_nodes, c->grp_flags, c->off, c->off + n,
                                                              kind == 0 ? GROUP_CELL_SPH : GROUP_CELL_GRAV,
                                                              (dist && kind == 0) ? c->cells_s : nullptr, (dist && kind == 0) ? c->d_ncells : nullptr,   /* halo marking works on the SPH group cells */
                                                              kind == 1 && c->grav_mode == 2 ? GRAV2_GROUP : 32);
        LAUNCH_CHECK();
        if (n > 0) {
            size_t tb = c->cub_tmp_bytes;
            CK(cub::DeviceSelect::Flagged(c->cub_tmp, tb, thrust::counting_iterator<int>(c->off), c->grp_flags,
                                          kind == 0 ? c->grp_start : c->grp_start_g, kind == 0 ? c->d_ngroups : c->d_ngroups_g, n, c->stream));
            ++c->launches;
        } else CK(cudaMemsetAsync(kind == 0 ? c->d_ngroups : c->d_ngroups_g, 0, sizeof(int), c->stream));
    }
    c->tree_valid = true;
    c->ksize_valid = false; c->hsoft_valid = false;
    c->last_counters.tree_nodes = (uint64_t)n_nodes;
    return 0;
}

// BHTree::set_kernel: leaf maxima over the own particles -> parents; multi-GPU: max across ranks; -> walk records
int set_kernel(sphb_ctx * c)
{
    const int nn = c->td.n_nodes;
    k_clear_kernel<<<cdiv(nn, 256), 256, 0, c->stream>>>(c->td); LAUNCH_CHECK();
    k_set_kernel<<<cdiv(nn, 256), 256, 0, c->stream>>>(c->td, c->gv.sml, c->off, c->off + c->n); LAUNCH_CHECK();
    if (c->world > 1) {
        Timer tx(c, SPHB_T_EXCHANGE), ty(c, SPHB_T_REDUCE);
        CKN(g_nccl.AllReduce(c->td.ksize, c->td.ksize, (size_t)nn, ncclFloat64, ncclMax, c->comm, c->stream));
    }
    k_apply_ksize<<<cdiv(nn, 256), 256, 0, c->stream>>>(c->td); LAUNCH_CHECK();
    c->ksize_valid = true;
    return 0;
}

// ---- multi-GPU halo: mark the remote leaves the own group cells can reach, then pull their records from the owners ------
// phase 0: gather search of initial_smoothing (radius h_guess), 1: gather search of PreInteraction (h_guess * kernel_ratio),
// 2: symmetric search of FluidForce + leaves the gravity walk may open.
template <int DIM> int halo_t(sphb_ctx * c, int phase)
{
    if (c->world == 1) return 0;
    Timer tx(c, SPHB_T_EXCHANGE), ty(c, SPHB_T_HALO);
    const int grid = c->sm_count * 4;
    if (phase < 2) {
        k_mark_halo<DIM><<<grid, 128, 0, c->stream>>>(c->td, c->P, c->cells_s, c->d_ncells + 0, 0, 1, phase == 0 ? 1.0 : c->P.kernel_ratio,
                                                       c->gv.mass, c->gv.dens, c->off, c->off + c->n, c->halo_flags, c->d_err);
        LAUNCH_CHECK();
    } else {
        // symmetric search and gravity opening in ONE descent per (small) group cell: the gravity criterion from the cube of a
        // <= 1024-particle cell opens fewer nodes than from a 4096-particle gravity cell and spreads over more warps
        k_mark_halo<DIM><<<grid, 128, 0, c->stream>>>(c->td, c->P, c->cells_s, c->d_ncells + 0, 1, c->P.use_gravity ? 3 : 1, 1.0,
                                                       c->gv.mass, c->gv.dens, c->off, c->off + c->n, c->halo_flags, c->d_err);
        LAUNCH_CHECK();
    }
    if (nccl_barrier(c)) return 1;                      // the owners' records of this phase are written
    const int need_sph = phase < 2 ? (PULL_POSM | PULL_VELC | PULL_THERMO_A) : (PULL_POSM | PULL_VELC | PULL_THERMO_B);
    const int need_grav = phase < 2 ? 0 : (PULL_POSM | PULL_HSOFT);
    k_pull_halo<<<cdiv(c->td.n_nodes, 128), 128, 0, c->stream>>>(c->td, c->pt, c->lay, c->slab, c->halo_flags, c->halo_have, need_sph, need_grav, c->d_err + 3);
    LAUNCH_CHECK();
    if (phase == 2) { if (nccl_barrier(c)) return 1; }  // all pulls of the step are done: the owners may rewrite their records
    return 0;
}
int halo(sphb_ctx * c, int phase) { return c->dim == 1 ? halo_t<1>(c, phase) : c->dim == 2 ? halo_t<2>(c, phase) : halo_t<3>(c, phase); }

} // namespace

namespace {

template <int DIM, int KT, int SPH> int pre_t(sphb_ctx * c)
{
    Timer tm(c, SPHB_T_PRE);
    GroupTable gt;
    if (c->first_pre) {
        // initial_smoothing (src/pre_interaction.cpp:171-215), first call only
        if (ensure_recs(c) || halo(c, 0) || group_table(c, gt)) return 1;
        k_initial_smoothing<DIM, KT><<<c->pre_grid, 128, 0, c->stream>>>(c->gv, c->rc, c->td, c->P, gt, c->d_err); LAUNCH_CHECK();
        c->first_pre = false;
        c->recs_dirty = true;             // dens changed
    }
    if (ensure_recs(c) || halo(c, 1)) return 1;
    k_set_scalars<<<1, 1, 0, c->stream>>>(c->d_scal, 0.0, DBL_MAX, 2); LAUNCH_CHECK();
    if (group_table(c, gt)) return 1;
    k_pre_interaction<DIM, KT, SPH><<<c->pre_grid, 128, 0, c->stream>>>(c->gv, c->rc, c->td, c->P, gt,
        c->scratch_r, c->scratch_m, c->scratch_j, c->d_scal + 0, c->d_scal + 1, c->d_err, c->counters_on ? c->d_cnt : nullptr);
    LAUNCH_CHECK();
    if (c->P.use_gravity && c->n > 0) { k_grav_pack<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->cur.sml, c->rc.hsoft + c->off, c->n); LAUNCH_CHECK(); }
    c->hsoft_valid = true;
    if (c->world > 1) {
        Timer tx(c, SPHB_T_EXCHANGE), ty(c, SPHB_T_REDUCE);
        CKN(g_nccl.AllReduce(c->d_scal + 1, c->d_scal + 1, 1, ncclFloat64, ncclMin, c->comm, c->stream));
    }
    if (set_kernel(c)) return 1;
    return halo(c, 2);
}

template <int DIM> int pre_d(sphb_ctx * c)
{
    const int kt = c->P.kernel, st = c->P.sph_type;
#define SPHB_PRE(K, S) if (kt == K && st == S) return pre_t<DIM, K, S>(c);
    SPHB_PRE(K_CUBIC, T_SSPH) SPHB_PRE(K_CUBIC, T_DISPH) SPHB_PRE(K_CUBIC, T_GSPH)
    if (DIM > 1) {
        constexpr int D2 = DIM > 1 ? DIM : 2;
        if (kt == K_WENDLAND && st == T_SSPH) return pre_t<D2, K_WENDLAND, T_SSPH>(c);
        if (kt == K_WENDLAND && st == T_DISPH) return pre_t<D2, K_WENDLAND, T_DISPH>(c);
        if (kt == K_WENDLAND && st == T_GSPH) return pre_t<D2, K_WENDLAND, T_GSPH>(c);
    }
#undef SPHB_PRE
    c->err = "unsupported kernel / SPH type"; return 1;
}

// stages called on their own (module mode) after an upload: records and kernel sizes may be stale
int refresh_for_forces(sphb_ctx * c)
{
    if (c->recs_dirty || !c->hsoft_valid) {
        if (ensure_recs(c)) return 1;
        if (c->P.use_gravity && c->n > 0) { k_grav_pack<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->cur.sml, c->rc.hsoft + c->off, c->n); LAUNCH_CHECK(); }
        c->hsoft_valid = true;
        c->ksize_valid = false;
    }
    if (!c->ksize_valid) {
        if (set_kernel(c)) return 1;
        if (halo(c, 2)) return 1;
    }
    return 0;
}

template <int DIM, int KT, int SPH> int force_t(sphb_ctx * c)
{
    Timer tm(c, SPHB_T_FLUID);
    if (refresh_for_forces(c)) return 1;
    GroupTable gt;
    if (group_table(c, gt)) return 1;
    k_fluid_force<DIM, KT, SPH><<<std::min(c->pre_grid, c->sm_count * FF_BLOCKS), 128, 0, c->stream>>>(c->gv, c->rc, c->td, c->P, gt,
        c->scratch_j, c->d_scal + 0, c->d_err, c->counters_on ? c->d_cnt : nullptr);
    LAUNCH_CHECK();
    return 0;
}
template <int DIM> int force_d(sphb_ctx * c)
{
    const int kt = c->P.kernel, st = c->P.sph_type;
    if (kt == K_CUBIC && st == T_SSPH) return force_t<DIM, K_CUBIC, T_SSPH>(c);
    if (kt == K_CUBIC && st == T_DISPH) return force_t<DIM, K_CUBIC, T_DISPH>(c);
    if (kt == K_CUBIC && st == T_GSPH) return force_t<DIM, K_CUBIC, T_GSPH>(c);
    if (DIM > 1) {
        constexpr int D2 = DIM > 1 ? DIM : 2;
        if (kt == K_WENDLAND && st == T_SSPH) return force_t<D2, K_WENDLAND, T_SSPH>(c);
        if (kt == K_WENDLAND && st == T_DISPH) return force_t<D2, K_WENDLAND, T_DISPH>(c);
        if (kt == K_WENDLAND && st == T_GSPH) return force_t<D2, K_WENDLAND, T_GSPH>(c);
    }
    c->err = "unsupported kernel / SPH type"; return 1;
}

template <int DIM> int gravity_t(sphb_ctx * c, bool direct, int k_targets = 0)
{
    Timer tm(c, SPHB_T_GRAVITY);
    if (direct) {
        if (c->world > 1) { c->err = "sphb_gravity_direct is single-GPU only"; return 1; }
        if (k_targets > 0 && k_targets < c->n) {
            // targets = particles 0..k-1 of the caller's buffer: list of their sorted-order indices in idx (tree scratch)
            int * cnt = c->d_grp_ctl;
            CK(cudaMemsetAsync(cnt, 0, sizeof(int), c->stream));
            k_select_targets<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->cur.orig, c->n, k_targets, c->idx, cnt); LAUNCH_CHECK();
            k_gravity_direct<DIM><<<cdiv(k_targets, 128), 128, 0, c->stream>>>(c->cur, c->P, c->n, c->idx, k_targets); LAUNCH_CHECK();
            return 0;
        }
        k_gravity_direct<DIM><<<cdiv(c->n, 128), 128, 0, c->stream>>>(c->cur, c->P, c->n, nullptr, c->n); LAUNCH_CHECK();
        return 0;
    }
    if (refresh_for_forces(c)) return 1;
    GroupTable gt;
    if (group_table(c, gt, true)) return 1;
    if (c->grav_mode == 2) {
        const int smem = (int)(4 * sizeof(Grav2Smem));
        const int grid = std::min(c->grav_grid, c->sm_count * GV2_BLOCKS);
#define SPHB_GRAV(PER, CNT) do { \
            bool & attr_set = c->grav_attr_set[PER ? 1 : 0][CNT ? 1 : 0];      /* per context: the attribute is per device */ \
            if (!attr_set) { CK(cudaFuncSetAttribute(k_gravity2<DIM, PER, CNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; } \
            k_gravity2<DIM, PER, CNT><<<grid, 128, smem, c->stream>>>(c->gv, c->td, c->P, gt, c->rc.posm, c->rc.hsoft, \
                c->grav_lq, c->grav_near, c->d_cnt, c->d_err); } while (0)
        if (c->P.periodic) { if (c->counters_on) SPHB_GRAV(true, true); else SPHB_GRAV(true, false); }
        else               { if (c->counters_on) SPHB_GRAV(false, true); else SPHB_GRAV(false, false); }
#undef SPHB_GRAV
        LAUNCH_CHECK();
        return 0;
    }
    const int smem = (int)(4 * sizeof(GravSmem));
#define SPHB_GRAV(PER, CNT) do { \
        bool & attr_set = c->grav_attr_set[PER ? 1 : 0][CNT ? 1 : 0];      /* per context: the attribute is per device */ \
        if (!attr_set) { CK(cudaFuncSetAttribute(k_gravity<DIM, PER, CNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; } \
        k_gravity<DIM, PER, CNT><<<c->grav_grid, 128, smem, c->stream>>>(c->gv, c->td, c->P, gt, c->rc.posm, c->rc.hsoft, \
            c->grav_lq, c->grav_near, c->d_cnt, c->d_err); } while (0)
    if (c->P.periodic) { if (c->counters_on) SPHB_GRAV(true, true); else SPHB_GRAV(true, false); }
    else               { if (c->counters_on) SPHB_GRAV(false, true); else SPHB_GRAV(false, false); }
#undef SPHB_GRAV
    LAUNCH_CHECK();
    return 0;
}

template <int DIM> int timestep_t(sphb_ctx * c)
{
    Timer tm(c, SPHB_T_TIMESTEP);
    // the force minimum is taken over the own particles and all-reduced (min)
    const double big = DBL_MAX;
    CK(cudaMemcpyAsync(c->d_scal + 2, &big, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (c->n > 0) {
        const int grid = std::min(cdiv(c->n, 256), 4 * c->sm_count);
        k_timestep_partial<DIM><<<grid, 256, 0, c->stream>>>(c->cur, 0, c->n, c->P.cfl_force, c->d_scal + 2);
        LAUNCH_CHECK();
    }
    if (c->world > 1) CKN(g_nccl.AllReduce(c->d_scal + 2, c->d_scal + 2, 1, ncclFloat64, ncclMin, c->comm, c->stream));
    k_timestep_final<<<1, 1, 0, c->stream>>>(c->d_scal + 2, c->d_scal + 1, c->P.cfl_sound, c->d_scal + 0); LAUNCH_CHECK();
    return 0;
}

template <int DIM> int predict_t(sphb_ctx * c)
{
    Timer tm(c, SPHB_T_PREDICT);
    if (c->n > 0) { k_predict<DIM><<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->cur, c->P, c->n, c->d_scal + 0); LAUNCH_CHECK(); }
    c->tree_valid = false;
    c->recs_dirty = true;
    return 0;
}
template <int DIM> int correct_t(sphb_ctx * c)
{
    Timer tm(c, SPHB_T_CORRECT);
    if (c->n > 0) { k_correct<DIM><<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->cur, c->P, c->n, c->d_scal + 0); LAUNCH_CHECK(); }
    c->recs_dirty = true;
    return 0;
}

#define DIM_SWITCH(c, EXPR1, EXPR2, EXPR3) ((c)->dim == 1 ? (EXPR1) : (c)->dim == 2 ? (EXPR2) : (EXPR3))

int make_tree(sphb_ctx * c) { return DIM_SWITCH(c, make_tree_t<1>(c), make_tree_t<2>(c), make_tree_t<3>(c)); }
int pre(sphb_ctx * c) { return DIM_SWITCH(c, pre_d<1>(c), pre_d<2>(c), pre_d<3>(c)); }
int force(sphb_ctx * c) { return DIM_SWITCH(c, force_d<1>(c), force_d<2>(c), force_d<3>(c)); }
int gravity(sphb_ctx * c, bool direct, int k = 0) { return DIM_SWITCH(c, gravity_t<1>(c, direct, k), gravity_t<2>(c, direct, k), gravity_t<3>(c, direct, k)); }
int timestep(sphb_ctx * c) { return DIM_SWITCH(c, timestep_t<1>(c), timestep_t<2>(c), timestep_t<3>(c)); }
int predict(sphb_ctx * c) { return DIM_SWITCH(c, predict_t<1>(c), predict_t<2>(c), predict_t<3>(c)); }
int correct(sphb_ctx * c) { return DIM_SWITCH(c, correct_t<1>(c), correct_t<2>(c), correct_t<3>(c)); }

// copy dt, h_per_v_sig and the error counters to the host; turn device-side errors into a status.
// collective = every rank is in this call (stage entry points): the error words are all-reduced first, so that all
// ranks leave with the same status instead of one returning early and the others waiting in the next collective.
int sync_scalars(sphb_ctx * c, bool collective = false)
{
    double s[2]; unsigned long long e[4];
    if (collective && c->world > 1) CKN(g_nccl.AllReduce(c->d_err + 1, c->d_err + 1, 2, ncclUint64, ncclMax, c->comm, c->stream));
    CK(cudaMemcpyAsync(s, c->d_scal, sizeof(s), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(e, c->d_err, sizeof(e), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->dt = s[0]; c->hpvs = s[1];
    c->nonconverged_total = e[0];
    c->halo_pulled = e[3];
    if (e[1]) {
        c->err = "neighbor list overflow: a particle has more than neighborNumber*20 candidates (include/defines.hpp:29)";
        CK(cudaMemsetAsync(c->d_err + 1, 0, sizeof(unsigned long long), c->stream));
        return 1;
    }
    if (e[2]) {
        c->err = (e[2] & WALK_ERR_GRAV_STACK) ? "gravity walk: node stack overflow" : "neighbour walk: node stack overflow";
        CK(cudaMemsetAsync(c->d_err + 2, 0, sizeof(unsigned long long), c->stream));
        return 1;
    }
    if (c->timers_on) read_timers(c);
    return 0;
}

int require_tree(sphb_ctx * c)
{
    if (!c->n_glob) { c->err = "no particles uploaded"; return 1; }
    if (!c->tree_valid) { c->err = "tree is not made for the current positions (call sphb_make_tree)"; return 1; }
    return 0;
}

} // namespace

// =====================================================================================================
extern "C" {

size_t sphb_sizeof_particle(int dim) { return rec_size(dim); }

int sphb_create(const sphb_params * hp, int dim, int device, sphb_ctx ** out)
{
    if (!hp || !out || dim < 1 || dim > 3) { g_create_error = "bad arguments"; return 1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_create_error = "no CUDA device (libsphb has no CPU fallback)"; return 1; }
    if (device < 0 || device >= ndev) { g_create_error = "bad device ordinal"; return 1; }
    if (hp->kernel == SPHB_WENDLAND && dim == 1) { g_create_error = "Wendland C4 is not defined for DIM == 1 (wendland_kernel.hpp:25-28)"; return 1; }
    if (hp->kernel != SPHB_CUBIC_SPLINE && hp->kernel != SPHB_WENDLAND) { g_create_error = "kernel is unknown."; return 1; }   // src/simulation.cpp:19
    if (hp->sph_type < 0 || hp->sph_type > 2) { g_create_error = "Unknown SPH type"; return 1; }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return 1; }
    sphb_ctx * c = new sphb_ctx;
    c->dim = dim; c->device = device; c->hp = *hp;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->sm_count = prop.multiProcessorCount;
    DevParams & P = c->P;
    P.sph_type = hp->sph_type; P.kernel = hp->kernel;
    P.cfl_sound = hp->cfl_sound; P.cfl_force = hp->cfl_force;
    P.av_alpha = hp->av_alpha; P.use_balsara = hp->use_balsara_switch; P.use_tdav = hp->use_time_dependent_av;
    P.alpha_max = hp->alpha_max; P.alpha_min = hp->alpha_min; P.epsilon_av = hp->epsilon_av;
    P.use_ac = hp->use_ac; P.alpha_ac = hp->alpha_ac;
    P.max_level = hp->max_tree_level; P.leaf_num = hp->leaf_particle_num; P.ngb = hp->neighbor_number;
    P.iterative = hp->iterative_sml; P.gamma = hp->gamma;
    P.periodic = hp->periodic; P.use_gravity = hp->use_gravity;
    for (int d = 0; d < 3; ++d) {
        P.rmax[d] = d < dim ? hp->range_max[d] : 0.0; P.rmin[d] = d < dim ? hp->range_min[d] : 0.0;
        P.range[d] = P.rmax[d] - P.rmin[d];
    }
    P.G = hp->G; P.theta = hp->theta; P.theta2 = hp->theta * hp->theta;
    P.gsph2 = hp->gsph_2nd_order;
    P.kernel_ratio = hp->iterative_sml ? 1.2 : 1.0;
    // the octree key holds maxTreeLevel * DIM bits (+ one spare bit for the multi-GPU migration): 20 (the default) fits
    // every DIM; deeper trees than the key can describe are refused instead of silently truncated
    if (hp->max_tree_level < 1 || (long long)hp->max_tree_level * dim > 63) {
        g_create_error = "maxTreeLevel * DIM must be in [1, 63] (the 64-bit octree key cannot describe a deeper tree)";
        delete c; return 1;
    }
    P.key_levels = hp->max_tree_level;
    P.list_cap = hp->neighbor_number * 20;
    if (const char * gm = std::getenv("SPHB_GRAVITY")) c->grav_mode = std::atoi(gm) == 2 ? 2 : 1;      // developer A/B switch
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    c->stream = c->own_stream;
    auto alloc = [&](auto ** p, size_t bytes) { void * q = nullptr; if (ok && cudaMalloc(&q, bytes) != cudaSuccess) ok = false; *p = static_cast<std::remove_reference_t<decltype(**p)> *>(q); };
    alloc(&c->d_root, 4 * sizeof(double));
    alloc(&c->d_scal, 16 * sizeof(double));
    alloc(&c->d_err, 4 * sizeof(unsigned long long));
    alloc(&c->d_cnt, sizeof(Counters));
    alloc(&c->d_lvl, (SPHB_MAX_LEVELS + 4) * sizeof(int));
    alloc(&c->d_lvl_bad, 2 * sizeof(int));
    alloc(&c->d_ngroups, sizeof(int));
    alloc(&c->d_ngroups_g, sizeof(int));
    alloc(&c->d_grp_ctl, 2 * sizeof(int));
    if (ok) {
        ok = cudaMemset(c->d_scal, 0, 16 * sizeof(double)) == cudaSuccess && cudaMemset(c->d_err, 0, 4 * sizeof(unsigned long long)) == cudaSuccess &&
             cudaMemset(c->d_cnt, 0, sizeof(Counters)) == cudaSuccess;
    }
    if (ok && P.periodic) {
        // BHTree::initialize, src/bhtree.cpp:19-31
        double root[4] = {0, 0, 0, 0};
        double l = 0.0;
        for (int d = 0; d < dim; ++d) {
            root[d] = (P.rmax[d] + P.rmin[d]) * 0.5;
            if (l < P.range[d]) l = P.range[d];
        }
        root[3] = l;
        ok = cudaMemcpy(c->d_root, root, sizeof(root), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    for (int k = 0; ok && k < SPHB_T_COUNT; ++k)
        for (int sgm = 0; ok && sgm < SPHB_TSEG; ++sgm)
            ok = cudaEventCreate(&c->ev[k][sgm][0]) == cudaSuccess && cudaEventCreate(&c->ev[k][sgm][1]) == cudaSuccess;
    if (!ok) {
        g_create_error = std::string("sphb_create: ") + cudaGetErrorString(cudaGetLastError());
        sphb_destroy(c);
        return 1;
    }
    *out = c;
    return 0;
}

void sphb_destroy(sphb_ctx * c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->comm && c->own_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    close_peers(c);
    free_bag(c->allocs); free_bag(c->node_allocs);
    cudaFree(c->d_root); cudaFree(c->d_scal); cudaFree(c->d_err); cudaFree(c->d_cnt); cudaFree(c->d_lvl); cudaFree(c->d_lvl_bad); cudaFree(c->d_ngroups); cudaFree(c->d_ngroups_g); cudaFree(c->d_grp_ctl);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (auto & k : c->ev) for (auto & sgm : k) for (auto & ev : sgm) if (ev) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char * sphb_last_error(const sphb_ctx * c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int sphb_set_stream(sphb_ctx * c, void * s)
{
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;          // NULL: back to the context's own stream
    return 0;
}
int sphb_synchronize(sphb_ctx * c) { CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream)); return 0; }
int sphb_dim(const sphb_ctx * c) { return c->dim; }
int sphb_particle_num(const sphb_ctx * c) { return c->n; }
long long sphb_global_particle_num(const sphb_ctx * c) { return c->n_glob; }
long long sphb_first_global_index(const sphb_ctx * c) { return c->off; }

int sphb_nccl_unique_id(void * out128)
{
    std::string err;
    if (!g_nccl.load(err)) { g_create_error = err; return 1; }
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) { g_create_error = "ncclGetUniqueId failed"; return 1; }
    std::memcpy(out128, &id, sizeof(id));
    return 0;
}

int sphb_set_distributed(sphb_ctx * c, int rank, int world, void * nccl_comm)
{
    if (world < 1 || rank < 0 || rank >= world) { c->err = "bad rank / world"; return 1; }
    if (c->n_glob) { c->err = "sphb_set_distributed must precede the first upload"; return 1; }
    if (world > 1) {
        if (!g_nccl.load(c->err)) return 1;
        if (!nccl_comm) { c->err = "nccl communicator required"; return 1; }
        c->comm = (ncclComm_t)nccl_comm; c->own_comm = false;
    }
    c->rank = rank; c->world = world;
    return 0;
}

int sphb_set_distributed_id(sphb_ctx * c, int rank, int world, const void * unique_id128)
{
    if (world < 1 || rank < 0 || rank >= world) { c->err = "bad rank / world"; return 1; }
    if (c->n_glob) { c->err = "sphb_set_distributed_id must precede the first upload"; return 1; }
    if (world > 1) {
        if (!g_nccl.load(c->err)) return 1;
        CK(cudaSetDevice(c->device));
        ncclUniqueId id;
        std::memcpy(&id, unique_id128, sizeof(id));
        ncclComm_t comm = nullptr;
        CKN(g_nccl.CommInitRank(&comm, world, id, rank));
        c->comm = comm; c->own_comm = true;
    }
    c->rank = rank; c->world = world;
    return 0;
}

// ---- state transfer -----------------------------------------------------------------------------------
// Device-visible address of a caller buffer when it is pinned / registered host memory (sphb_host_alloc,
// cudaHostRegister): the pack / unpack kernels then read and write the caller's SPHParticle records in place over
// PCIe, touching only the members in the field mask.  nullptr for pageable memory (staged copies instead).
constexpr int ZERO_COPY_MAX = 1 << 18;      // records up to which a masked transfer runs in place over PCIe
static void * mapped_host(const void * p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type == cudaMemoryTypeHost && a.devicePointer) return a.devicePointer;
    return nullptr;
}

int sphb_upload_aos(sphb_ctx * c, const void * particles, int n, size_t stride, uint32_t mask)
{
    CK(cudaSetDevice(c->device));
    const size_t rec = rec_size(c->dim);
    if (!particles || n <= 0 || stride < rec) { c->err = "bad upload arguments"; return 1; }
    const bool first = (c->n_glob == 0 || n != c->n);
    if (first) {
        if ((mask & SPHB_F_ALL) != SPHB_F_ALL) { c->err = "the first upload must use SPHB_F_ALL"; return 1; }
        if (alloc_particles(c, n)) {
            // out of memory (or a failed peer mapping) half-way: drop everything, the context is empty again
            close_peers(c);
            free_bag(c->allocs);
            c->cur = PSoA{}; c->alt = PSoA{}; c->rc = Recs{};
            c->n = 0; c->n_glob = 0; c->off = 0; c->tree_valid = false;
            return 1;
        }
    } else if (!c->orig_valid) {
        c->err = "multi-GPU mode: particles migrated since the last download; download (which renumbers the rank's particles) before uploading into them";
        return 1;
    }
    const bool full = (mask & SPHB_F_ALL) == SPHB_F_ALL;
    const char * src = (const char *)c->d_aos;
    size_t src_stride = rec;
    // partial mask from pinned memory: in place over PCIe only for small sets — above ZERO_COPY_MAX records the whole-record
    // DMA (51 GB/s) beats the per-member reads of the kernel (one PCIe transaction per member run of a 208-byte record)
    void * mp = (full || n > ZERO_COPY_MAX) ? nullptr : mapped_host(particles);
    if (mp) { src = (const char *)mp; src_stride = stride; }            // masked upload straight from pinned host memory
    else if (stride == rec) CK(cudaMemcpyAsync(c->d_aos, particles, rec * n, cudaMemcpyHostToDevice, c->stream));
    else CK(cudaMemcpy2DAsync(c->d_aos, rec, particles, stride, rec, n, cudaMemcpyHostToDevice, c->stream));
    switch (c->dim) {
    case 1: k_unpack<1><<<cdiv(n, 256), 256, 0, c->stream>>>(src, src_stride, c->cur, n, mask, first); break;
    case 2: k_unpack<2><<<cdiv(n, 256), 256, 0, c->stream>>>(src, src_stride, c->cur, n, mask, first); break;
    default: k_unpack<3><<<cdiv(n, 256), 256, 0, c->stream>>>(src, src_stride, c->cur, n, mask, first); break;
    }
    LAUNCH_CHECK();
    c->recs_dirty = true;
    if (mask & SPHB_F_SML) { c->ksize_valid = false; c->hsoft_valid = false; }
    if (mask & (SPHB_F_POS | SPHB_F_MASS)) c->tree_valid = false;
    CK(cudaStreamSynchronize(c->stream));        // the caller may reuse its buffer
    return 0;
}

int sphb_download_aos(sphb_ctx * c, void * particles, int n, size_t stride, uint32_t mask)
{
    CK(cudaSetDevice(c->device));
    const size_t rec = rec_size(c->dim);
    if (!particles || n != c->n || stride < rec) { c->err = "bad download arguments (n must be sphb_particle_num())"; return 1; }
    if (!c->orig_valid) {
        // multi-GPU: the rank's set changed by migration; the caller's record k is now the k-th own particle in tree order
        k_renumber_orig<<<cdiv(n, 256), 256, 0, c->stream>>>(c->cur.orig, n); LAUNCH_CHECK();
        c->orig_valid = true;
    }
    const bool full = (mask & SPHB_F_ALL) == SPHB_F_ALL;
    void * mp = (full || n > ZERO_COPY_MAX) ? nullptr : mapped_host(particles);
    char * dst = mp ? (char *)mp : (char *)c->d_aos;
    const size_t dst_stride = mp ? stride : rec;
    // staged path of a masked download: every member is packed, the host picks the masked ones out of the staging copy
    const uint32_t pack_mask = mp ? mask : (uint32_t)SPHB_F_ALL;
    switch (c->dim) {
    case 1: k_pack<1><<<cdiv(n, 256), 256, 0, c->stream>>>(dst, dst_stride, c->cur, n, pack_mask); break;
    case 2: k_pack<2><<<cdiv(n, 256), 256, 0, c->stream>>>(dst, dst_stride, c->cur, n, pack_mask); break;
    default: k_pack<3><<<cdiv(n, 256), 256, 0, c->stream>>>(dst, dst_stride, c->cur, n, pack_mask); break;
    }
    LAUNCH_CHECK();
    if (mp) { CK(cudaStreamSynchronize(c->stream)); return 0; }          // written in place over PCIe
    if (full) {
        if (stride == rec) CK(cudaMemcpyAsync(particles, c->d_aos, rec * n, cudaMemcpyDeviceToHost, c->stream));
        else CK(cudaMemcpy2DAsync(particles, stride, c->d_aos, rec, rec, n, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    // pageable destination, partial mask: staged copy of the records, then the masked members are copied by host threads
    if (c->h_stage_bytes < rec * n) {
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr; c->h_stage_bytes = 0;
        CK(cudaMallocHost(&c->h_stage, rec * n));
        c->h_stage_bytes = rec * n;
    }
    CK(cudaMemcpyAsync(c->h_stage, c->d_aos, rec * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int D = c->dim;
    // contiguous byte runs of the selected members within a record
    std::vector<std::pair<size_t, size_t>> runs;
    auto add = [&](size_t o, size_t len) { if (!runs.empty() && runs.back().first + runs.back().second == o) runs.back().second += len; else runs.push_back({o, len}); };
    const uint32_t vec_bits[4] = {SPHB_F_POS, SPHB_F_VEL, SPHB_F_VEL_P, SPHB_F_ACC};
    for (int v = 0; v < 4; ++v) if (mask & vec_bits[v]) add((size_t)v * D * 8, (size_t)D * 8);
    for (int k = 0; k < 12; ++k) if (mask & (SPHB_F_MASS << k)) add((size_t)(4 * D + k) * 8, 8);
    if (mask & SPHB_F_ID) add((size_t)(4 * D + 12) * 8, 4);
    if (mask & SPHB_F_NEIGHBOR) add((size_t)(4 * D + 12) * 8 + 4, 4);
    const int nth = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto work = [&](int t) {
        const long long i0 = (long long)n * t / nth, i1 = (long long)n * (t + 1) / nth;
        for (long long i = i0; i < i1; ++i) {
            const char * s_ = (const char *)c->h_stage + (size_t)i * rec;
            char * d_ = (char *)particles + (size_t)i * stride;
            for (const auto & r : runs) std::memcpy(d_ + r.first, s_ + r.first, r.second);
        }
    };
    if (nth == 1 || n < (1 << 16)) { for (int t = 0; t < nth; ++t) work(t); }
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nth; ++t) th.emplace_back(work, t);
        for (auto & t : th) t.join();
    }
    return 0;
}

static double ** vector_array_by_name(sphb_ctx * c, const char * name)
{
    if (c->P.sph_type != T_GSPH) { c->err = std::string("additional_vector_array does not have ") + name; return nullptr; }   // src/simulation.cpp:74
    PSoA & p = c->cur;
    if (!std::strcmp(name, "grad_density")) return p.grad_d;
    if (!std::strcmp(name, "grad_pressure")) return p.grad_p;
    if (!std::strncmp(name, "grad_velocity_", 14)) {
        const int k = name[14] - '0';
        if (k >= 0 && k < c->dim && name[15] == 0) return p.grad_v[k];
    }
    c->err = std::string("additional_vector_array does not have ") + name;
    return nullptr;
}

int sphb_get_vector_array(sphb_ctx * c, const char * name, double * out)
{
    CK(cudaSetDevice(c->device));
    double ** a = vector_array_by_name(c, name);
    if (!a) return 1;
    double * tmp = (double *)c->d_aos;      // n*dim doubles fit in the AoS staging buffer
    for (int k = 0; k < c->dim; ++k) { k_scatter_by_orig<<<cdiv(c->n, 256), 256, 0, c->stream>>>(a[k], c->cur.orig, tmp, c->n, c->dim, k); LAUNCH_CHECK(); }
    CK(cudaMemcpyAsync(out, tmp, (size_t)c->n * c->dim * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int sphb_set_vector_array(sphb_ctx * c, const char * name, const double * in)
{
    CK(cudaSetDevice(c->device));
    double ** a = vector_array_by_name(c, name);
    if (!a) return 1;
    double * tmp = (double *)c->d_aos;
    CK(cudaMemcpyAsync(tmp, in, (size_t)c->n * c->dim * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    for (int k = 0; k < c->dim; ++k) { k_gather_by_orig<<<cdiv(c->n, 256), 256, 0, c->stream>>>(tmp, c->cur.orig, a[k], c->n, c->dim, k); LAUNCH_CHECK(); }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int sphb_set_dt(sphb_ctx * c, double dt)
{
    CK(cudaSetDevice(c->device));
    c->dt = dt;
    k_set_scalars<<<1, 1, 0, c->stream>>>(c->d_scal, dt, 0.0, 1); LAUNCH_CHECK();
    return 0;
}
int sphb_get_dt(sphb_ctx * c, double * dt) { CK(cudaSetDevice(c->device)); if (sync_scalars(c)) return 1; *dt = c->dt; return 0; }
int sphb_set_h_per_v_sig(sphb_ctx * c, double v)
{
    CK(cudaSetDevice(c->device));
    c->hpvs = v;
    k_set_scalars<<<1, 1, 0, c->stream>>>(c->d_scal, 0.0, v, 2); LAUNCH_CHECK();
    return 0;
}
int sphb_get_h_per_v_sig(sphb_ctx * c, double * v) { CK(cudaSetDevice(c->device)); if (sync_scalars(c)) return 1; *v = c->hpvs; return 0; }

// ---- the hot path -----------------------------------------------------------------------------------------
// In the multi-GPU mode every entry point below is COLLECTIVE: all ranks call it, in the same order.
int sphb_init_state(sphb_ctx * c)
{
    CK(cudaSetDevice(c->device));
    if (!c->n_glob) { c->err = "no particles uploaded"; return 1; }
    if (c->n > 0) { k_init_state<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->cur, c->P, c->n); LAUNCH_CHECK(); }
    c->recs_dirty = true;
    return 0;
}

int sphb_make_tree(sphb_ctx * c)
{
    CK(cudaSetDevice(c->device));
    if (!c->n_glob) { c->err = "no particles uploaded"; return 1; }
    return make_tree(c);
}

int sphb_pre_interaction(sphb_ctx * c)
{
    CK(cudaSetDevice(c->device));
    if (require_tree(c)) return 1;
    if (pre(c)) return 1;
    return sync_scalars(c, true);
}

int sphb_fluid_force(sphb_ctx * c)
{
    CK(cudaSetDevice(c->device));
    if (require_tree(c)) return 1;
    return force(c);
}

int sphb_gravity_force(sphb_ctx * c)
{
    CK(cudaSetDevice(c->device));
    if (!c->P.use_gravity) return 0;                  // src/gravity_force.cpp:54-56
    if (require_tree(c)) return 1;
    return gravity(c, false);
}

int sphb_gravity_direct(sphb_ctx * c)
{
    CK(cudaSetDevice(c->device));
    if (!c->P.use_gravity) return 0;
    if (!c->n_glob) { c->err = "no particles uploaded"; return 1; }
    return gravity(c, true);
}

int sphb_gravity_direct_targets(sphb_ctx * c, int k)
{
    CK(cudaSetDevice(c->device));
    if (!c->P.use_gravity) return 0;
    if (!c->n_glob) { c->err = "no particles uploaded"; return 1; }
    if (k <= 0) { c->err = "bad target count"; return 1; }
    return gravity(c, true, k);
}

int sphb_timestep(sphb_ctx * c, double * dt)
{
    CK(cudaSetDevice(c->device));
    if (!c->n_glob) { c->err = "no particles uploaded"; return 1; }
    if (timestep(c)) return 1;
    if (sync_scalars(c, true)) return 1;
    if (dt) *dt = c->dt;
    return 0;
}

int sphb_predict(sphb_ctx * c) { CK(cudaSetDevice(c->device)); if (!c->n_glob) { c->err = "no particles uploaded"; return 1; } return predict(c); }
int sphb_correct(sphb_ctx * c) { CK(cudaSetDevice(c->device)); if (!c->n_glob) { c->err = "no particles uploaded"; return 1; } return correct(c); }

int sphb_initialize(sphb_ctx * c)
{
    CK(cudaSetDevice(c->device));
    if (!c->n_glob) { c->err = "no particles uploaded"; return 1; }
    if (sphb_init_state(c) || make_tree(c) || pre(c) || force(c)) return 1;
    if (c->P.use_gravity) { if (gravity(c, false)) return 1; }
    return sync_scalars(c, true);
}

int sphb_integrate(sphb_ctx * c, double * dt)
{
    CK(cudaSetDevice(c->device));
    if (!c->n_glob) { c->err = "no particles uploaded"; return 1; }
    if (timestep(c) || predict(c) || make_tree(c) || pre(c) || force(c)) return 1;
    if (c->P.use_gravity) { if (gravity(c, false)) return 1; }
    if (correct(c)) return 1;
    if (sync_scalars(c, true)) return 1;
    if (dt) *dt = c->dt;
    return 0;
}

int sphb_energy(sphb_ctx * c, double out[3])
{
    CK(cudaSetDevice(c->device));
    if (!c->n_glob) { c->err = "no particles uploaded"; return 1; }
    CK(cudaMemsetAsync(c->d_scal + 3, 0, 3 * sizeof(double), c->stream));
    if (c->n > 0) {
        const int grid = std::min(cdiv(c->n, 256), 4 * c->sm_count);
        switch (c->dim) {
        case 1: k_energy<1><<<grid, 256, 0, c->stream>>>(c->cur, c->n, c->d_scal + 3); break;
        case 2: k_energy<2><<<grid, 256, 0, c->stream>>>(c->cur, c->n, c->d_scal + 3); break;
        default: k_energy<3><<<grid, 256, 0, c->stream>>>(c->cur, c->n, c->d_scal + 3); break;
        }
        LAUNCH_CHECK();
    }
    if (c->world > 1) CKN(g_nccl.AllReduce(c->d_scal + 3, c->d_scal + 3, 3, ncclFloat64, ncclSum, c->comm, c->stream));
    CK(cudaMemcpyAsync(out, c->d_scal + 3, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- test / measurement hooks ---------------------------------------------------------------------------
int sphb_neighbor_lists(sphb_ctx * c, const double * h, int symmetric, int64_t * offsets, int32_t * ids, int64_t cap_total, int64_t * total)
{
    CK(cudaSetDevice(c->device));
    if (require_tree(c)) return 1;
    if (c->world > 1) { c->err = "sphb_neighbor_lists is a single-GPU test hook"; return 1; }
    const int n = c->n;
    if (ensure_recs(c)) return 1;
    if (symmetric) { if (set_kernel(c)) return 1; }
    std::vector<void *> bag;
    double * d_h = nullptr; int * d_counts = nullptr; long long * d_offs = nullptr; int * d_ids = nullptr;
    if (h) {
        double * d_h_orig = nullptr;
        if (dev_alloc(c, &d_h_orig, n, bag) || dev_alloc(c, &d_h, n, bag)) { free_bag(bag); return 1; }
        CK(cudaMemcpyAsync(d_h_orig, h, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        k_gather_by_orig<<<cdiv(n, 256), 256, 0, c->stream>>>(d_h_orig, c->cur.orig, d_h, n, 1, 0); LAUNCH_CHECK();
    }
    if (dev_alloc(c, &d_counts, n, bag) || dev_alloc(c, &d_offs, n + 1, bag)) { free_bag(bag); return 1; }
    GroupTable gt;
#define NL(D, FILL) if (group_table(c, gt)) { free_bag(bag); return 1; } k_neighbor_lists<D><<<c->pre_grid, 128, 0, c->stream>>>(c->cur, c->rc, c->td, c->P, gt, d_h, symmetric, FILL, d_counts, d_offs, d_ids, cap_total, c->d_err)
    switch (c->dim) { case 1: NL(1, 0); break; case 2: NL(2, 0); break; default: NL(3, 0); break; }
    LAUNCH_CHECK();
    std::vector<int> counts(n), orig(n);
    CK(cudaMemcpyAsync(counts.data(), d_counts, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(orig.data(), c->cur.orig, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    // CSR in the caller's particle order
    std::vector<long long> cnt_orig(n);
    for (int s = 0; s < n; ++s) cnt_orig[orig[s]] = counts[s];
    long long tot = 0;
    for (int k = 0; k < n; ++k) { offsets[k] = tot; tot += cnt_orig[k]; }
    offsets[n] = tot;
    if (total) *total = tot;
    if (ids && cap_total > 0) {
        std::vector<long long> offs_sorted(n + 1);
        for (int s = 0; s < n; ++s) offs_sorted[s] = offsets[orig[s]];
        offs_sorted[n] = tot;
        const long long cap = std::min<long long>(cap_total, tot);
        if (dev_alloc(c, &d_ids, (size_t)std::max<long long>(cap, 1), bag)) { free_bag(bag); return 1; }
        CK(cudaMemcpyAsync(d_offs, offs_sorted.data(), (size_t)(n + 1) * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
        switch (c->dim) { case 1: NL(1, 1); break; case 2: NL(2, 1); break; default: NL(3, 1); break; }
        LAUNCH_CHECK();
        CK(cudaMemcpyAsync(ids, d_ids, (size_t)cap * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < n; ++k) {
            const long long a = std::min<long long>(offsets[k], cap), b = std::min<long long>(offsets[k + 1], cap);
            std::sort(ids + a, ids + b);
        }
    }
#undef NL
    CK(cudaStreamSynchronize(c->stream));
    free_bag(bag);
    return 0;
}

int sphb_enable_counters(sphb_ctx * c, int enable)
{
    CK(cudaSetDevice(c->device));
    c->counters_on = enable != 0;
    CK(cudaMemsetAsync(c->d_cnt, 0, sizeof(Counters), c->stream));
    return 0;
}

int sphb_get_counters(sphb_ctx * c, sphb_counters * out)
{
    CK(cudaSetDevice(c->device));
    Counters h;
    CK(cudaMemcpyAsync(&h, c->d_cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemsetAsync(c->d_cnt, 0, sizeof(Counters), c->stream));
    sphb_counters & o = c->last_counters;
    o.n_particles = (uint64_t)c->n;
    o.newton_evals = h.newton_evals; o.newton_iters = h.newton_iters;
    o.pre_candidates = h.pre_candidates; o.pre_neighbors = h.pre_neighbors;
    o.force_pairs = h.force_pairs; o.grav_pp = h.grav_pp; o.grav_pc = h.grav_pc; o.grav_node_visits = h.grav_node_visits;
    o.grav_pc_group = h.grav_pc_group; o.grav_pp_group = h.grav_pp_group;
    CK(cudaMemcpy(&c->last_ngroups, c->d_ngroups, sizeof(int), cudaMemcpyDeviceToHost));
    o.n_groups = (uint64_t)c->last_ngroups;
    if (c->tree_valid) {
        std::vector<double2> nn((size_t)c->td.n_nodes * 4);
        CK(cudaMemcpy(nn.data(), c->td.nn, nn.size() * sizeof(double2), cudaMemcpyDeviceToHost));
        uint64_t leaves = 0;
        for (int k = 0; k < c->td.n_nodes; ++k) {
            long long bits; std::memcpy(&bits, &nn[(size_t)k * 4 + 2].y, 8);
            leaves += (bits >> 32) ? 0 : 1;          // {child0, nchild}: nchild == 0
        }
        o.tree_leaves = leaves; o.tree_nodes = (uint64_t)c->td.n_nodes;
    }
    *out = o;
    return 0;
}

int sphb_enable_timers(sphb_ctx * c, int enable) { c->timers_on = enable != 0; for (auto & u : c->nseg) u = 0; for (auto & m : c->ms) m = 0.f; return 0; }
int sphb_get_timers(sphb_ctx * c, float ms[SPHB_T_COUNT])
{
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    read_timers(c);
    for (int k = 0; k < SPHB_T_COUNT; ++k) ms[k] = c->ms[k];
    return 0;
}
uint64_t sphb_halo_records(const sphb_ctx * c) { return c->halo_pulled; }
uint64_t sphb_migrated(const sphb_ctx * c) { return c->migrated; }

uint64_t sphb_launch_count(const sphb_ctx * c) { return c->launches; }
uint64_t sphb_nonconverged(const sphb_ctx * c) { return c->nonconverged_total; }

void * sphb_host_alloc(size_t bytes) { void * p = nullptr; if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr; return p; }
void sphb_host_free(void * p) { if (p) cudaFreeHost(p); }

int sphb_bench_fp64(int device, double * tflops)
{
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double * out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return 1;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_fp64_peak<<<blocks, threads>>>(out, 1024);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k_fp64_peak<<<blocks, threads>>>(out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        best = std::min(best, ms);
    }
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

} // extern "C"
