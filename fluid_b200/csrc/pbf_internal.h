// pbf_internal.h — solver state shared by pbf_kernels.cu (device code) and pbf_api.cu (C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pbf_b200.h"
#include "pbf_device.cuh"

namespace pbf {

enum { ERRBIT_NBR_CAPACITY = 1, ERRBIT_NONFINITE = 2, ERRBIT_HALO_CAPACITY = 4, ERRBIT_MIGRATION = 8,
       ERRBIT_SLAB_CAPACITY = 16, ERRBIT_PEER_TIMEOUT = 32, ERRBIT_PEER_MISMATCH = 64 };

// Device-resident description of a slab handle's sorted arrays plus what its x-neighbours need to see of it (peer mode,
// include/pbf_b200_multi.h).  Written on the device after every sort, so the host never needs the counts and a step has no
// host round trip; the neighbours read b[] and write flag[] through peer-mapped pointers (same process: peer access,
// other process: CUDA IPC).
struct SlabLink {
  uint32_t b[8];          // b0, b1, b2, b3, n_sorted of the LAST sort: [0,b0) left ghosts | [b0,b3) owned | [b3,n) right ghosts
  uint32_t nb[2][8];      // copies of the x-neighbours' b[] for the same step (k_fetch_peer_ranges): [0] left, [1] right
  uint32_t flag[2];       // last epoch completed by the left [0] / right [1] neighbour (written by ITS k_exchange)
  uint32_t col_hist_valid, pad[5];
};
// where a solver pass also stores its boundary columns: the neighbours' ghost ranges of the same array (peer stores)
struct PushArgs { float4* dst[2]; };

// device-resident scalars (one allocation)
struct Scalars {
  unsigned long long nbr_cursor;   // rows (of 32 entries) handed out by the neighbour build
  int err; int pad;
  double rho_first, rho_final;     // sum of densities: first lambda pass / finalize pass
  unsigned int counters[8];        // slab packing cursors: emigrants L/R, ghosts L/R
  unsigned int bounds[8];          // slab: cell_start at the column boundaries after the sort
  unsigned int alert_count;        // particles below the neighbour-count alert threshold in the last step
  unsigned int alert_kept;         // ... of which this many have a record (ids below alert_bound)
  unsigned int alert_bound, pad2;
  unsigned int timeout_epoch, timeout_missing;   // peer mode: the exchange point that gave up, and who was missing (1 left, 2 right)
};

enum KernelId { K_PREDICT = 0, K_SCAN, K_SCATTER, K_CELLSORT, K_REORDER, K_NEIGHBORS, K_LAMBDA, K_DELTA,
                K_VELOCITY, K_VORT_XSPH, K_CONFINE, K_DENSITY, K_IO, K_SLAB, K_COUNT };
static const char* const kKernelNames[K_COUNT] = {
  "predict_collide_hash", "cell_scan", "scatter", "cell_sort", "reorder", "build_neighbors", "lambda",
  "delta_collide", "velocity", "vorticity_xsph", "confine_commit", "density_only", "io", "slab"};

struct Solver {
  int device = 0;
  cudaStream_t stream = nullptr;
  PbfParams hp;
  DevParams dp;
  std::string last_error;

  size_t n = 0, cap = 0;
  uint32_t ncell = 0;
  int cur = 0;                       // which of the double-buffered pos/vel/orig holds the state
  float4 *pos[2] = {nullptr, nullptr}, *vel[2] = {nullptr, nullptr};
  uint32_t* orig[2] = {nullptr, nullptr};
  float4 *xs_tmp = nullptr, *xs_a = nullptr, *xs_b = nullptr, *vtmp = nullptr, *omega = nullptr, *xpred = nullptr;
  float4* xv = nullptr;              // 2 float4 per particle: (x*, v) record gathered by the vorticity/XSPH pass
  float* rho = nullptr;
  uint32_t *cell_of = nullptr, *rank = nullptr, *perm = nullptr, *key = nullptr;
  uint32_t *cell_count = nullptr, *cell_start = nullptr, *block_sums = nullptr;
  uint32_t *nbr = nullptr, *slice_off = nullptr, *nbr_cnt = nullptr;
  size_t nbr_cap_rows = 0;           // capacity of nbr in rows of 32 entries
  // ranges of the cell-sorted arrays: [0, n_sorted) holds particles (ghost columns at both ends in
  // slab mode); neighbour lists / solver passes cover the owned range [r_i0, r_i0 + r_cnt)
  uint32_t n_sorted = 0, r_i0 = 0, r_cnt = 0;
  // x-slab decomposition (pbf_slab.inl)
  bool slab = false;
  bool own_stream = true;
  int has_left = 0, has_right = 0;
  size_t halo_cap = 0;               // particles per migration / ghost message
  float4 *mig_send[2] = {nullptr, nullptr}, *mig_recv[2] = {nullptr, nullptr};     // [0] left, [1] right; element 0 = header
  float4 *ghost_send[2] = {nullptr, nullptr}, *ghost_recv[2] = {nullptr, nullptr};
  uint32_t bounds[5] = {0, 0, 0, 0, 0};   // b0..b3, n_sorted of the last slab sort (see pbf_b200_slab.h)
  size_t n_in_cap() const { return slab ? cap : 0; }
  size_t append_base = 0;            // slab: immigrants / ghosts are appended at append_base + {0,1,2,3} * halo_cap (= particle_cap)
  int max_cols = 0;                  // slab: the cell arrays hold this many owned columns (+2 ghost columns): room for re-balancing
  // peer mode (pbf_b200_multi.h): device-side ranges, peer stores, flag synchronisation, no host sync inside a step
  bool p2p = false;
  SlabLink* link = nullptr;          // this handle's link block (device memory, visible to the neighbours)
  struct Peer {
    SlabLink* link = nullptr;        // the neighbour's link block
    float4 *xs_a = nullptr, *xs_b = nullptr, *xs_w = nullptr;     // its solver arrays (ghost ranges are written by us)
    float4 *mig_recv = nullptr, *ghost_recv = nullptr;           // its receive buffers for OUR side
    bool ipc = false;                // mapped with cudaIpcOpenMemHandle (closed in pbf_destroy)
    void* ipc_base[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  } peer[2];                         // [0] left, [1] right
  uint32_t epoch = 0;                // signals issued so far (identical sequence on every rank)
  long long wait_timeout_ns = 20000000000ll;
  uint32_t* col_hist_host = nullptr; // owned particles per global cell column after a sort (re-balancing): page-locked host memory the kernel stores into
  cudaEvent_t ev_hist = nullptr;
  int hist_every = 0;                // > 0: record the histogram at the sorts of steps that are multiples of it; 0: whenever the last one was delivered
  long long hist_step = -1;          // step index (steps_done at its sort) of the histogram in flight / delivered
  Scalars* sc = nullptr;             // device
  float* io_stage = nullptr;         // device staging for original-order fp32 AoS (7 floats / particle)
  void* host_extra = nullptr;        // pbf_api.cu's HandleExtra (pinned staging, registered host ranges)
  float4* tri_dev = nullptr;         // obstacle triangles (5 float4 each, leaf order), referenced by dp.tri
  float4* bvh_dev = nullptr;         // their bounding-volume hierarchy (2 float4 per node), dp.bvh
  uint32_t* tri_id_dev = nullptr;    // leaf order -> original triangle index, dp.tri_id
  size_t bvh_nodes = 0; int bvh_depth = 0;
  int capture_xpred = 0;
  // neighbour-count alert (Particle::initializeWithNewNeighbors, particles.cpp:165-173): records of the last step
  uint32_t alert_thr = 0; size_t alert_cap = 0;
  float4* alert_buf = nullptr;       // 2 float4 per record: (x*, id bits), (v, count bits); then the id histogram (k_alert_hist)
  bool have_neighbors = false;
  long long rebinned_at = -1;     // steps_done when the arrays were last re-binned by committed position
  // streaming read-back (pbf_set_readback): results leave for page-locked host buffers as soon as each is final
  double *rb_pos = nullptr, *rb_vel = nullptr, *rb_rho = nullptr;
  double* rb_stage = nullptr;        // device staging, 7 doubles per particle (owned by pbf_api.cu)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_rb[4] = {nullptr, nullptr, nullptr, nullptr};   // positions final, density final, velocity final, copies done
  bool rb_pending = false;        // read-back copies of the last step may still be in flight on copy_stream (ev_rb[3])

  // Launch-bound scenes (the reference's own p.xml / spheres_p.xml: 720 / 2106 particles, 36 launches of a few
  // microseconds each): the step's launch sequence is captured once per buffer parity (`cur` flips every
  // step) into a CUDA graph and replayed.  PBF_GRAPH=0/1 overrides the size policy (on up to 2^18 particles).
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  uint64_t graph_launches[2] = {0, 0};   // kernels inside each captured step (what launch_count() advances by per replay)
  int graph_policy = -1;                 // -1 not decided yet, 0 plain launches, 1 graphs
  void graph_invalidate() {
    for (int k = 0; k < 2; k++) if (graph_exec[k]) { cudaGraphExecDestroy(graph_exec[k]); graph_exec[k] = nullptr; }
    graph_policy = -1;
  }

  uint64_t launches = 0, steps_done = 0;
  double last_call_ms = 0.0;
  cudaEvent_t ev_call[2] = {nullptr, nullptr};
  bool call_timed = false;

  // optional per-kernel profiling with CUDA events on the launching stream
  bool profiling = false;
  struct ProfRec { int kid; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> event_pool;
  double prof_ms[K_COUNT] = {0};
  uint64_t prof_launches[K_COUNT] = {0};
  cudaEvent_t prof_pending = nullptr;
  cudaEvent_t get_event() {
    if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void prof_begin(int) { if (profiling) { prof_pending = get_event(); cudaEventRecord(prof_pending, stream); } }
  void prof_end(int kid) {
    if (profiling) { cudaEvent_t b = get_event(); cudaEventRecord(b, stream); prof_recs.push_back({kid, prof_pending, b}); }
  }
  void prof_collect() {   // call after a stream sync
    for (auto& r : prof_recs) {
      float ms = 0.f; cudaEventElapsedTime(&ms, r.a, r.b);
      prof_ms[r.kid] += ms; prof_launches[r.kid]++;
      event_pool.push_back(r.a); event_pool.push_back(r.b);
    }
    prof_recs.clear();
  }
};

void enqueue_step(Solver* h, bool readback = false);
void enqueue_readback_all(Solver* h);
void enqueue_estimate_densities(Solver* h);
void enqueue_predict_hash(Solver* h, int apply_forces);
void enqueue_sort(Solver* h, size_t n_in);
void enqueue_build(Solver* h, int include_self);
enum { PART_ALL = 0, PART_BOUNDARY = 1, PART_INTERIOR = 2 };
void enqueue_lambda(Solver* h, int first_iter, int part);
void enqueue_delta(Solver* h, int part);
void enqueue_velocity(Solver* h);
void enqueue_vorticity(Solver* h, int part);
void enqueue_confine(Solver* h);
int  alloc_particle_arrays(Solver* h, size_t cap);          // pbf_api.cu
int  fill_dev_params(const PbfParams& p, DevParams& d, std::string& err);
int  sync_and_check(Solver* h);
int  io_upload(Solver* h, size_t n, const double* pos_xyz, const double* vel_xyz);     // pbf_api.cu
int  io_download(Solver* h, double* pos_xyz, double* vel_xyz, double* density);
void enqueue_rebin(Solver* h);
void enqueue_density_at(Solver* h, uint32_t m, const float4* d_q, float* d_out);
void enqueue_import(Solver* h, const float* d_pos_xyz, const float* d_vel_xyz);
void enqueue_export3(Solver* h, const float4* src, float* dst_xyz);
void enqueue_export1(Solver* h, const float* src, float* dst);
void enqueue_export_w(Solver* h, const float4* src, float* dst);
void enqueue_digest(Solver* h, unsigned long long* digest, uint32_t* count);
void enqueue_import_f64(Solver* h, const double* d_pos_xyz, const double* d_vel_xyz);
void enqueue_export3_f64(Solver* h, const float4* src, double* dst_xyz);
void enqueue_export1_f64(Solver* h, const float* src, double* dst);

}  // namespace pbf

struct pbf_handle : public pbf::Solver {};
