// pbf_slab.inl — x-slab decomposition: device-side packing / absorbing of migrants and ghosts and
// the phase entry points of include/pbf_b200_slab.h.  Textually included by pbf_kernels.cu (same
// translation unit as the kernels it launches).  The library never communicates: the host driver
// (fluid_b200/slab.py, torch.distributed / NCCL) moves the message buffers between ranks.
#include "../../include/pbf_b200_slab.h"

#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace pbf {

// header of a message buffer: element 0, .x = count as uint32 bits
__global__ void k_write_headers(Scalars* sc, float4* m0, float4* m1, float4* g0, float4* g1, uint32_t cap, int which) {
  if (threadIdx.x || blockIdx.x) return;
  if (which == 0) {
    if (m0) m0[0] = make_float4(__uint_as_float(min(sc->counters[0], cap)), 0.f, 0.f, 0.f);
    if (m1) m1[0] = make_float4(__uint_as_float(min(sc->counters[1], cap)), 0.f, 0.f, 0.f);
  } else {
    if (g0) g0[0] = make_float4(__uint_as_float(min(sc->counters[2], cap)), 0.f, 0.f, 0.f);
    if (g1) g1[0] = make_float4(__uint_as_float(min(sc->counters[3], cap)), 0.f, 0.f, 0.f);
  }
}

// immigrants: append at base + k, hash by their (already predicted) x*
__global__ void __launch_bounds__(TPB)
k_absorb_migrants(const __grid_constant__ DevParams P, const float4* __restrict__ msg, uint32_t cap, uint32_t base,
                  float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ orig, float4* __restrict__ xs_tmp,
                  uint32_t* __restrict__ cell_of, uint32_t* __restrict__ rank, uint32_t* __restrict__ cell_count,
                  Scalars* __restrict__ sc) {
  const uint32_t k = blockIdx.x * TPB + threadIdx.x;
  const uint32_t cnt = min(__float_as_uint(msg[0].x), cap);
  if (k >= cnt) return;
  const float4 x = msg[1 + 3 * k], p = msg[2 + 3 * k], v = msg[3 + 3 * k];
  const uint32_t i = base + k;
  pos[i] = make_float4(x.x, x.y, x.z, 0.f);
  vel[i] = v;
  orig[i] = __float_as_uint(x.w);
  xs_tmp[i] = make_float4(p.x, p.y, p.z, 0.f);
  const int3 cg = cell_coords_global(P, p.x, p.y, p.z);
  if (cg.x < P.gx_lo || cg.x >= P.gx_hi) { atomicOr(&sc->err, ERRBIT_MIGRATION); return; }   // not ours: sender's bug or >1 hop
  const uint32_t c = cell_linear(P, make_int3(cg.x - P.cx_offset, cg.y, cg.z));
  cell_of[i] = c;
  rank[i] = atomicAdd(&cell_count[c], 1u);
}

// ghost layer to send: every owned particle (stayers and immigrants) whose cell lies in the first /
// last owned column.  Payload: x* with the global id in .w, committed x (for v = (x* - x)/dt).
__global__ void __launch_bounds__(TPB)
k_pack_ghosts(const __grid_constant__ DevParams P, uint32_t n_in, const uint32_t* __restrict__ cell_of,
              const float4* __restrict__ xs_tmp, const float4* __restrict__ pos, const uint32_t* __restrict__ orig,
              float4* __restrict__ gl, float4* __restrict__ gr, uint32_t cap, Scalars* __restrict__ sc) {
  const uint32_t i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n_in) return;
  const uint32_t c = cell_of[i];
  if (c == CELL_INVALID) return;
  const int col = (int)(c / (uint32_t)(P.gdim[1] * P.gdim[2]));
  const int last = P.gdim[0] - 2;
#pragma unroll
  for (int side = 0; side < 2; side++) {
    float4* dst = side ? gr : gl;
    if (dst == nullptr || col != (side ? last : 1)) continue;
    const uint32_t slot = atomicAdd(&sc->counters[2 + side], 1u);
    if (slot >= cap) { atomicOr(&sc->err, ERRBIT_HALO_CAPACITY); continue; }
    const float4 p = xs_tmp[i], x = pos[i];
    dst[1 + 2 * slot] = make_float4(p.x, p.y, p.z, __uint_as_float(orig[i]));
    dst[2 + 2 * slot] = make_float4(x.x, x.y, x.z, 0.f);
  }
}

// received ghosts: append at base + k; they must land in the ghost column of their side
__global__ void __launch_bounds__(TPB)
k_absorb_ghosts(const __grid_constant__ DevParams P, const float4* __restrict__ msg, uint32_t cap, uint32_t base, int side,
                float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ orig, float4* __restrict__ xs_tmp,
                uint32_t* __restrict__ cell_of, uint32_t* __restrict__ rank, uint32_t* __restrict__ cell_count,
                Scalars* __restrict__ sc) {
  const uint32_t k = blockIdx.x * TPB + threadIdx.x;
  const uint32_t cnt = min(__float_as_uint(msg[0].x), cap);
  if (k >= cnt) return;
  const float4 p = msg[1 + 2 * k], x = msg[2 + 2 * k];
  const uint32_t i = base + k;
  pos[i] = make_float4(x.x, x.y, x.z, 0.f);
  vel[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  orig[i] = __float_as_uint(p.w);
  xs_tmp[i] = make_float4(p.x, p.y, p.z, 0.f);
  const int3 cg = cell_coords_global(P, p.x, p.y, p.z);
  const int want = side ? P.gx_hi : P.gx_lo - 1;
  if (cg.x != want) { atomicOr(&sc->err, ERRBIT_MIGRATION); return; }
  const uint32_t c = cell_linear(P, make_int3(cg.x - P.cx_offset, cg.y, cg.z));
  cell_of[i] = c;
  rank[i] = atomicAdd(&cell_count[c], 1u);
}

// Column boundaries of the sorted arrays (x is the slowest cell axis): b0..b3, n_sorted.  Also left in the link block for
// the device-side consumers of peer mode, with the capacity checks the host does in the host-driven mode.
__global__ void k_gather_bounds(const uint32_t* __restrict__ cell_start, uint32_t c1, uint32_t c2, uint32_t cm, uint32_t cm1,
                                uint32_t ncell, Scalars* sc, SlabLink* lk, uint32_t max_sorted) {
  if (threadIdx.x || blockIdx.x) return;
  const uint32_t b0 = cell_start[c1], b1 = cell_start[c2], b2 = cell_start[cm], b3 = cell_start[cm1], n = cell_start[ncell];
  sc->bounds[0] = b0; sc->bounds[1] = b1; sc->bounds[2] = b2; sc->bounds[3] = b3; sc->bounds[4] = n;
  if (lk) { lk->b[0] = b0; lk->b[1] = b1; lk->b[2] = b2; lk->b[3] = b3; lk->b[4] = n; }
  if (n > max_sorted) atomicOr(&sc->err, ERRBIT_SLAB_CAPACITY);
}

// owned particles per GLOBAL cell column after the sort (what the re-balancing looks at): hist[gx] for gx in [gx_lo, gx_hi)
__global__ void __launch_bounds__(TPB)
k_column_hist(const uint32_t* __restrict__ cell_start, uint32_t gyz, int gx_lo, int n_owned_cols, int n_global_cols, uint32_t* __restrict__ hist) {
  const int c = blockIdx.x * TPB + threadIdx.x;
  if (c >= n_global_cols) return;
  const int local = c - gx_lo;                                    // owned columns are local columns 1 .. n_owned_cols
  hist[c] = (local >= 0 && local < n_owned_cols) ? cell_start[(uint32_t)(local + 2) * gyz] - cell_start[(uint32_t)(local + 1) * gyz] : 0u;
  __threadfence_system();                                         // hist is page-locked host memory
}

// ---- peer mode: flags and ranges through peer-mapped memory ---------------------------------------------------------
// A rank tells its x-neighbours that everything it enqueued before this point is complete (its kernels' stores into
// their ghost ranges / receive buffers included: stream order + a system-scope fence), by writing the epoch into the flag
// words they poll.  Every rank issues the same sequence of signals, so epochs agree without any host exchange.
// ... and then waits until both neighbours have reached `epoch`: ONE launch per exchange point (a step has 2 I + 5 of them).  The
// wait is bounded: after timeout_ns the error flag is set and this and every later exchange point of the handle return at
// once (the step finishes with garbage and pbf_sync reports PBF_ERR_CUDA).
__global__ void k_exchange(uint32_t* flag_at_left, uint32_t* flag_at_right, const uint32_t* flag_left, const uint32_t* flag_right, uint32_t epoch,
                           Scalars* sc, long long timeout_ns) {
  if (threadIdx.x || blockIdx.x) return;
  __threadfence_system();
  if (flag_at_left) *reinterpret_cast<volatile uint32_t*>(flag_at_left) = epoch;
  if (flag_at_right) *reinterpret_cast<volatile uint32_t*>(flag_at_right) = epoch;
  if (*reinterpret_cast<volatile int*>(&sc->err) & ERRBIT_PEER_TIMEOUT) return;
  unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    const bool l = !flag_left || (int)(*reinterpret_cast<const volatile uint32_t*>(flag_left) - epoch) >= 0;
    const bool r = !flag_right || (int)(*reinterpret_cast<const volatile uint32_t*>(flag_right) - epoch) >= 0;
    if (l && r) break;
    unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if ((long long)(t1 - t0) > timeout_ns) { sc->timeout_epoch = epoch; sc->timeout_missing = (l ? 0u : 1u) | (r ? 0u : 2u); atomicOr(&sc->err, ERRBIT_PEER_TIMEOUT); break; }
    __nanosleep(200);
  }
  __threadfence_system();
}
// after the neighbours' sorts: copy their ranges next to ours (the solver passes index the neighbours' ghost ranges with
// them) and check that our boundary columns are exactly as long as their ghost ranges
__global__ void k_fetch_peer_ranges(SlabLink* lk, const SlabLink* left, const SlabLink* right, Scalars* sc) {
  if (threadIdx.x || blockIdx.x) return;
  if (left) {
    for (int k = 0; k < 5; k++) lk->nb[0][k] = *reinterpret_cast<const volatile uint32_t*>(&left->b[k]);
    if (lk->nb[0][4] - lk->nb[0][3] != lk->b[1] - lk->b[0]) atomicOr(&sc->err, ERRBIT_PEER_MISMATCH);
  }
  if (right) {
    for (int k = 0; k < 5; k++) lk->nb[1][k] = *reinterpret_cast<const volatile uint32_t*>(&right->b[k]);
    if (lk->nb[1][0] != lk->b[3] - lk->b[2]) atomicOr(&sc->err, ERRBIT_PEER_MISMATCH);
  }
}

__global__ void __launch_bounds__(TPB)
k_import_ids(uint32_t n, const float* __restrict__ pos_xyz, const float* __restrict__ vel_xyz, const uint32_t* __restrict__ ids,
             float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ orig) {
  const uint32_t i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  pos[i] = make_float4(pos_xyz[3 * i], pos_xyz[3 * i + 1], pos_xyz[3 * i + 2], 0.f);
  vel[i] = make_float4(vel_xyz[3 * i], vel_xyz[3 * i + 1], vel_xyz[3 * i + 2], 0.f);
  orig[i] = ids[i];
}

}  // namespace pbf

using namespace pbf;

#define SCK(h, call)                                                                        \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      (h)->last_error = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
      return PBF_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

static int sfail(pbf_handle* h, int code, const char* msg) { if (h) h->last_error = msg; return code; }

extern "C" {

int pbf_grid_dims(const PbfParams* params, int dims_out[3]) {
  if (!params || !dims_out) return PBF_ERR_INVALID;
  DevParams d; std::string err;
  int rc = fill_dev_params(*params, d, err);
  if (rc != PBF_OK) return rc;
  for (int a = 0; a < 3; a++) dims_out[a] = d.gdim[a];
  return PBF_OK;
}

int pbf_cell_columns(const PbfParams* params, size_t n, const double* pos_xyz, int32_t* column_out) {
  if (!params || (n && (!pos_xyz || !column_out))) return PBF_ERR_INVALID;
  DevParams d; std::string err;
  int rc = fill_dev_params(*params, d, err);
  if (rc != PBF_OK) return rc;
  for (size_t i = 0; i < n; i++) {
    volatile float t = (float)pos_xyz[3 * i] - d.gmin[0];     // same two fp32 roundings as cell_coords_global
    volatile float u = t * d.inv_cell;
    int cx = (int)std::floor(u);
    column_out[i] = cx < 0 ? 0 : (cx > d.gdim[0] - 1 ? d.gdim[0] - 1 : cx);
  }
  return PBF_OK;
}

int pbf_set_stream(pbf_handle* h, void* cuda_stream) {
  if (!h) return PBF_ERR_INVALID;
  SCK(h, cudaSetDevice(h->device));
  SCK(h, cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)cuda_stream;
  h->own_stream = false;
  return PBF_OK;
}

// Owned columns [gx_lo, gx_hi) and the widths of the neighbours' slabs; sizes the local cell grid.  Shared by
// pbf_slab_configure and pbf_slab_set_columns (re-balancing: the next predict pass then emigrates every particle whose
// column changed hands, through the ordinary migration messages).
static int apply_columns(pbf_handle* h, int gx_lo, int gx_hi, int left_cols, int right_cols) {
  DevParams& d = h->dp;
  if (gx_lo < 0 || gx_hi > d.gdim_x_global || gx_hi <= gx_lo) return sfail(h, PBF_ERR_INVALID, "bad slab range");
  if (gx_hi - gx_lo > h->max_cols) return sfail(h, PBF_ERR_CAPACITY, "slab wider than max_cols (pbf_slab_configure_ex)");
  d.gx_lo = gx_lo; d.gx_hi = gx_hi; d.cx_offset = gx_lo - 1;
  d.hop_left = left_cols; d.hop_right = right_cols;
  d.gdim[0] = (gx_hi - gx_lo) + 2;                      // one ghost column on each side
  h->has_left = left_cols > 0; h->has_right = right_cols > 0;
  h->ncell = (uint32_t)((size_t)d.gdim[0] * d.gdim[1] * d.gdim[2]);
  return PBF_OK;
}

int pbf_slab_configure(pbf_handle* h, int gx_lo, int gx_hi, int left_cols, int right_cols, size_t particle_cap, size_t halo_cap) {
  return pbf_slab_configure_ex(h, gx_lo, gx_hi, left_cols, right_cols, particle_cap, halo_cap, 0);
}

int pbf_slab_configure_ex(pbf_handle* h, int gx_lo, int gx_hi, int left_cols, int right_cols, size_t particle_cap, size_t halo_cap, int max_cols) {
  if (!h) return PBF_ERR_INVALID;
  if (h->n != 0 || h->cap != 0) return sfail(h, PBF_ERR_INVALID, "pbf_slab_configure must precede any upload");
  if (halo_cap == 0 || particle_cap == 0) return sfail(h, PBF_ERR_INVALID, "bad slab capacities");
  if (h->hp.xsph_mode != PBF_XSPH_JACOBI) return sfail(h, PBF_ERR_INVALID, "slab mode supports PBF_XSPH_JACOBI only (reference order is a global sequential dependency)");
  SCK(h, cudaSetDevice(h->device));
  h->max_cols = std::max(max_cols, gx_hi - gx_lo);
  if (h->max_cols > h->dp.gdim_x_global) h->max_cols = h->dp.gdim_x_global;
  int rc = apply_columns(h, gx_lo, gx_hi, left_cols, right_cols);
  if (rc != PBF_OK) return rc;
  const size_t max_cells = (size_t)(h->max_cols + 2) * h->dp.gdim[1] * h->dp.gdim[2];
  if ((double)max_cells > 1.5e9) return sfail(h, PBF_ERR_INVALID, "slab grid too large");
  cudaFree(h->cell_count); cudaFree(h->cell_start); cudaFree(h->block_sums);
  SCK(h, cudaMalloc((void**)&h->cell_count, (max_cells + 1) * 4));
  SCK(h, cudaMalloc((void**)&h->cell_start, (max_cells + 2) * 4));
  SCK(h, cudaMalloc((void**)&h->block_sums, (max_cells / 2048 + 2) * 4));
  h->slab = true; h->halo_cap = halo_cap; h->append_base = particle_cap;
  rc = alloc_particle_arrays(h, particle_cap + 4 * halo_cap);
  if (rc != PBF_OK) return rc;
  for (int k = 0; k < 2; k++) {
    SCK(h, cudaMalloc((void**)&h->mig_send[k], (1 + 3 * halo_cap) * sizeof(float4)));
    SCK(h, cudaMalloc((void**)&h->mig_recv[k], (1 + 3 * halo_cap) * sizeof(float4)));
    SCK(h, cudaMalloc((void**)&h->ghost_send[k], (1 + 2 * halo_cap) * sizeof(float4)));
    SCK(h, cudaMalloc((void**)&h->ghost_recv[k], (1 + 2 * halo_cap) * sizeof(float4)));
    SCK(h, cudaMemset(h->mig_send[k], 0, sizeof(float4))); SCK(h, cudaMemset(h->mig_recv[k], 0, sizeof(float4)));
    SCK(h, cudaMemset(h->ghost_send[k], 0, sizeof(float4))); SCK(h, cudaMemset(h->ghost_recv[k], 0, sizeof(float4)));
  }
  SCK(h, cudaMalloc((void**)&h->link, sizeof(SlabLink)));
  SCK(h, cudaMemset(h->link, 0, sizeof(SlabLink)));
  SCK(h, cudaHostAlloc((void**)&h->col_hist_host, (size_t)h->dp.gdim_x_global * 4, cudaHostAllocPortable | cudaHostAllocMapped));
  std::memset(h->col_hist_host, 0, (size_t)h->dp.gdim_x_global * 4);
  SCK(h, cudaEventCreateWithFlags(&h->ev_hist, cudaEventDisableTiming));
  return PBF_OK;
}

int pbf_slab_set_columns(pbf_handle* h, int gx_lo, int gx_hi, int left_cols, int right_cols) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  // kernels in flight hold the old grid in their parameters (by value), so changing it for LATER launches needs no sync
  return apply_columns(h, gx_lo, gx_hi, left_cols, right_cols);
}

int pbf_slab_set_histogram_interval(pbf_handle* h, int every_k_steps) {
  if (!h || !h->slab || every_k_steps < 0) return PBF_ERR_INVALID;
  h->hist_every = every_k_steps;
  return PBF_OK;
}

int pbf_slab_columns(pbf_handle* h, int out[4]) {
  if (!h || !h->slab || !out) return PBF_ERR_INVALID;
  out[0] = h->dp.gx_lo; out[1] = h->dp.gx_hi; out[2] = h->dp.hop_left; out[3] = h->dp.hop_right;
  return PBF_OK;
}

// Owned particles per global cell column after the sort of the last COMPLETED step that recorded them (asynchronous:
// the copy is enqueued by the sort phase and polled here, so reading it never stalls the pipeline).  Returns
// PBF_ERR_INVALID until a first histogram has arrived.
int pbf_slab_column_histogram(pbf_handle* h, uint32_t* hist_out, size_t n_cols, int wait, long long* step_out) {
  if (!h || !h->slab || !hist_out || n_cols != (size_t)h->dp.gdim_x_global) return PBF_ERR_INVALID;
  if (h->hist_step < 0) return sfail(h, PBF_ERR_INVALID, "no histogram recorded yet");
  SCK(h, cudaSetDevice(h->device));
  if (wait) SCK(h, cudaEventSynchronize(h->ev_hist));
  else if (cudaEventQuery(h->ev_hist) != cudaSuccess) { cudaGetLastError(); return sfail(h, PBF_ERR_INVALID, "histogram copy still in flight"); }
  std::memcpy(hist_out, h->col_hist_host, n_cols * 4);
  if (step_out) *step_out = h->hist_step;
  return PBF_OK;
}

// fresh particles of a rank: link ranges = everything owned, nothing sorted yet
static int reset_link_ranges(pbf_handle* h, size_t n) {
  uint32_t b[8] = {0, 0, (uint32_t)n, (uint32_t)n, (uint32_t)n, 0, 0, 0};
  SCK(h, cudaMemcpyAsync(h->link->b, b, sizeof(b), cudaMemcpyHostToDevice, h->stream));
  SCK(h, cudaStreamSynchronize(h->stream));
  return PBF_OK;
}

int pbf_slab_upload(pbf_handle* h, size_t n, const double* pos_xyz, const double* vel_xyz, const uint32_t* ids) {
  if (!h || !h->slab || (n && (!pos_xyz || !vel_xyz || !ids))) return sfail(h, PBF_ERR_INVALID, "pbf_slab_upload: bad argument / not configured");
  if (n > h->append_base) return sfail(h, PBF_ERR_CAPACITY, "pbf_slab_upload: more particles than particle_cap");
  SCK(h, cudaSetDevice(h->device));
  SCK(h, cudaStreamSynchronize(h->stream));
  h->n = n; h->cur = 0; h->have_neighbors = false;
  h->r_i0 = 0; h->r_cnt = (uint32_t)n; h->n_sorted = (uint32_t)n;
  for (int k = 0; k < 5; k++) h->bounds[k] = 0;
  int rc = reset_link_ranges(h, n);
  if (rc != PBF_OK || n == 0) return rc;
  rc = io_upload(h, n, pos_xyz, vel_xyz);                // sets orig = identity ...
  if (rc != PBF_OK) return rc;
  SCK(h, cudaMemcpy(h->orig[0], ids, n * 4, cudaMemcpyHostToDevice));   // ... replaced by the global ids
  return PBF_OK;
}

// A. predict owned particles (apply_forces = 0: only re-bin the committed positions); emigrants go to the migration messages
static int slab_predict(pbf_handle* h, int apply_forces) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  SCK(h, cudaSetDevice(h->device));
  cudaMemsetAsync(&h->sc->rho_first, 0, 2 * sizeof(double), h->stream);
  enqueue_predict_hash(h, apply_forces);
  h->prof_begin(K_SLAB);
  k_write_headers<<<1, 32, 0, h->stream>>>(h->sc, h->has_left ? mig_out(h, 0) : nullptr, h->has_right ? mig_out(h, 1) : nullptr,
                                           nullptr, nullptr, (uint32_t)h->halo_cap, 0);
  h->prof_end(K_SLAB); h->launches++;
  SCK(h, cudaGetLastError());
  return PBF_OK;
}

int pbf_slab_phase_predict(pbf_handle* h) { return slab_predict(h, 1); }

// C. append immigrants (slots [append_base, append_base + 2*cap)), then pack the ghost layers to send
int pbf_slab_phase_migrate(pbf_handle* h) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  SCK(h, cudaSetDevice(h->device));
  const uint32_t cap = (uint32_t)h->halo_cap, base = (uint32_t)h->append_base;
  const int cur = h->cur;
  for (int side = 0; side < 2; side++) {
    if (!(side ? h->has_right : h->has_left)) continue;
    LAUNCH(h, K_SLAB, k_absorb_migrants, blocks_for(cap), h->dp, h->mig_recv[side], cap, base + side * cap, h->pos[cur], h->vel[cur],
           h->orig[cur], h->xs_tmp, h->cell_of, h->rank, h->cell_count, h->sc);
  }
  LAUNCH(h, K_SLAB, k_pack_ghosts, blocks_for((size_t)base + 2 * cap), h->dp, base + 2 * cap, h->cell_of, h->xs_tmp, h->pos[cur],
         h->orig[cur], h->has_left ? ghost_out(h, 0) : (float4*)nullptr, h->has_right ? ghost_out(h, 1) : (float4*)nullptr, cap, h->sc);
  h->prof_begin(K_SLAB);
  k_write_headers<<<1, 32, 0, h->stream>>>(h->sc, nullptr, nullptr, h->has_left ? ghost_out(h, 0) : nullptr,
                                           h->has_right ? ghost_out(h, 1) : nullptr, cap, 1);
  h->prof_end(K_SLAB); h->launches++;
  SCK(h, cudaGetLastError());
  return PBF_OK;
}

// E (device part). append ghosts (slots [append_base + 2*cap, append_base + 4*cap)), sort, leave the column boundaries and the
// column histogram on the device
static int enqueue_slab_sort(pbf_handle* h) {
  const uint32_t cap = (uint32_t)h->halo_cap, base = (uint32_t)h->append_base;
  const int cur = h->cur;
  for (int side = 0; side < 2; side++) {
    if (!(side ? h->has_right : h->has_left)) continue;
    LAUNCH(h, K_SLAB, k_absorb_ghosts, blocks_for(cap), h->dp, h->ghost_recv[side], cap, base + (2 + side) * cap, side, h->pos[cur],
           h->vel[cur], h->orig[cur], h->xs_tmp, h->cell_of, h->rank, h->cell_count, h->sc);
  }
  enqueue_sort(h, (size_t)base + 4 * cap);
  const uint32_t gyz = (uint32_t)(h->dp.gdim[1] * h->dp.gdim[2]);
  const uint32_t m = (uint32_t)(h->dp.gx_hi - h->dp.gx_lo);
  h->prof_begin(K_SLAB);
  k_gather_bounds<<<1, 32, 0, h->stream>>>(h->cell_start, gyz, 2 * gyz, m * gyz, (m + 1) * gyz, h->ncell, h->sc, h->link, (uint32_t)h->append_base);
  h->prof_end(K_SLAB); h->launches++;
  // hist_every > 0 (pbf_slab_set_histogram_interval, what pbf_multi uses): on the steps the driver names, which makes the
  // re-balancing a pure function of the step count; otherwise whenever the previous histogram has been delivered
  const bool record = h->hist_every > 0 ? (h->steps_done % (uint64_t)h->hist_every == 0) : (cudaEventQuery(h->ev_hist) == cudaSuccess);
  if (record) {
    // stored straight into page-locked host memory by the kernel: a copy-engine transfer here would queue behind the other
    // slabs' copies when several slabs share a GPU, and a slab waiting at an exchange point would then block its neighbour
    LAUNCH(h, K_SLAB, k_column_hist, blocks_for(h->dp.gdim_x_global), h->cell_start, gyz, h->dp.gx_lo, (int)m, h->dp.gdim_x_global, h->col_hist_host);
    cudaEventRecord(h->ev_hist, h->stream);
    h->hist_step = (long long)h->steps_done;             // the step whose sort it describes
  } else cudaGetLastError();
  return PBF_OK;
}

// E. host-driven mode: ... read the column boundaries back (synchronises), build the neighbour lists of the owned range
int pbf_slab_phase_sort(pbf_handle* h, uint32_t bounds_out[5]) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  if (h->p2p) return sfail(h, PBF_ERR_INVALID, "peer mode: use pbf_slab_step_p2p");
  SCK(h, cudaSetDevice(h->device));
  enqueue_slab_sort(h);
  int rc = sync_and_check(h);
  if (rc != PBF_OK) return rc;
  Scalars s;
  SCK(h, cudaMemcpy(&s, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 5; k++) h->bounds[k] = s.bounds[k];
  h->n_sorted = s.bounds[4];
  h->r_i0 = s.bounds[0]; h->r_cnt = s.bounds[3] - s.bounds[0];
  h->n = h->r_cnt;
  enqueue_build(h, 0);
  h->have_neighbors = true;
  if (bounds_out) for (int k = 0; k < 5; k++) bounds_out[k] = h->bounds[k];
  SCK(h, cudaGetLastError());
  return PBF_OK;
}

// ---- peer mode ----------------------------------------------------------------------------------------------------------
static void p2p_exchange_point(pbf_handle* h) {
  h->epoch++;
  h->prof_begin(K_SLAB);
  k_exchange<<<1, 32, 0, h->stream>>>(h->has_left ? &h->peer[0].link->flag[1] : nullptr, h->has_right ? &h->peer[1].link->flag[0] : nullptr,
                                      h->has_left ? &h->link->flag[0] : nullptr, h->has_right ? &h->link->flag[1] : nullptr, h->epoch, h->sc, h->wait_timeout_ns);
  h->prof_end(K_SLAB); h->launches++;
}

// One whole step of a slab in peer mode: the phases above with every message written straight into the neighbour's
// memory and every hand-over a flag, i.e. no host synchronisation and no copy engine work at all.  Asynchronous.
static int step_p2p_once(pbf_handle* h) {
  int rc;
  if ((rc = pbf_slab_phase_predict(h)) != PBF_OK) return rc;      // emigrants -> neighbours' mig_recv
  p2p_exchange_point(h);
  if ((rc = pbf_slab_phase_migrate(h)) != PBF_OK) return rc;      // absorb immigrants; ghost layers -> neighbours' ghost_recv
  p2p_exchange_point(h);
  enqueue_slab_sort(h);                                            // absorb ghosts, sort, ranges -> link->b
  p2p_exchange_point(h);
  h->prof_begin(K_SLAB);
  k_fetch_peer_ranges<<<1, 32, 0, h->stream>>>(h->link, h->has_left ? h->peer[0].link : nullptr, h->has_right ? h->peer[1].link : nullptr, h->sc);
  h->prof_end(K_SLAB); h->launches++;
  enqueue_build(h, 0);
  for (int it = 0; it < h->dp.iterations; it++) {
    enqueue_lambda(h, it == 0, PART_ALL); p2p_exchange_point(h);   // boundary (x*, lambda) already sits in the neighbours' xs_b
    enqueue_delta(h, PART_ALL); p2p_exchange_point(h);             // ... x* in their xs_a
  }
  enqueue_velocity(h);
  enqueue_vorticity(h, PART_ALL); p2p_exchange_point(h);           // ... (x*, |omega|) in their xs_w
  enqueue_confine(h);
  h->steps_done++;
  h->have_neighbors = true;
  return PBF_OK;
}

// Load-time densities (Particles::estimateDensities, particles.cpp:440-444) of a peer-mode slab: committed positions
// re-binned, ghosts exchanged, lists INCLUDING self, one poly6 pass.  Every rank must call it.  Asynchronous.
int pbf_slab_estimate_densities_p2p(pbf_handle* h) {
  if (!h || !h->slab || !h->p2p) return sfail(h, PBF_ERR_INVALID, "pbf_slab_estimate_densities_p2p: not a connected peer-mode slab");
  SCK(h, cudaSetDevice(h->device));
  int rc;
  if ((rc = slab_predict(h, 0)) != PBF_OK) return rc;
  p2p_exchange_point(h);
  if ((rc = pbf_slab_phase_migrate(h)) != PBF_OK) return rc;
  p2p_exchange_point(h);
  enqueue_slab_sort(h);
  p2p_exchange_point(h);
  h->prof_begin(K_SLAB);
  k_fetch_peer_ranges<<<1, 32, 0, h->stream>>>(h->link, h->has_left ? h->peer[0].link : nullptr, h->has_right ? h->peer[1].link : nullptr, h->sc);
  h->prof_end(K_SLAB); h->launches++;
  enqueue_build(h, 1);
  LAUNCH(h, K_DENSITY, k_density_only, blocks_for(owned_ub(h)), h->dp, 0u, 0u, h->xs_a, h->rho, h->nbr, h->slice_off, h->nbr_cnt, link_of(h));
  h->have_neighbors = true;
  SCK(h, cudaGetLastError());
  return PBF_OK;
}

int pbf_slab_step_p2p(pbf_handle* h, int n_steps) {
  if (!h || !h->slab || !h->p2p || n_steps < 0) return sfail(h, PBF_ERR_INVALID, "pbf_slab_step_p2p: not a connected peer-mode slab");
  SCK(h, cudaSetDevice(h->device));
  SCK(h, cudaEventRecord(h->ev_call[0], h->stream));
  for (int s = 0; s < n_steps; s++) { int rc = step_p2p_once(h); if (rc != PBF_OK) return rc; }
  SCK(h, cudaEventRecord(h->ev_call[1], h->stream));
  h->call_timed = true;
  SCK(h, cudaGetLastError());
  return PBF_OK;
}

// After a sync: bring the device-side ranges of the last sort to the host fields the download / digest paths use.
int pbf_slab_refresh_ranges(pbf_handle* h, uint32_t bounds_out[5]) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  int rc = sync_and_check(h);
  if (rc != PBF_OK) return rc;
  if (h->p2p) {
    SlabLink lk;
    SCK(h, cudaMemcpy(&lk, h->link, sizeof(SlabLink), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 5; k++) h->bounds[k] = lk.b[k];
    h->n_sorted = lk.b[4]; h->r_i0 = lk.b[0]; h->r_cnt = lk.b[3] - lk.b[0]; h->n = h->r_cnt;
  }
  if (bounds_out) for (int k = 0; k < 5; k++) bounds_out[k] = h->bounds[k];
  return PBF_OK;
}

}  // extern "C"

// CUDA loads a kernel's code at its first launch, and that load may have to wait until the device is idle: a rank that
// sits in k_exchange for a neighbour whose launches the (blocked) host thread has not issued yet would never be released.
// Everything a peer-mode step launches is therefore loaded before the first step (cudaFuncGetAttributes loads a function).
static void preload_step_kernels() {
  cudaFuncAttributes a;
#define PL(...) cudaFuncGetAttributes(&a, __VA_ARGS__)
  PL(k_predict_hash<true>); PL(k_predict_hash<false>); PL(k_scan_reduce); PL(k_scan_block_sums); PL(k_scan_apply); PL(k_scatter);
  PL(k_cell_sort); PL(k_reorder); PL(k_build_neighbors); PL(k_alert_hist); PL(k_alert_bound); PL(k_neighbor_alert); PL(k_set_sentinel);
  PL(k_lambda); PL(k_delta<4, true>); PL(k_delta<4, false>); PL(k_delta<-1, true>); PL(k_delta<-1, false>); PL(k_velocity);
  PL(k_vorticity_xsph); PL(k_confine_commit); PL(k_write_headers); PL(k_absorb_migrants); PL(k_pack_ghosts); PL(k_absorb_ghosts);
  PL(k_density_only); PL(k_gather_bounds); PL(k_column_hist); PL(k_exchange); PL(k_fetch_peer_ranges); PL(k_neighbor_digest);
  PL(k_export3_f64); PL(k_export1_f64); PL(k_export3); PL(k_export1); PL(k_export_w);
#undef PL
  cudaGetLastError();
}

extern "C" {

// Connect to the x-neighbours' memory.  Same process: the handles themselves (peer access is enabled here).
int pbf_slab_p2p_connect_local(pbf_handle* h, pbf_handle* left, pbf_handle* right) {
  if (!h || !h->slab || !h->link) return sfail(h, PBF_ERR_INVALID, "pbf_slab_p2p_connect_local: configure the slab first");
  SCK(h, cudaSetDevice(h->device));
  pbf_handle* nb[2] = {left, right};
  for (int s = 0; s < 2; s++) {
    pbf_handle* q = nb[s];
    if ((q != nullptr) != (s ? h->has_right : h->has_left)) return sfail(h, PBF_ERR_INVALID, "pbf_slab_p2p_connect_local: neighbours do not match the slab configuration");
    if (!q) continue;
    if (!q->slab || !q->link || q->halo_cap != h->halo_cap) return sfail(h, PBF_ERR_INVALID, "pbf_slab_p2p_connect_local: neighbour not configured alike");
    if (q->device != h->device) {
      int can = 0;
      SCK(h, cudaDeviceCanAccessPeer(&can, h->device, q->device));
      if (!can) return sfail(h, PBF_ERR_CUDA, "no peer access between the devices of adjacent slabs");
      cudaError_t e = cudaDeviceEnablePeerAccess(q->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { h->last_error = cudaGetErrorString(e); return PBF_ERR_CUDA; }
      cudaGetLastError();
    }
    Solver::Peer& P = h->peer[s];
    P.link = q->link; P.xs_a = q->xs_a; P.xs_b = q->xs_b; P.xs_w = q->xs_tmp;
    P.mig_recv = q->mig_recv[s ^ 1]; P.ghost_recv = q->ghost_recv[s ^ 1];      // we are on the OTHER side of the neighbour
    P.ipc = false;
  }
  preload_step_kernels();
  h->p2p = true;
  return PBF_OK;
}

// Other process (one process per GPU): CUDA IPC handles of the six allocations the neighbours touch.
struct P2pBlob { cudaIpcMemHandle_t mem[8]; uint64_t halo_cap; int32_t device; int32_t pad; };
size_t pbf_slab_p2p_blob_size(void) { return sizeof(P2pBlob); }
int pbf_slab_p2p_export(pbf_handle* h, void* blob_out) {
  if (!h || !h->slab || !h->link || !blob_out) return sfail(h, PBF_ERR_INVALID, "pbf_slab_p2p_export: configure the slab first");
  SCK(h, cudaSetDevice(h->device));
  P2pBlob b; std::memset(&b, 0, sizeof(b));
  void* ptr[8] = {h->link, h->xs_a, h->xs_b, h->xs_tmp, h->mig_recv[0], h->mig_recv[1], h->ghost_recv[0], h->ghost_recv[1]};
  for (int k = 0; k < 8; k++) SCK(h, cudaIpcGetMemHandle(&b.mem[k], ptr[k]));
  b.halo_cap = h->halo_cap; b.device = h->device;
  std::memcpy(blob_out, &b, sizeof(b));
  return PBF_OK;
}
int pbf_slab_p2p_connect_ipc(pbf_handle* h, const void* left_blob, const void* right_blob) {
  if (!h || !h->slab || !h->link) return sfail(h, PBF_ERR_INVALID, "pbf_slab_p2p_connect_ipc: configure the slab first");
  SCK(h, cudaSetDevice(h->device));
  const void* nb[2] = {left_blob, right_blob};
  for (int s = 0; s < 2; s++) {
    if ((nb[s] != nullptr) != (s ? h->has_right : h->has_left)) return sfail(h, PBF_ERR_INVALID, "pbf_slab_p2p_connect_ipc: neighbours do not match the slab configuration");
    if (!nb[s]) continue;
    P2pBlob b; std::memcpy(&b, nb[s], sizeof(b));
    if (b.halo_cap != h->halo_cap) return sfail(h, PBF_ERR_INVALID, "pbf_slab_p2p_connect_ipc: neighbour configured with another halo_cap");
    Solver::Peer& P = h->peer[s];
    const int want[6] = {0, 1, 2, 3, 4 + (s ^ 1), 6 + (s ^ 1)};      // link, xs_a, xs_b, xs_w, its mig_recv / ghost_recv for OUR side
    void* got[6];
    for (int k = 0; k < 6; k++) { SCK(h, cudaIpcOpenMemHandle(&got[k], b.mem[want[k]], cudaIpcMemLazyEnablePeerAccess)); P.ipc_base[k] = got[k]; }
    P.link = (SlabLink*)got[0]; P.xs_a = (float4*)got[1]; P.xs_b = (float4*)got[2]; P.xs_w = (float4*)got[3];
    P.mig_recv = (float4*)got[4]; P.ghost_recv = (float4*)got[5];
    P.ipc = true;
  }
  preload_step_kernels();
  h->p2p = true;
  return PBF_OK;
}
// Back to the host-driven protocol (a driver falls back to it when some rank could not map its neighbours).
int pbf_slab_p2p_disconnect(pbf_handle* h) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  SCK(h, cudaSetDevice(h->device));
  SCK(h, cudaStreamSynchronize(h->stream));
  for (int s = 0; s < 2; s++) {
    Solver::Peer& P = h->peer[s];
    if (P.ipc) for (void*& q : P.ipc_base) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    P = Solver::Peer();
  }
  h->p2p = false;
  return PBF_OK;
}
int pbf_slab_set_wait_timeout(pbf_handle* h, double seconds) {
  if (!h || !(seconds > 0)) return PBF_ERR_INVALID;
  h->wait_timeout_ns = (long long)(seconds * 1e9);
  return PBF_OK;
}

int pbf_slab_phase(pbf_handle* h, int phase) { return pbf_slab_phase_part(h, phase, PBF_PART_ALL); }

int pbf_slab_phase_part(pbf_handle* h, int phase, int part) {
  if (!h || !h->slab || part < PBF_PART_ALL || part > PBF_PART_INTERIOR) return PBF_ERR_INVALID;
  SCK(h, cudaSetDevice(h->device));
  switch (phase) {
    case PBF_PHASE_LAMBDA_FIRST: enqueue_lambda(h, 1, part); break;
    case PBF_PHASE_LAMBDA: enqueue_lambda(h, 0, part); break;
    case PBF_PHASE_DELTA: enqueue_delta(h, part); break;
    case PBF_PHASE_VELOCITY: enqueue_velocity(h); break;
    case PBF_PHASE_VORTICITY: enqueue_vorticity(h, part); break;
    case PBF_PHASE_CONFINE: enqueue_confine(h); h->steps_done++; break;
    default: return sfail(h, PBF_ERR_INVALID, "unknown phase");
  }
  SCK(h, cudaGetLastError());
  return PBF_OK;
}

int pbf_slab_stats(pbf_handle* h, double* rho_first_sum, double* rho_final_sum, uint64_t* n_owned) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  int rc = pbf_slab_refresh_ranges(h, nullptr);
  if (rc != PBF_OK) return rc;
  Scalars s;
  SCK(h, cudaMemcpy(&s, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost));
  if (rho_first_sum) *rho_first_sum = s.rho_first;
  if (rho_final_sum) *rho_final_sum = s.rho_final;
  if (n_owned) *n_owned = h->r_cnt;
  return PBF_OK;
}

int pbf_slab_download(pbf_handle* h, size_t cap, double* pos_xyz, double* vel_xyz, double* density, uint32_t* ids, size_t* n_out) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  SCK(h, cudaSetDevice(h->device));
  { int rc0 = pbf_slab_refresh_ranges(h, nullptr); if (rc0 != PBF_OK) return rc0; }
  const size_t n = h->r_cnt;
  if (n_out) *n_out = n;
  if (n > cap) return sfail(h, PBF_ERR_CAPACITY, "pbf_slab_download: output buffers too small");
  int rc = io_download(h, pos_xyz, vel_xyz, density);    // range order (orig == nullptr in slab mode)
  if (rc != PBF_OK) return rc;
  if (ids && n) SCK(h, cudaMemcpy(ids, h->orig[h->cur] + h->r_i0, n * 4, cudaMemcpyDeviceToHost));
  return PBF_OK;
}

int pbf_slab_neighbor_digest(pbf_handle* h, size_t cap, uint64_t* digest, uint32_t* count) {
  if (!h || !h->slab || !digest || !count) return PBF_ERR_INVALID;
  if (!h->have_neighbors) return sfail(h, PBF_ERR_INVALID, "no neighbour lists yet");
  SCK(h, cudaSetDevice(h->device));
  { int rc0 = pbf_slab_refresh_ranges(h, nullptr); if (rc0 != PBF_OK) return rc0; }
  const size_t n = h->r_cnt;
  if (n > cap) return sfail(h, PBF_ERR_CAPACITY, "output buffers too small");
  unsigned long long* dd = nullptr; uint32_t* dc = nullptr;
  SCK(h, cudaMalloc((void**)&dd, std::max<size_t>(n, 1) * 8)); SCK(h, cudaMalloc((void**)&dc, std::max<size_t>(n, 1) * 4));
  enqueue_digest(h, dd, dc);
  int rc = sync_and_check(h);
  if (rc == PBF_OK && n) {
    cudaMemcpy(digest, dd, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(count, dc, n * 4, cudaMemcpyDeviceToHost);
  }
  cudaFree(dd); cudaFree(dc);
  return rc;
}

void* pbf_slab_buffer(pbf_handle* h, int which, size_t* bytes_out) {
  if (!h || !h->slab) return nullptr;
  const size_t mig = (1 + 3 * h->halo_cap) * sizeof(float4), gh = (1 + 2 * h->halo_cap) * sizeof(float4), arr = h->cap * sizeof(float4);
  void* p = nullptr; size_t b = 0;
  switch (which) {
    case PBF_BUF_MIG_SEND_L: p = h->mig_send[0]; b = mig; break;
    case PBF_BUF_MIG_SEND_R: p = h->mig_send[1]; b = mig; break;
    case PBF_BUF_MIG_RECV_L: p = h->mig_recv[0]; b = mig; break;
    case PBF_BUF_MIG_RECV_R: p = h->mig_recv[1]; b = mig; break;
    case PBF_BUF_GHOST_SEND_L: p = h->ghost_send[0]; b = gh; break;
    case PBF_BUF_GHOST_SEND_R: p = h->ghost_send[1]; b = gh; break;
    case PBF_BUF_GHOST_RECV_L: p = h->ghost_recv[0]; b = gh; break;
    case PBF_BUF_GHOST_RECV_R: p = h->ghost_recv[1]; b = gh; break;
    case PBF_BUF_XS_A: p = h->xs_a; b = arr; break;
    case PBF_BUF_XS_B: p = h->xs_b; b = arr; break;
    case PBF_BUF_OMEGA: p = h->omega; b = arr; break;
    case PBF_BUF_XS_W: p = h->xs_tmp; b = arr; break;
    default: break;
  }
  if (bytes_out) *bytes_out = b;
  return p;
}

}  // extern "C"
