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

__global__ void k_gather_bounds(const uint32_t* __restrict__ cell_start, uint32_t c1, uint32_t c2, uint32_t cm, uint32_t cm1,
                                uint32_t ncell, Scalars* sc) {
  if (threadIdx.x || blockIdx.x) return;
  sc->bounds[0] = cell_start[c1]; sc->bounds[1] = cell_start[c2]; sc->bounds[2] = cell_start[cm];
  sc->bounds[3] = cell_start[cm1]; sc->bounds[4] = cell_start[ncell];
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

int pbf_slab_configure(pbf_handle* h, int gx_lo, int gx_hi, int left_cols, int right_cols, size_t particle_cap, size_t halo_cap) {
  if (!h) return PBF_ERR_INVALID;
  if (h->n != 0 || h->cap != 0) return sfail(h, PBF_ERR_INVALID, "pbf_slab_configure must precede any upload");
  DevParams& d = h->dp;
  if (gx_lo < 0 || gx_hi > d.gdim_x_global || gx_hi <= gx_lo || halo_cap == 0) return sfail(h, PBF_ERR_INVALID, "bad slab range");
  if (h->hp.xsph_mode != PBF_XSPH_JACOBI) return sfail(h, PBF_ERR_INVALID, "slab mode supports PBF_XSPH_JACOBI only (reference order is a global sequential dependency)");
  SCK(h, cudaSetDevice(h->device));
  d.gx_lo = gx_lo; d.gx_hi = gx_hi; d.cx_offset = gx_lo - 1;
  d.hop_left = left_cols; d.hop_right = right_cols;
  d.gdim[0] = (gx_hi - gx_lo) + 2;                      // one ghost column on each side
  h->has_left = left_cols > 0; h->has_right = right_cols > 0;
  h->ncell = (uint32_t)((size_t)d.gdim[0] * d.gdim[1] * d.gdim[2]);
  cudaFree(h->cell_count); cudaFree(h->cell_start); cudaFree(h->block_sums);
  SCK(h, cudaMalloc((void**)&h->cell_count, ((size_t)h->ncell + 1) * 4));
  SCK(h, cudaMalloc((void**)&h->cell_start, ((size_t)h->ncell + 2) * 4));
  SCK(h, cudaMalloc((void**)&h->block_sums, ((size_t)h->ncell / 2048 + 2) * 4));
  h->slab = true; h->halo_cap = halo_cap;
  int rc = alloc_particle_arrays(h, particle_cap + 4 * halo_cap);
  if (rc != PBF_OK) return rc;
  for (int k = 0; k < 2; k++) {
    SCK(h, cudaMalloc((void**)&h->mig_send[k], (1 + 3 * halo_cap) * sizeof(float4)));
    SCK(h, cudaMalloc((void**)&h->mig_recv[k], (1 + 3 * halo_cap) * sizeof(float4)));
    SCK(h, cudaMalloc((void**)&h->ghost_send[k], (1 + 2 * halo_cap) * sizeof(float4)));
    SCK(h, cudaMalloc((void**)&h->ghost_recv[k], (1 + 2 * halo_cap) * sizeof(float4)));
    SCK(h, cudaMemset(h->mig_send[k], 0, sizeof(float4))); SCK(h, cudaMemset(h->mig_recv[k], 0, sizeof(float4)));
    SCK(h, cudaMemset(h->ghost_send[k], 0, sizeof(float4))); SCK(h, cudaMemset(h->ghost_recv[k], 0, sizeof(float4)));
  }
  return PBF_OK;
}

int pbf_slab_upload(pbf_handle* h, size_t n, const double* pos_xyz, const double* vel_xyz, const uint32_t* ids) {
  if (!h || !h->slab || (n && (!pos_xyz || !vel_xyz || !ids))) return sfail(h, PBF_ERR_INVALID, "pbf_slab_upload: bad argument / not configured");
  if (n + 4 * h->halo_cap + 33 > h->cap) return sfail(h, PBF_ERR_CAPACITY, "pbf_slab_upload: more particles than particle_cap");
  SCK(h, cudaSetDevice(h->device));
  h->n = n; h->cur = 0; h->have_neighbors = false;
  h->r_i0 = 0; h->r_cnt = (uint32_t)n; h->n_sorted = (uint32_t)n;
  if (n == 0) return PBF_OK;
  int rc = io_upload(h, n, pos_xyz, vel_xyz);            // sets orig = identity ...
  if (rc != PBF_OK) return rc;
  SCK(h, cudaMemcpy(h->orig[0], ids, n * 4, cudaMemcpyHostToDevice));   // ... replaced by the global ids
  return PBF_OK;
}

// A. predict owned particles; emigrants go to the migration messages
int pbf_slab_phase_predict(pbf_handle* h) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  SCK(h, cudaSetDevice(h->device));
  if ((size_t)h->n_sorted + 4 * h->halo_cap + 33 > h->cap) return sfail(h, PBF_ERR_CAPACITY, "slab particle capacity exceeded");
  cudaMemsetAsync(&h->sc->rho_first, 0, 2 * sizeof(double), h->stream);
  enqueue_predict_hash(h, 1);
  h->prof_begin(K_SLAB);
  k_write_headers<<<1, 32, 0, h->stream>>>(h->sc, h->has_left ? h->mig_send[0] : nullptr, h->has_right ? h->mig_send[1] : nullptr,
                                           nullptr, nullptr, (uint32_t)h->halo_cap, 0);
  h->prof_end(K_SLAB); h->launches++;
  SCK(h, cudaGetLastError());
  return PBF_OK;
}

// C. append immigrants (slots [n_prev, n_prev + 2*cap)), then pack the ghost layers to send
int pbf_slab_phase_migrate(pbf_handle* h) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  SCK(h, cudaSetDevice(h->device));
  const uint32_t cap = (uint32_t)h->halo_cap, n_prev = h->n_sorted;
  const int cur = h->cur;
  for (int side = 0; side < 2; side++) {
    if (!(side ? h->has_right : h->has_left)) continue;
    LAUNCH(h, K_SLAB, k_absorb_migrants, blocks_for(cap), h->dp, h->mig_recv[side], cap, n_prev + side * cap, h->pos[cur], h->vel[cur],
           h->orig[cur], h->xs_tmp, h->cell_of, h->rank, h->cell_count, h->sc);
  }
  LAUNCH(h, K_SLAB, k_pack_ghosts, blocks_for((size_t)n_prev + 2 * cap), h->dp, n_prev + 2 * cap, h->cell_of, h->xs_tmp, h->pos[cur],
         h->orig[cur], h->has_left ? h->ghost_send[0] : (float4*)nullptr, h->has_right ? h->ghost_send[1] : (float4*)nullptr, cap, h->sc);
  h->prof_begin(K_SLAB);
  k_write_headers<<<1, 32, 0, h->stream>>>(h->sc, nullptr, nullptr, h->has_left ? h->ghost_send[0] : nullptr,
                                           h->has_right ? h->ghost_send[1] : nullptr, cap, 1);
  h->prof_end(K_SLAB); h->launches++;
  SCK(h, cudaGetLastError());
  return PBF_OK;
}

// E. append ghosts (slots [n_prev + 2*cap, n_prev + 4*cap)), sort, read the column boundaries back,
//    build the neighbour lists of the owned range
int pbf_slab_phase_sort(pbf_handle* h, uint32_t bounds_out[5]) {
  if (!h || !h->slab) return PBF_ERR_INVALID;
  SCK(h, cudaSetDevice(h->device));
  const uint32_t cap = (uint32_t)h->halo_cap, n_prev = h->n_sorted;
  const int cur = h->cur;
  for (int side = 0; side < 2; side++) {
    if (!(side ? h->has_right : h->has_left)) continue;
    LAUNCH(h, K_SLAB, k_absorb_ghosts, blocks_for(cap), h->dp, h->ghost_recv[side], cap, n_prev + (2 + side) * cap, side, h->pos[cur],
           h->vel[cur], h->orig[cur], h->xs_tmp, h->cell_of, h->rank, h->cell_count, h->sc);
  }
  enqueue_sort(h, (size_t)n_prev + 4 * cap);
  const uint32_t gyz = (uint32_t)(h->dp.gdim[1] * h->dp.gdim[2]);
  const uint32_t m = (uint32_t)(h->dp.gx_hi - h->dp.gx_lo);
  h->prof_begin(K_SLAB);
  k_gather_bounds<<<1, 32, 0, h->stream>>>(h->cell_start, gyz, 2 * gyz, m * gyz, (m + 1) * gyz, h->ncell, h->sc);
  h->prof_end(K_SLAB); h->launches++;
  int rc = sync_and_check(h);
  if (rc != PBF_OK) return rc;
  Scalars s;
  SCK(h, cudaMemcpy(&s, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 5; k++) h->bounds[k] = s.bounds[k];
  h->n_sorted = s.bounds[4];
  h->r_i0 = s.bounds[0]; h->r_cnt = s.bounds[3] - s.bounds[0];
  h->n = h->r_cnt;
  if ((size_t)h->n_sorted + 33 > h->cap) return sfail(h, PBF_ERR_CAPACITY, "slab particle capacity exceeded");   // +1 sentinel, +32 unconditional candidate groups
  enqueue_build(h, 0);
  h->have_neighbors = true;
  if (bounds_out) for (int k = 0; k < 5; k++) bounds_out[k] = h->bounds[k];
  SCK(h, cudaGetLastError());
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
  int rc = sync_and_check(h);
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
  const size_t n = h->r_cnt;
  if (n > cap) return sfail(h, PBF_ERR_CAPACITY, "output buffers too small");
  SCK(h, cudaSetDevice(h->device));
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
