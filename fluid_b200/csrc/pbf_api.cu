// pbf_api.cu — C ABI (include/pbf_b200.h) over the CUDA solver.  No torch types, no CPU fallback:
// without a CUDA device every entry point fails with PBF_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "pbf_internal.h"

using namespace pbf;

namespace {

#define CK(h, call)                                                                         \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      (h)->last_error = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
      return PBF_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

int fail(pbf_handle* h, int code, const char* msg) { if (h) h->last_error = msg; return code; }

template <class T> cudaError_t dmalloc(T** p, size_t count) { return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)); }

// Parameters in fp32, computed with plain IEEE float operations (this TU is host code compiled by
// the host compiler without fast-math; no contraction is possible on constants folded here).
}  // namespace
int pbf::fill_dev_params(const PbfParams& p, DevParams& d, std::string& err) {
  if (!(p.h > 0) || !(p.dt > 0) || !(p.rest_density > 0) || p.iterations < 0 || p.n_corr < 0) { err = "invalid PbfParams"; return PBF_ERR_INVALID; }
  if (p.xsph_mode != PBF_XSPH_JACOBI && p.xsph_mode != PBF_XSPH_REFERENCE_ORDER) { err = "unknown xsph_mode"; return PBF_ERR_INVALID; }
  volatile float h = (float)p.h;
  d.h = h; d.h2 = h * h; d.dt = (float)p.dt; d.inv_dt = 1.0f / d.dt;
  d.rho0 = (float)p.rest_density; d.inv_rho0 = 1.0f / d.rho0;
  d.eps_relax = (float)p.eps_relax; d.kcorr = (float)p.k_corr; d.visc_c = (float)p.visc_c;
  d.vort_dt_eps = d.dt * (float)p.vort_eps;
  d.gdt = (float)(-p.gravity_y) * d.dt;            // velocity.y -= 10 * delta_t
  d.eps_d = 1e-11f;
  const double hd = (double)(float)p.h;
  d.poly6_c = (float)(1.56668147106 / std::pow(hd, 9));
  d.spiky_c = (float)(-3.0 * 4.774648292756860 / std::pow(hd, 6));
  const double dq = p.dq_ratio * hd, t = hd * hd - dq * dq;
  const double wdq = 1.56668147106 * t * t * t / std::pow(hd, 9);          // poly6(0,0,dq)
  d.tscale_c = (float)((1.56668147106 / std::pow(hd, 9)) / wdq);
  d.n_corr = p.n_corr; d.iterations = p.iterations;
  d.enable_vorticity = p.enable_vorticity; d.enable_xsph = p.enable_xsph;
  const float cell = d.h * (1.0f + 1.0f / 256.0f);
  d.inv_cell = 1.0f / cell;
  d.zsub = 8;                                          // thin cells along z (measured: 1 / 4 / 8)
  if (const char* e = getenv("PBF_ZSUB")) d.zsub = std::min(16, std::max(1, atoi(e)));
  d.inv_cell_z = (float)d.zsub / cell;
  double ncell = 1;
  for (int a = 0; a < 3; a++) {
    if (!(p.box_max[a] > p.box_min[a])) { err = "empty box"; return PBF_ERR_INVALID; }
    d.bmin[a] = (float)p.box_min[a]; d.bmax[a] = (float)p.box_max[a];
    volatile float lo = d.bmin[a] + d.eps_d, hi = d.bmax[a] - d.eps_d;
    d.clo[a] = lo; d.chi[a] = hi;
    d.gmin[a] = d.bmin[a];
    d.gdim[a] = (int)std::floor((p.box_max[a] - p.box_min[a]) / (double)cell) + 1;
    if (a == 2) d.gdim[a] *= d.zsub;
    ncell *= d.gdim[a];
  }
  d.yl = (float)p.y_light; d.zf = (float)p.z_front;
  for (int a = 0; a < 3; a++) {   // collide fast path: inside these bounds (both ends of a move) nothing can be hit
    const float margin = 1e-4f * (1.0f + std::max(std::fabs(d.bmin[a]), std::fabs(d.bmax[a])));
    d.slo[a] = d.bmin[a] + margin; d.shi[a] = d.bmax[a] - margin;
  }
  d.shi[1] = std::min(d.shi[1], d.yl - 1e-4f * (1.0f + std::fabs(d.yl)));
  d.shi[2] = std::min(d.shi[2], d.zf - 1e-4f * (1.0f + std::fabs(d.zf)));
  if (ncell > 1.5e9) { err = "grid too large (box / h)"; return PBF_ERR_INVALID; }
  d.gdim_x_global = d.gdim[0]; d.cx_offset = 0; d.gx_lo = 0; d.gx_hi = d.gdim[0]; d.hop_left = 0; d.hop_right = 0;
  d.n_sph = 0; d.n_sm = 148; d.one = 1.0f; d.n_tri = 0; d.tri = nullptr; d.bvh = nullptr; d.tri_id = nullptr;
  {   // fp32 contact rules of the obstacle triangles: same expressions (in double) as Oracle<float>'s constructor
    double m = 0;
    for (int a = 0; a < 3; a++) m = std::max(m, std::max(std::fabs((double)p.box_min[a]), std::fabs((double)p.box_max[a])));
    const double ulp = m * 1.1920928955078125e-07;
    d.skin = (float)std::max(1e-5 * (double)p.h, 16.0 * ulp);
    d.tol_n = (float)(std::max(1e-4 * (double)p.h, 64.0 * ulp) + std::max(1e-5 * (double)p.h, 16.0 * ulp));
    d.tol_ray = (float)std::max(0.05 * (double)p.h, 16.0 * std::max(1e-4 * (double)p.h, 64.0 * ulp));
  }
  for (int a = 0; a < 3; a++) { d.tlo[a] = 0.f; d.thi[a] = 0.f; d.olo[a] = 1e30f; d.ohi[a] = -1e30f; }
  for (int k = 0; k < PBF_MAX_SPHERES; k++) { d.sph[k] = make_float4(0.f, 0.f, 0.f, 0.f); d.sph_r2[k] = 0.f; }
  return PBF_OK;
}
namespace {

void free_arrays(pbf_handle* h) {
  for (int b = 0; b < 2; b++) { cudaFree(h->pos[b]); cudaFree(h->vel[b]); cudaFree(h->orig[b]); h->pos[b] = h->vel[b] = nullptr; h->orig[b] = nullptr; }
  cudaFree(h->xs_tmp); cudaFree(h->xs_a); cudaFree(h->xs_b); cudaFree(h->vtmp); cudaFree(h->omega); cudaFree(h->xpred); cudaFree(h->xv);
  h->xv = nullptr;
  cudaFree(h->rho); cudaFree(h->cell_of); cudaFree(h->rank); cudaFree(h->perm); cudaFree(h->key);
  cudaFree(h->nbr); cudaFree(h->slice_off); cudaFree(h->nbr_cnt); cudaFree(h->io_stage);
  h->xs_tmp = h->xs_a = h->xs_b = h->vtmp = h->omega = h->xpred = nullptr; h->rho = nullptr;
  h->cell_of = h->rank = h->perm = h->key = h->nbr = h->slice_off = h->nbr_cnt = nullptr; h->io_stage = nullptr;
  h->cap = 0; h->nbr_cap_rows = 0;
}

}  // namespace
// capacity in particles (sorted entries incl. ghosts and append slots); +1 for the sentinel
int pbf::alloc_particle_arrays(Solver* hs, size_t n) {
  pbf_handle* h = static_cast<pbf_handle*>(hs);
  // +1: the sentinel particle lives at index n_sorted; +32: the neighbour build reads candidates in
  // unconditional groups of 8 and may run past the last particle (masked off)
  const size_t cap = (n + 1 + 31) / 32 * 32 + 32;
  if (cap <= h->cap) return PBF_OK;               // the PADDED requirement: a larger n that fits the old rounding must not lose the +32
  free_arrays(h);
  for (int b = 0; b < 2; b++) { CK(h, dmalloc(&h->pos[b], cap)); CK(h, dmalloc(&h->vel[b], cap)); CK(h, dmalloc(&h->orig[b], cap)); }
  CK(h, dmalloc(&h->xs_tmp, cap)); CK(h, dmalloc(&h->xs_a, cap)); CK(h, dmalloc(&h->xs_b, cap));
  CK(h, dmalloc(&h->vtmp, cap)); CK(h, dmalloc(&h->omega, cap)); CK(h, dmalloc(&h->rho, cap)); CK(h, dmalloc(&h->xv, 2 * cap));
  CK(h, dmalloc(&h->cell_of, cap)); CK(h, dmalloc(&h->rank, cap)); CK(h, dmalloc(&h->perm, cap)); CK(h, dmalloc(&h->key, cap));
  CK(h, dmalloc(&h->slice_off, cap / 32 + 1)); CK(h, dmalloc(&h->nbr_cnt, cap));
  CK(h, dmalloc(&h->io_stage, cap * 7));
  if (h->capture_xpred) CK(h, dmalloc(&h->xpred, cap));
  // neighbour rows: one row = 32 uint4 = 4 entries per lane of a slice.  Default 48 rows per slice
  // = room for 192 neighbours per particle (lattice spacing h/3 has 122); PBF_NBR_ROWS overrides.
  // Overflow is an error, never a truncation.
  size_t rows_per_slice = 48;
  if (const char* e = getenv("PBF_NBR_ROWS")) rows_per_slice = (size_t)std::max(4, atoi(e));
  h->nbr_cap_rows = (cap / 32) * rows_per_slice;
  CK(h, dmalloc(&h->nbr, h->nbr_cap_rows * 128));
  h->cap = cap;
  return PBF_OK;
}
namespace {
int ensure_capacity(pbf_handle* h, size_t n) { return alloc_particle_arrays(h, n); }

// double <-> float conversion of big host arrays, spread over a few host threads
template <class F> void parallel_for(size_t n, F fn) {
  unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
  if (n < (1u << 16)) nt = 1;
  if (nt == 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  const size_t chunk = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; t++) { size_t a = t * chunk, b = std::min(n, a + chunk); if (a < b) th.emplace_back(fn, a, b); }
  for (auto& t : th) t.join();
}

struct PinnedBuf {
  float* p = nullptr; size_t cap = 0;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost((void**)&p, n * sizeof(float));
    if (e == cudaSuccess) cap = n;
    return e;
  }
};

int check_device_errors(pbf_handle* h) {
  Scalars s;
  CK(h, cudaMemcpy(&s, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost));
  if (s.err) {
    int zero = 0;
    cudaMemcpy(&h->sc->err, &zero, sizeof(int), cudaMemcpyHostToDevice);
    if (s.err & ERRBIT_NONFINITE) return fail(h, PBF_ERR_DOMAIN, "non-finite particle position produced by the step");
    if (s.err & ERRBIT_PEER_TIMEOUT) {
      const std::string who = s.timeout_missing == 3 ? "both neighbours" : (s.timeout_missing == 1 ? "the left neighbour" : "the right neighbour");
      return fail(h, PBF_ERR_CUDA, ("peer mode: " + who + " did not reach exchange point " + std::to_string(s.timeout_epoch) + " in time").c_str());
    }
    if (s.err & ERRBIT_MIGRATION) return fail(h, PBF_ERR_DOMAIN, "particle left its slab by more than one cell column in one step");
    if (s.err & ERRBIT_NBR_CAPACITY) return fail(h, PBF_ERR_CAPACITY, "neighbour list capacity exceeded (raise PBF_NBR_ROWS); results of this step are invalid");
    if (s.err & ERRBIT_HALO_CAPACITY) return fail(h, PBF_ERR_CAPACITY, "halo / migration buffer capacity exceeded");
    if (s.err & ERRBIT_SLAB_CAPACITY) return fail(h, PBF_ERR_CAPACITY, "slab particle capacity exceeded (owned + ghost particles > particle_cap)");
    if (s.err & ERRBIT_PEER_MISMATCH) return fail(h, PBF_ERR_DOMAIN, "peer mode: boundary column and the neighbour's ghost range differ in length");
    return fail(h, PBF_ERR_CUDA, "unknown device error flag");
  }
  return PBF_OK;
}

}  // namespace

int pbf::sync_and_check(Solver* hs) {
  pbf_handle* h = static_cast<pbf_handle*>(hs);
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaStreamSynchronize(h->stream));
  if (h->copy_stream) CK(h, cudaStreamSynchronize(h->copy_stream));
  h->rb_pending = false;
  CK(h, cudaGetLastError());
  if (h->call_timed) { float ms = 0.f; cudaEventElapsedTime(&ms, h->ev_call[0], h->ev_call[1]); h->last_call_ms = ms; h->call_timed = false; }
  h->prof_collect();
  return check_device_errors(h);
}

// pinned staging lives outside the struct so pbf_internal.h stays CUDA-only
struct HostRange { char* p; size_t bytes; };
struct HandleExtra {
  PinnedBuf pin;
  std::vector<HostRange> regs;       // caller buffers page-locked through pbf_host_register
  double* stage64 = nullptr; size_t stage64_cap = 0;   // device staging for fp64 AoS (7 doubles / particle)
  bool registered(const void* q, size_t bytes) const {
    const char* c = (const char*)q;
    for (const HostRange& r : regs) if (c >= r.p && c + bytes <= r.p + r.bytes) return true;
    return false;
  }
  ~HandleExtra() { for (HostRange& r : regs) cudaHostUnregister(r.p); if (stage64) cudaFree(stage64); }
  cudaError_t ensure_stage64(size_t n) {
    if (n <= stage64_cap) return cudaSuccess;
    if (stage64) cudaFree(stage64);
    stage64 = nullptr; stage64_cap = 0;
    cudaError_t e = cudaMalloc((void**)&stage64, n * sizeof(double));
    if (e == cudaSuccess) stage64_cap = n;
    return e;
  }
};
// owned by the handle (Solver::host_extra), created on first use, freed in pbf_destroy: no global state
static HandleExtra* extra_of(pbf_handle* h) {
  if (!h->host_extra) h->host_extra = new HandleExtra();
  return static_cast<HandleExtra*>(h->host_extra);
}

extern "C" {

void pbf_default_params(PbfParams* p) {
  std::memset(p, 0, sizeof(*p));
  p->h = 0.3; p->dt = 0.016; p->rest_density = 1000.0; p->eps_relax = 2.0; p->k_corr = 0.0001;
  p->dq_ratio = 0.1; p->visc_c = 0.001; p->vort_eps = 0.001; p->gravity_y = -10.0;
  p->n_corr = 4; p->iterations = 12;
  p->box_min[0] = -1; p->box_min[1] = 0; p->box_min[2] = -1;
  p->box_max[0] = 1; p->box_max[1] = 1.49; p->box_max[2] = 1;
  p->y_light = 1.49; p->z_front = 1.0;
  p->xsph_mode = PBF_XSPH_JACOBI; p->enable_vorticity = 1; p->enable_xsph = 1;
}

int pbf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int pbf_create(const PbfParams* params, int device_id, pbf_handle** out) {
  if (!params || !out) return PBF_ERR_INVALID;
  *out = nullptr;
  const int ndev = pbf_device_count();
  if (ndev <= 0) return PBF_ERR_NO_DEVICE;
  if (device_id < 0 || device_id >= ndev) return PBF_ERR_INVALID;
  pbf_handle* h = new pbf_handle();
  h->device = device_id; h->hp = *params;
  std::string err;
  int rc = fill_dev_params(*params, h->dp, err);
  if (rc != PBF_OK) { fprintf(stderr, "pbf_create: %s\n", err.c_str()); delete h; return rc; }
  h->ncell = (uint32_t)((size_t)h->dp.gdim[0] * h->dp.gdim[1] * h->dp.gdim[2]);
  { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_id) == cudaSuccess && sms > 0) h->dp.n_sm = sms; }
  auto bail = [&](cudaError_t e) { fprintf(stderr, "pbf_create: %s\n", cudaGetErrorString(e)); delete h; return PBF_ERR_CUDA; };
  cudaError_t e;
  if ((e = cudaSetDevice(device_id)) != cudaSuccess) return bail(e);
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
  if ((e = dmalloc(&h->cell_count, (size_t)h->ncell + 1)) != cudaSuccess) return bail(e);
  if ((e = dmalloc(&h->cell_start, (size_t)h->ncell + 2)) != cudaSuccess) return bail(e);
  if ((e = dmalloc(&h->block_sums, (size_t)h->ncell / 2048 + 2)) != cudaSuccess) return bail(e);
  if ((e = dmalloc(&h->sc, 1)) != cudaSuccess) return bail(e);
  if ((e = cudaMemset(h->sc, 0, sizeof(Scalars))) != cudaSuccess) return bail(e);
  cudaEventCreate(&h->ev_call[0]); cudaEventCreate(&h->ev_call[1]);
  *out = h;
  return PBF_OK;
}

void pbf_destroy(pbf_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->prof_collect();
  h->graph_invalidate();
  for (auto e : h->event_pool) cudaEventDestroy(e);
  free_arrays(h);
  cudaFree(h->cell_count); cudaFree(h->cell_start); cudaFree(h->block_sums); cudaFree(h->sc);
  if (h->ev_call[0]) cudaEventDestroy(h->ev_call[0]);
  if (h->ev_call[1]) cudaEventDestroy(h->ev_call[1]);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); for (int k = 0; k < 4; k++) if (h->ev_rb[k]) cudaEventDestroy(h->ev_rb[k]); }
  if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
  for (int k = 0; k < 2; k++) { cudaFree(h->mig_send[k]); cudaFree(h->mig_recv[k]); cudaFree(h->ghost_send[k]); cudaFree(h->ghost_recv[k]); }
  for (int k = 0; k < 2; k++) if (h->peer[k].ipc) for (void* p : h->peer[k].ipc_base) if (p) cudaIpcCloseMemHandle(p);
  cudaFree(h->link);
  if (h->col_hist_host) cudaFreeHost(h->col_hist_host);
  if (h->ev_hist) cudaEventDestroy(h->ev_hist);
  delete static_cast<HandleExtra*>(h->host_extra);
  cudaFree(h->tri_dev); cudaFree(h->bvh_dev); cudaFree(h->tri_id_dev); cudaFree(h->alert_buf);
  delete h;
}

const char* pbf_last_error(pbf_handle* h) { return h ? h->last_error.c_str() : "null handle"; }
size_t pbf_num_particles(pbf_handle* h) { return h ? h->n : 0; }
uint64_t pbf_launch_count(pbf_handle* h) { return h ? h->launches : 0; }

}  // extern "C" (reopened below)

// host fp64 AoS -> pos/vel[cur] (orig = identity), page-locked fast path when the buffers are registered
int pbf::io_upload(Solver* hs, size_t n, const double* pos_xyz, const double* vel_xyz) {
  pbf_handle* h = static_cast<pbf_handle*>(hs);
  if (n == 0) return PBF_OK;
  HandleExtra* x = extra_of(h);
  if (x->registered(pos_xyz, 3 * n * sizeof(double)) && x->registered(vel_xyz, 3 * n * sizeof(double))) {
    // page-locked caller buffers: DMA the doubles as they are, convert on the device
    CK(h, x->ensure_stage64(7 * h->cap));
    CK(h, cudaMemcpyAsync(x->stage64, pos_xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(x->stage64 + 3 * n, vel_xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    enqueue_import_f64(h, x->stage64, x->stage64 + 3 * n);
  } else {
    CK(h, x->pin.ensure(6 * n));
    float* st = x->pin.p;
    parallel_for(3 * n, [&](size_t a, size_t b) {
      for (size_t i = a; i < b; i++) { st[i] = (float)pos_xyz[i]; st[3 * n + i] = (float)vel_xyz[i]; }
    });
    CK(h, cudaMemcpyAsync(h->io_stage, st, 6 * n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    enqueue_import(h, h->io_stage, h->io_stage + 3 * n);
  }
  CK(h, cudaStreamSynchronize(h->stream));
  CK(h, cudaGetLastError());
  return PBF_OK;
}

// owned range of pos/vel[cur] and rho -> host fp64 AoS (original order, or range order in slab mode)
int pbf::io_download(Solver* hs, double* pos_xyz, double* vel_xyz, double* density) {
  pbf_handle* h = static_cast<pbf_handle*>(hs);
  const size_t n = h->r_cnt;
  if (n == 0) return sync_and_check(h);
  HandleExtra* x = extra_of(h);
  if ((!pos_xyz || x->registered(pos_xyz, 3 * n * sizeof(double))) && (!vel_xyz || x->registered(vel_xyz, 3 * n * sizeof(double))) &&
      (!density || x->registered(density, n * sizeof(double)))) {
    // page-locked caller buffers: scatter to original order as fp64 on the device, DMA straight out
    CK(h, x->ensure_stage64(7 * h->cap));
    double* d64 = x->stage64;
    if (pos_xyz) { enqueue_export3_f64(h, h->pos[h->cur], d64); CK(h, cudaMemcpyAsync(pos_xyz, d64, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream)); }
    if (vel_xyz) { enqueue_export3_f64(h, h->vel[h->cur], d64 + 3 * n); CK(h, cudaMemcpyAsync(vel_xyz, d64 + 3 * n, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream)); }
    if (density) { enqueue_export1_f64(h, h->rho, d64 + 6 * n); CK(h, cudaMemcpyAsync(density, d64 + 6 * n, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream)); }
    return sync_and_check(h);
  }
  CK(h, x->pin.ensure(7 * n));
  float* st = x->pin.p;
  float* d = h->io_stage;
  if (pos_xyz) { enqueue_export3(h, h->pos[h->cur], d); CK(h, cudaMemcpyAsync(st, d, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, h->stream)); }
  if (vel_xyz) { enqueue_export3(h, h->vel[h->cur], d + 3 * n); CK(h, cudaMemcpyAsync(st + 3 * n, d + 3 * n, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, h->stream)); }
  if (density) { enqueue_export1(h, h->rho, d + 6 * n); CK(h, cudaMemcpyAsync(st + 6 * n, d + 6 * n, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream)); }
  int rc = sync_and_check(h);
  if (rc != PBF_OK) return rc;
  parallel_for(3 * n, [&](size_t a, size_t b) {
    if (pos_xyz) for (size_t i = a; i < b; i++) pos_xyz[i] = (double)st[i];
    if (vel_xyz) for (size_t i = a; i < b; i++) vel_xyz[i] = (double)st[3 * n + i];
  });
  if (density) parallel_for(n, [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) density[i] = (double)st[6 * n + i]; });
  return PBF_OK;
}

extern "C" {

// What every (re-)upload of a single-GPU handle resets before new particles arrive: capacity, copies still in flight on
// the read-back stream, read-back targets sized for another n, captured graphs, and every "derived from the old
// positions" marker (neighbour lists, the re-binned cell grid of pbf_density_at / pbf_extract_surface).
static int begin_upload(pbf_handle* h, size_t n, const char* who) {
  if (n > 0xFFFFFFF0ull) return fail(h, PBF_ERR_INVALID, "upload: more than 2^32 particles");
  if (h->slab) return fail(h, PBF_ERR_INVALID, "slab mode: use pbf_slab_upload");
  (void)who;
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaStreamSynchronize(h->stream));
  if (h->copy_stream) CK(h, cudaStreamSynchronize(h->copy_stream));
  h->rb_pending = false;
  int rc = ensure_capacity(h, n);
  if (rc != PBF_OK) return rc;
  if (n != h->n) h->rb_pos = h->rb_vel = h->rb_rho = nullptr;          // read-back targets were sized for the old n
  h->graph_invalidate();                              // grids and buffers are baked into the captured step
  h->n = n; h->cur = 0; h->have_neighbors = false; h->rebinned_at = -1;
  h->r_i0 = 0; h->r_cnt = (uint32_t)n; h->n_sorted = (uint32_t)n;
  return PBF_OK;
}

int pbf_upload(pbf_handle* h, size_t n, const double* pos_xyz, const double* vel_xyz) {
  if (!h || (n && (!pos_xyz || !vel_xyz))) return fail(h, PBF_ERR_INVALID, "pbf_upload: null argument");
  int rc = begin_upload(h, n, "pbf_upload");
  if (rc != PBF_OK) return rc;
  return io_upload(h, n, pos_xyz, vel_xyz);
}

int pbf_upload_device(pbf_handle* h, size_t n, const float* d_pos_xyz, const float* d_vel_xyz) {
  if (!h || (n && (!d_pos_xyz || !d_vel_xyz))) return fail(h, PBF_ERR_INVALID, "pbf_upload_device: null argument");
  int rc = begin_upload(h, n, "pbf_upload_device");
  if (rc != PBF_OK) return rc;
  enqueue_import(h, d_pos_xyz, d_vel_xyz);
  CK(h, cudaStreamSynchronize(h->stream));
  CK(h, cudaGetLastError());
  return PBF_OK;
}

// CUDA-graph replay of the step (see Solver::graph_exec): single GPU, no per-kernel profiling (its events would be
// captured).  A streaming read-back, if set, follows the last replayed step as plain launches.
static bool step_uses_graph(pbf_handle* h) {
  if (h->slab || h->profiling || h->n == 0) return false;
  if (h->graph_policy < 0) {
    const char* e = getenv("PBF_GRAPH");
    h->graph_policy = e ? (atoi(e) != 0) : (h->n <= (1u << 18));
  }
  return h->graph_policy == 1;
}

int pbf_step(pbf_handle* h, int n_steps) {
  if (!h || n_steps < 0) return fail(h, PBF_ERR_INVALID, "pbf_step: bad argument");
  if (h->slab) return fail(h, PBF_ERR_INVALID, "slab mode: drive the step through the pbf_slab_phase_* calls");
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaEventRecord(h->ev_call[0], h->stream));
  if (step_uses_graph(h)) {
    for (int s = 0; s < n_steps; s++) {
      if (h->rb_pending) { cudaStreamWaitEvent(h->stream, h->ev_rb[3], 0); h->rb_pending = false; }   // copies still read xs_a / rho / vel
      const int par = h->cur;
      if (!h->graph_exec[par]) {                       // first step with this buffer parity: capture instead of launching
        cudaGraph_t g = nullptr;
        const uint64_t l0 = h->launches;
        CK(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        enqueue_step(h, false);
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        h->cur = par; h->steps_done--;                 // capturing ran nothing: undo the host-side bookkeeping
        h->graph_launches[par] = h->launches - l0; h->launches = l0;
        if (e == cudaSuccess) e = cudaGraphInstantiate(&h->graph_exec[par], g, 0);
        if (g) cudaGraphDestroy(g);
        if (e != cudaSuccess) { cudaGetLastError(); h->graph_exec[par] = nullptr; h->graph_policy = 0; enqueue_step(h, false); continue; }
      }
      CK(h, cudaGraphLaunch(h->graph_exec[par], h->stream));
      h->cur = par ^ 1; h->steps_done++; h->launches += h->graph_launches[par];
    }
    if (n_steps > 0) enqueue_readback_all(h);
  } else {
    for (int s = 0; s < n_steps; s++) enqueue_step(h, s == n_steps - 1);   // streaming read-back (if set) after the last step
  }
  CK(h, cudaEventRecord(h->ev_call[1], h->stream));
  h->call_timed = true;
  if (n_steps > 0 && h->n > 0) h->have_neighbors = true;
  CK(h, cudaGetLastError());
  return PBF_OK;
}

// Obstacle spheres of the collision scene (the reference keeps them as StaticScene::Sphere primitives in
// Particles::bvh).  Kernel parameters carry them, so captured graphs are rebuilt.
// bounding box of all obstacles (the collide fast path skips moves that stay clear of it)
static void update_obstacle_box(pbf_handle* h) {
  DevParams& d = h->dp;
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (int k = 0; k < d.n_sph; k++) {
    const float c[3] = {d.sph[k].x, d.sph[k].y, d.sph[k].z};
    for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], c[a] - d.sph[k].w); hi[a] = std::max(hi[a], c[a] + d.sph[k].w); }
  }
  if (d.n_tri > 0) for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], d.tlo[a]); hi[a] = std::max(hi[a], d.thi[a]); }
  const float margin = 1e-2f * d.h + d.tol_ray + 2.f * d.skin;
  for (int a = 0; a < 3; a++) { d.olo[a] = lo[a] - margin - 1e-5f * std::fabs(lo[a]); d.ohi[a] = hi[a] + margin + 1e-5f * std::fabs(hi[a]); }
}

int pbf_set_obstacle_spheres(pbf_handle* h, size_t count, const double* s) {
  if (!h || (count && !s)) return fail(h, PBF_ERR_INVALID, "pbf_set_obstacle_spheres: null argument");
  if (count > PBF_MAX_SPHERES) return fail(h, PBF_ERR_CAPACITY, "pbf_set_obstacle_spheres: more than PBF_MAX_SPHERES spheres");
  for (size_t k = 0; k < count; k++)
    if (!(s[4 * k + 3] > 0) || !std::isfinite(s[4 * k]) || !std::isfinite(s[4 * k + 1]) || !std::isfinite(s[4 * k + 2]) || !std::isfinite(s[4 * k + 3]))
      return fail(h, PBF_ERR_INVALID, "pbf_set_obstacle_spheres: radius must be positive and all values finite");
  h->graph_invalidate();
  h->dp.n_sph = (int)count;
  for (size_t k = 0; k < count; k++) {
    volatile float r = (float)s[4 * k + 3];
    volatile float r2 = r * r;                         // one fp32 rounding, like Oracle<float>::set_spheres
    h->dp.sph[k] = make_float4((float)s[4 * k], (float)s[4 * k + 1], (float)s[4 * k + 2], r);
    h->dp.sph_r2[k] = r2;
  }
  update_obstacle_box(h);
  return PBF_OK;
}

// Bounding-volume hierarchy over the obstacle triangles, built on the host once per pbf_set_obstacle_triangles and
// walked on the device by ex_mesh_hit (the reference keeps its primitives in BVHAccel, bvh.cpp:48-140; this is a
// new hierarchy, not that one: median split of the centroids along the widest axis, leaves of <= 4 triangles,
// children of an inner node adjacent).  Node = 2 float4: (lo.xyz, a), (hi.xyz, b) with the integers stored as bits.
namespace {
struct BvhBuild {
  const std::vector<float>& tb;            // per triangle: lo[3], hi[3], centroid[3]
  std::vector<uint32_t> order;             // leaf order -> original triangle index
  std::vector<float> nodes;                // 8 floats per node
  int max_depth = 0;
  explicit BvhBuild(const std::vector<float>& boxes) : tb(boxes) {}
  static void put_int(float* dst, int v) { std::memcpy(dst, &v, sizeof(int)); }
  void build(uint32_t node, uint32_t first, uint32_t count, int depth) {
    max_depth = std::max(max_depth, depth);
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f}, clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
    for (uint32_t k = first; k < first + count; k++) {
      const float* b = &tb[9 * order[k]];
      for (int a = 0; a < 3; a++) {
        lo[a] = std::min(lo[a], b[a]); hi[a] = std::max(hi[a], b[3 + a]);
        clo[a] = std::min(clo[a], b[6 + a]); chi[a] = std::max(chi[a], b[6 + a]);
      }
    }
    for (int a = 0; a < 3; a++) { nodes[8 * node + a] = lo[a]; nodes[8 * node + 4 + a] = hi[a]; }
    if (count <= 4) { put_int(&nodes[8 * node + 3], (int)first); put_int(&nodes[8 * node + 7], (int)count); return; }
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    const uint32_t half = count / 2;
    std::nth_element(order.begin() + first, order.begin() + first + half, order.begin() + first + count,
                     [&](uint32_t x, uint32_t y) { const float cx = tb[9 * x + 6 + axis], cy = tb[9 * y + 6 + axis]; return cx < cy || (cx == cy && x < y); });
    const uint32_t left = (uint32_t)(nodes.size() / 8);
    nodes.resize(nodes.size() + 16);
    put_int(&nodes[8 * node + 3], (int)left); put_int(&nodes[8 * node + 7], 0);
    build(left, first, half, depth + 1);
    build(left + 1, first + half, count - half, depth + 1);
  }
};
// fp32 bounding boxes and centroids of the triangles (9 floats each: lo, hi, centre) and the hierarchy over them
static void bvh_over_triangles(size_t count, const double* q, std::vector<float>& tb, BvhBuild*& out) {
  tb.resize(9 * count);
  for (size_t k = 0; k < count; k++) {
    float* b = &tb[9 * k];
    for (int a = 0; a < 3; a++) {
      const float v0 = (float)q[18 * k + a], v1 = (float)q[18 * k + 3 + a], v2 = (float)q[18 * k + 6 + a];
      b[a] = std::min(v0, std::min(v1, v2)); b[3 + a] = std::max(v0, std::max(v1, v2)); b[6 + a] = 0.5f * (b[a] + b[3 + a]);
    }
  }
  out = new BvhBuild(tb);
  out->order.resize(count);
  for (size_t k = 0; k < count; k++) out->order[k] = (uint32_t)k;
  out->nodes.reserve(8 * (count + 1));
  out->nodes.resize(8);
  out->build(0, 0, (uint32_t)count, 1);
}
}  // namespace

// Obstacle triangles.  Edges, orientation and |e1 x e2| are precomputed in fp32 with the same single roundings as
// Oracle<float>::set_triangles; the records are stored in the leaf order of the hierarchy above.
int pbf_set_obstacle_triangles(pbf_handle* h, size_t count, const double* q) {
  if (!h || (count && !q)) return fail(h, PBF_ERR_INVALID, "pbf_set_obstacle_triangles: null argument");
  if (count > PBF_MAX_TRIANGLES) return fail(h, PBF_ERR_CAPACITY, "pbf_set_obstacle_triangles: more than PBF_MAX_TRIANGLES triangles");
  for (size_t k = 0; k < 18 * count; k++) if (!std::isfinite(q[k])) return fail(h, PBF_ERR_INVALID, "pbf_set_obstacle_triangles: non-finite value");
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaStreamSynchronize(h->stream));            // kernels in flight still read the old list
  h->graph_invalidate();
  if (h->tri_dev) { cudaFree(h->tri_dev); h->tri_dev = nullptr; }
  if (h->bvh_dev) { cudaFree(h->bvh_dev); h->bvh_dev = nullptr; }
  if (h->tri_id_dev) { cudaFree(h->tri_id_dev); h->tri_id_dev = nullptr; }
  h->dp.n_tri = 0; h->dp.tri = nullptr; h->dp.bvh = nullptr; h->dp.tri_id = nullptr;
  update_obstacle_box(h);
  if (count == 0) return PBF_OK;
  std::vector<float> t(20 * count), tb;
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (size_t k = 0; k < count; k++) {
    volatile float v[18];
    for (int a = 0; a < 18; a++) v[a] = (float)q[18 * k + a];
    volatile float e1[3], e2[3], ng[3];
    for (int a = 0; a < 3; a++) { e1[a] = v[3 + a] - v[a]; e2[a] = v[6 + a] - v[a]; }
    { volatile float m1 = e1[1] * e2[2], m2 = e1[2] * e2[1]; ng[0] = m1 - m2; }
    { volatile float m1 = e1[2] * e2[0], m2 = e1[0] * e2[2]; ng[1] = m1 - m2; }
    { volatile float m1 = e1[0] * e2[1], m2 = e1[1] * e2[0]; ng[2] = m1 - m2; }
    volatile float ns[3];
    for (int a = 0; a < 3; a++) { volatile float s1 = v[9 + a] + v[12 + a]; ns[a] = s1 + v[15 + a]; }
    volatile float d0 = ng[0] * ns[0], d1 = ng[1] * ns[1], d2 = ng[2] * ns[2];
    volatile float dsum = d0 + d1; dsum = dsum + d2;
    volatile float q0 = ng[0] * ng[0], q1 = ng[1] * ng[1], q2 = ng[2] * ng[2];
    volatile float qs = q0 + q1; qs = qs + q2;
    float* o = &t[20 * k];
    for (int a = 0; a < 3; a++) { o[a] = v[a]; o[3 + a] = e1[a]; o[6 + a] = e2[a]; }
    for (int a = 0; a < 9; a++) o[9 + a] = v[9 + a];
    o[18] = dsum < 0.f ? -1.f : 1.f;                    // orientation of e1 x e2 against the vertex normals
    o[19] = std::sqrt((float)qs);                       // |e1 x e2|
    for (int p3 = 0; p3 < 3; p3++) for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], (float)v[3 * p3 + a]); hi[a] = std::max(hi[a], (float)v[3 * p3 + a]); }
  }
  BvhBuild* Bp = nullptr;
  bvh_over_triangles(count, q, tb, Bp);
  std::unique_ptr<BvhBuild> Bown(Bp);
  BvhBuild& B = *Bp;
  if (B.max_depth > PBF_BVH_MAX_DEPTH) return fail(h, PBF_ERR_CAPACITY, "pbf_set_obstacle_triangles: hierarchy deeper than the traversal stack");
  const float margin = 1e-2f * h->dp.h + h->dp.tol_ray + 2.f * h->dp.skin;   // every accepted hit lies within tol_ray of the segment; plus the inflated edges and the rounding of a segment end
  for (size_t nd = 0; nd < B.nodes.size() / 8; nd++)
    for (int a = 0; a < 3; a++) {
      float& l = B.nodes[8 * nd + a]; float& u = B.nodes[8 * nd + 4 + a];
      l = l - margin - 1e-5f * std::fabs(l); u = u + margin + 1e-5f * std::fabs(u);
    }
  std::vector<float> ts(20 * count);
  for (size_t k = 0; k < count; k++) std::memcpy(&ts[20 * k], &t[20 * (size_t)B.order[k]], 20 * sizeof(float));
  CK(h, cudaMalloc((void**)&h->tri_dev, ts.size() * sizeof(float)));
  CK(h, cudaMemcpy(h->tri_dev, ts.data(), ts.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(h, cudaMalloc((void**)&h->bvh_dev, B.nodes.size() * sizeof(float)));
  CK(h, cudaMemcpy(h->bvh_dev, B.nodes.data(), B.nodes.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(h, cudaMalloc((void**)&h->tri_id_dev, count * sizeof(uint32_t)));
  CK(h, cudaMemcpy(h->tri_id_dev, B.order.data(), count * sizeof(uint32_t), cudaMemcpyHostToDevice));
  for (int a = 0; a < 3; a++) { h->dp.tlo[a] = lo[a] - margin - 1e-5f * std::fabs(lo[a]); h->dp.thi[a] = hi[a] + margin + 1e-5f * std::fabs(hi[a]); }
  h->dp.tri = h->tri_dev; h->dp.bvh = h->bvh_dev; h->dp.tri_id = h->tri_id_dev; h->dp.n_tri = (int)count;
  h->bvh_nodes = B.nodes.size() / 8; h->bvh_depth = B.max_depth;
  update_obstacle_box(h);
  return PBF_OK;
}

// Pure host function (no device needed): the hierarchy pbf_set_obstacle_triangles would build over these triangles,
// WITHOUT the safety margin on the node boxes.  nodes_out: 8 floats per node = lo.xyz, a (int bits), hi.xyz, b (int bits);
// order_out: leaf order -> original triangle index.  Returns PBF_ERR_CAPACITY when cap_nodes is too small (n_nodes is set).
int pbf_debug_build_bvh(size_t count, const double* q, float* nodes_out, size_t cap_nodes, uint32_t* order_out, size_t* n_nodes, int* depth) {
  if (!q || !n_nodes || !depth || count == 0 || count > PBF_MAX_TRIANGLES) return PBF_ERR_INVALID;
  std::vector<float> tb;
  BvhBuild* Bp = nullptr;
  bvh_over_triangles(count, q, tb, Bp);
  std::unique_ptr<BvhBuild> Bown(Bp);
  BvhBuild& B = *Bp;
  *n_nodes = B.nodes.size() / 8; *depth = B.max_depth;
  if (*n_nodes > cap_nodes || !nodes_out || !order_out) return PBF_ERR_CAPACITY;
  std::memcpy(nodes_out, B.nodes.data(), B.nodes.size() * sizeof(float));
  std::memcpy(order_out, B.order.data(), count * sizeof(uint32_t));
  return PBF_OK;
}

int pbf_sync(pbf_handle* h) {
  if (!h) return PBF_ERR_INVALID;
  return sync_and_check(h);
}

int pbf_estimate_densities(pbf_handle* h) {
  if (!h) return PBF_ERR_INVALID;
  if (h->slab) return fail(h, PBF_ERR_INVALID, "not available in slab mode");
  CK(h, cudaSetDevice(h->device));
  enqueue_estimate_densities(h);
  if (h->n > 0) h->have_neighbors = true;
  return pbf_sync(h);
}

int pbf_stats(pbf_handle* h, double* first, double* final_, double* last_call_ms) {
  int rc = pbf_sync(h);
  if (rc != PBF_OK) return rc;
  Scalars s;
  CK(h, cudaMemcpy(&s, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost));
  const double n = h->n ? (double)h->n : 1.0;
  if (first) *first = s.rho_first / n;
  if (final_) *final_ = s.rho_final / n;
  if (last_call_ms) *last_call_ms = h->last_call_ms;
  return PBF_OK;
}

int pbf_download_device(pbf_handle* h, float* d_pos_xyz, float* d_vel_xyz, float* d_density) {
  if (!h) return PBF_ERR_INVALID;
  CK(h, cudaSetDevice(h->device));
  if (d_pos_xyz) enqueue_export3(h, h->pos[h->cur], d_pos_xyz);
  if (d_vel_xyz) enqueue_export3(h, h->vel[h->cur], d_vel_xyz);
  if (d_density) enqueue_export1(h, h->rho, d_density);
  return pbf_sync(h);
}

int pbf_download(pbf_handle* h, double* pos_xyz, double* vel_xyz, double* density) {
  if (!h) return PBF_ERR_INVALID;
  if (h->slab) return fail(h, PBF_ERR_INVALID, "slab mode: use pbf_slab_download");
  CK(h, cudaSetDevice(h->device));
  return io_download(h, pos_xyz, vel_xyz, density);
}

// Particles::estimateDensityAt (particles.cpp:446-453) for m query points at once
int pbf_density_at(pbf_handle* h, size_t m, const double* query_xyz, double* density_out) {
  if (!h || (m && (!query_xyz || !density_out))) return PBF_ERR_INVALID;
  if (h->slab) return fail(h, PBF_ERR_INVALID, "pbf_density_at is single-GPU only");
  if (m > 0x7FFFFFFFull) return fail(h, PBF_ERR_INVALID, "too many query points in one call");
  if (m == 0) return PBF_OK;
  CK(h, cudaSetDevice(h->device));
  if (h->n == 0) { for (size_t i = 0; i < m; i++) density_out[i] = 0.0; return PBF_OK; }
  if (h->rebinned_at != (long long)h->steps_done) {     // cells of the last step belong to the predicted positions
    enqueue_rebin(h);
    h->have_neighbors = false;
    h->rebinned_at = (long long)h->steps_done;
  }
  std::vector<float> q(4 * m);
  parallel_for(m, [&](size_t a, size_t b) {
    for (size_t i = a; i < b; i++) { q[4*i] = (float)query_xyz[3*i]; q[4*i+1] = (float)query_xyz[3*i+1]; q[4*i+2] = (float)query_xyz[3*i+2]; q[4*i+3] = 0.f; }
  });
  float4* dq = nullptr; float* dout = nullptr;
  CK(h, dmalloc(&dq, m)); CK(h, dmalloc(&dout, m));
  cudaError_t e = cudaMemcpyAsync(dq, q.data(), m * sizeof(float4), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    enqueue_density_at(h, (uint32_t)m, dq, dout);
    std::vector<float> o(m);
    e = cudaMemcpyAsync(o.data(), dout, m * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    int rc = sync_and_check(h);
    cudaFree(dq); cudaFree(dout);
    if (rc != PBF_OK) return rc;
    if (e != cudaSuccess) { h->last_error = cudaGetErrorString(e); return PBF_ERR_CUDA; }
    for (size_t i = 0; i < m; i++) density_out[i] = (double)o[i];
    return PBF_OK;
  }
  cudaFree(dq); cudaFree(dout);
  h->last_error = cudaGetErrorString(e);
  return PBF_ERR_CUDA;
}

// Streaming read-back: after this call every pbf_step() sends the results of its last step to these
// page-locked (pbf_host_register) buffers on a second stream as soon as each becomes final —
// positions when the solver iterations end, density after the vorticity/XSPH pass, velocity after
// confinement — so most of the device-to-host time hides behind the finalize kernels.  pbf_sync()
// completes the copies.  Pass NULLs to switch it off.  Single-GPU handles only.
int pbf_set_readback(pbf_handle* h, double* pos_xyz, double* vel_xyz, double* density) {
  if (!h) return PBF_ERR_INVALID;
  if (h->slab) return fail(h, PBF_ERR_INVALID, "pbf_set_readback is single-GPU only");
  int rc = pbf_sync(h);
  if (rc != PBF_OK) return rc;
  h->rb_pos = h->rb_vel = h->rb_rho = nullptr;
  if (!pos_xyz && !vel_xyz && !density) return PBF_OK;
  if (h->n == 0 || h->cap == 0) return fail(h, PBF_ERR_INVALID, "pbf_set_readback: upload particles first");
  HandleExtra* x = extra_of(h);
  const size_t n = h->n;
  if ((pos_xyz && !x->registered(pos_xyz, 3 * n * sizeof(double))) || (vel_xyz && !x->registered(vel_xyz, 3 * n * sizeof(double))) ||
      (density && !x->registered(density, n * sizeof(double))))
    return fail(h, PBF_ERR_INVALID, "pbf_set_readback: buffers must be page-locked with pbf_host_register first");
  CK(h, cudaSetDevice(h->device));
  CK(h, x->ensure_stage64(7 * h->cap));
  h->rb_stage = x->stage64;
  if (!h->copy_stream) {
    CK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 4; k++) CK(h, cudaEventCreateWithFlags(&h->ev_rb[k], cudaEventDisableTiming));
  }
  h->rb_pos = pos_xyz; h->rb_vel = vel_xyz; h->rb_rho = density;
  return PBF_OK;
}

// Page-lock caller-owned host buffers (cudaHostRegister) so that pbf_upload / pbf_download can DMA
// to and from them directly.  The caller must unregister (or destroy the handle) before freeing.
int pbf_host_register(pbf_handle* h, void* ptr, size_t bytes) {
  if (!h || !ptr || !bytes) return PBF_ERR_INVALID;
  CK(h, cudaSetDevice(h->device));
  HandleExtra* x = extra_of(h);
  if (x->registered(ptr, bytes)) return PBF_OK;
  CK(h, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  x->regs.push_back(HostRange{(char*)ptr, bytes});
  return PBF_OK;
}

int pbf_host_unregister(pbf_handle* h, void* ptr) {
  if (!h || !ptr) return PBF_ERR_INVALID;
  HandleExtra* x = extra_of(h);
  for (size_t k = 0; k < x->regs.size(); k++)
    if (x->regs[k].p == (char*)ptr) { cudaHostUnregister(ptr); x->regs.erase(x->regs.begin() + k); return PBF_OK; }
  return fail(h, PBF_ERR_INVALID, "pbf_host_unregister: pointer was not registered");
}

// Neighbour-count alert of the reference (particles.cpp:32,165-173): see include/pbf_b200.h
int pbf_set_neighbor_alert(pbf_handle* h, int threshold, size_t max_records) {
  if (!h || threshold < 0) return fail(h, PBF_ERR_INVALID, "pbf_set_neighbor_alert: bad argument");
  if (h->slab) return fail(h, PBF_ERR_INVALID, "pbf_set_neighbor_alert is single-GPU only");
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaStreamSynchronize(h->stream));
  h->graph_invalidate();                               // the alert kernel is part of the captured step
  if (h->alert_buf) { cudaFree(h->alert_buf); h->alert_buf = nullptr; }
  h->alert_thr = 0; h->alert_cap = 0;
  if (threshold == 0 || max_records == 0) return PBF_OK;
  CK(h, dmalloc(&h->alert_buf, 2 * max_records + 4096 / 4));        // records + the 4096-bucket id histogram
  unsigned int zero[3] = {0, 0, 0};
  CK(h, cudaMemcpy(&h->sc->alert_count, zero, sizeof(zero), cudaMemcpyHostToDevice));
  h->alert_thr = (uint32_t)threshold; h->alert_cap = max_records;
  return PBF_OK;
}

int pbf_get_neighbor_alerts(pbf_handle* h, size_t cap, uint32_t* ids, uint32_t* counts, double* xpred_xyz, double* vel_xyz, size_t* n_written,
                            size_t* n_total) {
  if (!h || !n_total || !n_written) return PBF_ERR_INVALID;
  *n_total = 0; *n_written = 0;
  if (h->alert_thr == 0 || !h->alert_buf) return PBF_OK;
  int rc = pbf_sync(h);
  if (rc != PBF_OK) return rc;
  Scalars s;
  CK(h, cudaMemcpy(&s, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost));
  *n_total = s.alert_count;
  const size_t m = std::min<size_t>(s.alert_kept, h->alert_cap);
  if (m == 0 || cap == 0) return PBF_OK;
  *n_written = std::min(m, cap);
  std::vector<float4> rec(2 * m);
  CK(h, cudaMemcpy(rec.data(), h->alert_buf, 2 * m * sizeof(float4), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> order(m);
  for (size_t k = 0; k < m; k++) order[k] = (uint32_t)k;
  auto id_of = [&](uint32_t k) { uint32_t v; std::memcpy(&v, &rec[2 * (size_t)k].w, 4); return v; };
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return id_of(a) < id_of(b); });   // the reference warns in index order
  for (size_t q = 0; q < std::min(m, cap); q++) {
    const float4 x = rec[2 * (size_t)order[q]], v = rec[2 * (size_t)order[q] + 1];
    if (ids) ids[q] = id_of(order[q]);
    if (counts) std::memcpy(&counts[q], &v.w, 4);
    if (xpred_xyz) { xpred_xyz[3 * q] = x.x; xpred_xyz[3 * q + 1] = x.y; xpred_xyz[3 * q + 2] = x.z; }
    if (vel_xyz) { vel_xyz[3 * q] = v.x; vel_xyz[3 * q + 1] = v.y; vel_xyz[3 * q + 2] = v.z; }
  }
  return PBF_OK;
}

// ---- parity / debug ---------------------------------------------------------------------------
int pbf_debug_capture(pbf_handle* h, int on) {
  if (!h) return PBF_ERR_INVALID;
  h->capture_xpred = on ? 1 : 0;
  h->graph_invalidate();
  if (on && !h->xpred && h->cap) CK(h, dmalloc(&h->xpred, h->cap));
  return PBF_OK;
}

int pbf_debug_neighbor_digest(pbf_handle* h, uint64_t* digest, uint32_t* count) {
  if (!h || !digest || !count) return PBF_ERR_INVALID;
  if (!h->have_neighbors) return fail(h, PBF_ERR_INVALID, "no neighbour lists yet: call pbf_step first");
  const size_t n = h->r_cnt;
  CK(h, cudaSetDevice(h->device));
  unsigned long long* dd = nullptr; uint32_t* dc = nullptr;
  CK(h, dmalloc(&dd, n)); CK(h, dmalloc(&dc, n));
  enqueue_digest(h, dd, dc);
  int rc = pbf_sync(h);
  if (rc == PBF_OK) {
    cudaMemcpy(digest, dd, n * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    cudaMemcpy(count, dc, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  }
  cudaFree(dd); cudaFree(dc);
  return rc;
}

int pbf_debug_download_neighbors(pbf_handle* h, uint32_t* row_ptr, uint32_t* col_idx, size_t col_cap) {
  if (!h || !row_ptr) return PBF_ERR_INVALID;
  if (!h->have_neighbors) return fail(h, PBF_ERR_INVALID, "no neighbour lists yet: call pbf_step first");
  int rc = pbf_sync(h);
  if (rc != PBF_OK) return rc;
  if (h->slab) return fail(h, PBF_ERR_INVALID, "CSR download is single-GPU only (use the digest in slab mode)");
  const size_t n = h->n;
  Scalars s;
  CK(h, cudaMemcpy(&s, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> cnt(n), off(n / 32 + 1), orig(n), nb((size_t)s.nbr_cursor * 128);
  CK(h, cudaMemcpy(cnt.data(), h->nbr_cnt, n * 4, cudaMemcpyDeviceToHost));
  CK(h, cudaMemcpy(off.data(), h->slice_off, (n / 32 + 1) * 4, cudaMemcpyDeviceToHost));
  CK(h, cudaMemcpy(orig.data(), h->orig[h->cur], n * 4, cudaMemcpyDeviceToHost));
  if (!nb.empty()) CK(h, cudaMemcpy(nb.data(), h->nbr, nb.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<uint32_t> count_orig(n);
  for (size_t i = 0; i < n; i++) count_orig[orig[i]] = cnt[i];
  row_ptr[0] = 0;
  for (size_t i = 0; i < n; i++) row_ptr[i + 1] = row_ptr[i] + count_orig[i];
  if (!col_idx) return PBF_OK;
  if (row_ptr[n] > col_cap) return fail(h, PBF_ERR_CAPACITY, "col_idx too small");
  for (size_t i = 0; i < n; i++) {
    uint32_t* dst = col_idx + row_ptr[orig[i]];
    const size_t base = (size_t)off[i / 32] * 128 + (i % 32) * 4;
    for (uint32_t k = 0; k < cnt[i]; k++) dst[k] = orig[nb[base + (size_t)(k >> 2) * 128 + (k & 3)]];
    std::sort(dst, dst + cnt[i]);
  }
  return PBF_OK;
}

int pbf_debug_download_array(pbf_handle* h, int which, double* out) {
  if (!h || !out) return PBF_ERR_INVALID;
  const size_t n = h->n;
  if (n == 0) return PBF_OK;
  CK(h, cudaSetDevice(h->device));
  float* d = h->io_stage;
  size_t m = 3 * n;
  switch (which) {
    case PBF_ARRAY_XSTAR: enqueue_export3(h, h->xs_a, d); break;
    case PBF_ARRAY_LAMBDA: enqueue_export_w(h, h->xs_b, d); m = n; break;
    case PBF_ARRAY_VORTICITY: enqueue_export3(h, h->omega, d); break;
    case PBF_ARRAY_XPRED:
      if (!h->capture_xpred || !h->xpred) return fail(h, PBF_ERR_INVALID, "call pbf_debug_capture(h,1) before the step");
      enqueue_export3(h, h->xpred, d); break;
    default: return fail(h, PBF_ERR_INVALID, "unknown array id");
  }
  int rc = pbf_sync(h);
  if (rc != PBF_OK) return rc;
  std::vector<float> tmp(m);
  CK(h, cudaMemcpy(tmp.data(), d, m * sizeof(float), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < m; i++) out[i] = (double)tmp[i];
  return PBF_OK;
}

int pbf_profile_enable(pbf_handle* h, int on) {
  if (!h) return PBF_ERR_INVALID;
  int rc = pbf_sync(h);
  h->profiling = on != 0;
  for (int k = 0; k < K_COUNT; k++) { h->prof_ms[k] = 0; h->prof_launches[k] = 0; }
  return rc;
}

int pbf_profile_get(pbf_handle* h, int max, const char** names, double* total_ms, uint64_t* launches) {
  if (!h) return 0;
  pbf_sync(h);
  int k = 0;
  for (; k < K_COUNT && k < max; k++) {
    if (names) names[k] = kKernelNames[k];
    if (total_ms) total_ms[k] = h->prof_ms[k];
    if (launches) launches[k] = h->prof_launches[k];
  }
  return k;
}

}  // extern "C"
