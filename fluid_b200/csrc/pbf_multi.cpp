// pbf_multi.cpp — several GPUs behind ONE handle, driven by one host thread (include/pbf_b200_multi.h).
//
// Pure host code on top of the slab entry points of the C ABI (include/pbf_b200_slab.h): slab planning, particle
// distribution, peer-mode wiring, the per-step enqueue loop and the re-balancing policy.  The reference has no
// counterpart (it is one thread on one CPU, particles.cpp:250-297); SURVEY.md §8(b) / §8(e) specify this layer.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pbf_b200_multi.h"

// ---- planning (no device needed; unit-tested on the CPU) ----------------------------------------------------------------

extern "C" int pbf_partition_columns(const uint64_t* hist, int n_cols, int world, int* bounds_out) {
  if (!hist || !bounds_out || world < 1 || n_cols < world) return PBF_ERR_INVALID;
  std::vector<double> csum((size_t)n_cols + 1, 0.0);
  for (int c = 0; c < n_cols; c++) csum[c + 1] = csum[c] + (double)hist[c];
  const double total = csum[n_cols];
  bounds_out[0] = 0;
  for (int r = 1; r < world; r++) {
    const double target = total * r / world;
    int b = (int)(std::lower_bound(csum.begin(), csum.end(), target) - csum.begin());      // first boundary with count >= target
    if (b > n_cols) b = n_cols;
    if (b > 0 && std::fabs(csum[b - 1] - target) <= std::fabs(csum[b] - target)) b--;      // the nearer of the two
    b = std::max(b, bounds_out[r - 1] + 1);                                                 // every slab owns >= 1 column
    b = std::min(b, n_cols - (world - r));
    bounds_out[r] = b;
  }
  bounds_out[world] = n_cols;
  return PBF_OK;
}

// New column boundaries for a flowing fluid.  Boundary k may only move strictly inside (old[k-1], old[k+1]): a particle
// held by slab r then still belongs to r or to an ADJACENT slab, which is all one migration hop can do; and the particles
// of the columns that change hands must fit the migration message next to the ordinary emigrants (max_move).
extern "C" int pbf_plan_rebalance(const uint64_t* hist, int n_cols, int world, const int* old_bounds, uint64_t max_move, double threshold,
                                  int* new_bounds_out, double* imbalance_out) {
  if (!hist || !old_bounds || !new_bounds_out || world < 1 || n_cols < world) return -1;
  std::vector<double> owned(world, 0.0);
  double total = 0.0, mx = 0.0;
  for (int r = 0; r < world; r++) {
    if (old_bounds[r + 1] <= old_bounds[r] || old_bounds[r] < 0 || old_bounds[r + 1] > n_cols) return -1;
    for (int c = old_bounds[r]; c < old_bounds[r + 1]; c++) owned[r] += (double)hist[c];
    total += owned[r]; mx = std::max(mx, owned[r]);
  }
  const double imb = total > 0 ? mx / (total / world) : 1.0;
  if (imbalance_out) *imbalance_out = imb;
  for (int k = 0; k <= world; k++) new_bounds_out[k] = old_bounds[k];
  if (world == 1 || !(imb > threshold)) return 0;
  std::vector<int> ideal(world + 1);
  if (pbf_partition_columns(hist, n_cols, world, ideal.data()) != PBF_OK) return -1;
  int changed = 0;
  for (int k = 1; k < world; k++) {
    const int lo_lim = std::max(old_bounds[k - 1] + 1, new_bounds_out[k - 1] + 1), hi_lim = old_bounds[k + 1] - 1;
    const int target = std::min(std::max(ideal[k], lo_lim), hi_lim);
    int b = old_bounds[k];
    uint64_t moved = 0;
    while (b != target) {
      const int c = target > b ? b : b - 1;              // the column that changes hands next
      if (moved + hist[c] > max_move) break;
      moved += hist[c];
      b += target > b ? 1 : -1;
    }
    if (b < lo_lim) b = lo_lim;                           // only when the previous boundary moved right past it
    new_bounds_out[k] = b;
    if (b != old_bounds[k]) changed = 1;
  }
  return changed;
}

// ---- the handle -----------------------------------------------------------------------------------------------------------

struct pbf_multi {
  PbfParams params;
  std::vector<int> devices;
  std::vector<pbf_handle*> h;
  std::vector<int> col_bounds;          // world + 1
  int gdims[3] = {0, 0, 0};
  size_t n_total = 0, halo_cap = 0, particle_cap = 0;
  bool planned = false;
  std::string last_error;
  std::vector<double> spheres, tris;
  int rebalance_every = 8; double rebalance_threshold = 1.05;
  long long steps_done = 0, last_rebalance_at = 0; uint64_t n_rebalances = 0;
  std::vector<uint32_t> hist_tmp; std::vector<uint64_t> hist;
  double last_imbalance = 1.0;
  bool trace = getenv("PBF_MULTI_TRACE") != nullptr;
  // read-back staging, one set per slab, allocated and page-locked at the first pbf_multi_download and kept: the slabs'
  // copies then run as direct DMA, all devices at once, and every slab's scatter to original order runs on its own thread
  struct Stage { std::vector<double> p, v, r; std::vector<uint32_t> id; };
  std::vector<Stage> stage;
};

namespace {
int mfail(pbf_multi* m, int code, const std::string& msg) { if (m) m->last_error = msg; return code; }
int from_handle(pbf_multi* m, int d, int rc) {
  if (rc != PBF_OK) m->last_error = "device " + std::to_string(m->devices[d]) + ": " + pbf_last_error(m->h[d]);
  return rc;
}
void destroy_handles(pbf_multi* m) {
  for (pbf_handle* q : m->h) if (q) pbf_sync(q);       // nobody may still be writing into a neighbour that is about to go
  for (pbf_handle* q : m->h) if (q) pbf_destroy(q);
  m->h.clear(); m->planned = false;
  m->stage.clear();                                    // registered with the handles that just went (pbf_destroy unregisters)
}
int create_handles(pbf_multi* m) {
  m->h.assign(m->devices.size(), nullptr);
  for (size_t d = 0; d < m->devices.size(); d++) {
    int rc = pbf_create(&m->params, m->devices[d], &m->h[d]);
    if (rc != PBF_OK) { destroy_handles(m); return mfail(m, rc, "pbf_create failed on device " + std::to_string(m->devices[d])); }
  }
  return PBF_OK;
}
int apply_scene(pbf_multi* m) {
  for (size_t d = 0; d < m->h.size(); d++) {
    int rc = pbf_set_obstacle_spheres(m->h[d], m->spheres.size() / 4, m->spheres.data());
    if (rc == PBF_OK) rc = pbf_set_obstacle_triangles(m->h[d], m->tris.size() / 18, m->tris.data());
    if (rc != PBF_OK) return from_handle(m, (int)d, rc);
  }
  return PBF_OK;
}

// Move the boundaries if the slabs have drifted apart (called between two steps).  The histograms are those the sorts of
// step (steps_done - every) left behind: this thread waits for them, i.e. it runs at most `every` steps ahead of the devices,
// which still have the steps enqueued since then to work on.  The schedule is a function of the step count only.
void maybe_rebalance(pbf_multi* m) {
  const int world = (int)m->h.size(), ncol = m->gdims[0];
  if (world < 2 || m->rebalance_every <= 0 || m->steps_done == 0 || m->steps_done % m->rebalance_every != 0) return;
  m->hist.assign(ncol, 0); m->hist_tmp.resize(ncol);
  for (int d = 0; d < world; d++) {
    long long at = -1;
    if (pbf_slab_column_histogram(m->h[d], m->hist_tmp.data(), (size_t)ncol, 1, &at) != PBF_OK) return;
    if (at != m->steps_done - m->rebalance_every || at < m->last_rebalance_at) return;   // not the one expected (interval just changed) or taken under an older plan
    for (int c = m->col_bounds[d]; c < m->col_bounds[d + 1]; c++) m->hist[c] = m->hist_tmp[c];
  }
  std::vector<int> nb(world + 1);
  double imb = 1.0;
  const int changed = pbf_plan_rebalance(m->hist.data(), ncol, world, m->col_bounds.data(), (uint64_t)(0.4 * (double)m->halo_cap),
                                         m->rebalance_threshold, nb.data(), &imb);
  if (m->trace) {
    fprintf(stderr, "[pbf_multi] step %lld imbalance %.4f changed %d bounds", m->steps_done, imb, changed);
    for (int k = 0; k <= world; k++) fprintf(stderr, " %d->%d", m->col_bounds[k], nb[k]);
    fprintf(stderr, "\n");
  }
  m->last_imbalance = imb;
  if (changed != 1) return;
  for (int d = 0; d < world; d++) {
    const int left = d > 0 ? nb[d] - nb[d - 1] : 0, right = d + 1 < world ? nb[d + 2] - nb[d + 1] : 0;
    if (pbf_slab_set_columns(m->h[d], nb[d], nb[d + 1], left, right) != PBF_OK) return;   // wider than max_cols: keep the old plan everywhere
  }
  // (a failure above can only hit the first device that would exceed max_cols; earlier devices were already changed, so
  //  re-apply the old plan to keep every rank consistent)
  bool ok = true;
  for (int d = 0; d < world; d++) { int c[4]; if (pbf_slab_columns(m->h[d], c) != PBF_OK || c[0] != nb[d] || c[1] != nb[d + 1]) ok = false; }
  if (!ok) {
    for (int d = 0; d < world; d++) {
      const int left = d > 0 ? m->col_bounds[d] - m->col_bounds[d - 1] : 0, right = d + 1 < world ? m->col_bounds[d + 2] - m->col_bounds[d + 1] : 0;
      pbf_slab_set_columns(m->h[d], m->col_bounds[d], m->col_bounds[d + 1], left, right);
    }
    return;
  }
  m->col_bounds = nb; m->n_rebalances++; m->last_rebalance_at = m->steps_done;
}
}  // namespace

extern "C" {

int pbf_create_multi(const PbfParams* params, int n_devices, const int* device_ids, pbf_multi** out) {
  if (!params || !out || n_devices < 1) return PBF_ERR_INVALID;
  *out = nullptr;
  const int ndev = pbf_device_count();
  if (ndev <= 0) return PBF_ERR_NO_DEVICE;
  pbf_multi* m = new pbf_multi();
  m->params = *params;
  for (int d = 0; d < n_devices; d++) {
    const int id = device_ids ? device_ids[d] : d;
    if (id < 0 || id >= ndev) { delete m; return PBF_ERR_INVALID; }      // an id may repeat: several slabs on one GPU (testing on a 1-GPU box)
    m->devices.push_back(id);
  }
  if (pbf_grid_dims(params, m->gdims) != PBF_OK || m->gdims[0] < n_devices) { delete m; return PBF_ERR_INVALID; }   // every slab needs a column
  int rc = create_handles(m);
  if (rc != PBF_OK) { fprintf(stderr, "pbf_create_multi: %s\n", m->last_error.c_str()); delete m; return rc; }
  *out = m;
  return PBF_OK;
}

void pbf_multi_destroy(pbf_multi* m) {
  if (!m) return;
  destroy_handles(m);
  delete m;
}

const char* pbf_multi_last_error(pbf_multi* m) { return m ? m->last_error.c_str() : "null handle"; }
int pbf_multi_num_devices(pbf_multi* m) { return m ? (int)m->devices.size() : 0; }
size_t pbf_multi_num_particles(pbf_multi* m) { return m ? m->n_total : 0; }

int pbf_multi_set_obstacle_spheres(pbf_multi* m, size_t count, const double* s) {
  if (!m || (count && !s)) return PBF_ERR_INVALID;
  m->spheres.assign(s, s + 4 * count);
  return apply_scene(m);
}
int pbf_multi_set_obstacle_triangles(pbf_multi* m, size_t count, const double* t) {
  if (!m || (count && !t)) return PBF_ERR_INVALID;
  m->tris.assign(t, t + 18 * count);
  return apply_scene(m);
}

int pbf_multi_set_rebalance(pbf_multi* m, int every_k_steps, double threshold) {
  if (!m || every_k_steps < 0 || !(threshold >= 1.0)) return PBF_ERR_INVALID;
  m->rebalance_every = every_k_steps; m->rebalance_threshold = threshold;
  if (m->planned) for (pbf_handle* q : m->h) pbf_slab_set_histogram_interval(q, every_k_steps);
  return PBF_OK;
}

int pbf_multi_upload(pbf_multi* m, size_t n, const double* pos_xyz, const double* vel_xyz) {
  if (!m || (n && (!pos_xyz || !vel_xyz))) return mfail(m, PBF_ERR_INVALID, "pbf_multi_upload: null argument");
  if (n > 0xFFFFFFF0ull) return mfail(m, PBF_ERR_INVALID, "pbf_multi_upload: more than 2^32 particles");
  const int world = (int)m->devices.size(), ncol = m->gdims[0];
  if (m->planned) {                                       // slab handles are configured once: start from fresh ones
    destroy_handles(m);
    int rc = create_handles(m);
    if (rc != PBF_OK) return rc;
  }
  // plan: equal particle counts from the histogram of cell columns (same fp32 arithmetic as the device's binning)
  std::vector<int32_t> col(n);
  int rc = pbf_cell_columns(&m->params, n, pos_xyz, col.data());
  if (rc != PBF_OK) return mfail(m, rc, "pbf_cell_columns failed");
  std::vector<uint64_t> hist(ncol, 0);
  for (size_t i = 0; i < n; i++) hist[col[i]]++;
  m->col_bounds.assign(world + 1, 0);
  if (pbf_partition_columns(hist.data(), ncol, world, m->col_bounds.data()) != PBF_OK) return mfail(m, PBF_ERR_INVALID, "cannot partition the cell columns");
  std::vector<int> slab_of(ncol);
  for (int r = 0; r < world; r++) for (int c = m->col_bounds[r]; c < m->col_bounds[r + 1]; c++) slab_of[c] = r;
  std::vector<size_t> owned(world, 0);
  for (size_t i = 0; i < n; i++) owned[slab_of[col[i]]]++;
  const uint64_t per_col = std::max<uint64_t>(*std::max_element(hist.begin(), hist.end()), 1);
  m->halo_cap = (size_t)std::max<uint64_t>(4096, 4 * per_col);
  const size_t most = *std::max_element(owned.begin(), owned.end());
  // owned + two ghost columns, with 30 % of slack: re-balancing keeps the slabs within a few per cent of each other
  m->particle_cap = (size_t)(1.30 * (double)std::max<size_t>(most, (n + world - 1) / world)) + 2 * (size_t)per_col + 4096;
  // cell arrays: room for the slab to widen as the fluid spreads (all columns when that is cheap)
  const double cells_per_col = (double)m->gdims[1] * m->gdims[2];
  for (int r = 0; r < world; r++) {
    const int width = m->col_bounds[r + 1] - m->col_bounds[r];
    int max_cols = ncol;
    if (cells_per_col * ncol * 8.0 > 2e9) max_cols = std::min(ncol, 4 * width + 16);
    const int left = r > 0 ? m->col_bounds[r] - m->col_bounds[r - 1] : 0, right = r + 1 < world ? m->col_bounds[r + 2] - m->col_bounds[r + 1] : 0;
    rc = pbf_slab_configure_ex(m->h[r], m->col_bounds[r], m->col_bounds[r + 1], left, right, m->particle_cap, m->halo_cap, max_cols);
    if (rc != PBF_OK) return from_handle(m, r, rc);
  }
  for (int r = 0; r < world; r++) pbf_slab_set_histogram_interval(m->h[r], m->rebalance_every);
  for (int r = 0; r < world; r++) {
    rc = pbf_slab_p2p_connect_local(m->h[r], r > 0 ? m->h[r - 1] : nullptr, r + 1 < world ? m->h[r + 1] : nullptr);
    if (rc != PBF_OK) return from_handle(m, r, rc);
  }
  m->planned = true;
  if ((rc = apply_scene(m)) != PBF_OK) return rc;
  // distribute: one pass per device over the column array, staging sized for the largest slab
  std::vector<double> sp(3 * most), sv(3 * most); std::vector<uint32_t> sid(most);
  for (int r = 0; r < world; r++) {
    size_t k = 0;
    for (size_t i = 0; i < n; i++) {
      if (slab_of[col[i]] != r) continue;
      sp[3*k] = pos_xyz[3*i]; sp[3*k+1] = pos_xyz[3*i+1]; sp[3*k+2] = pos_xyz[3*i+2];
      sv[3*k] = vel_xyz[3*i]; sv[3*k+1] = vel_xyz[3*i+1]; sv[3*k+2] = vel_xyz[3*i+2];
      sid[k++] = (uint32_t)i;
    }
    rc = pbf_slab_upload(m->h[r], k, sp.data(), sv.data(), sid.data());
    if (rc != PBF_OK) return from_handle(m, r, rc);
  }
  m->n_total = n; m->steps_done = 0; m->last_rebalance_at = 0;
  return PBF_OK;
}

int pbf_multi_step(pbf_multi* m, int n_steps) {
  if (!m || n_steps < 0) return PBF_ERR_INVALID;
  if (!m->planned) return mfail(m, PBF_ERR_INVALID, "pbf_multi_step: upload particles first");
  // steps outermost: every device's queue advances in lock-step, and a waiting device always finds the signal of its
  // neighbour already enqueued or about to be (this thread never blocks on a device inside the loop)
  for (int s = 0; s < n_steps; s++) {
    maybe_rebalance(m);
    for (size_t d = 0; d < m->h.size(); d++) {
      int rc = pbf_slab_step_p2p(m->h[d], 1);
      if (rc != PBF_OK) return from_handle(m, (int)d, rc);
    }
    m->steps_done++;
  }
  return PBF_OK;
}

int pbf_multi_estimate_densities(pbf_multi* m) {
  if (!m) return PBF_ERR_INVALID;
  if (!m->planned) return mfail(m, PBF_ERR_INVALID, "pbf_multi_estimate_densities: upload particles first");
  for (size_t d = 0; d < m->h.size(); d++) {
    int rc = pbf_slab_estimate_densities_p2p(m->h[d]);
    if (rc != PBF_OK) return from_handle(m, (int)d, rc);
  }
  return PBF_OK;
}

int pbf_multi_sync(pbf_multi* m) {
  if (!m) return PBF_ERR_INVALID;
  int first = PBF_OK; std::string all;
  for (size_t d = 0; d < m->h.size(); d++) {             // wait for ALL devices even after an error, and report every one of them
    int rc = pbf_sync(m->h[d]);
    if (rc == PBF_OK) continue;
    if (first == PBF_OK) first = rc;
    all += (all.empty() ? "" : "; ") + std::string("slab ") + std::to_string(d) + " (device " + std::to_string(m->devices[d]) + "): " + pbf_last_error(m->h[d]);
  }
  if (first != PBF_OK) m->last_error = all;
  return first;
}

int pbf_multi_stats(pbf_multi* m, double* first, double* final_, double* last_call_ms) {
  if (!m) return PBF_ERR_INVALID;
  int rc = pbf_multi_sync(m);
  if (rc != PBF_OK) return rc;
  double a = 0, b = 0, ms = 0; uint64_t n = 0;
  for (size_t d = 0; d < m->h.size(); d++) {
    double ad = 0, bd = 0, msd = 0; uint64_t nd = 0;
    rc = pbf_slab_stats(m->h[d], &ad, &bd, &nd);
    if (rc == PBF_OK) rc = pbf_stats(m->h[d], nullptr, nullptr, &msd);
    if (rc != PBF_OK) return from_handle(m, (int)d, rc);
    a += ad; b += bd; n += nd; ms = std::max(ms, msd);
  }
  const double nn = n ? (double)n : 1.0;
  if (first) *first = a / nn;
  if (final_) *final_ = b / nn;
  if (last_call_ms) *last_call_ms = ms;
  return PBF_OK;
}

int pbf_multi_download(pbf_multi* m, double* pos_xyz, double* vel_xyz, double* density) {
  if (!m) return PBF_ERR_INVALID;
  if (!m->planned) return m->n_total == 0 ? PBF_OK : mfail(m, PBF_ERR_INVALID, "nothing uploaded");
  int rc = pbf_multi_sync(m);
  if (rc != PBF_OK) return rc;
  const size_t cap = m->particle_cap, nd = m->h.size();
  if (m->stage.size() != nd) m->stage.assign(nd, pbf_multi::Stage());
  for (size_t d = 0; d < nd; d++) {                    // staging of a slab: sized once, page-locked once (a refused registration only costs speed)
    pbf_multi::Stage& st = m->stage[d];
    auto grow = [&](auto& v, size_t count) {           // (re)allocate and page-lock; an old, smaller block is unlocked before it is freed
      if (v.size() >= count) return;
      if (!v.empty()) pbf_host_unregister(m->h[d], v.data());
      v.assign(count, 0);
      pbf_host_register(m->h[d], v.data(), count * sizeof(v[0]));
    };
    if (pos_xyz) grow(st.p, 3 * cap);
    if (vel_xyz) grow(st.v, 3 * cap);
    if (density) grow(st.r, cap);
    grow(st.id, cap);
  }
  std::vector<int> rcs(nd, PBF_OK); std::vector<size_t> got(nd, 0); std::vector<char> bad_id(nd, 0);
  auto one = [&](size_t d) {
    pbf_multi::Stage& st = m->stage[d];
    size_t k = 0;
    rcs[d] = pbf_slab_download(m->h[d], cap, pos_xyz ? st.p.data() : nullptr, vel_xyz ? st.v.data() : nullptr, density ? st.r.data() : nullptr, st.id.data(), &k);
    if (rcs[d] != PBF_OK) return;
    got[d] = k;
    const size_t n_total = m->n_total;
    for (size_t q = 0; q < k; q++) {                   // global ids are disjoint between slabs: the threads never write the same element
      const size_t i = st.id[q];
      if (i >= n_total) { bad_id[d] = 1; return; }
      if (pos_xyz) { pos_xyz[3*i] = st.p[3*q]; pos_xyz[3*i+1] = st.p[3*q+1]; pos_xyz[3*i+2] = st.p[3*q+2]; }
      if (vel_xyz) { vel_xyz[3*i] = st.v[3*q]; vel_xyz[3*i+1] = st.v[3*q+1]; vel_xyz[3*i+2] = st.v[3*q+2]; }
      if (density) density[i] = st.r[q];
    }
  };
  if (nd == 1) one(0);
  else {
    std::vector<std::thread> th;
    for (size_t d = 0; d < nd; d++) th.emplace_back(one, d);
    for (std::thread& t : th) t.join();
  }
  size_t seen = 0;
  for (size_t d = 0; d < nd; d++) {
    if (rcs[d] != PBF_OK) return from_handle(m, (int)d, rcs[d]);
    if (bad_id[d]) return mfail(m, PBF_ERR_CUDA, "corrupt particle id in a slab");
    seen += got[d];
  }
  if (seen != m->n_total) return mfail(m, PBF_ERR_CUDA, "particles lost or duplicated between slabs: " + std::to_string(seen) + " of " + std::to_string(m->n_total));
  return PBF_OK;
}

int pbf_multi_neighbor_digest(pbf_multi* m, uint64_t* digest, uint32_t* count) {
  if (!m || !digest || !count || !m->planned) return PBF_ERR_INVALID;
  int rc = pbf_multi_sync(m);
  if (rc != PBF_OK) return rc;
  const size_t cap = m->particle_cap;
  std::vector<uint64_t> sd(cap); std::vector<uint32_t> sc(cap), sid(cap);
  for (size_t d = 0; d < m->h.size(); d++) {
    size_t k = 0;
    rc = pbf_slab_download(m->h[d], cap, nullptr, nullptr, nullptr, sid.data(), &k);
    if (rc == PBF_OK) rc = pbf_slab_neighbor_digest(m->h[d], cap, sd.data(), sc.data());
    if (rc != PBF_OK) return from_handle(m, (int)d, rc);
    for (size_t q = 0; q < k; q++) { digest[sid[q]] = sd[q]; count[sid[q]] = sc[q]; }
  }
  return PBF_OK;
}

int pbf_multi_plan(pbf_multi* m, int* col_bounds_out, uint64_t* owned_out, uint64_t* n_rebalances_out) {
  if (!m || !m->planned) return PBF_ERR_INVALID;
  if (col_bounds_out) for (size_t k = 0; k < m->col_bounds.size(); k++) col_bounds_out[k] = m->col_bounds[k];
  if (n_rebalances_out) *n_rebalances_out = m->n_rebalances;
  if (owned_out) {
    for (size_t d = 0; d < m->h.size(); d++) {
      uint64_t nd = 0;
      int rc = pbf_slab_stats(m->h[d], nullptr, nullptr, &nd);
      if (rc != PBF_OK) return from_handle(m, (int)d, rc);
      owned_out[d] = nd;
    }
  }
  return PBF_OK;
}

uint64_t pbf_multi_launch_count(pbf_multi* m) {
  uint64_t s = 0;
  if (m) for (pbf_handle* q : m->h) s += pbf_launch_count(q);
  return s;
}

}  // extern "C"
