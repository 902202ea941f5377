// TEST INFRASTRUCTURE ONLY — C entry points of the CPU oracle (see pbf_oracle.hpp) for ctypes.
// Built by oracle/Makefile into oracle/liboracle.so.  Never linked into the product library.
#include "pbf_oracle.hpp"

#include <type_traits>

#include <chrono>
#include <omp.h>

using namespace pbf_oracle;

namespace {
struct Handle {
  int precision;  // 32 or 64
  Oracle<float>* f = nullptr;
  Oracle<double>* d = nullptr;
  double last_ms = 0;
};
template <class F> auto visit(Handle* h, F&& fn) { return h->precision == 32 ? fn(*h->f) : fn(*h->d); }
}  // namespace

extern "C" {

void oracle_default_params(PbfParams* p) {
  std::memset(p, 0, sizeof(*p));
  p->h = 0.3; p->dt = 0.016; p->rest_density = 1000.0; p->eps_relax = 2.0; p->k_corr = 0.0001;
  p->dq_ratio = 0.1; p->visc_c = 0.001; p->vort_eps = 0.001; p->gravity_y = -10.0;
  p->n_corr = 4; p->iterations = 12;
  p->box_min[0] = -1; p->box_min[1] = 0; p->box_min[2] = -1;
  p->box_max[0] = 1; p->box_max[1] = 1.49; p->box_max[2] = 1;
  p->y_light = 1.49; p->z_front = 1.0;
  p->xsph_mode = PBF_XSPH_JACOBI; p->enable_vorticity = 1; p->enable_xsph = 1;
}

void* oracle_create(const PbfParams* p, int precision, int collision_mode, int search_mode) {
  Handle* h = new Handle();
  h->precision = precision == 32 ? 32 : 64;
  if (h->precision == 32) h->f = new Oracle<float>(*p, collision_mode, search_mode);
  else h->d = new Oracle<double>(*p, collision_mode, search_mode);
  return h;
}

void oracle_destroy(void* hv) {
  Handle* h = (Handle*)hv;
  delete h->f; delete h->d; delete h;
}

// obstacle spheres (cx, cy, cz, r) x count; literal mode rebuilds the reference's BVH over walls + spheres
void oracle_set_spheres(void* hv, size_t count, const double* cxcyczr) {
  visit((Handle*)hv, [&](auto& o) { o.set_spheres(count, cxcyczr); return 0; });
}

// obstacle triangles, 18 doubles each: p1, p2, p3, n1, n2, n3 (MarchingTriangle primitives in the reference's BVH)
void oracle_set_triangles(void* hv, size_t count, const double* pn) {
  visit((Handle*)hv, [&](auto& o) { o.set_triangles(count, pn); return 0; });
}

void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_max_threads() { return omp_get_max_threads(); }

void oracle_upload(void* hv, size_t n, const double* pos, const double* vel) {
  visit((Handle*)hv, [&](auto& o) { o.upload(n, pos, vel); return 0; });
}

void oracle_estimate_densities(void* hv) {
  visit((Handle*)hv, [&](auto& o) { o.estimate_densities(); return 0; });
}

void oracle_step(void* hv, int steps) {
  Handle* h = (Handle*)hv;
  auto t0 = std::chrono::steady_clock::now();
  visit(h, [&](auto& o) { for (int s = 0; s < steps; s++) o.step(); return 0; });
  h->last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void oracle_download(void* hv, double* pos, double* vel, double* dens) {
  visit((Handle*)hv, [&](auto& o) {
    for (size_t i = 0; i < o.n; i++) {
      if (pos) { pos[3*i] = o.pos[i].x; pos[3*i+1] = o.pos[i].y; pos[3*i+2] = o.pos[i].z; }
      if (vel) { vel[3*i] = o.vel[i].x; vel[3*i+1] = o.vel[i].y; vel[3*i+2] = o.vel[i].z; }
      if (dens) dens[i] = o.dens[i];
    }
    return 0;
  });
}

// which: PBF_ARRAY_* of include/pbf_b200.h
void oracle_download_array(void* hv, int which, double* out) {
  visit((Handle*)hv, [&](auto& o) {
    for (size_t i = 0; i < o.n; i++) {
      switch (which) {
        case PBF_ARRAY_XSTAR: out[3*i] = o.npos[i].x; out[3*i+1] = o.npos[i].y; out[3*i+2] = o.npos[i].z; break;
        case PBF_ARRAY_LAMBDA: out[i] = o.lam[i]; break;
        case PBF_ARRAY_VORTICITY: out[3*i] = o.vort[i].x; out[3*i+1] = o.vort[i].y; out[3*i+2] = o.vort[i].z; break;
        case PBF_ARRAY_XPRED: out[3*i] = o.xpred[i].x; out[3*i+1] = o.xpred[i].y; out[3*i+2] = o.xpred[i].z; break;
      }
    }
    return 0;
  });
}

// Particles::estimateDensityAt (particles.cpp:446-453) at m query points
void oracle_density_at(void* hv, size_t m, const double* q, double* out) {
  visit((Handle*)hv, [&](auto& o) {
    typedef typename std::remove_reference<decltype(o)>::type O;
    typedef typename O::V V;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)m; i++)
      out[i] = (double)o.density_at(V((decltype(o.H))q[3*i], (decltype(o.H))q[3*i+1], (decltype(o.H))q[3*i+2]));
    return 0;
  });
}

// Particles::getSurfacePrims restated (fp64 handle only): returns the number of triangles, writes at most cap (18 doubles each)
size_t oracle_surface(void* hv, const double* lo, const double* hi, double iso, double step, double eps, size_t cap, double* out) {
  Handle* h = (Handle*)hv;
  if (!h->d) return 0;
  std::vector<double> t = h->d->surface(lo, hi, iso, step, eps);
  const size_t nt = t.size() / 18;
  if (out) std::memcpy(out, t.data(), sizeof(double) * 18 * std::min(nt, cap));
  return nt;
}
// marching.cpp polygonise for one sign pattern on the unit cell with corner values 0 (inside, bit set) / 1 and iso 0.5:
// every vertex is an edge midpoint.  Returns the number of triangles (<= 5), 9 doubles each.
int oracle_polygonise_case(int cube, double* out) {
  typedef Oracle<double>::V V;
  const V p[8] = {V(0, 0, 0), V(0, 1, 0), V(1, 1, 0), V(1, 0, 0), V(0, 0, 1), V(0, 1, 1), V(1, 1, 1), V(1, 0, 1)};
  double val[8];
  for (int i = 0; i < 8; i++) val[i] = ((cube >> i) & 1) ? 0.0 : 1.0;
  std::vector<double> t;
  Oracle<double>::polygonise(p, val, 0.5, t);
  std::memcpy(out, t.data(), sizeof(double) * t.size());
  return (int)(t.size() / 9);
}

size_t oracle_num_pairs(void* hv) {
  return visit((Handle*)hv, [&](auto& o) { return (size_t)o.col.size(); });
}

void oracle_neighbors(void* hv, uint32_t* row_ptr, uint32_t* col) {
  visit((Handle*)hv, [&](auto& o) {
    std::copy(o.row_ptr.begin(), o.row_ptr.end(), row_ptr);
    for (size_t k = 0; k < o.col.size(); k++) col[k] = (uint32_t)o.col[k];
    return 0;
  });
}

void oracle_neighbor_digest(void* hv, uint64_t* digest, uint32_t* count) {
  visit((Handle*)hv, [&](auto& o) {
    for (size_t i = 0; i < o.n; i++) {
      uint64_t d = 0;
      for (uint32_t k = o.row_ptr[i]; k < o.row_ptr[i + 1]; k++) d += mix64((uint64_t)o.col[k]);
      digest[i] = d; count[i] = o.row_ptr[i + 1] - o.row_ptr[i];
    }
    return 0;
  });
}

void oracle_stats(void* hv, double* first, double* final_, double* ms) {
  Handle* h = (Handle*)hv;
  visit(h, [&](auto& o) { if (first) *first = o.avg_rho_first; if (final_) *final_ = o.avg_rho_final; return 0; });
  if (ms) *ms = h->last_ms;
}

}  // extern "C"
