// pbf_surface.inl — marching-cubes surface of the fluid on the GPU: Particles::getSurfacePrims
// (particles.cpp:352-391) with generateGridCell (309-324), estimateDensityAt (446-453), getVertexNormal (407-418) and
// polygonise / vertexInterp (marching.cpp:17-403).  Textually included by pbf_kernels.cu.
//
// The reference evaluates the density at a lattice corner as a sum over ALL particles, eight times per lattice cell
// and six more times per triangle vertex: O(cells x N).  Here every evaluation visits only the 3 x 3 cell columns
// around the point in the solver's own cell grid (re-binned by the committed positions), and the whole pipeline —
// corner densities, cube index, triangle count, ordered output offsets, vertices, normals — stays on the device.
// Arithmetic is fp64 written with __d*_rn intrinsics (never contracted into FMAs), operation for operation the
// reference's, on the fp32 particle state: one density term has the reference's bits, only the ORDER of the sum
// differs (cell order instead of particle order), so the triangle soup equals the reference's on the same state to
// ~1e-13 and comes out in the reference's order (cells ix / iy / iz, triangles in table order).
#include "pbf_mc_table.h"

namespace pbf {

__constant__ uint8_t c_mc_tri[256 * 16];      // pbf_mc::expand(): edges of case c at [16 c ..], 0xFF-terminated
__constant__ uint8_t c_mc_edge[12 * 2];

struct SurfGrid {
  double lo[3], hi[3], step, iso, eps;
  double h2, h9;                   // H2 = H * H, H^9 = intpow<9>(H) (particles.cpp:28,134-141)
  int n1[3];                       // cells per axis = steps + 1 (the loops of particles.cpp:366-368 are inclusive)
  unsigned long long ncell;
};

// Particles::estimateDensityAt(q): sum of poly6(x_p - q); poly6 = 1.56668147106 * intpow<3>(H2 - r2) / H^9 for r2 < H2
__device__ double surf_density(const DevParams& P, const SurfGrid& G, double qx, double qy, double qz,
                               const float4* __restrict__ xs, const uint32_t* __restrict__ cell_start) {
  const float fx = (float)qx, fy = (float)qy, fz = (float)qz;
  const float cxf = floorf((fx - P.gmin[0]) * P.inv_cell), cyf = floorf((fy - P.gmin[1]) * P.inv_cell);
  if (!(cxf >= -1.f && cxf <= (float)P.gdim[0] && cyf >= -1.f && cyf <= (float)P.gdim[1])) return 0.0;   // further than a cell from the grid (or NaN)
  const int cx = (int)cxf, cy = (int)cyf;
  const float reach = P.h * (1.f + 1e-3f) + 1e-6f * fabsf(fz);
  const float zl = floorf((fz - reach - P.gmin[2]) * P.inv_cell_z), zh = floorf((fz + reach - P.gmin[2]) * P.inv_cell_z);
  if (!(zh >= 0.f && zl <= (float)(P.gdim[2] - 1))) return 0.0;
  const int zlo = (int)fmaxf(zl, 0.f), zhi = (int)fminf(zh, (float)(P.gdim[2] - 1));
  double sum = 0.0;
#pragma unroll 1
  for (int k = 0; k < 9; k++) {
    const int ccx = cx + (k / 3) - 1, ccy = cy + (k % 3) - 1;
    if (ccx < 0 || ccx >= P.gdim[0] || ccy < 0 || ccy >= P.gdim[1]) continue;
    const uint32_t base = (uint32_t)((ccx * P.gdim[1] + ccy) * P.gdim[2]);
    const uint32_t jb = cell_start[base + zlo], je = cell_start[base + zhi + 1];
    for (uint32_t j = jb; j < je; j++) {
      const float4 pj = __ldg(&xs[j]);
      const double dx = __dsub_rn((double)pj.x, qx), dy = __dsub_rn((double)pj.y, qy), dz = __dsub_rn((double)pj.z, qz);
      const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      if (r2 < G.h2) {
        const double t = __dsub_rn(G.h2, r2);
        sum = __dadd_rn(sum, __ddiv_rn(__dmul_rn(1.56668147106, __dmul_rn(__dmul_rn(t, t), t)), G.h9));
      }
    }
  }
  return sum;
}

// corner `i` of lattice cell (ix, iy, iz): generateGridCell's numbering; the upper corner is clipped to hi
__device__ __forceinline__ void surf_corner(const SurfGrid& G, int ix, int iy, int iz, int i, double& x, double& y, double& z) {
  const double x1 = __dadd_rn(G.lo[0], __dmul_rn((double)ix, G.step)), y1 = __dadd_rn(G.lo[1], __dmul_rn((double)iy, G.step)),
               z1 = __dadd_rn(G.lo[2], __dmul_rn((double)iz, G.step));
  const double x2 = fmin(G.hi[0], __dadd_rn(x1, G.step)), y2 = fmin(G.hi[1], __dadd_rn(y1, G.step)), z2 = fmin(G.hi[2], __dadd_rn(z1, G.step));
  x = (i == 2 || i == 3 || i == 6 || i == 7) ? x2 : x1;
  y = (i == 1 || i == 2 || i == 5 || i == 6) ? y2 : y1;
  z = (i >= 4) ? z2 : z1;
}
__device__ __forceinline__ void surf_cell_of(const SurfGrid& G, unsigned long long c, int& ix, int& iy, int& iz) {
  iz = (int)(c % (unsigned long long)G.n1[2]);
  iy = (int)((c / (unsigned long long)G.n1[2]) % (unsigned long long)G.n1[1]);
  ix = (int)(c / ((unsigned long long)G.n1[2] * (unsigned long long)G.n1[1]));
}

// pass 1: one lane per (cell, corner): corner density, cube index by ballot, triangle count of the cell
__global__ void __launch_bounds__(TPB)
k_surf_corners(const __grid_constant__ DevParams P, const __grid_constant__ SurfGrid G, const float4* __restrict__ xs,
               const uint32_t* __restrict__ cell_start, double* __restrict__ vals, uint32_t* __restrict__ counts) {
  const unsigned long long t = (unsigned long long)blockIdx.x * TPB + threadIdx.x;
  const unsigned long long c = t >> 3;
  const int corner = (int)(t & 7);
  double v = 0.0;
  const bool live = c < G.ncell;
  if (live) {
    int ix, iy, iz; surf_cell_of(G, c, ix, iy, iz);
    double x, y, z; surf_corner(G, ix, iy, iz, corner, x, y, z);
    v = surf_density(P, G, x, y, z, xs, cell_start);
    vals[8 * c + corner] = v;
  }
  const unsigned b = __ballot_sync(0xffffffffu, live && v < G.iso);
  if (live && corner == 0) {
    const unsigned cube = (b >> (threadIdx.x & 24)) & 0xFFu;
    uint32_t n = 0;
    while (n < 15 && c_mc_tri[16 * cube + n] != 0xFF) n++;
    counts[c] = n / 3;
  }
}

// marching.cpp:381-403
__device__ __forceinline__ void surf_interp(double iso, const double* p1, const double* p2, double v1, double v2, double* out) {
  if (fabs(__dsub_rn(iso, v1)) < 0.00001) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
  if (fabs(__dsub_rn(iso, v2)) < 0.00001) { out[0] = p2[0]; out[1] = p2[1]; out[2] = p2[2]; return; }
  if (fabs(__dsub_rn(v1, v2)) < 0.00001) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
  const double mu = __ddiv_rn(__dsub_rn(iso, v1), __dsub_rn(v2, v1));
  for (int a = 0; a < 3; a++) out[a] = __dadd_rn(p1[a], __dmul_rn(mu, __dsub_rn(p2[a], p1[a])));
}

// pass 2: one thread per lattice cell that emits triangles: vertices on the cut edges, written at the cell's offset
__global__ void __launch_bounds__(TPB)
k_surf_emit(const __grid_constant__ SurfGrid G, const double* __restrict__ vals, const uint32_t* __restrict__ counts,
            const uint32_t* __restrict__ start, double* __restrict__ tris) {
  const unsigned long long c = (unsigned long long)blockIdx.x * TPB + threadIdx.x;
  if (c >= G.ncell || counts[c] == 0) return;
  int ix, iy, iz; surf_cell_of(G, c, ix, iy, iz);
  double val[8], p[8][3];
  unsigned cube = 0;
  for (int i = 0; i < 8; i++) {
    val[i] = vals[8 * c + i];
    if (val[i] < G.iso) cube |= 1u << i;
    surf_corner(G, ix, iy, iz, i, p[i][0], p[i][1], p[i][2]);
  }
  double* out = tris + 18ull * start[c];
  for (int k = 0; k < 15 && c_mc_tri[16 * cube + k] != 0xFF; k++) {
    const int e = c_mc_tri[16 * cube + k], a = c_mc_edge[2 * e], b = c_mc_edge[2 * e + 1];
    double q[3];
    surf_interp(G.iso, p[a], p[b], val[a], val[b], q);
    double* dst = out + 18 * (k / 3) + 3 * (k % 3);
    dst[0] = q[0]; dst[1] = q[1]; dst[2] = q[2];
  }
}

// pass 3: vertex normals (particles.cpp:407-418): 8 lanes per vertex, 6 of them evaluate the field at +-eps along an axis
__global__ void __launch_bounds__(TPB)
k_surf_normals(const __grid_constant__ DevParams P, const __grid_constant__ SurfGrid G, const float4* __restrict__ xs,
               const uint32_t* __restrict__ cell_start, unsigned long long nvert, double* __restrict__ tris) {
  const unsigned long long t = (unsigned long long)blockIdx.x * TPB + threadIdx.x;
  const unsigned long long v = t >> 3;
  const int lane8 = (int)(t & 7);
  const bool live = v < nvert;
  double d = 0.0;
  double* tri = tris + 18ull * (live ? v / 3 : 0);
  const int k = live ? (int)(v % 3) : 0;
  if (live && lane8 < 6) {
    double q[3] = {tri[3 * k], tri[3 * k + 1], tri[3 * k + 2]};
    const int axis = lane8 >> 1;
    q[axis] = (lane8 & 1) ? __dadd_rn(q[axis], G.eps) : __dsub_rn(q[axis], G.eps);
    d = surf_density(P, G, q[0], q[1], q[2], xs, cell_start);
  }
  const int base = threadIdx.x & 24;
  double dm[6];
  for (int i = 0; i < 6; i++) dm[i] = __shfl_sync(0xffffffffu, d, base + i);
  if (live && lane8 == 0) {
    double nx = __dsub_rn(dm[0], dm[1]), ny = __dsub_rn(dm[2], dm[3]), nz = __dsub_rn(dm[4], dm[5]);
    const double nn = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
    if (nn > 0.0) {                                 // unit() multiplies by 1 / norm (vector3D.h:121-124)
      const double rc = __ddiv_rn(1.0, nn);
      nx = __dmul_rn(nx, rc); ny = __dmul_rn(ny, rc); nz = __dmul_rn(nz, rc);
    }
    tri[9 + 3 * k] = nx; tri[9 + 3 * k + 1] = ny; tri[9 + 3 * k + 2] = nz;
  }
}

}  // namespace pbf

extern "C" int pbf_extract_surface(pbf_handle* h, const double* lo, const double* hi, double isolevel, double step, double grad_eps,
                                   size_t cap_triangles, double* tris_out, size_t* n_triangles) {
  using namespace pbf;
  if (!h || !lo || !hi || !n_triangles || (cap_triangles && !tris_out)) return PBF_ERR_INVALID;
  auto bad = [&](int code, const char* msg) { h->last_error = msg; return code; };
  *n_triangles = 0;
  if (h->slab) return bad(PBF_ERR_INVALID, "pbf_extract_surface is single-GPU only");
  if (!(step > 0) || !std::isfinite(step) || !std::isfinite(isolevel) || !std::isfinite(grad_eps)) return bad(PBF_ERR_INVALID, "pbf_extract_surface: bad step / isolevel / eps");
  SurfGrid G;
  unsigned long long ncell = 1;
  for (int a = 0; a < 3; a++) {
    if (!std::isfinite(lo[a]) || !std::isfinite(hi[a]) || !(hi[a] >= lo[a])) return bad(PBF_ERR_INVALID, "pbf_extract_surface: bad lattice bounds");
    G.lo[a] = lo[a]; G.hi[a] = hi[a];
    const double steps = (hi[a] - lo[a]) / step;              // int xsteps = (int)((xmax-xmin)/fStepSize), particles.cpp:359-361
    if (!(steps < 1e6)) return bad(PBF_ERR_CAPACITY, "pbf_extract_surface: lattice too fine");
    G.n1[a] = (int)steps + 1;
    ncell *= (unsigned long long)G.n1[a];
  }
  if (ncell > (1ull << 27)) return bad(PBF_ERR_CAPACITY, "pbf_extract_surface: more than 2^27 lattice cells; extract the surface in tiles");
  G.ncell = ncell; G.step = step; G.iso = isolevel; G.eps = grad_eps;
  { volatile double H = h->hp.h, H2 = H * H, r = 1.0; for (int i = 0; i < 9; i++) r = r * H; G.h2 = H2; G.h9 = r; }
  if (cudaSetDevice(h->device) != cudaSuccess) return bad(PBF_ERR_CUDA, "cudaSetDevice failed");
  if (h->n == 0) return PBF_OK;
  if (h->rebinned_at != (long long)h->steps_done) {     // cells of the last step belong to the predicted positions
    enqueue_rebin(h);
    h->have_neighbors = false;
    h->rebinned_at = (long long)h->steps_done;
  }
  static bool tables_ready[64] = {false};
  if (h->device < 64 && !tables_ready[h->device]) {
    uint8_t tab[256][16]; pbf_mc::expand(tab);
    if (cudaMemcpyToSymbol(c_mc_tri, tab, sizeof(tab)) != cudaSuccess || cudaMemcpyToSymbol(c_mc_edge, pbf_mc::kEdgeCorner, 24) != cudaSuccess)
      return bad(PBF_ERR_CUDA, "pbf_extract_surface: cannot upload the marching-cubes table");
    tables_ready[h->device] = true;
  }
  double *vals = nullptr, *tris = nullptr; uint32_t *counts = nullptr, *start = nullptr, *bsums = nullptr;
  const unsigned sb = blocks_for((uint32_t)ncell, SCAN_TILE);
  auto release = [&]() { cudaFree(vals); cudaFree(tris); cudaFree(counts); cudaFree(start); cudaFree(bsums); };
  if (cudaMalloc((void**)&vals, 8 * ncell * sizeof(double)) != cudaSuccess || cudaMalloc((void**)&counts, ncell * sizeof(uint32_t)) != cudaSuccess ||
      cudaMalloc((void**)&start, (ncell + 1) * sizeof(uint32_t)) != cudaSuccess || cudaMalloc((void**)&bsums, (sb + 1) * sizeof(uint32_t)) != cudaSuccess) {
    cudaGetLastError(); release();
    return bad(PBF_ERR_CAPACITY, "pbf_extract_surface: out of device memory for the lattice");
  }
  LAUNCH(h, K_DENSITY, k_surf_corners, (unsigned)((8 * ncell + TPB - 1) / TPB), h->dp, G, h->xs_a, h->cell_start, vals, counts);
  LAUNCH(h, K_SCAN, k_scan_reduce, sb, (uint32_t)ncell, counts, bsums);
  LAUNCH(h, K_SCAN, k_scan_block_sums, 1, sb, bsums);
  LAUNCH(h, K_SCAN, k_scan_apply, sb, (uint32_t)ncell, counts, bsums, start);
  uint32_t total = 0;
  cudaError_t e = cudaMemcpyAsync(&total, start + ncell, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream);
  int rc = sync_and_check(h);
  if (rc != PBF_OK || e != cudaSuccess) { release(); if (rc == PBF_OK) { h->last_error = cudaGetErrorString(e); rc = PBF_ERR_CUDA; } return rc; }
  *n_triangles = total;
  if (total == 0) { release(); return PBF_OK; }
  if (cudaMalloc((void**)&tris, 18ull * total * sizeof(double)) != cudaSuccess) { cudaGetLastError(); release(); return bad(PBF_ERR_CAPACITY, "pbf_extract_surface: out of device memory for the triangles"); }
  LAUNCH(h, K_DENSITY, k_surf_emit, (unsigned)((ncell + TPB - 1) / TPB), G, vals, counts, start, tris);
  const unsigned long long nvert = 3ull * total;
  LAUNCH(h, K_DENSITY, k_surf_normals, (unsigned)((8 * nvert + TPB - 1) / TPB), h->dp, G, h->xs_a, h->cell_start, nvert, tris);
  const size_t ncopy = std::min<size_t>(total, cap_triangles);
  if (ncopy) e = cudaMemcpyAsync(tris_out, tris, 18 * ncopy * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  rc = sync_and_check(h);
  release();
  if (rc != PBF_OK) return rc;
  if (e != cudaSuccess) { h->last_error = cudaGetErrorString(e); return PBF_ERR_CUDA; }
  return PBF_OK;
}
