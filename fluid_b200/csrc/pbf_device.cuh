// pbf_device.cuh — device-side math of the B200 PBF step (sm_100a, fp32).
//
// Two arithmetic regimes live here, on purpose:
//   * EXACT regime (predict + box collision + neighbour predicate): every operation is a single
//     IEEE round-to-nearest fp32 op written with __f*_rn intrinsics, which nvcc never contracts
//     into FMAs.  The result is bit-identical to the same expression compiled for the host with
//     -ffp-contract=off, which is how the fp32 oracle is built, so predicted positions x* and the
//     frozen neighbour SETS can be compared without tolerance.
//   * FAST regime (lambda, delta-p, XSPH/vorticity sums): FMA contraction, rsqrtf, coefficients
//     hoisted out of the pair loop.  Checked against the oracle with stated tolerances.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PBF_BVH_MAX_DEPTH 40     // traversal stack; the host build refuses deeper hierarchies

namespace pbf {

// Parameters in working precision; filled on the host (pbf_api.cu) in plain IEEE float arithmetic.
struct DevParams {
  float h, h2, dt, inv_dt, rho0, inv_rho0, eps_relax, kcorr, visc_c, vort_dt_eps, gdt, eps_d;
  float poly6_c;      // 1.56668147106 / h^9
  float spiky_c;      // -3 * 4.774648292756860 / h^6
  float tscale_c;     // poly6_c * tensile scale: (W * tscale) == tscale_c * (h2-r2)^3
  int   n_corr, iterations, enable_vorticity, enable_xsph;
  float bmin[3], bmax[3];     // collision box
  float clo[3], chi[3];       // hard clamp bounds: bmin + eps_d, bmax - eps_d
  float slo[3], shi[3];       // "clear of every plane" bounds: box and virtual planes shrunk by a margin >> rounding
  float olo[3], ohi[3];       // bounding box of all obstacles (spheres, triangles) plus a margin; empty when there are none
  float yl, zf;               // virtual planes
  // uniform grid (cell edge slightly larger than h => 27-cell search is conservative)
  // Along z (the fastest axis of the linear cell id) every cell is split into `zsub` thin cells of edge cell/zsub:
  // the 27-cell neighbourhood is still 9 contiguous z-runs, but a run can be trimmed to the particle's reach at
  // cell/zsub granularity instead of whole cells (35 % fewer candidates in the neighbour search at zsub = 8).
  float inv_cell; float gmin[3];
  float inv_cell_z; int zsub;
  int   gdim[3];              // LOCAL cells per axis (z in thin cells); linear id = (cx*gdim[1] + cy)*gdim[2] + cz  (z fastest)
  // slab decomposition along x: the grid is global (same gmin / inv_cell on every rank, so a
  // position maps to the same cell everywhere); this rank stores columns cx_offset .. cx_offset+gdim[0]-1
  // and owns the global columns [gx_lo, gx_hi).  Single GPU: cx_offset 0, owns everything.
  int   gdim_x_global, cx_offset, gx_lo, gx_hi, hop_left, hop_right;
  // obstacle spheres (pbf_set_obstacle_spheres): centre xyz, radius in .w; r^2 = r*r rounded once
  int   n_sm;                 // SMs of the device (tile order of the gather kernels)
  float one;                  // 1.0f, a value the compiler cannot see (ex_is_neighbor_x2)
  // obstacle triangles (pbf_set_obstacle_triangles): 5 float4 each = p1, e1 = p2-p1, e2 = p3-p1, n1, n2, n3, sg, ngl
  // (sg = +-1 orientation of e1 x e2 against the vertex normals, ngl = |e1 x e2|), stored in the leaf order of the
  // bounding-volume hierarchy `bvh` (see ex_mesh_hit); tri_id = original index of each; tlo / thi = bounding box + margin
  const float4* tri;
  const float4* bvh;
  const uint32_t* tri_id;
  int   n_tri;
  float skin, tol_n, tol_ray;  // fp32 contact rules of the triangles (Oracle<float>: SKIN, TOL_N, TOL_RAY)
  float tlo[3], thi[3];
  int   n_sph;
  float4 sph[8];
  float sph_r2[8];
};

// ---- EXACT regime -------------------------------------------------------------------------------
__device__ __forceinline__ float ex_norm2(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// Frozen-neighbour predicate of the reference: (x_i - x_j).norm2() <= H2 (particles.cpp:260).
__device__ __forceinline__ bool ex_is_neighbor(float3 a, float3 b, float h2) {
  float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
  return ex_norm2(dx, dy, dz) <= h2;
}

// The same predicate for TWO candidates b0, b1 at once with the packed fp32x2 instructions of sm_100a (FADD2 / FMUL2 / FFMA2): every
// half of a packed operation is one IEEE round-to-nearest fp32 operation and the sequence of roundings is that of
// ex_is_neighbor, so the two result bits are bit-identical to two scalar evaluations; the pair costs 8 issue slots
// instead of 16.  ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 although both carry an explicit .rn
// (it does not for the scalar forms) -- one rounding less, no longer the reference's predicate.  The sums are therefore
// written as fma(a, one, b) with `one` = 1.0f read from the kernel parameters: a value ptxas cannot see, a product that is
// exact, hence the rounding of a + b and nothing to contract.  Returns bit 0 for b0, bit 1 for b1.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ uint32_t ex_is_neighbor_x2(unsigned long long ax2, unsigned long long ay2, unsigned long long az2,
                                                      float b0x, float b1x, float b0y, float b1y, float b0z, float b1z, float h2,
                                                      unsigned long long one2) {
  unsigned long long dx, dy, dz, n2;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(ax2), "l"(pack_f32x2(b0x, b1x)));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(ay2), "l"(pack_f32x2(b0y, b1y)));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(az2), "l"(pack_f32x2(b0z, b1z)));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(dx) : "l"(dx));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(dy) : "l"(dy));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(dz) : "l"(dz));
  asm("fma.rn.f32x2 %0, %1, %3, %2;" : "=l"(n2) : "l"(dx), "l"(dy), "l"(one2));
  asm("fma.rn.f32x2 %0, %1, %3, %2;" : "=l"(n2) : "l"(n2), "l"(dz), "l"(one2));
  float n0, n1;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(n0), "=f"(n1) : "l"(n2));
  return (n0 <= h2 ? 1u : 0u) | (n1 <= h2 ? 2u : 0u);
}

__device__ __forceinline__ float comp(const float3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// One-sided analytic walls x-, x+, y-, z- (the Cornell box of particles.cpp:59-83 has an open
// front and a light plane instead of z+ / y+ walls).  Same visiting order and "t <= max_t"
// acceptance as the oracle's box_hit().
__device__ __forceinline__ bool ex_box_hit(const DevParams& P, float3 o, float3 d, float& max_t,
                                           int& axis, int& side, int skip_axis, int skip_side) {
  bool hit = false;
#pragma unroll
  for (int w = 0; w < 4; w++) {
    const int a = (w < 2) ? 0 : (w == 2 ? 1 : 2);
    const int s = (w == 1) ? 1 : 0;
    if (a == skip_axis && s == skip_side) continue;
    const float plane = s ? P.bmax[a] : P.bmin[a];
    const float da = comp(d, a);
    if (s ? (da > 0.f) : (da < 0.f)) {
      float t = __fdiv_rn(__fsub_rn(plane, comp(o, a)), da);
      if (t < 0.f) t = 0.f;
      if (t <= max_t) { max_t = t; hit = true; axis = a; side = s; }
    }
  }
  return hit;
}

// One-sided obstacle spheres: a sphere blocks only motion INTO it (d . (o - c) < 0), the entry root is clamped to
// t >= 0 and an origin on or inside the surface is in contact now (t = 0).  For an origin outside the sphere this
// is the reference's Sphere::test (sphere.cpp:10-41).  `skip` = the sphere being slid on.  Mirrors
// Oracle<float>::sphere_hit_onesided operation for operation.
__device__ __forceinline__ float ex_dot(float3 a, float3 b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ bool ex_sphere_hit(const DevParams& P, float3 o, float3 d, float& max_t, int& which, float3& nrm, int skip) {
  bool hit = false;
  for (int k = 0; k < P.n_sph; k++) {
    if (k == skip) continue;
    const float4 S = P.sph[k];
    const float3 m = make_float3(__fsub_rn(o.x, S.x), __fsub_rn(o.y, S.y), __fsub_rn(o.z, S.z));
    const float b = __fmul_rn(2.f, ex_dot(m, d));
    if (!(b < 0.f)) continue;
    const float c = __fsub_rn(ex_dot(m, m), P.sph_r2[k]);
    float t;
    if (c <= 0.f) t = 0.f;
    else {
      const float a = ex_dot(d, d);
      const float delta = __fsub_rn(__fmul_rn(b, b), __fmul_rn(__fmul_rn(4.f, a), c));
      if (delta < 0.f) continue;
      t = __fdiv_rn(__fsub_rn(-b, __fsqrt_rn(delta)), __fmul_rn(2.f, a));
      if (t < 0.f) t = 0.f;
    }
    if (t <= max_t) {
      max_t = t; which = k; hit = true;
      const float3 nn = make_float3(__fsub_rn(__fadd_rn(o.x, __fmul_rn(t, d.x)), S.x), __fsub_rn(__fadd_rn(o.y, __fmul_rn(t, d.y)), S.y),
                                    __fsub_rn(__fadd_rn(o.z, __fmul_rn(t, d.z)), S.z));
      const float rn = __fdiv_rn(1.f, __fsqrt_rn(ex_norm2(nn.x, nn.y, nn.z)));
      nrm = make_float3(__fmul_rn(rn, nn.x), __fmul_rn(rn, nn.y), __fmul_rn(rn, nn.z));
    }
  }
  return hit;
}

// One-sided obstacle triangles (Moller-Trumbore as marching_triangle.cpp:21-73): mirrors
// Oracle<float>::mesh_hit_onesided operation for operation.  `slid` = the triangle being slid on (re-tested, but it
// blocks only a direction that dips below its plane by more than TAN: the slide direction comes from the
// interpolated vertex normals and may point into the surface).
__device__ __forceinline__ float3 ex_cross(float3 u, float3 v) {
  return make_float3(__fsub_rn(__fmul_rn(u.y, v.z), __fmul_rn(u.z, v.y)), __fsub_rn(__fmul_rn(u.z, v.x), __fmul_rn(u.x, v.z)),
                     __fsub_rn(__fmul_rn(u.x, v.y), __fmul_rn(u.y, v.x)));
}
// One triangle of the leaf order against the current best hit.  Acceptance = the oracle's sequential scan in ascending
// triangle index with "t <= max_t" (later index wins a tie), written order-independently: strictly nearer, or equally
// near with a larger ORIGINAL index, so the BVH's visiting order cannot change the result.
__device__ __forceinline__ bool ex_tri_test(const DevParams& P, const float4* __restrict__ T, int id, float3 o, float3 d, float& max_t,
                                            int& best, float3& nrm, int slid) {
  const float BT = 1e-6f, ONE_BT = __fadd_rn(1.f, 1e-6f), TOL_T = __fmul_rn(1e-4f, P.h), TAN = 1e-5f;
  const float4 a0 = __ldg(T), a1 = __ldg(T + 1), a2 = __ldg(T + 2);
  const float3 p1 = make_float3(a0.x, a0.y, a0.z), e1 = make_float3(a0.w, a1.x, a1.y), e2 = make_float3(a1.z, a1.w, a2.x);
  const float3 s = make_float3(__fsub_rn(o.x, p1.x), __fsub_rn(o.y, p1.y), __fsub_rn(o.z, p1.z));
  const float3 s1 = ex_cross(d, e2), s2 = ex_cross(s, e1);
  const float dd = ex_dot(s1, e1);
  const float4 a3 = __ldg(T + 3), a4 = __ldg(T + 4);
  const float sg = a4.z, ngl = a4.w;
  if (!(__fmul_rn(sg, dd) > (id == slid ? __fmul_rn(TAN, ngl) : 0.f))) return false;
  float t = __fdiv_rn(ex_dot(s2, e2), dd);
  const float add = fabsf(dd);
  const float shift = __fdiv_rn(__fmul_rn(P.skin, ngl), add);     // rest a skin in front of the plane (fp32 contact rule, see the oracle)
  t = __fsub_rn(t, shift < P.tol_ray ? shift : P.tol_ray);
  if (t < 0.f) {
    if (t >= -TOL_T) t = 0.f;
    else if (t >= -P.tol_ray && __fmul_rn(-t, add) <= __fmul_rn(P.tol_n, ngl)) t = 0.f;   // at most tol_n behind the skin
    else return false;
  }
  if (t > max_t || (t == max_t && best >= 0 && id < best)) return false;
  const float u = __fdiv_rn(ex_dot(s1, s), dd), v = __fdiv_rn(ex_dot(s2, d), dd), w = __fsub_rn(__fsub_rn(1.f, u), v);
  if ((u < -BT) || (u > ONE_BT) || (v < -BT) || (v > ONE_BT) || (w < -BT) || (w > ONE_BT)) return false;
  const float3 n1 = make_float3(a2.y, a2.z, a2.w), n2 = make_float3(a3.x, a3.y, a3.z), n3 = make_float3(a3.w, a4.x, a4.y);
  max_t = t; best = id;
  nrm = make_float3(__fadd_rn(__fadd_rn(__fmul_rn(w, n1.x), __fmul_rn(u, n2.x)), __fmul_rn(v, n3.x)),
                    __fadd_rn(__fadd_rn(__fmul_rn(w, n1.y), __fmul_rn(u, n2.y)), __fmul_rn(v, n3.y)),
                    __fadd_rn(__fadd_rn(__fmul_rn(w, n1.z), __fmul_rn(u, n2.z)), __fmul_rn(v, n3.z)));
  return true;
}

// Nearest one-sided hit of the segment [o, o + max_t d] with the obstacle mesh.  The triangles sit in the leaf order of a
// bounding-volume hierarchy built on the host (pbf_set_obstacle_triangles: 2 float4 per node = (lo, a), (hi, b); a leaf
// holds triangles [a, a+b) of the leaf order, an inner node (b = 0) has its children at a and a+1).  A subtree is
// skipped when the segment's bounding box misses the node's box; node boxes carry a margin above tol_ray (every hit
// the scan accepts lies within tol_ray of the segment), the inflated edges and any rounding of the segment end, so no
// triangle the scan would accept is skipped
// and the result equals the oracle's scan over ALL triangles bit for bit (the reference keeps the same primitives in
// BVHAccel, bvh.cpp:48-192).  `which` = ORIGINAL triangle index.
__device__ __forceinline__ bool ex_mesh_hit(const DevParams& P, float3 o, float3 d, float& max_t, int& which, float3& nrm, int slid) {
  float3 q = make_float3(o.x + max_t * d.x, o.y + max_t * d.y, o.z + max_t * d.z);
  float3 lo = make_float3(fminf(o.x, q.x), fminf(o.y, q.y), fminf(o.z, q.z));
  float3 hi = make_float3(fmaxf(o.x, q.x), fmaxf(o.y, q.y), fmaxf(o.z, q.z));
  if (hi.x < P.tlo[0] || lo.x > P.thi[0] || hi.y < P.tlo[1] || lo.y > P.thi[1] || hi.z < P.tlo[2] || lo.z > P.thi[2]) return false;
  int stack[PBF_BVH_MAX_DEPTH];
  int sp = 0, node = 0, best = -1;
  for (;;) {
    const float4 n0 = __ldg(P.bvh + 2 * node), n1 = __ldg(P.bvh + 2 * node + 1);
    bool descend = false;
    if (!(hi.x < n0.x || lo.x > n1.x || hi.y < n0.y || lo.y > n1.y || hi.z < n0.z || lo.z > n1.z)) {
      const int a = __float_as_int(n0.w), b = __float_as_int(n1.w);
      if (b > 0) {
        bool any = false;
        for (int k = a; k < a + b; k++)
          if (ex_tri_test(P, P.tri + 5 * k, (int)__ldg(P.tri_id + k), o, d, max_t, best, nrm, slid)) any = true;
        if (any) {       // the segment got shorter
          q = make_float3(o.x + max_t * d.x, o.y + max_t * d.y, o.z + max_t * d.z);
          lo = make_float3(fminf(o.x, q.x), fminf(o.y, q.y), fminf(o.z, q.z));
          hi = make_float3(fmaxf(o.x, q.x), fmaxf(o.y, q.y), fmaxf(o.z, q.z));
        }
      } else {
        stack[sp++] = a + 1; node = a; descend = true;
      }
    }
    if (!descend) {
      if (sp == 0) break;
      node = stack[--sp];
    }
  }
  if (best >= 0) which = best;
  return best >= 0;
}

// Swept move of p by delta against the box: clamp() (respond=false, particles.cpp:51-84) and
// clamp_response() (respond=true: slide once along the wall, particles.cpp:87-132), with the
// fp32 contact rules of SURVEY.md §7.3-4 (one-sided planes, exact axis normals, sticky virtual
// planes).  Mirrors Oracle<float>::collide(COLLIDE_ANALYTIC_BOX) operation for operation.
// SPH: obstacles (spheres, triangles) compiled in (the kernels are instantiated both ways and the box-only build is
// launched when the scene has none, so box-only scenes pay nothing for them).
template <bool SPH>
__device__ __forceinline__ float3 ex_collide(const DevParams& P, float3 p, float3 delta, bool respond) {
  {
    // Fast path for the bulk of the fluid: start and end point both clear of every wall and virtual plane by a
    // margin far above any rounding (and, with obstacles, the move's bounding box clear of the obstacles' box), so
    // the general path below would find no hit and return exactly p + delta (its hard clamp is then the
    // identity).  |delta| <= eps_d (returns p unchanged) goes the general way.
    const float3 q = make_float3(__fadd_rn(p.x, delta.x), __fadd_rn(p.y, delta.y), __fadd_rn(p.z, delta.z));
    bool clear = p.x > P.slo[0] && p.x < P.shi[0] && q.x > P.slo[0] && q.x < P.shi[0] &&
                 p.y > P.slo[1] && p.y < P.shi[1] && q.y > P.slo[1] && q.y < P.shi[1] &&
                 p.z > P.slo[2] && p.z < P.shi[2] && q.z > P.slo[2] && q.z < P.shi[2];
    if (SPH)
      clear = clear && (fmaxf(p.x, q.x) < P.olo[0] || fminf(p.x, q.x) > P.ohi[0] || fmaxf(p.y, q.y) < P.olo[1] || fminf(p.y, q.y) > P.ohi[1] ||
                        fmaxf(p.z, q.z) < P.olo[2] || fminf(p.z, q.z) > P.ohi[2]);
    if (clear && ex_norm2(delta.x, delta.y, delta.z) > 1e-20f) return q;
  }
  const float total_l = __fsqrt_rn(ex_norm2(delta.x, delta.y, delta.z));
  float l = total_l;
  if (!(l <= P.eps_d)) {
    const float rc = __fdiv_rn(1.f, l);
    float3 d = make_float3(__fmul_rn(rc, delta.x), __fmul_rn(rc, delta.y), __fmul_rn(rc, delta.z));
    bool virt = false;
    if (d.z > 0.f) { float pt = __fdiv_rn(__fsub_rn(P.zf, p.z), d.z); if (pt >= 0.f && pt < l) { l = pt; virt = true; } }
    if (d.y > 0.f) { float pt = __fdiv_rn(__fsub_rn(P.yl, p.y), d.y); if (pt >= 0.f && pt < l) { l = pt; virt = true; } }
    float max_t = l; int axis = -1, side = 0, sph = -1, tri = -1;
    float3 sn = make_float3(0.f, 0.f, 0.f);
    bool hit = ex_box_hit(P, p, d, max_t, axis, side, -1, 0);
    if (SPH && ex_sphere_hit(P, p, d, max_t, sph, sn, -1)) hit = true;           // nearest of walls, spheres and triangles
    if (SPH && P.n_tri > 0 && ex_mesh_hit(P, p, d, max_t, tri, sn, -1)) { hit = true; sph = -1; }
    if (hit || virt) {
      const float s = __fsub_rn(max_t, P.eps_d);
      p.x = __fadd_rn(p.x, __fmul_rn(s, d.x));
      p.y = __fadd_rn(p.y, __fmul_rn(s, d.y));
      p.z = __fadd_rn(p.z, __fmul_rn(s, d.z));
      if (respond && hit && !virt) {
        float dn; float3 tg = delta;
        if (SPH && (sph >= 0 || tri >= 0)) {              // radial / interpolated normal; tangent = delta - (delta . n) n
          dn = ex_dot(d, sn);
          const float dd = ex_dot(delta, sn);
          tg = make_float3(__fsub_rn(delta.x, __fmul_rn(dd, sn.x)), __fsub_rn(delta.y, __fmul_rn(dd, sn.y)), __fsub_rn(delta.z, __fmul_rn(dd, sn.z)));
        } else {
          dn = side ? -comp(d, axis) : comp(d, axis);
          if (axis == 0) tg.x = 0.f; else if (axis == 1) tg.y = 0.f; else tg.z = 0.f;
        }
        const float t2 = ex_norm2(tg.x, tg.y, tg.z);
        if (dn > -1.f && t2 > 0.f) {
          const float rn = __fdiv_rn(1.f, __fsqrt_rn(t2));
          float3 d2 = make_float3(__fmul_rn(rn, tg.x), __fmul_rn(rn, tg.y), __fmul_rn(rn, tg.z));
          float mt = __fmul_rn(__fsub_rn(total_l, max_t), 0.5f);
          int a2 = -1, s2 = 0, k2 = -1; float3 n2;
          ex_box_hit(P, p, d2, mt, a2, s2, (SPH && (sph >= 0 || tri >= 0)) ? -1 : axis, side);   // the wall / sphere being slid on is never re-tested
          if (SPH) ex_sphere_hit(P, p, d2, mt, k2, n2, sph);
          if (SPH && P.n_tri > 0) ex_mesh_hit(P, p, d2, mt, k2, n2, tri);
          const float s3 = __fsub_rn(mt, P.eps_d);
          p.x = __fadd_rn(p.x, __fmul_rn(s3, d2.x));
          p.y = __fadd_rn(p.y, __fmul_rn(s3, d2.y));
          p.z = __fadd_rn(p.z, __fmul_rn(s3, d2.z));
        }
      }
    } else {
      p.x = __fadd_rn(p.x, delta.x); p.y = __fadd_rn(p.y, delta.y); p.z = __fadd_rn(p.z, delta.z);
    }
    // hard clamp, std::min/std::max semantics of particles.cpp:81-83
    float t;
    t = (p.x < P.chi[0]) ? p.x : P.chi[0]; p.x = (P.clo[0] < t) ? t : P.clo[0];
    t = (p.y < P.chi[1]) ? p.y : P.chi[1]; p.y = (P.clo[1] < t) ? t : P.clo[1];
    t = (p.z < P.chi[2]) ? p.z : P.chi[2]; p.z = (P.clo[2] < t) ? t : P.clo[2];
  }
  return p;
}

// Cell coordinates of a position (clamped into the grid).  Conservative for the 27-cell search
// because the cell edge is h*(1+2^-8) and fp32 rounding of (x-gmin)*inv_cell is << 2^-9 cells
// for the domains we support (<= 2^12 cells per axis).  z is in thin cells (see DevParams::zsub); the search
// derives its z-range from the particle's z, not from +-1 cell.
__device__ __forceinline__ int3 cell_coords_global(const DevParams& P, float x, float y, float z) {
  int cx = (int)floorf((x - P.gmin[0]) * P.inv_cell);
  int cy = (int)floorf((y - P.gmin[1]) * P.inv_cell);
  int cz = (int)floorf((z - P.gmin[2]) * P.inv_cell_z);
  cx = min(max(cx, 0), P.gdim_x_global - 1);
  cy = min(max(cy, 0), P.gdim[1] - 1);
  cz = min(max(cz, 0), P.gdim[2] - 1);
  return make_int3(cx, cy, cz);
}
// local cell coordinates (x column relative to this rank's first stored column)
__device__ __forceinline__ int3 cell_coords(const DevParams& P, float x, float y, float z) {
  int3 c = cell_coords_global(P, x, y, z);
  c.x = min(max(c.x - P.cx_offset, 0), P.gdim[0] - 1);
  return c;
}
__device__ __forceinline__ uint32_t cell_linear(const DevParams& P, int3 c) {
  return (uint32_t)((c.x * P.gdim[1] + c.y) * P.gdim[2] + c.z);
}

// ---- FAST regime: pair terms -------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_ftz(float x) {   // one MUFU.RSQ, no denormal fix-up code
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// For r = x_i - x_j returns (via refs) the un-scaled poly6 term t^3 (t = h2 - r2, 0 outside the
// support; particles.cpp:134-141) and the un-scaled spiky gradient magnitude g such that
// grad W = spiky_c * g * r_vec, g = (h - r)^2 / r (0 for r >= h; particles.cpp:143-149).
// Branch-free: the cut-offs are max(.,0) clamps (ALU pipe) instead of compare+select.  The
// reference's "r < 1e-11 -> 0" guard (coincident particles) becomes a clamp of r2 from below:
// g stays finite and multiplies a zero r_vec, so the contribution is 0 as in the reference.
__device__ __forceinline__ void pair_terms(const DevParams& P, float dx, float dy, float dz,
                                           float& r2, float& w3, float& g) {
  r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  const float t = fmaxf(P.h2 - r2, 0.f);
  w3 = t * t * t;
  const float rinv = rsqrt_ftz(fmaxf(r2, 1e-22f));
  const float r = r2 * rinv;
  const float hr = fmaxf(P.h - r, 0.f);
  g = hr * hr * rinv;
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t j) {   // splitmix64 finaliser (the digest function the header documents)
  uint64_t z = j + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace pbf
