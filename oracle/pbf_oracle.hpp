// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference PBF step (SsnL/Fluid).
//
// Nothing under fluid_b200/ (the product) includes, links or calls this file.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
// only as the checker / the CPU baseline.
//
// Parity status: PINNED.  The reference has no tests or golden vectors for this path
// (SURVEY.md §4), so the pin is the unmodified reference itself, compiled headlessly by
// oracle/Makefile into oracle/_ref/ref_harness: Oracle<double>(XSPH reference order, triangle
// walls) reproduces its full state and ordered neighbour lists bit-for-bit
// (tests/test_oracle_golden.py; committed fixtures in tests/golden/).
//
// Each function cites the reference lines it restates.  Arithmetic ORDER is part of the
// contract (SURVEY.md §8c "rules"): Vector3D / scalar multiplies by the reciprocal
// (CGL/include/CGL/vector3D.h:79-82,100-102), unit() multiplies by 1/sqrt (121-124), intpow<e>
// is ((1*b)*b)... (src/bsdf.h:49-57), norm2 is x*x+y*y+z*z left to right (114-116).
// Compile with -ffp-contract=off.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/pbf_b200.h"
#include "../fluid_b200/csrc/pbf_mc_table.h"   // the marching-cubes table (data only; pinned by the reference fixture)

namespace pbf_oracle {

// ---- CGL::Vector3D semantics (vector3D.h:54-150), generic in the scalar type ------------------
template <class R>
struct V3 {
  R x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(R x_, R y_, R z_) : x(x_), y(y_), z(z_) {}
  V3 operator-() const { return V3(-x, -y, -z); }
  V3 operator+(const V3& v) const { return V3(x + v.x, y + v.y, z + v.z); }
  V3 operator-(const V3& v) const { return V3(x - v.x, y - v.y, z - v.z); }
  V3 operator*(R c) const { return V3(x * c, y * c, z * c); }           // right scalar mult
  V3 operator/(R c) const { const R rc = R(1) / c; return V3(rc * x, rc * y, rc * z); }
  void operator+=(const V3& v) { x += v.x; y += v.y; z += v.z; }
  void operator*=(R c) { x *= c; y *= c; z *= c; }
  void operator/=(R c) { (*this) *= (R(1) / c); }
  R norm() const { return std::sqrt(x * x + y * y + z * z); }
  R norm2() const { return x * x + y * y + z * z; }
  V3 unit() const { R rn = R(1) / std::sqrt(x * x + y * y + z * z); return V3(rn * x, rn * y, rn * z); }
  R& operator[](int i) { return (&x)[i]; }
  const R& operator[](int i) const { return (&x)[i]; }
};
template <class R> inline V3<R> operator*(R c, const V3<R>& v) { return V3<R>(c * v.x, c * v.y, c * v.z); }
template <class R> inline R dot(const V3<R>& u, const V3<R>& v) { return u.x * v.x + u.y * v.y + u.z * v.z; }
template <class R> inline V3<R> cross(const V3<R>& u, const V3<R>& v) {
  return V3<R>(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
template <class R> inline R intpow(R b, int e) { R r = R(1); for (int i = 0; i < e; i++) r = r * b; return r; }

// std::min/std::max as the reference uses them (NaN behaviour included)
template <class R> inline R rmin(R a, R b) { return (b < a) ? b : a; }
template <class R> inline R rmax(R a, R b) { return (a < b) ? b : a; }

inline uint64_t mix64(uint64_t j) {   // digest of a neighbour index (order-independent sum)
  uint64_t z = j + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

enum { COLLIDE_REFERENCE_TRIANGLES = 0, COLLIDE_ANALYTIC_BOX = 1 };
enum { SEARCH_BRUTE = 0, SEARCH_GRID = 1 };

// ---- wall triangles of the reference scene (dae/sky/CBempty.dae:229-444 as exact quads) --------
template <class R>
struct Tri { V3<R> p1, p2, p3, n1, n2, n3; R sg, ngl; };   // vertex normals (marching_triangle.h); sg, ngl: see mesh_hit_onesided
// ---- obstacle sphere (static_scene/sphere.h:23-24: r2 = r*r), e.g. the two r = 0.3 spheres of the
// CBspheres scenes (dae/sky/CBspheres_lambertian.dae:291-305,575-594) -------------------------------------------
template <class R>
struct Sph { V3<R> c; R r, r2; };

template <class R>
struct Oracle {
  typedef V3<R> V;
  PbfParams P;
  int collision_mode, search_mode;
  // parameters in working precision (particles.cpp:24-44)
  R H, H2, H6, H9, DT, RHO0, EPS_RELAX, KCORR, VISC, VORT_EPS, GRAV, EPS_D, TSCALE;
  int NCORR, ITERS;
  V bmin, bmax; R YL, ZF;
  R SKIN, TOL_N, TOL_RAY;          // fp32-only contact rules of the obstacle triangles (0 in fp64)
  std::vector<Tri<R>> tris;       // the five walls (10 triangles)
  std::vector<Sph<R>> spheres;
  std::vector<Tri<R>> mesh;       // obstacle triangles (any MarchingTriangle / Triangle primitives of the scene's BVH)

  size_t n = 0;
  std::vector<V> pos, npos, vel, vort, xpred;
  std::vector<R> dens, lam;
  std::vector<uint32_t> row_ptr;
  std::vector<int32_t> col;
  double avg_rho_first = 0, avg_rho_final = 0, sim_time = 0;

  Oracle(const PbfParams& p, int cmode, int smode) : P(p), collision_mode(cmode), search_mode(smode) {
    H = R(p.h); H2 = H * H;   // literal H2 0.09 == 0.3*0.3 in fp64 and fp32 (tests assert it)
    H6 = intpow(H, 6); H9 = intpow(H, 9);   // intpow<6>(H), intpow<9>(H): hoisted, same bits
    DT = R(p.dt); RHO0 = R(p.rest_density); EPS_RELAX = R(p.eps_relax); KCORR = R(p.k_corr);
    VISC = R(p.visc_c); VORT_EPS = R(p.vort_eps); GRAV = R(p.gravity_y); EPS_D = R(1e-11);  // misc.h:11
    NCORR = p.n_corr; ITERS = p.iterations;
    bmin = V(R(p.box_min[0]), R(p.box_min[1]), R(p.box_min[2]));
    bmax = V(R(p.box_max[0]), R(p.box_max[1]), R(p.box_max[2]));
    YL = R(p.y_light); ZF = R(p.z_front);
    // fp32 contact rules of the obstacle triangles scale with the resolution of a coordinate (see mesh_hit_onesided)
    SKIN = TOL_N = R(0); TOL_RAY = R(0);
    if (sizeof(R) == 4 && !std::getenv("PBF_ORACLE_NO_FP32_CONTACT_RULES")) {   // the switch exists for the test that shows what they are for
      double m = 0;
      for (int a = 0; a < 3; a++) m = std::max(m, std::max(std::fabs((double)p.box_min[a]), std::fabs((double)p.box_max[a])));
      const double ulp = m * 1.1920928955078125e-07;
      SKIN = R(std::max(1e-5 * (double)p.h, 16.0 * ulp));
      TOL_N = R(std::max(1e-4 * (double)p.h, 64.0 * ulp) + std::max(1e-5 * (double)p.h, 16.0 * ulp));
      TOL_RAY = R(std::max(0.05 * (double)p.h, 16.0 * std::max(1e-4 * (double)p.h, 64.0 * ulp)));
    }
    // particles.cpp:151  tensile_instability_scale = 1 / poly6(0,0,0.1*H)
    TSCALE = R(1) / poly6(V(R(0), R(0), R(p.dq_ratio) * H));
    build_walls();
    build_bvh();
  }

  void add_quad(V a, V b, V c, V d, V nn) {
    tris.push_back(Tri<R>{a, b, c, nn, nn, nn, R(1), R(0)});
    tris.push_back(Tri<R>{a, c, d, nn, nn, nn, R(1), R(0)});
  }
  // Walls generalised from the Cornell box: ceiling 0.01 above y_light (1.5 vs 1.49), floor,
  // x-, x+, back (z-); the front (z+) is open (virtual plane).  For the default params these are
  // exactly the quads of oracle/ref_harness/harness.cpp.
  void build_walls() {
    R x0 = bmin.x, x1 = bmax.x, y0 = bmin.y, y1 = R(P.box_max[1] + 0.01), z0 = bmin.z, z1 = bmax.z;
    add_quad(V(x1, y1, z0), V(x0, y1, z0), V(x0, y1, z1), V(x1, y1, z1), V(0, -1, 0));
    add_quad(V(x1, y0, z0), V(x1, y0, z1), V(x0, y0, z1), V(x0, y0, z0), V(0, 1, 0));
    add_quad(V(x0, y1, z0), V(x0, y0, z0), V(x0, y0, z1), V(x0, y1, z1), V(1, 0, 0));
    add_quad(V(x1, y1, z1), V(x1, y0, z1), V(x1, y0, z0), V(x1, y1, z0), V(-1, 0, 0));
    add_quad(V(x1, y1, z0), V(x1, y0, z0), V(x0, y0, z0), V(x0, y1, z0), V(0, 0, 1));
  }

  // ---- kernels: particles.cpp:134-149 ----------------------------------------------------------
  R poly6(const V& r) const {
    R r2 = r.norm2();
    if (r2 >= H2) return R(0);
    R t = H2 - r2;
    return R(1.56668147106) * intpow(t, 3) / H9;
  }
  V grad_spiky(const V& r) const {
    R rl = r.norm();
    if (rl >= H || rl < EPS_D) return V();
    // -3 * 4.77... * intpow<2>(H - r_l) * r / (intpow<6>(H) * r_l)
    R s = R(-3 * 4.774648292756860) * intpow(H - rl, 2);
    return (s * r) / (H6 * rl);
  }

  // ---- ray / triangle (static_scene/marching_triangle.cpp:21-73), segment [0, max_t] -----------
  bool tri_hit(const Tri<R>& T, const V& o, const V& d, R& max_t, V* nrm) const {
    V e1 = T.p2 - T.p1, e2 = T.p3 - T.p1, s = o - T.p1;
    V s1 = cross(d, e2), s2 = cross(s, e1);
    R dd = dot(s1, e1);
    if (dd == 0) return false;
    R t = dot(s2, e2) / dd;
    if ((t < R(0)) || (t > max_t)) return false;
    R u = dot(s1, s) / dd, v = dot(s2, d) / dd, w = R(1) - u - v;
    if ((u < 0) || (u > 1) || (v < 0) || (v > 1) || (w < 0) || (w > 1)) return false;
    max_t = t;
    if (nrm) *nrm = w * T.n1 + u * T.n2 + v * T.n3;
    return true;
  }
  // ---- ray / sphere, literal (static_scene/sphere.cpp:10-41,43-76), segment [0, max_t] ----------------------
  // Quirk kept: when the near root is out of range the FAR root is accepted, so a ray that starts inside
  // a sphere "hits" the surface from within.
  bool sphere_hit_literal(const Sph<R>& S, const V& o, const V& d, R& max_t, V* nrm) const {
    R a = dot(d, d);
    V m = o - S.c;
    R b = R(2) * dot(m, d);
    R c = dot(m, m) - S.r2;
    R delta = b * b - R(4) * a * c;
    if (delta < R(0)) return false;
    R t = (-b - std::sqrt(delta)) / (R(2) * a), t1;
    if ((t >= R(0)) && (t <= max_t)) t1 = t;
    else {
      t = (-b + std::sqrt(delta)) / (R(2) * a);
      if ((t >= R(0)) && (t <= max_t)) t1 = t; else return false;
    }
    max_t = t1;
    if (nrm) { V nn = o + t1 * d - S.c; *nrm = nn.unit(); }
    return true;
  }

  // ---- the reference's BVH (bvh.cpp:48-142, bbox.h:29-122, bbox.cpp:10-30), restated so that primitives are
  // visited in the reference's own order.  The order matters with obstacles: clamp() uses the ANY-hit query
  // (particles.cpp:76, bvh.cpp:142-163), which returns the first primitive found, not the nearest one.
  struct Box {
    V mn, mx, ext;
    Box() : mn(inf(), inf(), inf()), mx(-inf(), -inf(), -inf()) { ext = mx - mn; }
    explicit Box(const V& p) : mn(p), mx(p) { ext = mx - mn; }
    Box(const V& a, const V& b) : mn(a), mx(b) { ext = mx - mn; }
    static R inf() { return std::numeric_limits<R>::infinity(); }
    void expand(const Box& b) {
      mn.x = rmin(mn.x, b.mn.x); mn.y = rmin(mn.y, b.mn.y); mn.z = rmin(mn.z, b.mn.z);
      mx.x = rmax(mx.x, b.mx.x); mx.y = rmax(mx.y, b.mx.y); mx.z = rmax(mx.z, b.mx.z);
      ext = mx - mn;
    }
    void expand(const V& p) { expand(Box(p)); }
    V centroid() const { return (mn + mx) / R(2); }
    R area() const {
      if (mn.x > mx.x || mn.y > mx.y || mn.z > mx.z) return R(0);
      return R(2) * (ext.x * ext.z + ext.x * ext.y + ext.y * ext.z);
    }
    bool hit(const V& o, const V& d, R& t0, R& t1) const {   // bbox.cpp:10-30 (division by a zero component included)
      R tmin = -inf(), tmax = inf();
      for (int i = 0; i < 3; i++) {
        R ta = (mn[i] - o[i]) / d[i], tb = (mx[i] - o[i]) / d[i];
        if (ta > tb) { R tt = tb; tb = ta; ta = tt; }
        tmin = rmax(tmin, ta); tmax = rmin(tmax, tb);
        if (tmax < tmin) return false;
      }
      t0 = tmin; t1 = tmax;
      return true;
    }
  };
  struct Node { Box bb; int l = -1, r = -1; size_t lo = 0, hi = 0; };
  std::vector<int> prim_order;     // primitive ids in leaf order; ids: walls, then spheres, then obstacle triangles (the harness's push order)
  std::vector<Node> nodes;
  int bvh_root = -1;

  const Tri<R>* prim_tri(int id) const {
    if (id < (int)tris.size()) return &tris[id];
    id -= (int)(tris.size() + spheres.size());
    return id >= 0 ? &mesh[id] : nullptr;
  }
  Box prim_box(int id) const {
    if (const Tri<R>* T = prim_tri(id)) { Box b(T->p1); b.expand(T->p2); b.expand(T->p3); return b; }
    const Sph<R>& S = spheres[id - (int)tris.size()];
    return Box(S.c - V(S.r, S.r, S.r), S.c + V(S.r, S.r, S.r));
  }
  int build_range(size_t l, size_t r, const Box& bbox, size_t max_leaf) {   // bvh.cpp:48-127
    const size_t cnt = r - l;
    const int me = (int)nodes.size();
    nodes.push_back(Node()); nodes[me].bb = bbox;
    if (cnt <= max_leaf) { nodes[me].lo = l; nodes[me].hi = r; return me; }
    int axis;
    if (bbox.ext[0] > bbox.ext[1] && bbox.ext[0] > bbox.ext[2]) axis = 0;
    else if (bbox.ext[1] > bbox.ext[0] && bbox.ext[1] > bbox.ext[2]) axis = 1;
    else axis = 2;
    std::sort(prim_order.begin() + l, prim_order.begin() + r,
              [&](int a, int b) { return prim_box(a).centroid()[axis] < prim_box(b).centroid()[axis]; });
    std::vector<Box> right(cnt);
    Box acc;
    for (size_t i = r; i-- > l;) { acc.expand(prim_box(prim_order[i])); right[i - l] = acc; }
    Box left = prim_box(prim_order[l]), lbest, rbest;
    R best = Box::inf();
    size_t split = l + 1;
    for (size_t i = l + 1, nl = 1; i < r; i++, nl++) {
      const R cost = left.area() * R((double)nl) + right[nl].area() * R((double)(cnt - nl));
      if (cost < best) { best = cost; split = i; lbest = left; rbest = right[nl]; }
      left.expand(prim_box(prim_order[i]));
    }
    const int lc = build_range(l, split, lbest, max_leaf);
    const int rc = build_range(split, r, rbest, max_leaf);
    nodes[me].l = lc; nodes[me].r = rc;
    return me;
  }
  void build_bvh() {   // bvh.cpp:130-139, default max_leaf_size 4 (bvh.h:71)
    nodes.clear(); prim_order.clear();
    const int np = (int)(tris.size() + spheres.size() + mesh.size());
    Box all;
    for (int i = 0; i < np; i++) { prim_order.push_back(i); all.expand(prim_box(i)); }
    bvh_root = build_range(0, (size_t)np, all, 4);
  }
  bool prim_hit(int id, const V& o, const V& d, R& max_t, V* nrm) const {
    if (const Tri<R>* T = prim_tri(id)) return tri_hit(*T, o, d, max_t, nrm);
    return sphere_hit_literal(spheres[id - (int)tris.size()], o, d, max_t, nrm);
  }
  // nrm == nullptr: any hit (bvh.cpp:142-163, first primitive found wins and sets max_t);
  // nrm != nullptr: nearest hit (bvh.cpp:165-192: every primitive the boxes let through, each hit shrinks max_t)
  bool bvh_hit(int node, const V& o, const V& d, R& max_t, V* nrm) const {
    const Node& N = nodes[node];
    R t0, t1;
    if (!N.bb.hit(o, d, t0, t1)) return false;
    if (t1 < R(0) || t0 > max_t) return false;
    if (N.l < 0) {
      bool hit = false;
      for (size_t i = N.lo; i < N.hi; i++)
        if (prim_hit(prim_order[i], o, d, max_t, nrm)) { if (!nrm) return true; hit = true; }
      return hit;
    }
    if (!nrm) return bvh_hit(N.l, o, d, max_t, nrm) || bvh_hit(N.r, o, d, max_t, nrm);
    bool hit = bvh_hit(N.l, o, d, max_t, nrm);
    hit |= bvh_hit(N.r, o, d, max_t, nrm);
    return hit;
  }
  bool scene_hit(const V& o, const V& d, R& max_t, V* nrm) const { return bvh_hit(bvh_root, o, d, max_t, nrm); }

  // obstacle triangles: 18 doubles each (p1, p2, p3, n1, n2, n3); n == nullptr: geometric normals
  void set_triangles(size_t count, const double* pn) {
    mesh.clear();
    for (size_t k = 0; k < count; k++) {
      const double* q = pn + 18 * k;
      auto vec = [&](int a) { return V(R(q[3 * a]), R(q[3 * a + 1]), R(q[3 * a + 2])); };
      Tri<R> T{vec(0), vec(1), vec(2), vec(3), vec(4), vec(5), R(1), R(0)};
      const V ng = cross(T.p2 - T.p1, T.p3 - T.p1);
      T.sg = dot(ng, T.n1 + T.n2 + T.n3) < R(0) ? R(-1) : R(1);   // +1: the vertex order winds counter-clockwise seen from the normals' side
      T.ngl = ng.norm();                                            // |e1 x e2| = twice the area
      mesh.push_back(T);
    }
    build_bvh();
  }

  void set_spheres(size_t count, const double* cxcyczr) {
    spheres.clear();
    for (size_t k = 0; k < count; k++) {
      const R r = R(cxcyczr[4 * k + 3]);
      spheres.push_back(Sph<R>{V(R(cxcyczr[4 * k]), R(cxcyczr[4 * k + 1]), R(cxcyczr[4 * k + 2])), r, r * r});
    }
    build_bvh();
  }

  // ---- analytic box walls with the fp32 contact rules ------------------------
  // one-sided planes (hit only when moving into the wall), exact axis normals, t >= 0.
  // skip_axis/skip_side: the plane currently being slid on is never re-tested.
  bool box_hit(const V& o, const V& d, R& max_t, int* axis, int* side, int skip_axis, int skip_side) const {
    bool hit = false;
    // order x-, x+, y-, z- ; "t <= max_t" like the reference's acceptance test
    const int ax[4] = {0, 0, 1, 2}; const int sd[4] = {0, 1, 0, 0};
    for (int w = 0; w < 4; w++) {
      int a = ax[w], s = sd[w];
      if (a == skip_axis && s == skip_side) continue;
      R plane = s ? bmax[a] : bmin[a];
      R da = d[a];
      if (s ? (da > 0) : (da < 0)) {
        R t = (plane - o[a]) / da;
        if (t < R(0)) t = R(0);
        if (t <= max_t) { max_t = t; hit = true; *axis = a; *side = s; }
      }
    }
    return hit;
  }

  // One-sided obstacle spheres, the sphere analogue of the wall rules above: a sphere blocks only motion INTO
  // it (d . (o - c) < 0); the entry root is clamped to t >= 0 and an origin on or inside the surface is in
  // contact now (t = 0); motion away from the centre is never blocked.  For an origin outside the sphere
  // this is exactly sphere_hit_literal (both roots share the sign of -b), so in fp64 the two modes agree as
  // long as no particle is inside a sphere.  `skip`: the sphere currently being slid on.
  bool sphere_hit_onesided(const V& o, const V& d, R& max_t, int* which, V* nrm, int skip) const {
    bool hit = false;
    for (int k = 0; k < (int)spheres.size(); k++) {
      if (k == skip) continue;
      const Sph<R>& S = spheres[k];
      V m = o - S.c;
      R b = R(2) * dot(m, d);
      if (!(b < R(0))) continue;
      R c = dot(m, m) - S.r2, t;
      if (c <= R(0)) t = R(0);
      else {
        R a = dot(d, d);
        R delta = b * b - R(4) * a * c;
        if (delta < R(0)) continue;
        t = (-b - std::sqrt(delta)) / (R(2) * a);
        if (t < R(0)) t = R(0);
      }
      if (t <= max_t) { max_t = t; *which = k; V nn = o + t * d - S.c; *nrm = nn.unit(); hit = true; }
    }
    return hit;
  }

  // One-sided obstacle triangles (same Moller-Trumbore expressions as tri_hit): a triangle blocks only motion
  // against its oriented normal; an origin a hair behind its plane (fp32 landing error) is in contact now; the
  // barycentric test is inflated by BT in fp32 so that a ray through a shared edge cannot slip between two
  // triangles.  `slid` = the triangle being slid on: the slide direction comes from the INTERPOLATED vertex normals
  // (particles.cpp:120, marching_triangle.cpp:66-68), which need not be the geometric normal, so it may point into
  // the surface and the triangle must be re-tested; it blocks the slide only when the direction dips below the
  // tangent plane by more than TAN (rounding of an exactly tangent direction must not freeze the particle).
  // In fp64 (BT = 0) this differs from the reference's two-sided test only for origins behind a triangle.
  // fp32 only, so that nothing leaks where a coordinate resolves no better than ~1e-5 (|x| ~ 100): the hit distance is
  // shortened so that a particle comes to rest SKIN = max(1e-5 h, 16 ulp) in front of the plane (measured along the
  // normal) instead of +-1 ulp around it, and an origin up to TOL_N = max(1e-4 h, 64 ulp) + SKIN behind that skin
  // surface (again along the normal, not along the ray: a grazing ray reaches far for a small depth) is still in
  // contact.  Both reach at most TOL_RAY along the ray, so every accepted hit point lies within TOL_RAY of the
  // segment [o, o + max_t d] (what a bounding-volume hierarchy over the triangles may rely on).
  // (ulp = 2^-23 x the largest box coordinate.)  The barycentric test stays that of the UNSHIFTED plane: shifted
  // copies of the triangles would leave gaps along convex edges.
  bool mesh_hit_onesided(const V& o, const V& d, R& max_t, int* which, V* nrm, int slid) const {
    const R BT = sizeof(R) == 4 ? R(1e-6) : R(0), TOL_T = R(1e-4) * H, TAN = R(1e-5);
    bool hit = false;
    for (int k = 0; k < (int)mesh.size(); k++) {
      const Tri<R>& T = mesh[k];
      V e1 = T.p2 - T.p1, e2 = T.p3 - T.p1, s = o - T.p1;
      V s1 = cross(d, e2), s2 = cross(s, e1);
      R dd = dot(s1, e1);                         // = -d . (e1 x e2)
      if (!(T.sg * dd > (k == slid ? TAN * T.ngl : R(0)))) continue;   // moving away from / parallel to the front side
      R t = dot(s2, e2) / dd;
      if (sizeof(R) == 4) {                       // stop a skin in front of the plane: SKIN along the normal = SKIN |n| / |dd| along the ray
        const R shift = (SKIN * T.ngl) / std::fabs(dd);
        t = t - (shift < TOL_RAY ? shift : TOL_RAY);
      }
      if (t < R(0)) {
        if (t >= -TOL_T) t = R(0);
        else if (sizeof(R) == 4 && t >= -TOL_RAY && (-t) * std::fabs(dd) <= TOL_N * T.ngl) t = R(0);   // at most TOL_N behind the skin
        else continue;
      }
      if (t > max_t) continue;
      R u = dot(s1, s) / dd, v = dot(s2, d) / dd, w = R(1) - u - v;
      if ((u < -BT) || (u > R(1) + BT) || (v < -BT) || (v > R(1) + BT) || (w < -BT) || (w > R(1) + BT)) continue;
      max_t = t; *which = k; *nrm = w * T.n1 + u * T.n2 + v * T.n3; hit = true;
    }
    return hit;
  }

  void hard_clamp(V& p) const {   // particles.cpp:81-83 / 129-131
    p.x = rmax(bmin.x + EPS_D, rmin(bmax.x - EPS_D, p.x));
    p.y = rmax(bmin.y + EPS_D, rmin(bmax.y - EPS_D, p.y));
    p.z = rmax(bmin.z + EPS_D, rmin(bmax.z - EPS_D, p.z));
  }

  // particles.cpp:51-84 (respond=false) and 87-132 (respond=true).
  void collide(V& p, const V& delta_p, bool respond) const {
    R total_l = delta_p.norm(), l = total_l;
    if (l <= EPS_D) return;
    V d = delta_p / l;
    bool virt = false;
    if (collision_mode == COLLIDE_REFERENCE_TRIANGLES) {
      if (d.z != R(0)) { R pt = (ZF - p.z) / d.z; if (pt > R(0) && pt < l) { l = pt; virt = true; } }
      if (d.y != R(0)) { R pt = (YL - p.y) / d.y; if (pt > R(0) && pt < l) { l = pt; virt = true; } }
      R max_t = l; V nrm;
      bool hit = scene_hit(p, d, max_t, respond ? &nrm : nullptr);
      if (hit || virt) {
        p += (max_t - EPS_D) * d;
        if (respond && dot(d, nrm) > R(-1) && !virt) {   // slide once (particles.cpp:118-124)
          d = (delta_p - dot(delta_p, nrm) * nrm).unit();
          R mt = (total_l - max_t) * R(0.5);
          V n2;
          scene_hit(p, d, mt, &n2);
          p += (mt - EPS_D) * d;
        }
      } else {
        p += delta_p;
      }
    } else {
      // sticky virtual planes: d>0 && pt>=0 (fp64 never has pt==0; fp32 does)
      if (d.z > R(0)) { R pt = (ZF - p.z) / d.z; if (pt >= R(0) && pt < l) { l = pt; virt = true; } }
      if (d.y > R(0)) { R pt = (YL - p.y) / d.y; if (pt >= R(0) && pt < l) { l = pt; virt = true; } }
      R max_t = l; int axis = -1, side = 0, sph = -1, tri = -1; V sn;
      bool hit = box_hit(p, d, max_t, &axis, &side, -1, 0);
      if (sphere_hit_onesided(p, d, max_t, &sph, &sn, -1)) hit = true;      // nearest of walls, spheres and triangles
      if (mesh_hit_onesided(p, d, max_t, &tri, &sn, -1)) { hit = true; sph = -1; }
      if (hit || virt) {
        p += (max_t - EPS_D) * d;
        if (respond && hit && !virt) {
          R dn; V tang;
          if (sph >= 0 || tri >= 0) {                      // radial normal (sphere.cpp:69-70) / interpolated vertex normals; slide as particles.cpp:118-124
            dn = dot(d, sn);
            tang = delta_p - dot(delta_p, sn) * sn;
          } else {
            // exact axis normal n = +-e_axis: dot(d,n) > -1  <=>  d is not exactly anti-normal
            dn = side ? -d[axis] : d[axis];
            tang = delta_p; tang[axis] = R(0);             // delta - dot(delta,n) n, exactly
          }
          if (dn > R(-1) && tang.norm2() > R(0)) {
            V d2 = tang.unit();
            R mt = (total_l - max_t) * R(0.5);
            int a2 = -1, s2 = 0, k2 = -1; V n2;
            box_hit(p, d2, mt, &a2, &s2, (sph >= 0 || tri >= 0) ? -1 : axis, side);   // the surface being slid on is never re-tested
            sphere_hit_onesided(p, d2, mt, &k2, &n2, sph);
            mesh_hit_onesided(p, d2, mt, &k2, &n2, tri);
            p += (mt - EPS_D) * d2;
          }
        }
      } else {
        p += delta_p;
      }
    }
    hard_clamp(p);
  }

  // ---- state ------------------------------------------------------------------------------------
  void upload(size_t n_, const double* p, const double* v) {
    n = n_;
    pos.resize(n); npos.resize(n); vel.resize(n); vort.assign(n, V()); xpred.resize(n);
    dens.assign(n, R(0)); lam.assign(n, R(0));
    for (size_t i = 0; i < n; i++) {
      pos[i] = V(R(p[3*i]), R(p[3*i+1]), R(p[3*i+2]));
      npos[i] = pos[i];
      vel[i] = V(R(v[3*i]), R(v[3*i+1]), R(v[3*i+2]));
    }
    row_ptr.assign(n + 1, 0); col.clear();
  }

  // particles.cpp:440-444 + 158-163: density over ALL particles including self (Q1)
  void estimate_densities() {
    build_neighbors(pos, /*include_self=*/true);
    #pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; i++) {
      R d = R(0);
      for (uint32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) d += poly6(pos[col[k]] - pos[i]);
      dens[i] = d;
    }
  }

  // particles.cpp:446-453 estimateDensityAt: sum of poly6(x_p - q) over ALL particles, index order
  R density_at(const V& q) const {
    R d = R(0);
    for (size_t i = 0; i < n; i++) d += poly6(pos[i] - q);
    return d;
  }

  // ---- marching-cubes surface -------------------------------------------------
  // marching.cpp:381-403 vertexInterp
  static V vertex_interp(R iso, const V& p1, const V& p2, R v1, R v2) {
    if (std::abs(iso - v1) < R(0.00001)) return p1;
    if (std::abs(iso - v2) < R(0.00001)) return p2;
    if (std::abs(v1 - v2) < R(0.00001)) return p1;
    const R mu = (iso - v1) / (v2 - v1);
    return V(p1.x + mu * (p2.x - p1.x), p1.y + mu * (p2.y - p1.y), p1.z + mu * (p2.z - p1.z));
  }
  // marching.cpp:17-380 polygonise: corner i is "inside" when val[i] < iso; appends 9 R per triangle
  static void polygonise(const V p[8], const R val[8], R iso, std::vector<R>& out) {
    static uint8_t table[256][16];
    static bool ready = false;
    if (!ready) {
#pragma omp critical(pbf_mc_table)
      { if (!ready) { pbf_mc::expand(table); ready = true; } }
    }
    int cube = 0;
    for (int i = 0; i < 8; i++) if (val[i] < iso) cube |= 1 << i;
    if (cube == 0 || cube == 255) return;            // edgeTable[cube] == 0
    V vert[12];
    for (int e = 0; e < 12; e++) {
      const int a = pbf_mc::kEdgeCorner[e][0], b = pbf_mc::kEdgeCorner[e][1];
      if (((cube >> a) & 1) != ((cube >> b) & 1)) vert[e] = vertex_interp(iso, p[a], p[b], val[a], val[b]);
    }
    for (int k = 0; table[cube][k] != 0xFF; k++) {
      const V& q = vert[table[cube][k]];
      out.push_back(q.x); out.push_back(q.y); out.push_back(q.z);
    }
  }
  // particles.cpp:407-418 getVertexNormal: central differences of the density field, not divided by 2 eps
  V vertex_normal(const V& q, R eps) const {
    V nn(density_at(V(q.x - eps, q.y, q.z)) - density_at(V(q.x + eps, q.y, q.z)),
         density_at(V(q.x, q.y - eps, q.z)) - density_at(V(q.x, q.y + eps, q.z)),
         density_at(V(q.x, q.y, q.z - eps)) - density_at(V(q.x, q.y, q.z + eps)));
    if (nn.norm() > R(0)) return nn.unit();
    return nn;
  }
  // particles.cpp:352-391 getSurfacePrims over the lattice [lo, hi] (the reference hard-codes (-1,0,-1)..(1,1.5,1),
  // particles.cpp:326-350): cells in ix / iy / iz order, the last cell of every axis clipped to hi (and of zero
  // width when the step divides the extent); 18 R per triangle: p1 p2 p3 n1 n2 n3.
  std::vector<R> surface(const R lo[3], const R hi[3], R iso, R step, R eps) const {
    const int xs = (int)((hi[0] - lo[0]) / step), ys = (int)((hi[1] - lo[1]) / step), zs = (int)((hi[2] - lo[2]) / step);
    const long long ncell = (long long)(xs + 1) * (ys + 1) * (zs + 1);
    std::vector<std::vector<R>> per(ncell);
#pragma omp parallel for schedule(dynamic, 16)
    for (long long c = 0; c < ncell; c++) {
      const int iz = (int)(c % (zs + 1)), iy = (int)((c / (zs + 1)) % (ys + 1)), ix = (int)(c / ((long long)(zs + 1) * (ys + 1)));
      const R x1 = lo[0] + ix * step, x2 = rmin(hi[0], x1 + step);
      const R y1 = lo[1] + iy * step, y2 = rmin(hi[1], y1 + step);
      const R z1 = lo[2] + iz * step, z2 = rmin(hi[2], z1 + step);
      const V p[8] = {V(x1, y1, z1), V(x1, y2, z1), V(x2, y2, z1), V(x2, y1, z1), V(x1, y1, z2), V(x1, y2, z2), V(x2, y2, z2), V(x2, y1, z2)};
      R val[8];
      for (int i = 0; i < 8; i++) val[i] = density_at(p[i]);
      std::vector<R> tri;
      polygonise(p, val, iso, tri);
      for (size_t t = 0; t + 9 <= tri.size(); t += 9) {
        for (int k = 0; k < 9; k++) per[c].push_back(tri[t + k]);
        for (int v = 0; v < 3; v++) {
          const V nn = vertex_normal(V(tri[t + 3 * v], tri[t + 3 * v + 1], tri[t + 3 * v + 2]), eps);
          per[c].push_back(nn.x); per[c].push_back(nn.y); per[c].push_back(nn.z);
        }
      }
    }
    std::vector<R> out;
    for (long long c = 0; c < ncell; c++) out.insert(out.end(), per[c].begin(), per[c].end());
    return out;
  }

  // particles.cpp:258-265: inclusive predicate on predicted positions, ascending index lists,
  // self excluded (the reference's i<j double loop).  Grid search = conservative binning + the
  // same predicate + ascending sort, so both searches give identical lists.
  // NOTE estimate_densities() sums poly6 over all particles; particles with r2 >= H2 add exactly
  // 0, so restricting to r2 <= H2 (and including self) leaves the sum bit-identical.
  void build_neighbors(const std::vector<V>& x, bool include_self) {
    std::vector<uint32_t> cnt(n + 1, 0);
    if (search_mode == SEARCH_BRUTE) {
      std::vector<std::vector<int32_t>> lists(n);
      for (size_t i = 0; i < n; i++) {
        if (include_self) lists[i].push_back((int32_t)i);
        for (size_t j = i + 1; j < n; j++)
          if ((x[i] - x[j]).norm2() <= H2) { lists[i].push_back((int32_t)j); lists[j].push_back((int32_t)i); }
      }
      col.clear();
      for (size_t i = 0; i < n; i++) {
        std::sort(lists[i].begin(), lists[i].end());
        row_ptr[i] = (uint32_t)col.size();
        col.insert(col.end(), lists[i].begin(), lists[i].end());
      }
      row_ptr[n] = (uint32_t)col.size();
      return;
    }
    // uniform grid, cell edge slightly larger than H, binning in double
    const double cs = double(H) * 1.001;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t i = 0; i < n; i++) for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], (double)x[i][a]); hi[a] = std::max(hi[a], (double)x[i][a]);
    }
    long dim[3];
    for (int a = 0; a < 3; a++) dim[a] = (long)std::floor((hi[a] - lo[a]) / cs) + 1;
    const size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
    std::vector<uint32_t> cstart(ncell + 1, 0), cellof(n);
    auto cell_of = [&](const V& p, long c[3]) {
      for (int a = 0; a < 3; a++) {
        c[a] = (long)std::floor(((double)p[a] - lo[a]) / cs);
        c[a] = std::max(0l, std::min(dim[a] - 1, c[a]));
      }
    };
    for (size_t i = 0; i < n; i++) {
      long c[3]; cell_of(x[i], c);
      cellof[i] = (uint32_t)((c[0] * dim[1] + c[1]) * dim[2] + c[2]);
      cstart[cellof[i] + 1]++;
    }
    for (size_t c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
    std::vector<uint32_t> fill(cstart.begin(), cstart.end() - 1), items(n);
    for (size_t i = 0; i < n; i++) items[fill[cellof[i]]++] = (uint32_t)i;   // ascending within a cell

    auto visit = [&](size_t i, std::vector<int32_t>& out) {
      out.clear();
      long c[3]; cell_of(x[i], c);
      for (long dx = -1; dx <= 1; dx++) for (long dy = -1; dy <= 1; dy++) for (long dz = -1; dz <= 1; dz++) {
        long cx = c[0] + dx, cy = c[1] + dy, cz = c[2] + dz;
        if (cx < 0 || cy < 0 || cz < 0 || cx >= dim[0] || cy >= dim[1] || cz >= dim[2]) continue;
        size_t cc = (size_t)((cx * dim[1] + cy) * dim[2] + cz);
        for (uint32_t k = cstart[cc]; k < cstart[cc + 1]; k++) {
          size_t j = items[k];
          if (j == i) { if (include_self) out.push_back((int32_t)j); continue; }
          // the reference evaluates ps[min]-ps[max]; the square is symmetric
          V d = (i < j) ? (x[i] - x[j]) : (x[j] - x[i]);
          if (d.norm2() <= H2) out.push_back((int32_t)j);
        }
      }
      std::sort(out.begin(), out.end());
    };
    #pragma omp parallel
    {
      std::vector<int32_t> tmp;
      #pragma omp for schedule(dynamic, 256)
      for (long i = 0; i < (long)n; i++) { visit((size_t)i, tmp); cnt[i + 1] = (uint32_t)tmp.size(); }
    }
    row_ptr[0] = 0;
    for (size_t i = 0; i < n; i++) row_ptr[i + 1] = row_ptr[i] + cnt[i + 1];
    col.resize(row_ptr[n]);
    #pragma omp parallel
    {
      std::vector<int32_t> tmp;
      #pragma omp for schedule(dynamic, 256)
      for (long i = 0; i < (long)n; i++) {
        visit((size_t)i, tmp);
        std::copy(tmp.begin(), tmp.end(), col.begin() + row_ptr[i]);
      }
    }
  }

  // ---- one time step: particles.cpp:250-297 -----------------------------------------------------
  void step() {
    sim_time += double(DT);
    // A: applyForceVelocity (175-183)
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
      vel[i].y -= R(-GRAV) * DT;          // velocity.y -= 10 * delta_t   (GRAV = -10)
      npos[i] = pos[i];
      collide(npos[i], vel[i] * DT, true);
      xpred[i] = npos[i];
    }
    // B: frozen neighbour lists on predicted positions (258-265)
    build_neighbors(npos, false);
    // D: Newton iterations (271-284)
    std::vector<V> snap(n);
    for (int it = 0; it < ITERS; it++) {
      // D1 newtonStepCalculateLambda (185-204)
      #pragma omp parallel for schedule(dynamic, 256)
      for (long i = 0; i < (long)n; i++) {
        R density = R(0), denom = R(0);
        V grad_i;
        for (uint32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
          V r = npos[i] - npos[col[k]];
          R w = poly6(r);
          V g = grad_spiky(r);
          density += w;
          g /= RHO0;
          grad_i += g;
          denom += g.norm2();
        }
        R c_i = density / RHO0 - R(1);
        denom += grad_i.norm2();
        dens[i] = density;
        lam[i] = -c_i / (denom + EPS_RELAX);
      }
      if (it == 0) {
        double ds = 0; for (size_t i = 0; i < n; i++) ds += double(dens[i]);
        avg_rho_first = n ? ds / double(n) : 0.0;
      }
      // D2 newtonStepUpdatePosition (206-213).  The reference reads per-neighbour caches
      // (s_corr, grad_w) filled in D1 from the pre-update positions; recomputing them from a
      // snapshot of those positions gives the same bits without 28 B/pair of cache.
      snap = npos;
      #pragma omp parallel for schedule(dynamic, 256)
      for (long i = 0; i < (long)n; i++) {
        V dp;
        for (uint32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
          int j = col[k];
          V r = snap[i] - snap[j];
          R w = poly6(r);
          R s_corr = -KCORR * intpow(w * TSCALE, NCORR);
          V g = grad_spiky(r);
          dp += (lam[i] + lam[j] + s_corr) * g;
        }
        dp /= RHO0;
        collide(npos[i], dp, false);
      }
    }
    // E: updateVelocity + calculateVorticityApplyXSPHViscosity (215-234, loop 285-288)
    std::vector<V> gradw;   // not cached: recomputed in F from the same positions
    if (P.xsph_mode == PBF_XSPH_REFERENCE_ORDER) {
      for (size_t i = 0; i < n; i++) {   // strictly sequential (quirk Q11)
        vel[i] = (npos[i] - pos[i]) / DT;
        pass_e(i, vel, vel);
      }
    } else {
      std::vector<V> vnew(n);
      #pragma omp parallel for schedule(static)
      for (long i = 0; i < (long)n; i++) vel[i] = (npos[i] - pos[i]) / DT;
      vnew = vel;
      #pragma omp parallel for schedule(dynamic, 256)
      for (long i = 0; i < (long)n; i++) pass_e((size_t)i, vel, vnew);
      vel.swap(vnew);
    }
    // F: applyVorticity + updatePosition (236-248, loop 290-294)
    #pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; i++) {
      if (P.enable_vorticity) {
        V gv;
        for (uint32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
          int j = col[k];
          gv += vort[j].norm() * grad_spiky(npos[i] - npos[j]);
        }
        if (gv.norm() > EPS_D) vel[i] += (DT * VORT_EPS) * cross(gv.unit(), vort[i]);
      }
    }
    double ds = 0;
    for (size_t i = 0; i < n; i++) { ds += double(dens[i]); pos[i] = npos[i]; }
    avg_rho_final = n ? ds / double(n) : 0.0;
  }

  // body of calculateVorticityApplyXSPHViscosity for particle i: reads vin, writes vout[i]
  void pass_e(size_t i, const std::vector<V>& vin, std::vector<V>& vout) {
    V w_i, visc;
    R density = R(0);
    const V vi = vin[i];
    for (uint32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
      int j = col[k];
      V v_ij = vin[j] - vi;
      V r = npos[i] - npos[j];
      V g = grad_spiky(r);
      w_i += cross(v_ij, g);
      R w = poly6(r);
      visc += v_ij * w;
      density += w;
    }
    vort[i] = w_i;
    dens[i] = density;
    V out = vi;
    if (P.enable_xsph) out += VISC * visc;
    vout[i] = out;
  }
};

}  // namespace pbf_oracle
