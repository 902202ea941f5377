/* pbf_b200.h — C ABI of the B200-native Position Based Fluids step.
 *
 * This is the drop-in boundary for the reference's per-timestep hot path.  The reference
 * (SsnL/Fluid) has no FFI layer; the seam is `struct Particles` (src/particles.h:105-140):
 *   Particles::Particles(rest_density)      -> pbf_create        (particles.h:114-116)
 *   Particles::addParticle(pos, v) x N      -> pbf_upload        (particles.h:118-120)
 *   Particles::timeStep() / timeStep(dt)    -> pbf_step          (particles.cpp:250-301)
 *   ps[i]->getPosition()/velocity/
 *        getLatestDensityEstimate()         -> pbf_download      (particles.h:21,36-42)
 *   cout "avg rho: a => b" per step         -> pbf_stats         (particles.cpp:267,279,295)
 *   Particles::estimateDensities()          -> pbf_estimate_densities (particles.cpp:440-444)
 * INTEGRATION.md shows the ~30-line patch a maintainer of the reference would apply.
 *
 * Conventions: plain C, opaque handle, int status codes (0 = PBF_OK), no exceptions cross the
 * boundary, no global state, one handle = one CUDA device + one stream, a handle is not
 * thread-safe, the caller owns every host buffer.  Host-side vectors are AoS xyz doubles in
 * ORIGINAL particle order (the order of addParticle calls), which is what Particles holds.
 * All device arithmetic is fp32 (see DESIGN.md for the parity contract).
 */
#ifndef PBF_B200_H
#define PBF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pbf_handle pbf_handle;

enum {
  PBF_OK = 0,
  PBF_ERR_INVALID = 1,       /* bad argument / call order                                  */
  PBF_ERR_CUDA = 2,          /* a CUDA runtime call failed; see pbf_last_error              */
  PBF_ERR_NO_DEVICE = 3,     /* no usable CUDA device: there is NO CPU fallback             */
  PBF_ERR_CAPACITY = 4,      /* neighbour / cell / halo capacity exceeded (never truncates) */
  PBF_ERR_DOMAIN = 5         /* non-finite particle state, or migration beyond one slab     */
};

/* XSPH ordering.  The reference applies updateVelocity + XSPH fused per
 * particle in index order (particles.cpp:285-288, quirk Q11); the parallel path is Jacobi.
 * REFERENCE_ORDER is a single-GPU validation mode: it reproduces the reference's velocities by
 * fixed-point sweeps over the lower-triangular dependency (~10x the cost of the finalize pass). */
enum { PBF_XSPH_JACOBI = 0, PBF_XSPH_REFERENCE_ORDER = 1 };

/* Runtime parameters.  Defaults (pbf_default_params) are the reference's private macros,
 * particles.cpp:10-44, and the hard-coded Cornell box of clamp()/clamp_response()
 * (particles.cpp:59-83,96-131). */
typedef struct PbfParams {
  double h;             /* SPH support radius H                 0.3    particles.cpp:26      */
  double dt;            /* DEFAULT_DELTA_T                      0.016  particles.cpp:24      */
  double rest_density;  /* rho0 (from XML <density>, via stof)  1000   particles.h:114       */
  double eps_relax;     /* EPSILON in lambda denominator        2      particles.cpp:36      */
  double k_corr;        /* K, artificial pressure coefficient   1e-4   particles.cpp:38      */
  double dq_ratio;      /* |dq| / H of the tensile term         0.1    particles.cpp:151     */
  double visc_c;        /* C, XSPH viscosity                    1e-3   particles.cpp:44      */
  double vort_eps;      /* VORTICITY_EPSILON                    1e-3   particles.cpp:42      */
  double gravity_y;     /* velocity.y += gravity_y*dt           -10    particles.cpp:180     */
  int32_t n_corr;       /* N, artificial pressure exponent      4      particles.cpp:40      */
  int32_t iterations;   /* NEWTON_NUM_STEPS                     12     particles.cpp:34      */
  double box_min[3];    /* collision box, hard clamp lower      {-1,0,-1}   particles.cpp:81-83 */
  double box_max[3];    /* hard clamp upper                     {1,1.49,1}  particles.cpp:81-83 */
  double y_light;       /* virtual plane y (light)              1.49   particles.cpp:69,106  */
  double z_front;       /* virtual plane z (open front)         1.0    particles.cpp:60,97   */
  int32_t xsph_mode;    /* PBF_XSPH_*; JACOBI is the performance path                        */
  int32_t enable_vorticity; /* 1: vorticity confinement (particles.cpp:236-244)              */
  int32_t enable_xsph;      /* 1: XSPH viscosity       (particles.cpp:229,233)               */
  int32_t reserved;
} PbfParams;

void pbf_default_params(PbfParams* p);

/* ---- lifecycle ------------------------------------------------------------------------- */
int  pbf_create(const PbfParams* params, int device_id, pbf_handle** out);
void pbf_destroy(pbf_handle* h);
const char* pbf_last_error(pbf_handle* h);      /* valid until the next call on h            */
int  pbf_device_count(void);                    /* 0 when no CUDA device is visible          */

/* ---- collision scene beyond the box ------------------------------------------------------- */
/* Obstacle spheres, rows (cx, cy, cz, r): the StaticScene::Sphere primitives of the BVH the reference collides
 * against (Particles::bvh, particles.cpp:76,112-122; sphere.cpp:10-76), e.g. the two r = 0.3 spheres of the
 * CBspheres scenes.  Replaces the previous set; count 0 removes them; at most PBF_MAX_SPHERES.  Takes effect
 * from the next pbf_step.  Rules: nearest hit of walls and spheres, a sphere blocks only motion
 * into it, one slide along the tangent in the predict pass. */
#define PBF_MAX_SPHERES 8
int  pbf_set_obstacle_spheres(pbf_handle* h, size_t count, const double* cx_cy_cz_r);
/* Obstacle triangles, 18 doubles each (p1, p2, p3, then the vertex normals n1, n2, n3): the triangle primitives
 * (StaticScene::Triangle / MarchingTriangle, triangle.cpp:21-80, marching_triangle.cpp:21-73) of the BVH the
 * reference collides against (bvh.cpp:48-192).  A bounding-volume hierarchy over them is built on the host here and
 * walked on the device; its result is, bit for bit, that of testing every triangle in index order.  Rules (DESIGN.md
 * §2): a triangle blocks only motion against its oriented normal (orientation = the side the vertex normals point
 * to), nearest hit of walls, spheres and triangles (equally near: the larger index), one slide along the plane
 * given by the interpolated vertex normals in the predict pass.  Replaces the previous set; count 0 removes it. */
#define PBF_MAX_TRIANGLES (1u << 22)
int  pbf_set_obstacle_triangles(pbf_handle* h, size_t count, const double* p1_p2_p3_n1_n2_n3);

/* ---- state in / out (host buffers, original order, doubles) ----------------------------- */
int  pbf_upload(pbf_handle* h, size_t n, const double* pos_xyz, const double* vel_xyz);
int  pbf_download(pbf_handle* h, double* pos_xyz, double* vel_xyz, double* density); /* syncs; any pointer may be NULL */
size_t pbf_num_particles(pbf_handle* h);
/* Optional: page-lock caller-owned host buffers so upload/download DMA them directly (fp64 on the
 * wire, conversion on the device).  Unregister, or destroy the handle, before freeing them. */
int  pbf_host_register(pbf_handle* h, void* ptr, size_t bytes);
int  pbf_host_unregister(pbf_handle* h, void* ptr);
/* Streaming read-back into page-locked buffers: every pbf_step() then sends the state after its last
 * step (what pbf_download would return) on a second stream as soon as each array is final, hidden
 * behind the finalize kernels; pbf_sync() completes it.  NULLs switch it off. */
int  pbf_set_readback(pbf_handle* h, double* pos_xyz, double* vel_xyz, double* density);

/* ---- the hot path ------------------------------------------------------------------------ */
int  pbf_step(pbf_handle* h, int n_steps);      /* enqueues n_steps * timeStep(dt); asynchronous */
int  pbf_sync(pbf_handle* h);                   /* waits; reports deferred device-side errors    */
int  pbf_estimate_densities(pbf_handle* h);     /* load-time density incl. self (Q1, a15)        */
/* avg density after the first lambda pass and after the finalize pass of the LAST step
 * (the two numbers the reference prints), and the device time of the last pbf_step call. */
int  pbf_stats(pbf_handle* h, double* avg_rho_first_iter, double* avg_rho_final, double* last_call_ms);

/* Neighbour-count alert = the cerr warning of Particle::initializeWithNewNeighbors (particles.cpp:165-173: every
 * particle with fewer than NUM_NEIGHBOR_ALERT_THRESHOLD = 18 neighbours, particles.cpp:32, is reported with its
 * predicted position and velocity).  With a threshold > 0 every step compacts those particles on the device (one
 * small kernel over the counts the neighbour build leaves behind); pbf_get_neighbor_alerts returns the records of the
 * LAST step sorted by particle id (the reference's order; *n_written of them) and the full count in *n_total.  When more than
 * max_records particles qualify, the records are those with the smallest ids, cut at a multiple of
 * 2^ceil(log2(n / 4096)) ids so that at most max_records remain: a deterministic prefix of the reference's output.  Any output pointer may be NULL.  threshold 0 switches it off (the default).  Single GPU. */
int  pbf_set_neighbor_alert(pbf_handle* h, int threshold, size_t max_records);
int  pbf_get_neighbor_alerts(pbf_handle* h, size_t cap, uint32_t* ids, uint32_t* counts, double* xpred_xyz, double* vel_xyz,
                             size_t* n_written, size_t* n_total);

/* SPH density at m arbitrary points from the committed positions = Particles::estimateDensityAt
 * (particles.cpp:446-453), the scalar field the marching-cubes surfacer samples
 * (particles.cpp:350-418).  Host fp64 AoS in, fp64 out; single GPU. */
int  pbf_density_at(pbf_handle* h, size_t m, const double* query_xyz, double* density_out);

/* Marching-cubes surface of the committed positions = Particles::getSurfacePrims (particles.cpp:352-391, called by
 * updateSurface 393-402 with isolevel 0.95 rho0 and step 0.5 H) with marching.cpp's polygonise / vertexInterp and
 * getVertexNormal (particles.cpp:407-418; grad_eps 0.001): lattice cells over [lo, hi] in ix / iy / iz order, the
 * last cell of every axis clipped to hi.  Writes 18 doubles per triangle (p1 p2 p3 n1 n2 n3 = the arguments of the
 * reference's MarchingTriangle constructor) for the first min(*n_triangles, cap_triangles) triangles in the
 * reference's order; *n_triangles = the full count (call with cap 0 to size the buffer).  Everything runs on the
 * device in fp64 on the fp32 state, every density term with the reference's operations; only the order of the
 * sum differs, so the soup matches the reference's on the same state to ~1e-13.  Single GPU. */
int  pbf_extract_surface(pbf_handle* h, const double lo[3], const double hi[3], double isolevel, double step, double grad_eps,
                         size_t cap_triangles, double* tris_out, size_t* n_triangles);

/* ---- device-resident I/O (bench "value" leg; fp32 xyz AoS device pointers, original order) */
int  pbf_upload_device(pbf_handle* h, size_t n, const float* d_pos_xyz, const float* d_vel_xyz);
int  pbf_download_device(pbf_handle* h, float* d_pos_xyz, float* d_vel_xyz, float* d_density);

/* ---- parity / debug ---------------------------------------------------------------------- */
/* Order-independent 64-bit digest and size of each particle's frozen neighbour set of the last
 * step, in ORIGINAL indices: sum over neighbours j of mix64(j), mix64 = splitmix64 finaliser of
 * (j + 0x9E3779B97F4A7C15). */
int  pbf_debug_neighbor_digest(pbf_handle* h, uint64_t* per_particle_digest, uint32_t* per_particle_count);
/* CSR of the frozen neighbour sets, original indices, ascending within a row (small N). */
int  pbf_debug_download_neighbors(pbf_handle* h, uint32_t* row_ptr /*n+1*/, uint32_t* col_idx, size_t col_cap);
/* Intermediate per-particle arrays of the last step, original order, as doubles. */
enum { PBF_ARRAY_XSTAR = 0 /*n*3 predicted/corrected positions*/, PBF_ARRAY_LAMBDA = 1 /*n*/,
       PBF_ARRAY_VORTICITY = 2 /*n*3*/, PBF_ARRAY_XPRED = 3 /*n*3 x* right after predict+collide*/ };
int  pbf_debug_download_array(pbf_handle* h, int which, double* out);
/* Keep a copy of x* right after predict+collide (PBF_ARRAY_XPRED) during subsequent steps. */
int  pbf_debug_capture(pbf_handle* h, int on);
/* development probe of the gather kernels on the lists of the last step (fluid_b200/csrc/pbf_probe.inl): average ms per
 * launch over `reps` launches of probe `variant`; out4 (n x 4 floats, sorted order) optional */
int  pbf_debug_probe(pbf_handle* h, int variant, int reps, double* ms_out, float* out4);
/* Host only (no device): the bounding-volume hierarchy pbf_set_obstacle_triangles builds over these triangles, without the
 * safety margin on the node boxes.  8 floats per node: lo.xyz, a, hi.xyz, b with the integers a, b stored as bits; a leaf
 * (b > 0) holds triangles [a, a + b) of the leaf order, an inner node (b = 0) has its children at a and a + 1.
 * order_out maps the leaf order to original triangle indices.  PBF_ERR_CAPACITY when cap_nodes is too small. */
int  pbf_debug_build_bvh(size_t count, const double* p1_p2_p3_n1_n2_n3, float* nodes_out, size_t cap_nodes, uint32_t* order_out,
                         size_t* n_nodes, int* depth);
/* Number of kernel launches issued by this handle so far (bench "gpu_launches"). */
uint64_t pbf_launch_count(pbf_handle* h);
/* Per-kernel device time (CUDA events) accumulated since the last reset; names are static.
 * Enabling profiling serialises the stream with events; off by default. */
int  pbf_profile_enable(pbf_handle* h, int on);
int  pbf_profile_get(pbf_handle* h, int max, const char** names, double* total_ms, uint64_t* launches);

/* ---- slab decomposition (one process per GPU; x-slabs) ---------------------
 * See fluid_b200/slab.py for the protocol; declared in pbf_b200_slab.h. */

#ifdef __cplusplus
}
#endif
#endif /* PBF_B200_H */
