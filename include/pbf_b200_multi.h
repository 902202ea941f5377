/* pbf_b200_multi.h — the PBF step on several GPUs of one box from ONE host process (C ABI).
 *
 * SURVEY.md §8(b) asks for `pbf_create(params, n_devices, device_ids, out)`: the reference calls the step from C++
 * (pathtracer.cpp:464-469 -> Particles::timeStep, particles.cpp:299-301) through one object, `struct Particles`
 * (particles.h:105-140), so the multi-GPU path has to sit behind the same kind of handle.  pbf_multi is that handle:
 * it owns one slab handle (pbf_b200_slab.h) per device, plans the x-slabs from the particles' cell columns, connects the
 * slabs in peer mode (messages and per-iteration boundary values are stored straight into the neighbours' memory over
 * NVLink, hand-overs are flag words: no host synchronisation inside a step, no copy engine, no communication library,
 * no Python) and re-balances the slabs while the fluid flows.  An N-device run reproduces the 1-device run bit for
 * bit (same global cell grid, same in-cell order, same summation order).
 *
 *   Particles::Particles(rho0)         -> pbf_create_multi         particles.h:114-116
 *   addParticle x N                    -> pbf_multi_upload         particles.h:118-120
 *   estimateDensities()                -> pbf_multi_estimate_densities   particles.cpp:440-444
 *   timeStep()                         -> pbf_multi_step           particles.cpp:250-301
 *   ps[i]->getPosition() ...           -> pbf_multi_download       particles.h:21,36-42
 *   "avg rho: a => b"                  -> pbf_multi_stats          particles.cpp:267,279,295
 * Host vectors are AoS xyz doubles in ORIGINAL particle order, as in pbf_b200.h.  Not thread-safe per handle.
 */
#ifndef PBF_B200_MULTI_H
#define PBF_B200_MULTI_H

#include "pbf_b200_slab.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pbf_multi pbf_multi;

/* device_ids == NULL: devices 0 .. n_devices-1.  n_devices == 1 is allowed (one slab, no neighbours), and an id may
 * appear more than once (several slabs sharing one GPU: no speed-up, but the whole protocol runs on a 1-GPU box). */
int  pbf_create_multi(const PbfParams* params, int n_devices, const int* device_ids, pbf_multi** out);
void pbf_multi_destroy(pbf_multi* m);
const char* pbf_multi_last_error(pbf_multi* m);
int  pbf_multi_num_devices(pbf_multi* m);

/* Global scene data, replicated on every device (pbf_set_obstacle_spheres / _triangles). */
int  pbf_multi_set_obstacle_spheres(pbf_multi* m, size_t count, const double* cx_cy_cz_r);
int  pbf_multi_set_obstacle_triangles(pbf_multi* m, size_t count, const double* p1_p2_p3_n1_n2_n3);

/* Plans the slabs (equal particle counts per device from the histogram of cell columns), distributes the particles
 * and connects the slabs.  A later upload re-plans from scratch. */
int  pbf_multi_upload(pbf_multi* m, size_t n, const double* pos_xyz, const double* vel_xyz);
int  pbf_multi_step(pbf_multi* m, int n_steps);          /* asynchronous on every device */
int  pbf_multi_estimate_densities(pbf_multi* m);         /* Particles::estimateDensities (particles.cpp:440-444); asynchronous */
int  pbf_multi_sync(pbf_multi* m);                       /* waits for all devices; reports deferred device-side errors */
int  pbf_multi_download(pbf_multi* m, double* pos_xyz, double* vel_xyz, double* density);   /* original order; syncs */
size_t pbf_multi_num_particles(pbf_multi* m);
/* averages over ALL particles (the two numbers the reference prints) and the slowest device's time of the last call */
int  pbf_multi_stats(pbf_multi* m, double* avg_rho_first_iter, double* avg_rho_final, double* last_call_ms);

/* Re-balancing.  Every `every_k_steps` steps the per-column particle counts the devices leave behind
 * are read (without stalling the devices); when max/mean of the owned counts exceeds `threshold` the slab boundaries
 * move towards the equal-count partition, by at most what one migration message can carry per step.  Defaults: 8 steps,
 * 1.05.  every_k_steps = 0 switches it off. */
int  pbf_multi_set_rebalance(pbf_multi* m, int every_k_steps, double threshold);
/* Current plan: n_devices + 1 column boundaries, owned particle counts per device (syncs), number of re-balancing moves so far */
int  pbf_multi_plan(pbf_multi* m, int* col_bounds_out, uint64_t* owned_out, uint64_t* n_rebalances_out);

/* parity: digest / size of every particle's frozen neighbour set of the last step, original order (pbf_b200.h) */
int  pbf_multi_neighbor_digest(pbf_multi* m, uint64_t* per_particle_digest, uint32_t* per_particle_count);
uint64_t pbf_multi_launch_count(pbf_multi* m);

#ifdef __cplusplus
}
#endif
#endif /* PBF_B200_MULTI_H */
