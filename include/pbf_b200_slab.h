/* pbf_b200_slab.h — C ABI of the x-slab decomposition (one process per GPU).
 *
 * The reference is a single process; this part of the boundary has no reference counterpart.  It
 * splits Particles::timeStep (particles.cpp:250-297) into phases so that a host-side driver
 * (fluid_b200/slab.py, torch.distributed over NCCL) can exchange particles between x-adjacent
 * ranks in between.  The library itself never communicates: it fills / consumes device message
 * buffers whose addresses pbf_slab_buffer() hands out.
 *
 * Geometry.  The cell grid is GLOBAL (same origin and cell edge on every rank).  Rank r owns the
 * cell columns [gx_lo, gx_hi) and stores one ghost column on each side.  After the per-step sort
 * the arrays are ordered (column, cy, cz, global id), so with x slowest
 *      [0,b0) left ghosts | [b0,b3) owned | [b3,n) right ghosts
 * and the owner's boundary column [b0,b1) (resp. [b2,b3)) has the SAME order as the neighbour's
 * ghost range: per-iteration halo refreshes are plain contiguous copies, no pack / unpack.
 * Every floating-point sum runs in the same order as on one GPU, so an N-slab run reproduces the
 * single-GPU run bit for bit.
 *
 * One step (driver's view):
 *   pbf_slab_phase_predict   predict + collide owned particles; emigrants -> MIG_SEND_L/R
 *   <exchange MIG_SEND_* -> neighbour's MIG_RECV_*>                       (fixed-size messages)
 *   pbf_slab_phase_migrate   append immigrants; boundary columns -> GHOST_SEND_L/R
 *   <exchange GHOST_SEND_* -> GHOST_RECV_*>
 *   pbf_slab_phase_sort      append ghosts, counting sort, neighbour lists; returns b0..b3,n (syncs)
 *   iterations x { pbf_slab_phase(LAMBDA) ; <copy XS_B[b0,b1) -> left's XS_B[b3',n'), XS_B[b2,b3) -> right's XS_B[0,b0')> ;
 *                  pbf_slab_phase(DELTA)  ; <same for XS_A> }
 *   pbf_slab_phase(VELOCITY) ; pbf_slab_phase(VORTICITY) ; <same copy for XS_W = (x*,|omega|)> ; pbf_slab_phase(CONFINE)
 */
#ifndef PBF_B200_SLAB_H
#define PBF_B200_SLAB_H

#include "pbf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Cells per axis of the global grid for these parameters (cell edge = h * (1 + 2^-8)). */
int pbf_grid_dims(const PbfParams* params, int dims_out[3]);
/* Global cell column of x positions (fp32 arithmetic identical to the device's). */
int pbf_cell_columns(const PbfParams* params, size_t n, const double* pos_xyz, int32_t* column_out);

/* Run all work of this handle on the caller's CUDA stream (e.g. torch's current stream). */
int pbf_set_stream(pbf_handle* h, void* cuda_stream);

/* This handle owns global cell columns [gx_lo, gx_hi); left_cols / right_cols = number of columns
 * owned by the x-neighbours (0 = no neighbour on that side).  particle_cap bounds owned + ghost
 * particles; halo_cap bounds each migration / ghost message (particles).  Call before uploading. */
int pbf_slab_configure(pbf_handle* h, int gx_lo, int gx_hi, int left_cols, int right_cols,
                       size_t particle_cap, size_t halo_cap);
/* Same, with room for re-balancing: the cell arrays are sized for max_cols owned columns (0 = gx_hi - gx_lo), so
 * pbf_slab_set_columns may later move the slab's boundaries within that width. */
int pbf_slab_configure_ex(pbf_handle* h, int gx_lo, int gx_hi, int left_cols, int right_cols,
                          size_t particle_cap, size_t halo_cap, int max_cols);
/* Re-balancing ("re-balanced only when imbalance > few %"): change the owned columns between two steps.
 * Every rank must be given consistent ranges.  Nothing moves here: the next predict pass finds every particle whose
 * column changed hands outside [gx_lo, gx_hi) and sends it through the ordinary migration message, so a boundary may
 * move by at most as many columns as one message holds (halo_cap particles) and only to the adjacent slab. */
int pbf_slab_set_columns(pbf_handle* h, int gx_lo, int gx_hi, int left_cols, int right_cols);
int pbf_slab_columns(pbf_handle* h, int gx_lo_hi_left_right_out[4]);
/* Owned particles per GLOBAL cell column (n_cols = pbf_grid_dims()[0]) after the sort of a recent step.  The sort phase
 * enqueues the copy whenever the previous one has been delivered, so with wait = 0 this never stalls the pipeline
 * (PBF_ERR_INVALID while a copy is in flight); wait = 1 blocks until the pending copy has arrived.  *step_out = how many
 * steps the handle had completed before the step whose sort the histogram describes. */
int pbf_slab_column_histogram(pbf_handle* h, uint32_t* hist_out, size_t n_cols, int wait, long long* step_out);
/* every_k_steps > 0: record the histogram at the sorts of steps 0, k, 2k, ... instead (the caller reads each one, with
 * wait = 1, before it enqueues the step that records the next): the schedule of a re-balancing driver then depends on
 * the step count only, not on how far the host runs ahead of the devices.  0 restores the default above. */
int pbf_slab_set_histogram_interval(pbf_handle* h, int every_k_steps);

/* ---- peer mode: the exchanges inside the library, over peer-mapped memory -------------------------------------------
 * After pbf_slab_p2p_connect_* a slab handle writes its migration / ghost messages and, from inside the solver kernels,
 * the per-iteration boundary values DIRECTLY into its x-neighbours' device memory (NVLink peer stores), learns the
 * neighbours' ranges from their device-resident link block, and hands over with flag words instead of host
 * synchronisation: pbf_slab_step_p2p enqueues whole steps with no host round trip, no copy-engine work and no
 * communication library.  Every rank must enqueue the same number of steps.  Results are bit-identical to the
 * host-driven protocol and to one GPU.
 *   same process (pbf_create_multi):   pbf_slab_p2p_connect_local(h, left_handle, right_handle)
 *   one process per GPU:               pbf_slab_p2p_export(h, blob) on every rank, exchange the blobs (any transport),
 *                                      pbf_slab_p2p_connect_ipc(h, left_blob, right_blob)            (CUDA IPC)
 * NULL = no neighbour on that side.  A neighbour that never arrives at an exchange point makes the waiting rank give up
 * after pbf_slab_set_wait_timeout seconds (default 20) with an error at the next pbf_sync, never a hang. */
int pbf_slab_p2p_connect_local(pbf_handle* h, pbf_handle* left, pbf_handle* right);
size_t pbf_slab_p2p_blob_size(void);
int pbf_slab_p2p_export(pbf_handle* h, void* blob_out);
int pbf_slab_p2p_connect_ipc(pbf_handle* h, const void* left_blob, const void* right_blob);
int pbf_slab_p2p_disconnect(pbf_handle* h);                           /* back to the host-driven phases (unmaps the neighbours) */
int pbf_slab_set_wait_timeout(pbf_handle* h, double seconds);
int pbf_slab_step_p2p(pbf_handle* h, int n_steps);                    /* asynchronous */
/* Particles::estimateDensities (particles.cpp:440-444) across the slabs: densities of the committed positions, self included.
 * Every rank calls it; asynchronous. */
int pbf_slab_estimate_densities_p2p(pbf_handle* h);
/* Synchronises and returns b0, b1, b2, b3, n of the last sort (peer mode keeps them on the device during the step). */
int pbf_slab_refresh_ranges(pbf_handle* h, uint32_t bounds_out[5]);

/* Owned particles of this rank with their GLOBAL ids (host, fp64 AoS). */
int pbf_slab_upload(pbf_handle* h, size_t n, const double* pos_xyz, const double* vel_xyz, const uint32_t* ids);
/* Owned particles, in the rank's current sorted order (fp64 AoS) + ids; *n_out = count. */
int pbf_slab_download(pbf_handle* h, size_t cap, double* pos_xyz, double* vel_xyz, double* density, uint32_t* ids, size_t* n_out);
/* Neighbour digests of the owned particles, same order as pbf_slab_download. */
int pbf_slab_neighbor_digest(pbf_handle* h, size_t cap, uint64_t* digest, uint32_t* count);

int pbf_slab_phase_predict(pbf_handle* h);
int pbf_slab_phase_migrate(pbf_handle* h);
int pbf_slab_phase_sort(pbf_handle* h, uint32_t bounds_out[5]);   /* b0, b1, b2, b3, n ; synchronises */
enum { PBF_PHASE_LAMBDA_FIRST = 0, PBF_PHASE_LAMBDA = 1, PBF_PHASE_DELTA = 2, PBF_PHASE_VELOCITY = 3,
       PBF_PHASE_VORTICITY = 4, PBF_PHASE_CONFINE = 5 };
int pbf_slab_phase(pbf_handle* h, int phase);
/* Same, restricted to a part of the owned range, so a driver can overlap the halo exchange of a
 * pass with the bulk of its compute: BOUNDARY = the first and last owned cell columns (rounded
 * outwards to whole 32-particle slices) = exactly what the x-neighbours receive as ghosts;
 * INTERIOR = the rest.  LAMBDA*, DELTA and VORTICITY honour `part`; the other phases ignore it. */
enum { PBF_PART_ALL = 0, PBF_PART_BOUNDARY = 1, PBF_PART_INTERIOR = 2 };
int pbf_slab_phase_part(pbf_handle* h, int phase, int part);
/* sums over the owned particles of the last step: density after the first lambda pass / final, count */
int pbf_slab_stats(pbf_handle* h, double* rho_first_sum, double* rho_final_sum, uint64_t* n_owned);

/* Device buffers.  Message buffers are float4 arrays: element 0 is a header (count in .x as uint32
 * bits), then 3 float4 per migrant (x|id, x*, v) or 2 per ghost (x*|id, x). */
enum { PBF_BUF_MIG_SEND_L = 0, PBF_BUF_MIG_SEND_R = 1, PBF_BUF_MIG_RECV_L = 2, PBF_BUF_MIG_RECV_R = 3,
       PBF_BUF_GHOST_SEND_L = 4, PBF_BUF_GHOST_SEND_R = 5, PBF_BUF_GHOST_RECV_L = 6, PBF_BUF_GHOST_RECV_R = 7,
       PBF_BUF_XS_A = 8, PBF_BUF_XS_B = 9, PBF_BUF_OMEGA = 10,
       PBF_BUF_XS_W = 11 /* (x*, |omega|) written by the vorticity pass, gathered by confinement */ };
void* pbf_slab_buffer(pbf_handle* h, int which, size_t* bytes_out);

#ifdef __cplusplus
}
#endif
#endif /* PBF_B200_SLAB_H */
