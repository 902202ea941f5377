// pbf_kernels.cu — hand-written sm_100a kernels of the PBF step and their launch sequence.
//
// Replaces Particles::timeStep (reference src/particles.cpp:250-297).  Pass structure:
//   predict+collide+hash -> counting sort by cell (hist / scan / scatter / canonical in-cell order /
//   reorder to cell-sorted float4 SoA) -> frozen neighbour lists (SELL-32, built once per step from
//   the predicted positions, like the reference's all-pairs loop 258-265) -> I x (lambda, delta-p +
//   collide) -> velocity, vorticity + XSPH + density -> confinement + commit.
// No tensor cores: this is a gather-bound stencil, not a contraction.
#include "pbf_internal.h"

namespace pbf {

static constexpr int TPB = 256;          // threads per block for per-particle kernels (512 measured slower: 79.4 vs 76.6 ms/step)
static constexpr int SCAN_ITEMS = 8;     // items per thread in the cell scan
static constexpr int SCAN_TILE = TPB * SCAN_ITEMS;
static constexpr uint32_t CELL_INVALID = 0xFFFFFFFFu;

__device__ __forceinline__ float3 xyz(const float4& v) { return make_float3(v.x, v.y, v.z); }

// Peer mode: the ranges of the sorted arrays live on the device (SlabLink::b), launches are sized for the capacity, and
// a pass that produces values its x-neighbours gather as ghosts stores them straight into the neighbours' arrays
// (peer-mapped pointers over NVLink): the owner's boundary column and the neighbour's ghost range have the same order
// (include/pbf_b200_slab.h), so entry t of the left boundary column goes to entry b3' + t of the left neighbour and
// entry k of the right boundary column to entry k of the right neighbour.  lk == nullptr: ranges by value, no push.
__device__ __forceinline__ void push_boundary(const SlabLink* __restrict__ lk, const PushArgs& pa, uint32_t t, float4 v) {
  if (lk == nullptr) return;
  const uint32_t cnt = lk->b[3] - lk->b[0], nl = lk->b[1] - lk->b[0], nr = lk->b[3] - lk->b[2];
  if (pa.dst[0] != nullptr && t < nl) pa.dst[0][lk->nb[0][3] + t] = v;
  if (pa.dst[1] != nullptr && t >= cnt - nr) pa.dst[1][t - (cnt - nr)] = v;
}

// ------------------------------------------------------------------------------------------------
// A. predict + collide + hash      (applyForceVelocity + clamp_response, particles.cpp:175-183,87-132)
// EXACT regime: x* is bit-identical to the fp32 oracle.
// ------------------------------------------------------------------------------------------------
// Slab mode: a particle whose predicted position leaves the owned cell columns [gx_lo, gx_hi) is an
// emigrant: it is packed (3 float4: x with the global id in .w, x*, v) into the message for the
// x-neighbour and dropped from the local sort (cell_of stays INVALID).  Single-GPU mode owns every
// column, so nothing ever emigrates.
template <bool SPH>
__global__ void __launch_bounds__(TPB)
k_predict_hash(const __grid_constant__ DevParams P, uint32_t i0, uint32_t n, const float4* __restrict__ pos,
               float4* __restrict__ vel, const uint32_t* __restrict__ orig, float4* __restrict__ xs_tmp,
               uint32_t* __restrict__ cell_of, uint32_t* __restrict__ rank, uint32_t* __restrict__ cell_count,
               int apply_forces, float4* __restrict__ mig_left, float4* __restrict__ mig_right, uint32_t mig_cap,
               Scalars* __restrict__ sc, const SlabLink* __restrict__ lk) {
  const uint32_t t = blockIdx.x * TPB + threadIdx.x;
  if (lk) { i0 = lk->b[0]; n = lk->b[3] - lk->b[0]; }            // owned range of the previous sort
  if (t >= n) return;
  const uint32_t i = i0 + t;
  const float4 x = pos[i];
  float3 p = xyz(x);
  float4 v = vel[i];
  // The reference's hard clamp (std::min/max, particles.cpp:129-131) silently turns a NaN coordinate into
  // the box corner; report non-finite state instead of hiding it.
  if (!(isfinite(x.x) && isfinite(x.y) && isfinite(x.z) && isfinite(v.x) && isfinite(v.y) && isfinite(v.z))) atomicOr(&sc->err, ERRBIT_NONFINITE);
  if (apply_forces) {
    v.y = __fsub_rn(v.y, P.gdt);                                   // velocity.y -= 10 * delta_t
    const float3 delta = make_float3(__fmul_rn(v.x, P.dt), __fmul_rn(v.y, P.dt), __fmul_rn(v.z, P.dt));
    p = ex_collide<SPH>(P, p, delta, true);
    vel[i] = v;                                                    // clamp_response leaves v untouched (Q9)
  }
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) { atomicOr(&sc->err, ERRBIT_NONFINITE); p = make_float3(P.clo[0], P.clo[1], P.clo[2]); }
  xs_tmp[i] = make_float4(p.x, p.y, p.z, 0.f);
  const int3 cg = cell_coords_global(P, p.x, p.y, p.z);
  if (cg.x < P.gx_lo || cg.x >= P.gx_hi) {
    const int side = cg.x < P.gx_lo ? 0 : 1;
    float4* dst = side ? mig_right : mig_left;
    // one hop only: the neighbour slab must own the landing column
    if (dst == nullptr || cg.x < P.gx_lo - P.hop_left || cg.x >= P.gx_hi + P.hop_right) { atomicOr(&sc->err, ERRBIT_MIGRATION); return; }
    const uint32_t slot = atomicAdd(&sc->counters[side], 1u);
    if (slot >= mig_cap) { atomicOr(&sc->err, ERRBIT_HALO_CAPACITY); return; }
    dst[1 + 3 * slot + 0] = make_float4(x.x, x.y, x.z, __uint_as_float(orig[i]));
    dst[1 + 3 * slot + 1] = make_float4(p.x, p.y, p.z, 0.f);
    dst[1 + 3 * slot + 2] = v;
    return;                                                        // cell_of[i] stays INVALID
  }
  const uint32_t c = cell_linear(P, make_int3(cg.x - P.cx_offset, cg.y, cg.z));
  cell_of[i] = c;
  rank[i] = atomicAdd(&cell_count[c], 1u);
}

// ------------------------------------------------------------------------------------------------
// B. exclusive scan of the cell histogram -> cell_start[0..ncell]   (3 launches)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t& block_total) {
  __shared__ uint32_t warp_sums[TPB / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t ws = (lane < TPB / 32) ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < TPB / 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += t; }
    if (lane < TPB / 32) warp_sums[lane] = ws;   // inclusive over warps
  }
  __syncthreads();
  const uint32_t warp_off = wid ? warp_sums[wid - 1] : 0u;
  block_total = warp_sums[TPB / 32 - 1];
  __syncthreads();
  return warp_off + inc - v;
}

__global__ void __launch_bounds__(TPB)
k_scan_reduce(uint32_t ncell, const uint32_t* __restrict__ count, uint32_t* __restrict__ block_sums) {
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < ncell) s += count[base + k];
  uint32_t total;
  block_exclusive_scan(s, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(TPB)
k_scan_block_sums(uint32_t nblocks, uint32_t* __restrict__ block_sums) {
  uint32_t running = 0;
  for (uint32_t base = 0; base < nblocks; base += TPB) {
    const uint32_t idx = base + threadIdx.x;
    const uint32_t v = idx < nblocks ? block_sums[idx] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, total);
    if (idx < nblocks) block_sums[idx] = running + ex;
    running += total;
  }
}

__global__ void __launch_bounds__(TPB)
k_scan_apply(uint32_t ncell, const uint32_t* __restrict__ count, const uint32_t* __restrict__ block_sums,
             uint32_t* __restrict__ cell_start) {
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS]; uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < ncell) ? count[base + k] : 0u; s += v[k]; }
  uint32_t total;
  uint32_t off = block_exclusive_scan(s, total) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < ncell) cell_start[base + k] = off; off += v[k]; }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == TPB - 1) cell_start[ncell] = off;   // == n
}

// ------------------------------------------------------------------------------------------------
// C. scatter to cell order, canonical order inside each cell (ascending original id), reorder
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_scatter(uint32_t n, const uint32_t* __restrict__ cell_of, const uint32_t* __restrict__ rank,
          const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ orig,
          uint32_t* __restrict__ perm, uint32_t* __restrict__ key) {
  const uint32_t i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = cell_of[i];
  if (c == CELL_INVALID) return;            // emigrant, stale ghost or unused append slot
  const uint32_t slot = cell_start[c] + rank[i];
  perm[slot] = i;
  key[slot] = orig[i];
}

// The atomic rank above is arrival order, i.e. not reproducible.  Sorting each cell's few particles
// by their original id makes the whole layout (and therefore every floating-point sum downstream)
// a pure function of the particle state: runs are bit-reproducible, and a slab decomposition sees
// the same order as one GPU does.
// One thread per (thin) cell.  The cell's (id, slot) pairs are loaded into registers as 64-bit words and sorted by a
// fully unrolled N-input bitonic network (compile-time indices, no memory traffic), N = 8 / 16 / 32 by occupancy
// (thin cells of a lattice at spacing h/3 hold 3-4 particles); unused inputs hold +inf.  Fuller cells fall back to
// an insertion sort in global memory.
template <int N>
__device__ __forceinline__ void cell_sort_network(uint32_t s, uint32_t cnt, uint32_t* __restrict__ perm, uint32_t* __restrict__ key) {
  unsigned long long v[N];
#pragma unroll
  for (int k = 0; k < N; k++)
    v[k] = (uint32_t)k < cnt ? ((unsigned long long)key[s + k] << 32) | perm[s + k] : ~0ull;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
      for (int i = 0; i < N; i++) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const unsigned long long a = v[i], b = v[l];
          const bool sw = (a > b) == up;
          v[i] = sw ? b : a; v[l] = sw ? a : b;
        }
      }
#pragma unroll
  for (int k = 0; k < N; k++)
    if ((uint32_t)k < cnt) { key[s + k] = (uint32_t)(v[k] >> 32); perm[s + k] = (uint32_t)v[k]; }
}

__global__ void __launch_bounds__(128)
k_cell_sort(uint32_t ncell, const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ perm,
            uint32_t* __restrict__ key) {
  const uint32_t c = blockIdx.x * 128 + threadIdx.x;
  if (c >= ncell) return;
  const uint32_t s = cell_start[c], e = cell_start[c + 1];
  const uint32_t cnt = e - s;
  if (cnt <= 1u) return;
  if (cnt <= 8u) { cell_sort_network<8>(s, cnt, perm, key); return; }
  if (cnt <= 16u) { cell_sort_network<16>(s, cnt, perm, key); return; }
  if (cnt <= 32u) { cell_sort_network<32>(s, cnt, perm, key); return; }
  for (uint32_t a = s + 1; a < e; a++) {
    const uint32_t k = key[a], p = perm[a];
    uint32_t b = a;
    while (b > s && key[b - 1] > k) { key[b] = key[b - 1]; perm[b] = perm[b - 1]; b--; }
    key[b] = k; perm[b] = p;
  }
}

__global__ void __launch_bounds__(TPB)
k_reorder(uint32_t n_upper, const uint32_t* __restrict__ n_sorted_ptr, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ key,
          const float4* __restrict__ pos_in, const float4* __restrict__ vel_in, const float4* __restrict__ xs_tmp,
          float4* __restrict__ pos_out, float4* __restrict__ vel_out, float4* __restrict__ xs_out,
          uint32_t* __restrict__ orig_out) {
  const uint32_t s = blockIdx.x * TPB + threadIdx.x;
  if (s >= n_upper || s >= *n_sorted_ptr) return;     // n_sorted = cell_start[ncell], known on the device
  const uint32_t i = perm[s];
  pos_out[s] = pos_in[i];
  vel_out[s] = vel_in[i];
  xs_out[s] = xs_tmp[i];
  orig_out[s] = key[s];
}

// ------------------------------------------------------------------------------------------------
// D. frozen neighbour lists  (all-pairs loop particles.cpp:258-265 -> 27-cell search, EXACT predicate)
// Layout SELL-32x4: the 32 particles of a warp form a slice; a row holds one uint4 (4 neighbour
// indices) per lane, so entry s of lane l lives at nbr[(slice_off + s/4)*128 + 4*l + s%4] and
// every later pass reads its indices as fully coalesced 16-byte loads.  Lists are padded to a
// multiple of 4 with the sentinel index n (a particle parked far outside the domain, so it adds
// exactly 0 to every sum).  No truncation, ever: overflow of the row pool is an error.
// ------------------------------------------------------------------------------------------------
// Single pass over the candidates: the predicate results are kept as bitmasks (one word per 32
// candidates of a z-run, parked in thread-local memory: one store per 32 candidates), the counts
// come from popc, rows are allocated with one atomic per slice, and the fill phase just walks
// the set bits.  Candidates are evaluated in unconditional groups of 8 (reads past the end of a
// run are in-bounds of the padded arrays and masked off), so the inner loop has no bounds checks.
static constexpr int NB_TOTW = 48;      // cached mask words per particle (a lattice at spacing h/3 needs about 20 with thin z-cells);
                                        // crowded neighbourhoods recompute the words beyond that in the fill phase

__device__ __forceinline__ uint32_t nb_eval_word(const float4* __restrict__ xs, float3 pi, float h2, uint32_t wb, uint32_t lim) {
  uint32_t m = 0;
  for (uint32_t g = 0; g < lim; g += 8) {
    const float4* src = xs + wb + g;
    uint32_t m8 = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const float4 q = __ldg(src + u);
      if (ex_is_neighbor(pi, make_float3(q.x, q.y, q.z), h2)) m8 |= (1u << u);   // predicated OR with an immediate
    }
    m |= m8 << g;
  }
  if (lim < 32u) m &= (1u << lim) - 1u;
  return m;
}

// Candidate range [b, e) of z-run k (k = 3*(dx+1) + (dy+1)) for a particle at pi in cell c, with
// conservative culling: the run is dropped when the (dx,dy) cell column is farther than h in the
// xy-plane, and its z-extent is trimmed to the thin cells that can still reach the particle.  `slack`
// covers the fp32 rounding of the cell assignment (a particle may sit a few ulp outside the
// nominal bounds of its cell), so no true neighbour is ever culled; the exact predicate decides.
__device__ __forceinline__ void nb_run_range(const DevParams& P, const uint32_t* __restrict__ cell_start, float3 pi, int3 c, int k,
                                             float cell, uint32_t& b, uint32_t& e) {
  b = 0; e = 0;
  const int ox = (k / 3) - 1, oy = (k % 3) - 1;
  const int cx = c.x + ox, cy = c.y + oy;
  if (cx < 0 || cx >= P.gdim[0] || cy < 0 || cy >= P.gdim[1]) return;
  const float slack = 1e-3f * P.h + 1e-6f * fmaxf(fabsf(pi.x), fmaxf(fabsf(pi.y), fabsf(pi.z)));
  // distance from the particle to the neighbouring column along x and y (0 for the own column)
  const float x_lo = P.gmin[0] + (float)(c.x + P.cx_offset) * cell, y_lo = P.gmin[1] + (float)c.y * cell;
  float dx = ox < 0 ? pi.x - x_lo : (ox > 0 ? (x_lo + cell) - pi.x : 0.f);
  float dy = oy < 0 ? pi.y - y_lo : (oy > 0 ? (y_lo + cell) - pi.y : 0.f);
  dx = fmaxf(dx - slack, 0.f); dy = fmaxf(dy - slack, 0.f);
  const float rz2 = P.h2 - (dx * dx + dy * dy);
  if (rz2 < 0.f) return;
  // z-extent of the particle's reach inside this column, in thin cells (monotonic in z, so every candidate with
  // |z_j - z_i| <= rz lies in [zlo, zhi]; the slack covers the roundings of rz and of the cell index)
  const float rz = sqrtf(rz2) + slack;
  const int zlo = max((int)floorf((pi.z - rz - P.gmin[2]) * P.inv_cell_z), 0);
  const int zhi = min((int)floorf((pi.z + rz - P.gmin[2]) * P.inv_cell_z), P.gdim[2] - 1);
  const uint32_t base = (uint32_t)((cx * P.gdim[1] + cy) * P.gdim[2]);
  b = cell_start[base + zlo];
  e = cell_start[base + zhi + 1];
}

__global__ void __launch_bounds__(TPB, 5)      // 48 registers: 5 CTAs per SM (50 would round up to 56 and lose one)
k_build_neighbors(const __grid_constant__ DevParams P, uint32_t i0, uint32_t cnt_range, uint32_t sentinel,
                  const float4* __restrict__ xs,
                  const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ nbr,
                  uint32_t* __restrict__ slice_off, uint32_t* __restrict__ nbr_cnt,
                  unsigned long long cap_rows, int include_self, Scalars* __restrict__ sc, int use_per_lane,
                  const SlabLink* __restrict__ lk) {
  if (lk) { i0 = lk->b[0]; cnt_range = lk->b[3] - lk->b[0]; sentinel = lk->b[4]; }
  if (blockIdx.x * TPB >= cnt_range) return;             // peer mode sizes the grid for the capacity: nothing to do, and no zero-row atomics on the cursor
  const uint32_t t = blockIdx.x * TPB + threadIdx.x;     // index inside the owned range
  const uint32_t i = i0 + t;                             // index in the cell-sorted arrays
  const int lane = threadIdx.x & 31;
  const bool valid = t < cnt_range;
  uint32_t masks[NB_TOTW], wbase[NB_TOTW];
  uint32_t cnt = 0, nwords = 0, cnt_cached = 0;
  float3 pi = make_float3(0.f, 0.f, 0.f);
  int3 c = make_int3(0, 0, 0);
  const float cell = 1.0f / P.inv_cell;
  if (valid) {
    pi = xyz(xs[i]);
    c = cell_coords(P, pi.x, pi.y, pi.z);
  }
  // Warp-uniform evaluation: when all particles of the warp sit in the same (x, y) cell column (all but the one warp in
  // ~56 that straddles two columns), their candidate ranges of a z-run are shifted copies of each other, so the warp
  // walks the UNION of the 32 ranges together: 32 candidates per step are loaded once, coalesced, parked in shared
  // memory (coordinate-major) and read back as 128-bit broadcasts of 4 candidates' x, y or z, and every lane tests every one of them against its own particle.  Candidates
  // outside a lane's own (conservative) range cannot be neighbours, so the exact predicate alone yields the same
  // bits; all 32 lanes stay busy for the same number of steps.  Only non-zero words are kept.
  __shared__ __align__(16) float stage[TPB / 32][3][32];
  const int wid = threadIdx.x >> 5;
  const int col = valid ? c.x * P.gdim[1] + c.y : -1;
  const int col0 = __shfl_sync(0xffffffffu, col, 0);
  bool uniform = __all_sync(0xffffffffu, !valid || col == col0) != 0 && !use_per_lane;
  const unsigned long long PX2 = pack_f32x2(pi.x, pi.x), PY2 = pack_f32x2(pi.y, pi.y), PZ2 = pack_f32x2(pi.z, pi.z), ONE2 = pack_f32x2(P.one, P.one);
  if (uniform) {
#pragma unroll 1
    for (int k = 0; k < 9; k++) {
      uint32_t b0 = 0, e0 = 0;
      if (valid) nb_run_range(P, cell_start, pi, c, k, cell, b0, e0);
      const bool has = e0 > b0;
      const uint32_t ub = __reduce_min_sync(0xffffffffu, has ? b0 : 0xffffffffu), ue = __reduce_max_sync(0xffffffffu, has ? e0 : 0u);
      for (uint32_t wb = ub; wb < ue; wb += 32u) {
        const uint32_t lim = min(32u, ue - wb);
        {
          const float4 q = (uint32_t)lane < lim ? __ldg(xs + wb + lane) : make_float4(1e18f, 1e18f, 1e18f, 0.f);
          stage[wid][0][lane] = q.x; stage[wid][1][lane] = q.y; stage[wid][2][lane] = q.z;
        }
        __syncwarp();
        uint32_t m = 0;
        for (uint32_t g = 0; g < lim; g += 8) {       // coordinate-major staging: one 128-bit broadcast read brings 4 candidates' x (y, z)
          const float4 xa = *reinterpret_cast<const float4*>(&stage[wid][0][g]), xb = *reinterpret_cast<const float4*>(&stage[wid][0][g + 4]);
          const float4 ya = *reinterpret_cast<const float4*>(&stage[wid][1][g]), yb = *reinterpret_cast<const float4*>(&stage[wid][1][g + 4]);
          const float4 za = *reinterpret_cast<const float4*>(&stage[wid][2][g]), zb = *reinterpret_cast<const float4*>(&stage[wid][2][g + 4]);
          // two candidates per packed instruction (FADD2 / FMUL2 / FFMA2 with a unit factor, each half rounded like the scalar __f*_rn: the EXACT regime holds)
          uint32_t m8 = ex_is_neighbor_x2(PX2, PY2, PZ2, xa.x, xa.y, ya.x, ya.y, za.x, za.y, P.h2, ONE2);
          m8 |= ex_is_neighbor_x2(PX2, PY2, PZ2, xa.z, xa.w, ya.z, ya.w, za.z, za.w, P.h2, ONE2) << 2;
          m8 |= ex_is_neighbor_x2(PX2, PY2, PZ2, xb.x, xb.y, yb.x, yb.y, zb.x, zb.y, P.h2, ONE2) << 4;
          m8 |= ex_is_neighbor_x2(PX2, PY2, PZ2, xb.z, xb.w, yb.z, yb.w, zb.z, zb.w, P.h2, ONE2) << 6;
          m |= m8 << g;
        }
        __syncwarp();
        if (!include_self && i - wb < 32u) m &= ~(1u << (i - wb));      // only the own column's run contains i
        if (!valid) m = 0;
        if (m) {
          cnt += __popc(m);
          if (nwords < (uint32_t)NB_TOTW) { masks[nwords] = m; wbase[nwords] = wb; cnt_cached = cnt; }
          nwords++;
        }
      }
    }
    if (__any_sync(0xffffffffu, nwords > (uint32_t)NB_TOTW)) { uniform = false; cnt = 0; nwords = 0; cnt_cached = 0; }   // crowded: per-lane path below
  }
  if (!uniform && valid) {
#pragma unroll 1
    for (int k = 0; k < 9; k++) {
      uint32_t b0, e0;
      nb_run_range(P, cell_start, pi, c, k, cell, b0, e0);
      for (uint32_t wb = b0; wb < e0; wb += 32u) {
        uint32_t m = nb_eval_word(xs, pi, P.h2, wb, min(32u, e0 - wb));
        if (!include_self && i - wb < 32u) m &= ~(1u << (i - wb));      // only the own column's run contains i
        cnt += __popc(m);
        if (nwords < (uint32_t)NB_TOTW) { masks[nwords] = m; wbase[nwords] = wb; cnt_cached = cnt; }
        nwords++;
      }
    }
  }
  const uint32_t rows = (__reduce_max_sync(0xffffffffu, cnt) + 3u) >> 2;    // rows of uint4
  unsigned long long off = 0;
  if (lane == 0) {
    off = atomicAdd(&sc->nbr_cursor, (unsigned long long)rows);
    if (off + rows > cap_rows) { atomicOr(&sc->err, ERRBIT_NBR_CAPACITY); off = ~0ull; }
  }
  off = __shfl_sync(0xffffffffu, off, 0);
  if (!valid) return;
  if (off == ~0ull) { nbr_cnt[t] = 0; if (lane == 0) slice_off[t >> 5] = 0; return; }
  if (lane == 0) slice_off[t >> 5] = (uint32_t)off;
  nbr_cnt[t] = cnt;
  uint32_t* out = nbr + off * 128ull + lane * 4;
  // fill: one flat loop over the set bits of the cached words (trip count = this lane's count)
  uint32_t s = 0, wi = 0, m = 0, wb = 0;
  for (; s < cnt_cached; s++) {
    while (m == 0) { m = masks[wi]; wb = wbase[wi]; wi++; }
    const uint32_t bit = __ffs(m) - 1;
    m &= m - 1;
    out[(size_t)(s >> 2) * 128u + (s & 3u)] = wb + bit;
  }
  if (nwords > (uint32_t)NB_TOTW) {                        // rare: recompute the words that did not fit
    uint32_t widx = 0;
#pragma unroll 1
    for (int k = 0; k < 9; k++) {
      uint32_t b0, e0;
      nb_run_range(P, cell_start, pi, c, k, cell, b0, e0);
      for (uint32_t w0 = b0; w0 < e0; w0 += 32u, widx++) {
        if (widx < (uint32_t)NB_TOTW) continue;
        uint32_t mm = nb_eval_word(xs, pi, P.h2, w0, min(32u, e0 - w0));
        if (!include_self && i - w0 < 32u) mm &= ~(1u << (i - w0));
        while (mm) {
          const uint32_t bit = __ffs(mm) - 1;
          mm &= mm - 1;
          out[(size_t)(s >> 2) * 128u + (s & 3u)] = w0 + bit;
          s++;
        }
      }
    }
  }
  for (; s & 3u; s++) out[(size_t)(s >> 2) * 128u + (s & 3u)] = sentinel;   // sentinel padding
}

// Particle::initializeWithNewNeighbors (particles.cpp:165-173) warns about every particle with fewer than
// NUM_NEIGHBOR_ALERT_THRESHOLD neighbours, printing its predicted position and velocity, in index order.  The counts
// are already on the device.  Three small kernels keep the output a pure function of the state although the
// compaction uses atomics: (1) histogram of the flagged particles over ALERT_BUCKETS ranges of the original id,
// (2) the largest bucket boundary B with at most `cap` flagged ids below it, (3) records of the flagged particles with
// id < B (the host sorts them by id): "the first m <= cap warnings in index order", plus the total.
static constexpr uint32_t ALERT_BUCKETS = 4096;
__global__ void __launch_bounds__(TPB)
k_alert_hist(uint32_t n, uint32_t thr, uint32_t shift, const uint32_t* __restrict__ nbr_cnt, const uint32_t* __restrict__ orig,
             uint32_t* __restrict__ hist) {
  const uint32_t i = blockIdx.x * TPB + threadIdx.x;
  if (i < n && nbr_cnt[i] < thr) atomicAdd(&hist[min(orig[i] >> shift, ALERT_BUCKETS - 1u)], 1u);
}
__global__ void k_alert_bound(uint32_t cap, uint32_t shift, const uint32_t* __restrict__ hist, Scalars* __restrict__ sc) {
  if (threadIdx.x || blockIdx.x) return;
  uint32_t total = 0, kept = 0, bound = 0;
  bool open_ = true;
  for (uint32_t b = 0; b < ALERT_BUCKETS; b++) {
    const uint32_t c = hist[b];
    if (open_ && kept + c <= cap) { kept += c; bound = b + 1; } else open_ = false;
    total += c;
  }
  sc->alert_count = total;
  sc->alert_kept = 0;
  sc->alert_bound = bound >= ALERT_BUCKETS ? 0xFFFFFFFFu : bound << shift;     // ids below this are reported
}
__global__ void __launch_bounds__(TPB)
k_neighbor_alert(uint32_t n, uint32_t thr, const uint32_t* __restrict__ nbr_cnt, const uint32_t* __restrict__ orig,
                 const float4* __restrict__ xs, const float4* __restrict__ vel, float4* __restrict__ out, uint32_t cap,
                 Scalars* __restrict__ sc) {
  const uint32_t i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = nbr_cnt[i];
  if (c >= thr) return;
  const uint32_t id = orig[i];
  if (id >= sc->alert_bound) return;
  const uint32_t slot = atomicAdd(&sc->alert_kept, 1u);
  if (slot >= cap) return;                                  // cannot happen: the bound admits at most cap
  const float4 x = xs[i], v = vel[i];
  out[2 * (size_t)slot] = make_float4(x.x, x.y, x.z, __uint_as_float(id));
  out[2 * (size_t)slot + 1] = make_float4(v.x, v.y, v.z, __uint_as_float(c));
}

// the sentinel particle (index n): far outside every support radius, zero velocity / vorticity
__global__ void k_set_sentinel(uint32_t n, float4* a, float4* b, float4* w, float4* vtmp, float4* omega, float4* xv, const SlabLink* lk) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (lk) n = lk->b[4];
    xv[2 * (size_t)n] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
    xv[2 * (size_t)n + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    a[n] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
    b[n] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
    w[n] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
    vtmp[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    omega[n] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Visit the frozen neighbours of the calling thread's particle: body(p_j float4, extra...) is
// applied to 4 gathered neighbours per coalesced uint4 index load, with the next index row
// prefetched while the current one is processed.
// The list rows are read exactly once per pass: they bypass L1 allocation (LDG.E.NA) so that they do not evict the
// neighbour positions, which ARE reused (measured: -2.5 % per step; an L2 evict-first policy on the same loads or
// evict-last on the gathers did not help, profiles/r01_cache_policy_ab.txt).
__device__ __forceinline__ uint4 ld_list_row(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
#define PBF_FOR_NEIGHBORS(t, BODY)                                                                   \
  {                                                                                                  \
    const uint4* lst_ = reinterpret_cast<const uint4*>(nbr) + (size_t)slice_off[(t) >> 5] * 32u + (threadIdx.x & 31); \
    const uint32_t rows_ = (nbr_cnt[t] + 3u) >> 2;                                                   \
    uint4 nx_ = rows_ ? ld_list_row(lst_) : make_uint4(0, 0, 0, 0);                                  \
    for (uint32_t r_ = 0; r_ < rows_; r_++) {                                                        \
      const uint4 jj_ = nx_;                                                                         \
      if (r_ + 1 < rows_) nx_ = ld_list_row(lst_ + (size_t)(r_ + 1) * 32u);                          \
      BODY(jj_.x) BODY(jj_.y) BODY(jj_.z) BODY(jj_.w)                                                \
    }                                                                                                \
  }

// ------------------------------------------------------------------------------------------------
// E. solver iteration: lambda pass (newtonStepCalculateLambda, particles.cpp:185-204)
//    reads xs_in.xyz, writes xs_out = (xyz unchanged, w = lambda_i)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_to_double(float v, double* target) {
  // warp shuffle reduce, one double atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(target, (double)v);
  return v;
}

__device__ __forceinline__ float lambda_particle(const DevParams& P, uint32_t t, uint32_t i, const float4* __restrict__ xs_in,
                                                 float4* __restrict__ xs_out, const uint32_t* __restrict__ nbr,
                                                 const uint32_t* __restrict__ slice_off, const uint32_t* __restrict__ nbr_cnt,
                                                 float* __restrict__ rho_out, const SlabLink* __restrict__ lk, const PushArgs& push) {
  const float4 pi = xs_in[i];
  float w3s = 0.f, gx = 0.f, gy = 0.f, gz = 0.f, dsum = 0.f;
#define BODY_L(J)                                                          \
  {                                                                        \
    const float4 pj = __ldg(&xs_in[J]);                                  \
    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;      \
    float r2, w3, g;                                                       \
    pair_terms(P, dx, dy, dz, r2, w3, g);                                  \
    w3s += w3;                                                             \
    gx = fmaf(g, dx, gx); gy = fmaf(g, dy, gy); gz = fmaf(g, dz, gz);      \
    dsum = fmaf(g * g, r2, dsum);                                          \
  }
  PBF_FOR_NEIGHBORS(t, BODY_L)
#undef BODY_L
  const float rho = P.poly6_c * w3s;
  const float gs = P.spiky_c * P.inv_rho0;              // grad_j C_i = gs * g * r_vec
  const float Gx = gs * gx, Gy = gs * gy, Gz = gs * gz;
  const float denom = gs * gs * dsum + (Gx * Gx + Gy * Gy + Gz * Gz);
  const float c_i = rho * P.inv_rho0 - 1.f;              // not clamped at 0 (Q7)
  const float lambda = -c_i / (denom + P.eps_relax);
  const float4 out = make_float4(pi.x, pi.y, pi.z, lambda);
  xs_out[i] = out;
  push_boundary(lk, push, t, out);                       // (x*, lambda) of the boundary columns -> the neighbours' ghost ranges
  if (rho_out) rho_out[i] = rho;
  return rho;
}

// Tile order of the gather kernels.  Blocks are handed to the SMs round-robin, so with the identity mapping SM s
// holds tiles s, s + n_sm, s + 2 n_sm, ... of the active window.  Inside every group of n_sm * 6 blocks (6 = resident
// CTAs per SM) the tile index is transposed, so that the CTAs resident on one SM work on ADJACENT tiles and share
// their neighbourhood lines in L1: lambda 2.56 -> 2.42 ms (the other gather kernels are unchanged; any group
// width 2..48 gives the same gain, profiles/r01_cache_policy_ab.txt).  Particles are independent within a pass,
// so the order has no effect on the results.
__device__ __forceinline__ uint32_t tile_of_block(const DevParams& P) {
  const uint32_t C = 6u, G = (uint32_t)P.n_sm * C, b = blockIdx.x, g = b / G, r = b % G;
  if ((g + 1u) * G > gridDim.x) return b;                 // tail group stays as is
  return g * G + (r % (uint32_t)P.n_sm) * C + (r / (uint32_t)P.n_sm);
}

__global__ void __launch_bounds__(TPB)      // 40 registers, 6 CTAs per SM; forcing 32 registers / 8 CTAs measured no faster
k_lambda(const __grid_constant__ DevParams P, uint32_t i0, uint32_t t0 /* multiple of 32 */, uint32_t n /* end of the t range */,
         const float4* __restrict__ xs_in,
         float4* __restrict__ xs_out, const uint32_t* __restrict__ nbr, const uint32_t* __restrict__ slice_off,
         const uint32_t* __restrict__ nbr_cnt, float* __restrict__ rho_out, double* __restrict__ rho_sum,
         const SlabLink* __restrict__ lk, const PushArgs push) {
  if (lk) { i0 = lk->b[0]; n = lk->b[3] - lk->b[0]; }
  const uint32_t t = t0 + tile_of_block(P) * TPB + threadIdx.x;
  float rho = 0.f;
  if (t < n) rho = lambda_particle(P, t, i0 + t, xs_in, xs_out, nbr, slice_off, nbr_cnt, rho_out, lk, push);
  // peer mode sizes the grid for the capacity: a block past the last particle must not queue its zeros on the one address every block adds to
  if (rho_sum && t - threadIdx.x < n) block_sum_to_double(rho, rho_sum);
}

// ------------------------------------------------------------------------------------------------
// F. solver iteration: delta-p + collide (newtonStepUpdatePosition + clamp, particles.cpp:206-213,51-84)
//    reads xs_in = (x*, lambda), writes xs_out.xyz = corrected position
//    NCORR: artificial-pressure exponent known at compile time (4 = reference), or -1 = runtime.
// ------------------------------------------------------------------------------------------------
template <int NCORR, bool SPH>
__device__ __forceinline__ void delta_particle(const DevParams& P, uint32_t t, uint32_t i, const float4* __restrict__ xs_in,
                                               float4* __restrict__ xs_out, const uint32_t* __restrict__ nbr,
                                               const uint32_t* __restrict__ slice_off, const uint32_t* __restrict__ nbr_cnt,
                                               const SlabLink* __restrict__ lk, const PushArgs& push) {
  const float4 pi = xs_in[i];
  float ax = 0.f, ay = 0.f, az = 0.f;
#define BODY_D(J)                                                          \
  {                                                                        \
    const float4 pj = __ldg(&xs_in[J]);                                  \
    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;      \
    float r2, w3, g;                                                       \
    pair_terms(P, dx, dy, dz, r2, w3, g);                                  \
    const float q = P.tscale_c * w3;                 /* W / W(dq) */       \
    float qn;                                                              \
    if (NCORR == 4) { const float q2 = q * q; qn = q2 * q2; }              \
    else { qn = 1.f; for (int e = 0; e < P.n_corr; e++) qn *= q; }         \
    const float f = (pi.w + pj.w - P.kcorr * qn) * g;                      \
    ax = fmaf(f, dx, ax); ay = fmaf(f, dy, ay); az = fmaf(f, dz, az);      \
  }
  PBF_FOR_NEIGHBORS(t, BODY_D)
#undef BODY_D
  const float sc = P.spiky_c * P.inv_rho0;
  const float3 dp = make_float3(sc * ax, sc * ay, sc * az);
  const float3 p = ex_collide<SPH>(P, make_float3(pi.x, pi.y, pi.z), dp, false);
  const float4 out = make_float4(p.x, p.y, p.z, 0.f);
  xs_out[i] = out;
  push_boundary(lk, push, t, out);
}

template <int NCORR, bool SPH>
__global__ void __launch_bounds__(TPB, SPH ? 6 : 1)   // obstacle instantiation: capped at 40 registers = 6 CTAs per SM like the box-only one (38)
k_delta(const __grid_constant__ DevParams P, uint32_t i0, uint32_t t0, uint32_t n, const float4* __restrict__ xs_in,
        float4* __restrict__ xs_out, const uint32_t* __restrict__ nbr, const uint32_t* __restrict__ slice_off,
        const uint32_t* __restrict__ nbr_cnt, const SlabLink* __restrict__ lk, const PushArgs push) {
  if (lk) { i0 = lk->b[0]; n = lk->b[3] - lk->b[0]; }
  const uint32_t t = t0 + tile_of_block(P) * TPB + threadIdx.x;
  if (t >= n) return;
  delta_particle<NCORR, SPH>(P, t, i0 + t, xs_in, xs_out, nbr, slice_off, nbr_cnt, lk, push);
}

// ------------------------------------------------------------------------------------------------
// G. finalize: velocity update (215-217), vorticity + XSPH + density (219-234, Jacobi, §7.3-3),
//    confinement + commit (236-248)
// ------------------------------------------------------------------------------------------------
// Also writes the 32-byte record (x*, v) per particle that the vorticity/XSPH pass gathers with ONE 256-bit load
// per pair (LDG.E.ENL2.256): a 32-byte gather costs 1.37x a 16-byte one, two 16-byte gathers from separate arrays
// cost 2x (scripts/ubench/ffma2.cu, profiles/r01_ubench_ffma2_gather.txt).
__global__ void __launch_bounds__(TPB)
k_velocity(const __grid_constant__ DevParams P, uint32_t n, const float4* __restrict__ xs,
           const float4* __restrict__ pos, float4* __restrict__ vtmp, float4* __restrict__ xv, const SlabLink* __restrict__ lk) {
  const uint32_t i = blockIdx.x * TPB + threadIdx.x;
  if (lk) n = lk->b[4];
  if (i >= n) return;
  const float4 a = xs[i], b = pos[i];
  const float4 v = make_float4(P.inv_dt * (a.x - b.x), P.inv_dt * (a.y - b.y), P.inv_dt * (a.z - b.z), 0.f);
  vtmp[i] = v;
  xv[2 * (size_t)i] = make_float4(a.x, a.y, a.z, 0.f);
  xv[2 * (size_t)i + 1] = v;
}

// one 256-bit read-only gather of record j: (x*, v)
__device__ __forceinline__ void ld_xv(const float4* __restrict__ xv, uint32_t j, float4& x, float4& v) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w), "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(xv + 2 * (size_t)j));
}

__global__ void __launch_bounds__(TPB)
k_vorticity_xsph(const __grid_constant__ DevParams P, uint32_t i0, uint32_t t0, uint32_t n, const float4* __restrict__ xs,
                 float4* __restrict__ xs_w, const float4* __restrict__ vtmp, const float4* __restrict__ xv, float4* __restrict__ vel_out, float4* __restrict__ omega,
                 float* __restrict__ rho_out, const uint32_t* __restrict__ nbr,
                 const uint32_t* __restrict__ slice_off, const uint32_t* __restrict__ nbr_cnt,
                 double* __restrict__ rho_sum, const SlabLink* __restrict__ lk, const PushArgs push) {
  if (lk) { i0 = lk->b[0]; n = lk->b[3] - lk->b[0]; }
  const uint32_t t = t0 + tile_of_block(P) * TPB + threadIdx.x;
  const uint32_t i = i0 + t;
  float rho = 0.f;
  if (t < n) {
    const float4 pi = xs[i];
    const float4 vi = vtmp[i];
    float w3s = 0.f, ox = 0.f, oy = 0.f, oz = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
#define BODY_V(J)                                                                     \
    {                                                                                 \
      float4 pj, vj;                                                                  \
      ld_xv(xv, J, pj, vj);                                                           \
      const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;               \
      const float ux = vj.x - vi.x, uy = vj.y - vi.y, uz = vj.z - vi.z;               \
      float r2, w3, g;                                                                \
      pair_terms(P, dx, dy, dz, r2, w3, g);                                           \
      ox = fmaf(g, uy * dz - uz * dy, ox);  /* omega += v_ij x grad W */              \
      oy = fmaf(g, uz * dx - ux * dz, oy);                                            \
      oz = fmaf(g, ux * dy - uy * dx, oz);                                            \
      sx = fmaf(w3, ux, sx); sy = fmaf(w3, uy, sy); sz = fmaf(w3, uz, sz);            \
      w3s += w3;                                                                      \
    }
    PBF_FOR_NEIGHBORS(t, BODY_V)
#undef BODY_V
    rho = P.poly6_c * w3s;
    ox *= P.spiky_c; oy *= P.spiky_c; oz *= P.spiky_c;
    const float on = sqrtf(ox * ox + oy * oy + oz * oz);
    omega[i] = make_float4(ox, oy, oz, on);
    xs_w[i] = make_float4(pi.x, pi.y, pi.z, on);        // (x*, |omega|): ONE 16-byte gather per pair in the confinement pass
    push_boundary(lk, push, t, make_float4(pi.x, pi.y, pi.z, on));
    const float xc = P.enable_xsph ? P.visc_c * P.poly6_c : 0.f;   // v += C * sum v_ij W  (not density-normalised, Q12)
    vel_out[i] = make_float4(fmaf(xc, sx, vi.x), fmaf(xc, sy, vi.y), fmaf(xc, sz, vi.z), 0.f);
    rho_out[i] = rho;                                              // the density the visualiser reads
  }
  if (rho_sum && t - threadIdx.x < n) block_sum_to_double(rho, rho_sum);   // not from the empty blocks of a capacity-sized grid (see k_lambda)
}

__global__ void __launch_bounds__(TPB)
k_confine_commit(const __grid_constant__ DevParams P, uint32_t i0, uint32_t n, const float4* __restrict__ xs /* (x*, |omega|) */,
                 const float4* __restrict__ omega, float4* __restrict__ vel, float4* __restrict__ pos,
                 const uint32_t* __restrict__ nbr, const uint32_t* __restrict__ slice_off,
                 const uint32_t* __restrict__ nbr_cnt, const SlabLink* __restrict__ lk) {
  if (lk) { i0 = lk->b[0]; n = lk->b[3] - lk->b[0]; }
  const uint32_t t = tile_of_block(P) * TPB + threadIdx.x;
  const uint32_t i = i0 + t;
  if (t >= n) return;
  const float4 pi = xs[i];
  if (P.enable_vorticity) {
    float ex = 0.f, ey = 0.f, ez = 0.f;
#define BODY_C(J)                                                          \
    {                                                                      \
      const float4 pj = __ldg(&xs[J]);                                 \
      const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;    \
      float r2, w3, g;                                                     \
      pair_terms(P, dx, dy, dz, r2, w3, g);                                \
      const float f = pj.w * g;    /* |omega_j| * grad W (no own term, Q13) */ \
      ex = fmaf(f, dx, ex); ey = fmaf(f, dy, ey); ez = fmaf(f, dz, ez);    \
    }
    PBF_FOR_NEIGHBORS(t, BODY_C)
#undef BODY_C
    ex *= P.spiky_c; ey *= P.spiky_c; ez *= P.spiky_c;
    const float en = sqrtf(ex * ex + ey * ey + ez * ez);
    if (en > P.eps_d) {
      const float rn = 1.f / en;
      const float nx = rn * ex, ny = rn * ey, nz = rn * ez;
      const float4 w = omega[i];
      float4 v = vel[i];
      v.x = fmaf(P.vort_dt_eps, ny * w.z - nz * w.y, v.x);
      v.y = fmaf(P.vort_dt_eps, nz * w.x - nx * w.z, v.y);
      v.z = fmaf(P.vort_dt_eps, nx * w.y - ny * w.x, v.z);
      vel[i] = v;
    }
  }
  pos[i] = make_float4(pi.x, pi.y, pi.z, 0.f);            // updatePosition (246-248)
}

// ------------------------------------------------------------------------------------------------
// G'. validation mode PBF_XSPH_REFERENCE_ORDER.  The reference runs
// updateVelocity + XSPH fused per particle in index order (particles.cpp:285-288, quirk Q11):
// particle i sees V_j = final velocity for j < i and the STALE pre-solve velocity s_j for j > i
// (original indices).  With u = (x* - x)/dt this is the strictly lower-triangular linear system
//     v_i = u_i + C * sum_j (V_j - u_i) W_ij ,   V_j = v_j (j < i)  |  s_j (j > i)
// solved here by fixed-point sweeps (contraction factor C*sum W ~ 0.7; every sweep makes one more
// level of the dependency chain exact).  FINAL also produces omega_i = sum (V_j - u_i) x grad W and
// the density.  Not on the performance path: three gathers per pair, ~48 sweeps.
// ------------------------------------------------------------------------------------------------
template <bool FINAL>
__global__ void __launch_bounds__(TPB)
k_xsph_reference(const __grid_constant__ DevParams P, uint32_t n, const float4* __restrict__ xs, const float4* __restrict__ u,
                 const float4* __restrict__ stale, const float4* __restrict__ v_in, float4* __restrict__ v_out,
                 const uint32_t* __restrict__ orig, float4* __restrict__ xs_w, float4* __restrict__ omega, float* __restrict__ rho_out,
                 const uint32_t* __restrict__ nbr, const uint32_t* __restrict__ slice_off, const uint32_t* __restrict__ nbr_cnt,
                 uint32_t sentinel, double* __restrict__ rho_sum) {
  const uint32_t t = blockIdx.x * TPB + threadIdx.x;
  float rho = 0.f;
  if (t < n) {
    const uint32_t i = t;
    const float4 pi = xs[i];
    const float4 ui = u[i];
    const uint32_t oi = orig[i];
    float w3s = 0.f, ox = 0.f, oy = 0.f, oz = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
#define BODY_X(J)                                                                     \
    {                                                                                 \
      const uint32_t j_ = (J);                                                        \
      const float4 pj = __ldg(&xs[j_]);                                               \
      const bool done_ = j_ != sentinel && __ldg(&orig[j_]) < oi;                     \
      const float4 vj = done_ ? v_in[j_] : __ldg(&stale[j_]);                         \
      const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;               \
      const float ux = vj.x - ui.x, uy = vj.y - ui.y, uz = vj.z - ui.z;               \
      float r2, w3, g;                                                                \
      pair_terms(P, dx, dy, dz, r2, w3, g);                                           \
      if (FINAL) {                                                                    \
        ox = fmaf(g, uy * dz - uz * dy, ox);                                          \
        oy = fmaf(g, uz * dx - ux * dz, oy);                                          \
        oz = fmaf(g, ux * dy - uy * dx, oz);                                          \
        w3s += w3;                                                                    \
      }                                                                               \
      sx = fmaf(w3, ux, sx); sy = fmaf(w3, uy, sy); sz = fmaf(w3, uz, sz);            \
    }
    PBF_FOR_NEIGHBORS(t, BODY_X)
#undef BODY_X
    const float xc = P.enable_xsph ? P.visc_c * P.poly6_c : 0.f;
    v_out[i] = make_float4(fmaf(xc, sx, ui.x), fmaf(xc, sy, ui.y), fmaf(xc, sz, ui.z), 0.f);
    if (FINAL) {
      rho = P.poly6_c * w3s;
      ox *= P.spiky_c; oy *= P.spiky_c; oz *= P.spiky_c;
      const float on = sqrtf(ox * ox + oy * oy + oz * oz);
      omega[i] = make_float4(ox, oy, oz, on);
      xs_w[i] = make_float4(pi.x, pi.y, pi.z, on);
      rho_out[i] = rho;
    }
  }
  if (FINAL && rho_sum) block_sum_to_double(rho, rho_sum);
}

// load-time density: sum of poly6 over the frozen set INCLUDING self (particles.cpp:158-163,440-444)
__global__ void __launch_bounds__(TPB)
k_density_only(const __grid_constant__ DevParams P, uint32_t i0, uint32_t n, const float4* __restrict__ xs,
               float* __restrict__ rho_out, const uint32_t* __restrict__ nbr,
               const uint32_t* __restrict__ slice_off, const uint32_t* __restrict__ nbr_cnt, const SlabLink* __restrict__ lk) {
  if (lk) { i0 = lk->b[0]; n = lk->b[3] - lk->b[0]; }
  const uint32_t t = blockIdx.x * TPB + threadIdx.x;
  const uint32_t i = i0 + t;
  if (t >= n) return;
  const float4 pi = xs[i];
  float w3s = 0.f;
#define BODY_R(J)                                                          \
  {                                                                        \
    const float4 pj = __ldg(&xs[J]);                                       \
    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;      \
    float r2, w3, g;                                                       \
    pair_terms(P, dx, dy, dz, r2, w3, g);                                  \
    w3s += w3;                                                             \
  }
  PBF_FOR_NEIGHBORS(t, BODY_R)
#undef BODY_R
  rho_out[i] = P.poly6_c * w3s;
}

// SPH density at arbitrary query points: Particles::estimateDensityAt (particles.cpp:446-453), the field
// the marching-cubes surfacer samples (particles.cpp:350-418).  xs = committed positions sorted by
// their own cells (the caller re-bins first); the reference's sum over ALL particles equals the sum
// over the 27-cell neighbourhood because poly6 vanishes beyond h.
__global__ void __launch_bounds__(TPB)
k_density_at(const __grid_constant__ DevParams P, uint32_t m, const float4* __restrict__ q, const float4* __restrict__ xs,
             const uint32_t* __restrict__ cell_start, float* __restrict__ out) {
  const uint32_t t = blockIdx.x * TPB + threadIdx.x;
  if (t >= m) return;
  const float4 pi = q[t];
  const int3 c = cell_coords(P, pi.x, pi.y, pi.z);
  const float reach = P.h * (1.f + 1e-3f) + 1e-6f * fabsf(pi.z);
  const int zlo = max((int)floorf((pi.z - reach - P.gmin[2]) * P.inv_cell_z), 0);
  const int zhi = min((int)floorf((pi.z + reach - P.gmin[2]) * P.inv_cell_z), P.gdim[2] - 1);
  float w3s = 0.f;
#pragma unroll 1
  for (int k = 0; k < 9; k++) {
    const int cx = c.x + (k / 3) - 1, cy = c.y + (k % 3) - 1;
    if (cx < 0 || cx >= P.gdim[0] || cy < 0 || cy >= P.gdim[1]) continue;
    const uint32_t base = (uint32_t)((cx * P.gdim[1] + cy) * P.gdim[2]);
    const uint32_t jb = cell_start[base + zlo], je = cell_start[base + zhi + 1];
#pragma unroll 4
    for (uint32_t j = jb; j < je; j++) {
      const float4 pj = __ldg(&xs[j]);
      const float dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
      const float tt = fmaxf(P.h2 - fmaf(dz, dz, fmaf(dy, dy, dx * dx)), 0.f);
      w3s = fmaf(tt * tt, tt, w3s);
    }
  }
  out[t] = P.poly6_c * w3s;
}

// ------------------------------------------------------------------------------------------------
// H. I/O helpers: original order <-> sorted order, fp32 AoS xyz <-> float4 SoA
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_import(uint32_t n, const float* __restrict__ pos_xyz, const float* __restrict__ vel_xyz,
         float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ orig) {
  const uint32_t i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  pos[i] = make_float4(pos_xyz[3 * i], pos_xyz[3 * i + 1], pos_xyz[3 * i + 2], 0.f);
  vel[i] = make_float4(vel_xyz[3 * i], vel_xyz[3 * i + 1], vel_xyz[3 * i + 2], 0.f);
  orig[i] = i;
}

// fp64 host layout converted on the device (same round-to-nearest cast as the host path)
__global__ void __launch_bounds__(TPB)
k_import_f64(uint32_t n, const double* __restrict__ pos_xyz, const double* __restrict__ vel_xyz,
             float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ orig) {
  const uint32_t i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  pos[i] = make_float4((float)pos_xyz[3 * i], (float)pos_xyz[3 * i + 1], (float)pos_xyz[3 * i + 2], 0.f);
  vel[i] = make_float4((float)vel_xyz[3 * i], (float)vel_xyz[3 * i + 1], (float)vel_xyz[3 * i + 2], 0.f);
  orig[i] = i;
}
__global__ void __launch_bounds__(TPB)
k_export3_f64(uint32_t n, const float4* __restrict__ src, const uint32_t* __restrict__ orig, double* __restrict__ dst_xyz) {
  const uint32_t s = blockIdx.x * TPB + threadIdx.x;
  if (s >= n) return;
  const float4 v = src[s];
  const size_t o = orig ? orig[s] : s;
  dst_xyz[3 * o] = (double)v.x; dst_xyz[3 * o + 1] = (double)v.y; dst_xyz[3 * o + 2] = (double)v.z;
}
__global__ void __launch_bounds__(TPB)
k_export1_f64(uint32_t n, const float* __restrict__ src, const uint32_t* __restrict__ orig, double* __restrict__ dst) {
  const uint32_t s = blockIdx.x * TPB + threadIdx.x;
  if (s >= n) return;
  dst[orig ? orig[s] : s] = (double)src[s];
}

__global__ void __launch_bounds__(TPB)
k_export3(uint32_t n, const float4* __restrict__ src, const uint32_t* __restrict__ orig, float* __restrict__ dst_xyz) {
  const uint32_t s = blockIdx.x * TPB + threadIdx.x;
  if (s >= n) return;
  const float4 v = src[s];
  const size_t o = orig ? orig[s] : s;
  dst_xyz[3 * o] = v.x; dst_xyz[3 * o + 1] = v.y; dst_xyz[3 * o + 2] = v.z;
}

__global__ void __launch_bounds__(TPB)
k_export1(uint32_t n, const float* __restrict__ src, const uint32_t* __restrict__ orig, float* __restrict__ dst) {
  const uint32_t s = blockIdx.x * TPB + threadIdx.x;
  if (s >= n) return;
  dst[orig ? orig[s] : s] = src[s];
}

__global__ void __launch_bounds__(TPB)
k_export_w(uint32_t n, const float4* __restrict__ src, const uint32_t* __restrict__ orig, float* __restrict__ dst) {
  const uint32_t s = blockIdx.x * TPB + threadIdx.x;
  if (s >= n) return;
  dst[orig ? orig[s] : s] = src[s].w;
}

__global__ void __launch_bounds__(TPB)
k_neighbor_digest(uint32_t i0, uint32_t n, const uint32_t* __restrict__ nbr, const uint32_t* __restrict__ slice_off,
                  const uint32_t* __restrict__ nbr_cnt, const uint32_t* __restrict__ orig,
                  unsigned long long* __restrict__ digest, uint32_t* __restrict__ count, int by_orig) {
  const uint32_t t = blockIdx.x * TPB + threadIdx.x;
  const uint32_t i = i0 + t;
  if (t >= n) return;
  const uint32_t cnt = nbr_cnt[t];
  const uint32_t* lst = nbr + (size_t)slice_off[t >> 5] * 128u + (threadIdx.x & 31) * 4u;
  unsigned long long d = 0;
  for (uint32_t s = 0; s < cnt; s++) d += mix64((uint64_t)orig[lst[(size_t)(s >> 2) * 128u + (s & 3u)]]);
  const uint32_t o = by_orig ? orig[i] : t;   // slab mode: ids are global, so write in range order
  digest[o] = d;
  count[o] = cnt;
}

// ------------------------------------------------------------------------------------------------
// launch sequence
// ------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(size_t n, int per_block = TPB) { return (unsigned)((n + per_block - 1) / per_block); }
// Peer mode: the counts live on the device, so launches are sized for the capacity (blocks beyond the real count exit
// at once) and the kernels read the ranges from the link block.
static inline const SlabLink* link_of(const Solver* h) { return h->p2p ? h->link : nullptr; }
static inline size_t owned_ub(const Solver* h) { return h->p2p ? h->append_base : h->r_cnt; }
static inline size_t sorted_ub(const Solver* h) { return h->p2p ? h->append_base : h->n_sorted; }
static inline PushArgs push_to(const Solver* h, int which /*0 xs_a, 1 xs_b, 2 xs_w*/) {
  PushArgs pa; pa.dst[0] = pa.dst[1] = nullptr;
  if (!h->p2p) return pa;
  for (int s = 0; s < 2; s++) pa.dst[s] = which == 0 ? h->peer[s].xs_a : (which == 1 ? h->peer[s].xs_b : h->peer[s].xs_w);
  return pa;
}

#define LAUNCH(h, kid, kern, grid, ...)                                   \
  do {                                                                    \
    if ((grid) > 0) {                                                     \
      (h)->prof_begin(kid);                                               \
      kern<<<(grid), TPB, 0, (h)->stream>>>(__VA_ARGS__);                 \
      (h)->prof_end(kid);                                                 \
      (h)->launches++;                                                    \
    }                                                                     \
  } while (0)

// Phase 1 of the sort: clear the histogram, predict the owned range [r_i0, r_i0 + r_cnt) of the
// current buffers (or just re-bin committed positions when apply_forces == 0) and hash it.
// where emigrants / ghosts for side s are packed: the own send buffer (a host driver moves it), or in peer mode the
// neighbour's receive buffer itself
static inline float4* mig_out(const Solver* h, int s) { return h->p2p ? h->peer[s].mig_recv : h->mig_send[s]; }
static inline float4* ghost_out(const Solver* h, int s) { return h->p2p ? h->peer[s].ghost_recv : h->ghost_send[s]; }

void enqueue_predict_hash(Solver* h, int apply_forces) {
  const int cur = h->cur;
  // every path that re-sorts (step, estimate_densities, re-binning for the density field / surfacer) starts here: the
  // copy stream may still be exporting xs_a / rho / vel through the old orig[] permutation (pbf_step is asynchronous)
  if (h->rb_pending) { cudaStreamWaitEvent(h->stream, h->ev_rb[3], 0); h->rb_pending = false; }
  cudaMemsetAsync(h->cell_count, 0, sizeof(uint32_t) * h->ncell, h->stream);
  cudaMemsetAsync(h->cell_of, 0xFF, sizeof(uint32_t) * h->n_in_cap(), h->stream);
  cudaMemsetAsync(&h->sc->nbr_cursor, 0, sizeof(unsigned long long), h->stream);
  cudaMemsetAsync(h->sc->counters, 0, sizeof(h->sc->counters), h->stream);
#define LAUNCH_PREDICT(KERN)                                                                                                      \
  LAUNCH(h, K_PREDICT, KERN, blocks_for(owned_ub(h)), h->dp, h->r_i0, h->r_cnt, h->pos[cur], h->vel[cur], h->orig[cur], h->xs_tmp, h->cell_of, \
         h->rank, h->cell_count, apply_forces, h->slab && h->has_left ? mig_out(h, 0) : (float4*)nullptr,                            \
         h->slab && h->has_right ? mig_out(h, 1) : (float4*)nullptr, (uint32_t)h->halo_cap, h->sc, link_of(h))
  if (h->dp.n_sph > 0 || h->dp.n_tri > 0) LAUNCH_PREDICT(k_predict_hash<true>); else LAUNCH_PREDICT(k_predict_hash<false>);
#undef LAUNCH_PREDICT
}

// Phase 2: scan, scatter the n_in entries of the unsorted arrays that carry a valid cell, canonical
// in-cell order, reorder into buffers cur^1 / xs_a.  Leaves n_sorted on the device (cell_start[ncell]).
void enqueue_sort(Solver* h, size_t n_in) {
  const uint32_t ncell = h->ncell;
  const int cur = h->cur, nxt = cur ^ 1;
  const unsigned sb = blocks_for(ncell, SCAN_TILE);
  LAUNCH(h, K_SCAN, k_scan_reduce, sb, ncell, h->cell_count, h->block_sums);
  LAUNCH(h, K_SCAN, k_scan_block_sums, 1, sb, h->block_sums);
  LAUNCH(h, K_SCAN, k_scan_apply, sb, ncell, h->cell_count, h->block_sums, h->cell_start);
  LAUNCH(h, K_SCATTER, k_scatter, blocks_for(n_in), (uint32_t)n_in, h->cell_of, h->rank, h->cell_start, h->orig[cur], h->perm, h->key);
  h->prof_begin(K_CELLSORT);
  k_cell_sort<<<blocks_for(ncell, 128), 128, 0, h->stream>>>(ncell, h->cell_start, h->perm, h->key);
  h->prof_end(K_CELLSORT); h->launches++;
  LAUNCH(h, K_REORDER, k_reorder, blocks_for(n_in), (uint32_t)n_in, h->cell_start + ncell, h->perm, h->key, h->pos[cur], h->vel[cur],
         h->xs_tmp, h->pos[nxt], h->vel[nxt], h->xs_a, h->orig[nxt]);
  h->cur = nxt;
}

static int nb_per_lane() {      // PBF_NB_PER_LANE=1: the per-lane candidate walk for every warp (A/B and fallback testing)
  static int v = -1;
  if (v < 0) { const char* e = getenv("PBF_NB_PER_LANE"); v = (e && atoi(e)) ? 1 : 0; }
  return v;
}

// Phase 3: frozen neighbour lists for the range [r_i0, r_i0 + r_cnt) of the n_sorted sorted particles.
void enqueue_build(Solver* h, int include_self) {
  h->prof_begin(K_REORDER);
  k_set_sentinel<<<1, 32, 0, h->stream>>>(h->n_sorted, h->xs_a, h->xs_b, h->xs_tmp, h->vtmp, h->omega, h->xv, link_of(h));
  h->prof_end(K_REORDER); h->launches++;
  LAUNCH(h, K_NEIGHBORS, k_build_neighbors, blocks_for(owned_ub(h)), h->dp, h->r_i0, h->r_cnt, h->n_sorted, h->xs_a, h->cell_start,
         h->nbr, h->slice_off, h->nbr_cnt, (unsigned long long)h->nbr_cap_rows, include_self, h->sc, nb_per_lane(), link_of(h));
}

// Sub-ranges of the owned range for overlapping halo exchange with compute (slab mode):
// PART_BOUNDARY = the first / last owned cell column rounded outwards to whole slices (what the
// x-neighbours need as ghosts), PART_INTERIOR = the rest, PART_ALL = everything.
static int part_ranges(const Solver* h, int part, uint32_t rng[2][2]) {
  const uint32_t cnt = (uint32_t)owned_ub(h);
  if (part == PART_ALL || !h->slab || h->p2p) { rng[0][0] = 0; rng[0][1] = cnt; return part == PART_INTERIOR ? 0 : 1; }
  const uint32_t nl = h->has_left ? h->bounds[1] - h->bounds[0] : 0u, nr = h->has_right ? h->bounds[3] - h->bounds[2] : 0u;
  uint32_t l_end = std::min(cnt, (nl + 31u) & ~31u), r_begin = (cnt - std::min(cnt, nr)) & ~31u;
  if (r_begin < l_end) r_begin = l_end;                   // thin slab: the two boundary parts meet
  if (part == PART_INTERIOR) { rng[0][0] = l_end; rng[0][1] = r_begin; return r_begin > l_end ? 1 : 0; }
  int k = 0;
  if (l_end > 0) { rng[k][0] = 0; rng[k][1] = l_end; k++; }
  if (cnt > r_begin) { rng[k][0] = r_begin; rng[k][1] = cnt; k++; }
  return k;
}

void enqueue_lambda(Solver* h, int first_iter, int part) {
  uint32_t rng[2][2];
  const int k = part_ranges(h, part, rng);
  for (int q = 0; q < k; q++)
    LAUNCH(h, K_LAMBDA, k_lambda, blocks_for(rng[q][1] - rng[q][0]), h->dp, h->r_i0, rng[q][0], rng[q][1], h->xs_a, h->xs_b, h->nbr,
           h->slice_off, h->nbr_cnt, (float*)nullptr, first_iter ? &h->sc->rho_first : (double*)nullptr, link_of(h), push_to(h, 1));
}
void enqueue_delta(Solver* h, int part) {
  uint32_t rng[2][2];
  const int k = part_ranges(h, part, rng);
  for (int q = 0; q < k; q++) {
    const unsigned g = blocks_for(rng[q][1] - rng[q][0]);
#define LAUNCH_DELTA(KERN) LAUNCH(h, K_DELTA, KERN, g, h->dp, h->r_i0, rng[q][0], rng[q][1], h->xs_b, h->xs_a, h->nbr, h->slice_off, h->nbr_cnt, link_of(h), push_to(h, 0))
    const bool sph = h->dp.n_sph > 0 || h->dp.n_tri > 0;   // box-only scenes run the instantiation without the obstacle code
    if (h->dp.n_corr == 4) { if (sph) LAUNCH_DELTA((k_delta<4, true>)); else LAUNCH_DELTA((k_delta<4, false>)); }
    else { if (sph) LAUNCH_DELTA((k_delta<-1, true>)); else LAUNCH_DELTA((k_delta<-1, false>)); }
#undef LAUNCH_DELTA
  }
}
void enqueue_velocity(Solver* h) {   // every sorted particle, ghosts included (their x and x* are bit-identical to the owner's)
  LAUNCH(h, K_VELOCITY, k_velocity, blocks_for(sorted_ub(h)), h->dp, h->n_sorted, h->xs_a, h->pos[h->cur], h->vtmp, h->xv, link_of(h));
}
void enqueue_vorticity(Solver* h, int part) {
  uint32_t rng[2][2];
  const int k = part_ranges(h, part, rng);
  for (int q = 0; q < k; q++)
    LAUNCH(h, K_VORT_XSPH, k_vorticity_xsph, blocks_for(rng[q][1] - rng[q][0]), h->dp, h->r_i0, rng[q][0], rng[q][1], h->xs_a, h->xs_tmp, h->vtmp, h->xv,
           h->vel[h->cur], h->omega, h->rho, h->nbr, h->slice_off, h->nbr_cnt, &h->sc->rho_final, link_of(h), push_to(h, 2));
}
// velocity + XSPH + vorticity in the reference's sequential order, by fixed-point sweeps (single GPU)
void enqueue_vorticity_reference_order(Solver* h) {
  const uint32_t n = h->r_cnt;
  if (n == 0) return;
  const int cur = h->cur, oth = cur ^ 1;
  int sweeps = 48;
  if (const char* e = getenv("PBF_XSPH_REF_SWEEPS")) sweeps = std::max(1, atoi(e));
  float4* a = h->vel[oth];            // the other halves of the ping-pong buffers are free after the reorder
  float4* b = h->pos[oth];
  const float4* stale = h->vel[cur];  // pre-solve velocity (old v + gravity*dt): what j > i still holds in the reference
  cudaMemcpyAsync(a, h->vtmp, sizeof(float4) * n, cudaMemcpyDeviceToDevice, h->stream);     // v^(0) = u
  for (int k = 0; k < sweeps; k++) {
    LAUNCH(h, K_VORT_XSPH, k_xsph_reference<false>, blocks_for(n), h->dp, n, h->xs_a, h->vtmp, stale, a, b, h->orig[cur], (float4*)nullptr,
           (float4*)nullptr, (float*)nullptr, h->nbr, h->slice_off, h->nbr_cnt, h->n_sorted, (double*)nullptr);
    std::swap(a, b);
  }
  LAUNCH(h, K_VORT_XSPH, k_xsph_reference<true>, blocks_for(n), h->dp, n, h->xs_a, h->vtmp, stale, a, b, h->orig[cur], h->xs_tmp, h->omega,
         h->rho, h->nbr, h->slice_off, h->nbr_cnt, h->n_sorted, &h->sc->rho_final);
  cudaMemcpyAsync(h->vel[cur], b, sizeof(float4) * n, cudaMemcpyDeviceToDevice, h->stream);  // only now may the stale velocities go
}

void enqueue_confine(Solver* h) {
  LAUNCH(h, K_CONFINE, k_confine_commit, blocks_for(owned_ub(h)), h->dp, h->r_i0, h->r_cnt, h->xs_tmp, h->omega, h->vel[h->cur],
         h->pos[h->cur], h->nbr, h->slice_off, h->nbr_cnt, link_of(h));
}

// single-GPU step: every particle is owned, n never changes, nothing needs a host round trip
// one piece of the streaming read-back: wait (on the copy stream) until `ready` fires on the main stream,
// scatter to original order as fp64 and DMA to the caller's page-locked buffer
static void readback_piece(Solver* h, cudaEvent_t ready, int what) {
  const size_t n = h->n;
  cudaEventRecord(ready, h->stream);
  cudaStream_t main_stream = h->stream;
  cudaStreamWaitEvent(h->copy_stream, ready, 0);
  h->stream = h->copy_stream;                               // LAUNCH() targets h->stream
  if (what == 0 && h->rb_pos) { enqueue_export3_f64(h, h->xs_a, h->rb_stage); cudaMemcpyAsync(h->rb_pos, h->rb_stage, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream); }
  if (what == 1 && h->rb_rho) { enqueue_export1_f64(h, h->rho, h->rb_stage + 6 * n); cudaMemcpyAsync(h->rb_rho, h->rb_stage + 6 * n, n * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream); }
  if (what == 2 && h->rb_vel) { enqueue_export3_f64(h, h->vel[h->cur], h->rb_stage + 3 * n); cudaMemcpyAsync(h->rb_vel, h->rb_stage + 3 * n, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream); }
  h->stream = main_stream;
  if (what == 2) { cudaEventRecord(h->ev_rb[3], h->copy_stream); h->rb_pending = true; }
}

// all three pieces at once, after a complete step (the graph-replay path of launch-bound scenes: nothing to hide behind)
void enqueue_readback_all(Solver* h) {
  if (!h->copy_stream || h->n == 0 || !(h->rb_pos || h->rb_vel || h->rb_rho)) return;
  for (int what = 0; what < 3; what++) readback_piece(h, h->ev_rb[what], what);
}

void enqueue_step(Solver* h, bool readback) {
  const uint32_t n = (uint32_t)h->n;
  if (n == 0) return;
  cudaMemsetAsync(&h->sc->rho_first, 0, 2 * sizeof(double), h->stream);   // rho_first, rho_final
  h->r_i0 = 0; h->r_cnt = n; h->n_sorted = n;
  enqueue_predict_hash(h, 1);
  enqueue_sort(h, n);
  enqueue_build(h, 0);
  if (h->capture_xpred) cudaMemcpyAsync(h->xpred, h->xs_a, sizeof(float4) * n, cudaMemcpyDeviceToDevice, h->stream);
  if (h->alert_thr > 0 && h->alert_buf) {
    uint32_t* hist = reinterpret_cast<uint32_t*>(h->alert_buf + 2 * h->alert_cap);
    uint32_t shift = 0;
    while (((uint64_t)ALERT_BUCKETS << shift) < (uint64_t)n) shift++;
    cudaMemsetAsync(hist, 0, ALERT_BUCKETS * sizeof(uint32_t), h->stream);
    LAUNCH(h, K_IO, k_alert_hist, blocks_for(n), n, h->alert_thr, shift, h->nbr_cnt, h->orig[h->cur], hist);
    h->prof_begin(K_IO);
    k_alert_bound<<<1, 32, 0, h->stream>>>((uint32_t)h->alert_cap, shift, hist, h->sc);
    h->prof_end(K_IO); h->launches++;
    LAUNCH(h, K_IO, k_neighbor_alert, blocks_for(n), n, h->alert_thr, h->nbr_cnt, h->orig[h->cur], h->xs_a, h->vel[h->cur], h->alert_buf,
           (uint32_t)h->alert_cap, h->sc);
  }
  for (int it = 0; it < h->dp.iterations; it++) { enqueue_lambda(h, it == 0, PART_ALL); enqueue_delta(h, PART_ALL); }
  readback = readback && h->copy_stream && (h->rb_pos || h->rb_vel || h->rb_rho);
  if (readback) readback_piece(h, h->ev_rb[0], 0);          // positions are final once the iterations end
  enqueue_velocity(h);
  if (h->hp.xsph_mode == PBF_XSPH_REFERENCE_ORDER) enqueue_vorticity_reference_order(h);
  else enqueue_vorticity(h, PART_ALL);
  if (readback) readback_piece(h, h->ev_rb[1], 1);          // density is final after the vorticity/XSPH pass
  enqueue_confine(h);
  if (readback) readback_piece(h, h->ev_rb[2], 2);          // velocity last
  h->steps_done++;
}

void enqueue_estimate_densities(Solver* h) {
  const uint32_t n = (uint32_t)h->n;
  if (n == 0) return;
  h->r_i0 = 0; h->r_cnt = n; h->n_sorted = n;
  enqueue_predict_hash(h, 0);
  enqueue_sort(h, n);
  enqueue_build(h, 1);
  LAUNCH(h, K_DENSITY, k_density_only, blocks_for(n), h->dp, 0u, n, h->xs_a, h->rho, h->nbr, h->slice_off, h->nbr_cnt, (const SlabLink*)nullptr);
}

// re-bin the particles by their COMMITTED positions (after a step the cells are those of the predicted
// positions): afterwards xs_a = sorted committed positions and cell_start matches them
void enqueue_rebin(Solver* h) {
  const uint32_t n = (uint32_t)h->n;
  if (n == 0) return;
  h->r_i0 = 0; h->r_cnt = n; h->n_sorted = n;
  enqueue_predict_hash(h, 0);
  enqueue_sort(h, n);
}
void enqueue_density_at(Solver* h, uint32_t m, const float4* d_q, float* d_out) {
  LAUNCH(h, K_DENSITY, k_density_at, blocks_for(m), h->dp, m, d_q, h->xs_a, h->cell_start, d_out);
}

void enqueue_import(Solver* h, const float* d_pos_xyz, const float* d_vel_xyz) {
  const uint32_t n = (uint32_t)h->n;
  if (n == 0) return;
  LAUNCH(h, K_IO, k_import, blocks_for(n), n, d_pos_xyz, d_vel_xyz, h->pos[h->cur], h->vel[h->cur], h->orig[h->cur]);
}

void enqueue_import_f64(Solver* h, const double* d_pos_xyz, const double* d_vel_xyz) {
  const uint32_t n = (uint32_t)h->n;
  if (n == 0) return;
  LAUNCH(h, K_IO, k_import_f64, blocks_for(n), n, d_pos_xyz, d_vel_xyz, h->pos[h->cur], h->vel[h->cur], h->orig[h->cur]);
}
void enqueue_export3_f64(Solver* h, const float4* src, double* dst_xyz) {
  if (h->r_cnt) LAUNCH(h, K_IO, k_export3_f64, blocks_for(h->r_cnt), h->r_cnt, src + h->r_i0, h->slab ? (const uint32_t*)nullptr : h->orig[h->cur] + h->r_i0, dst_xyz);
}
void enqueue_export1_f64(Solver* h, const float* src, double* dst) {
  if (h->r_cnt) LAUNCH(h, K_IO, k_export1_f64, blocks_for(h->r_cnt), h->r_cnt, src + h->r_i0, h->slab ? (const uint32_t*)nullptr : h->orig[h->cur] + h->r_i0, dst);
}

// exports act on the owned range; dst index = original id (single GPU) or position in the range (slab)
void enqueue_export3(Solver* h, const float4* src, float* dst_xyz) {
  if (h->r_cnt) LAUNCH(h, K_IO, k_export3, blocks_for(h->r_cnt), h->r_cnt, src + h->r_i0, h->slab ? (const uint32_t*)nullptr : h->orig[h->cur] + h->r_i0, dst_xyz);
}
void enqueue_export1(Solver* h, const float* src, float* dst) {
  if (h->r_cnt) LAUNCH(h, K_IO, k_export1, blocks_for(h->r_cnt), h->r_cnt, src + h->r_i0, h->slab ? (const uint32_t*)nullptr : h->orig[h->cur] + h->r_i0, dst);
}
void enqueue_export_w(Solver* h, const float4* src, float* dst) {
  if (h->r_cnt) LAUNCH(h, K_IO, k_export_w, blocks_for(h->r_cnt), h->r_cnt, src + h->r_i0, h->slab ? (const uint32_t*)nullptr : h->orig[h->cur] + h->r_i0, dst);
}
void enqueue_digest(Solver* h, unsigned long long* digest, uint32_t* count) {
  if (h->r_cnt) LAUNCH(h, K_IO, k_neighbor_digest, blocks_for(h->r_cnt), h->r_i0, h->r_cnt, h->nbr, h->slice_off, h->nbr_cnt, h->orig[h->cur], digest, count, h->slab ? 0 : 1);
}

}  // namespace pbf

#include "pbf_slab.inl"
#include "pbf_surface.inl"
#include "pbf_probe.inl"
