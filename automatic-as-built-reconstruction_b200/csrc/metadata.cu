// metadata.cu - the integer core on the GPU: coordinate packing, hash-grid insert/lookup,
// input-layer numbering, submanifold / strided rulebooks, tile books, SparseToDense rules.
//
// Reference semantics reproduced (all under SparseConvNet/sparseconvnet/SCN/Metadata/):
//   IOLayersRules.h:19-125            input layer: site id = order of first occurrence
//   SubmanifoldConvolutionRules.h:13-87  k = row-major offset (last dim fastest), pairs (in,out)
//   ConvolutionRules.h:12-105, RectangularRegions.h:97-119  strided rules + output grid
//   ConvolutionRules.h:110-151        SparseToDense linear offsets
//   Metadata.cpp:149-168              getSpatialLocations
// The reference numbers conv-created sites in sparsehash iteration order (unpinned); here the
// order is first touch in input-row order, made batch-contiguous ascending like the reference's
// per-sample loop (ConvolutionRules.h:80-88).
#include "metadata.cuh"
#include "../../include/scn_b200.h"
#include <algorithm>
#include <stdlib.h>
#include <mutex>

namespace scn {

static bool g_tile_grouping = true;

// ---------------------------------------------------------------------------------------
// hash grid
// ---------------------------------------------------------------------------------------
static uint32_t table_capacity(int64_t n) {
  uint64_t c = 64;
  while (c < (uint64_t)(2 * n + 2)) c <<= 1;
  return (uint32_t)c;
}

static int grid_alloc_table(Grid *g, int64_t n_keys, int fill_val_byte, cudaStream_t s) {
  g->hcap = table_capacity(n_keys);
  SCN_TRY(dev_alloc_t(&g->hkeys, g->hcap, s));
  SCN_TRY(dev_alloc_t(&g->hvals, g->hcap, s));
  SCN_CUDA(cudaMemsetAsync(g->hkeys, 0xFF, (size_t)g->hcap * 8, s));
  SCN_CUDA(cudaMemsetAsync(g->hvals, fill_val_byte, (size_t)g->hcap * 4, s));
  return 0;
}

static void grid_free(Grid *g, cudaStream_t s) {
  dev_free(g->coords, s);
  dev_free(g->hkeys, s);
  dev_free(g->hvals, s);
  delete g;
}

static void tilebook_free(TileBook &tb, cudaStream_t s) {
  if (tb.next) {
    tilebook_free(*tb.next, s);
    delete tb.next;
  }
  dev_free(tb.perm, s);
  dev_free(tb.tile_mask, s);
  dev_free(tb.tile_off, s);
  dev_free(tb.order, s);
  dev_free(tb.entries, s);
  tb = TileBook();
}

static void rulebook_free(RuleBook *rb, cudaStream_t s) {
  dev_free(rb->pairs, s);
  dev_free(rb->t_out, s);
  dev_free(rb->t_in, s);
  dev_free(rb->dw_work, s);
  tilebook_free(rb->tb_out, s);
  tilebook_free(rb->tb_in, s);
  delete rb;
}

// A rulebook (and the output grid of a strided one) is registered in the metadata before it is built, because the
// build's helpers look it up; this guard takes a half-built entry out again when the build fails (out of memory, a
// size check), so that a later find_rulebook / find_grid never returns it.
struct BuildGuard {
  scn_metadata *m;
  RuleBook *rb;
  Grid *grid;
  cudaStream_t s;
  bool ok = false;
  ~BuildGuard() {
    if (ok) return;
    for (size_t i = 0; i < m->rulebooks.size(); ++i)
      if (m->rulebooks[i] == rb) { m->rulebooks.erase(m->rulebooks.begin() + i); rulebook_free(rb, s); break; }
    if (grid)
      for (size_t i = 0; i < m->grids.size(); ++i)
        if (m->grids[i] == grid) { m->grids.erase(m->grids.begin() + i); grid_free(grid, s); break; }
  }
};

static bool same3(const int64_t *a, const int64_t *b) {
  return a[0] == b[0] && a[1] == b[1] && a[2] == b[2];
}

Grid *find_grid(scn_metadata *m, const int64_t *ss) {
  for (Grid *g : m->grids)
    if (same3(g->ss, ss)) return g;
  return nullptr;
}

static void metadata_clear(scn_metadata *m, cudaStream_t s) {
  for (Grid *g : m->grids) grid_free(g, s);
  m->grids.clear();
  for (RuleBook *rb : m->rulebooks) rulebook_free(rb, s);
  m->rulebooks.clear();
  dev_free(m->input.point_row, s);
  dev_free(m->input.csr_off, s);
  dev_free(m->input.members, s);
  dev_free(m->input.stat, s);
  m->input = InputRules();
  m->batch_size = 0;
}

// ---------------------------------------------------------------------------------------
// input layer (IOLayersRules.h:19-125)
// ---------------------------------------------------------------------------------------
// stat[0] error flag, stat[1] max batch index, stat[2] "batch column not ascending"
__global__ void k_in_insert(const int64_t *__restrict__ coords, long long n, int ncols,
                            uint64_t *hk, int32_t *hv, uint32_t mask, int32_t *pslot,
                            int32_t *stat, int take_max) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t *c = coords + i * ncols;
  const long long x = c[0], y = c[1], z = c[2], b = (ncols == 4) ? c[3] : 0;
  if (x < 0 || y < 0 || z < 0 || x > 65535 || y > 65535 || z > 65535 || b < 0 || b > 32767) {
    atomicOr(&stat[0], 1);
    pslot[i] = 0;
    return;
  }
  if (ncols == 4) {
    atomicMax(&stat[1], (int)b);
    if (i > 0 && coords[(i - 1) * ncols + 3] > b) atomicOr(&stat[2], 1);
  }
  const uint32_t slot = hash_insert(hk, mask, pack_key((int)x, (int)y, (int)z, (int)b));
  pslot[i] = (int32_t)slot;
  if (take_max) atomicMax(&hv[slot], (int)i);
  else atomicMin(&hv[slot], (int)i);
}

__global__ void k_in_flag(const int32_t *__restrict__ pslot, const int32_t *__restrict__ hv,
                          int32_t *__restrict__ flag, long long n) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (hv[pslot[i]] == (int)i) ? 1 : 0;
}

// first occurrences take their rank as site id, publish coordinates, and relabel the table
__global__ void k_in_assign(const int64_t *__restrict__ coords, long long n, int ncols,
                            const int32_t *__restrict__ rank, const int32_t *__restrict__ pslot,
                            int32_t *hv, int32_t *__restrict__ site_coords) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rank[i];
  if (rank[i + 1] == r) return;  // not a first occurrence
  const int64_t *c = coords + i * ncols;
  int4 v = make_int4((int)c[0], (int)c[1], (int)c[2], ncols == 4 ? (int)c[3] : 0);
  reinterpret_cast<int4 *>(site_coords)[r] = v;
  hv[pslot[i]] = r;
}

__global__ void k_in_rows(const int32_t *__restrict__ pslot, const int32_t *__restrict__ hv,
                          int32_t *__restrict__ prow, int32_t *cnt, long long n) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = hv[pslot[i]];
  prow[i] = r;
  atomicAdd(&cnt[r], 1);
}

__global__ void k_in_fill(const int32_t *__restrict__ prow, const int32_t *__restrict__ csr_off,
                          int32_t *cursor, int32_t *__restrict__ members, long long n) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = prow[i];
  const int pos = atomicAdd(&cursor[r], 1);
  members[csr_off[r] + pos] = (int)i;
}

// point lists come out of the atomics in arbitrary order; the reference lists them in
// ascending point order (IOLayersRules.h:92) - restore that (lists are 1-3 long in practice)
__global__ void k_in_sort_members(const int32_t *__restrict__ csr_off, int32_t *members,
                                  long long n_rows, int32_t *max_active) {
  pdl_sync();
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int b = csr_off[r], e = csr_off[r + 1];
  for (int i = b + 1; i < e; ++i) {
    const int v = members[i];
    int j = i - 1;
    while (j >= b && members[j] > v) { members[j + 1] = members[j]; --j; }
    members[j + 1] = v;
  }
  if (e - b > 1) atomicMax(max_active, e - b);
}

__global__ void k_iota(int32_t *a, long long n) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int)i;
}

__global__ void k_in_mode0(const int64_t *__restrict__ coords, long long n, int ncols,
                           int32_t *__restrict__ site_coords) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t *c = coords + i * ncols;
  reinterpret_cast<int4 *>(site_coords)[i] =
      make_int4((int)c[0], (int)c[1], (int)c[2], ncols == 4 ? (int)c[3] : 0);
}

static int input_layer_prepare(scn_metadata *m, const int64_t *ss, const int64_t *coords,
                               int64_t n, int ncols, int on_device, int64_t batch_size,
                               int mode, cudaStream_t s, int64_t *n_active_out) {
  SCN_CHECK(m->dim == 3, "only dimension 3 is implemented (Metadata_3)");
  SCN_CHECK(ncols == 3 || ncols == 4, "coords must have 3 or 4 columns, got %d", ncols);
  SCN_CHECK(mode >= 0 && mode <= 4, "InputLayer mode %d not in 0..4", mode);
  SCN_CHECK(n >= 0 && n < (1LL << 30), "too many input points: %lld", (long long)n);
  metadata_clear(m, s);
  m->last_stream = s;
  prof_begin(PROF_RULES, s);
  Grid *g = new Grid();
  memcpy(g->ss, ss, sizeof(g->ss));
  m->grids.push_back(g);
  InputRules &ir = m->input;
  ir.mode = mode;
  ir.n_points = n;
  if (n == 0) {
    SCN_TRY(grid_alloc_table(g, 0, 0x7F, s));
    SCN_TRY(dev_alloc_t(&g->coords, 4, s));
    SCN_TRY(dev_alloc_t(&ir.point_row, 1, s));
    SCN_TRY(dev_alloc_t(&ir.csr_off, 1, s));
    SCN_CUDA(cudaMemsetAsync(ir.csr_off, 0, 4, s));
    SCN_TRY(dev_alloc_t(&ir.members, 1, s));
    ir.built = true;
    m->batch_size = batch_size;
    *n_active_out = 0;
    return 0;
  }
  const int64_t *dcoords = coords;
  int64_t *staged = nullptr;
  if (!on_device) {
    SCN_TRY(dev_alloc_t(&staged, (size_t)n * ncols, s));
    SCN_CUDA(cudaMemcpyAsync(staged, coords, (size_t)n * ncols * 8, cudaMemcpyHostToDevice, s));
    dcoords = staged;
  }
  int32_t *pslot = nullptr, *stat = nullptr, *rank = nullptr;
  SCN_TRY(dev_alloc_t(&pslot, (size_t)n, s));
  SCN_TRY(dev_alloc_t(&stat, 8, s));
  SCN_CUDA(cudaMemsetAsync(stat, 0, 32, s));
  SCN_TRY(grid_alloc_table(g, n, mode == 0 ? 0x80 : 0x7F, s));
  const int T = 256, nb = cdiv(n, T);
  SCN_LAUNCH(k_in_insert, nb, T, 0, s, dcoords, n, ncols, g->hkeys, g->hvals, g->hcap - 1, pslot, stat,
                               mode == 0);
  SCN_LAUNCHED();
  int64_t *hs = host_scratch(16);
  SCN_CHECK(hs, "pinned host scratch allocation failed");
  int32_t *hs32 = (int32_t *)hs;
  if (mode == 0) {
    // guaranteed-unique mode: rows are the points themselves (IOLayersRules.h:29-58)
    SCN_CUDA(cudaMemcpyAsync(hs32, stat, 16, cudaMemcpyDeviceToHost, s));
    SCN_CUDA(cudaStreamSynchronize(s));
    SCN_CHECK(hs32[0] == 0, "InputLayer: coordinate outside [0,65535] or batch outside [0,32767]");
    g->n_active = n;
    SCN_TRY(dev_alloc_t(&g->coords, (size_t)n * 4, s));
    SCN_LAUNCH(k_in_mode0, nb, T, 0, s, dcoords, n, ncols, g->coords);
    SCN_LAUNCHED();
    SCN_TRY(dev_alloc_t(&ir.point_row, (size_t)n, s));
    SCN_TRY(dev_alloc_t(&ir.members, (size_t)n, s));
    SCN_TRY(dev_alloc_t(&ir.csr_off, (size_t)n + 1, s));
    SCN_LAUNCH(k_iota, nb, T, 0, s, ir.point_row, n);
    SCN_LAUNCHED();
    SCN_LAUNCH(k_iota, nb, T, 0, s, ir.members, n);
    SCN_LAUNCHED();
    SCN_LAUNCH(k_iota, cdiv(n + 1, T), T, 0, s, ir.csr_off, n + 1);
    SCN_LAUNCHED();
    ir.n_active = n;
    ir.max_active = 1;
  } else {
    SCN_TRY(dev_alloc_t(&rank, (size_t)n + 1, s));
    SCN_LAUNCH(k_in_flag, nb, T, 0, s, pslot, g->hvals, rank, n);
    SCN_LAUNCHED();
    SCN_TRY(exclusive_scan_i32(rank, rank, n, s));
    SCN_CUDA(cudaMemcpyAsync(hs32, stat, 16, cudaMemcpyDeviceToHost, s));
    SCN_CUDA(cudaMemcpyAsync(hs32 + 4, rank + n, 4, cudaMemcpyDeviceToHost, s));
    SCN_CUDA(cudaStreamSynchronize(s));  // the one documented read-back of this call
    SCN_CHECK(hs32[0] == 0, "InputLayer: coordinate outside [0,65535] or batch outside [0,32767]");
    const int64_t na = hs32[4];
    g->n_active = na;
    ir.n_active = na;
    SCN_TRY(dev_alloc_t(&g->coords, (size_t)na * 4, s));
    SCN_LAUNCH(k_in_assign, nb, T, 0, s, dcoords, n, ncols, rank, pslot, g->hvals, g->coords);
    SCN_LAUNCHED();
    int32_t *cnt = nullptr;
    SCN_TRY(dev_alloc_t(&ir.point_row, (size_t)n, s));
    SCN_TRY(dev_alloc_t(&ir.csr_off, (size_t)na + 1, s));
    SCN_TRY(dev_alloc_t(&ir.members, (size_t)n, s));
    SCN_TRY(dev_alloc_t(&cnt, (size_t)na + 2, s));
    SCN_CUDA(cudaMemsetAsync(cnt, 0, ((size_t)na + 2) * 4, s));
    SCN_LAUNCH(k_in_rows, nb, T, 0, s, pslot, g->hvals, ir.point_row, cnt, n);
    SCN_LAUNCHED();
    SCN_TRY(exclusive_scan_i32(cnt, ir.csr_off, na, s));
    SCN_CUDA(cudaMemsetAsync(cnt, 0, ((size_t)na + 2) * 4, s));
    SCN_LAUNCH(k_in_fill, nb, T, 0, s, ir.point_row, ir.csr_off, cnt, ir.members, n);
    SCN_LAUNCHED();
    int32_t *mx = cnt + na + 1;  // zeroed above, untouched by k_in_fill
    SCN_LAUNCH(k_in_sort_members, cdiv(na, T), T, 0, s, ir.csr_off, ir.members, na, mx);
    SCN_LAUNCHED();
    SCN_CUDA(cudaMemcpyAsync(stat + 4, mx, 4, cudaMemcpyDeviceToDevice, s));
    ir.max_active = -1;  // stat[4] on the device; fetched lazily (scn_input_rulebook_header)
    dev_free(cnt, s);
    dev_free(rank, s);
  }
  g->batch_sorted = (hs32[2] == 0);
  const int64_t bs_seen = (ncols == 4) ? (int64_t)hs32[1] + 1 : 1;
  m->batch_size = std::max<int64_t>(batch_size, bs_seen);
  ir.built = true;
  ir.stat = stat;  // owned by the InputRules from here on
  dev_free(pslot, s);
  dev_free(staged, s);
  prof_end(PROF_RULES, s, (double)n * ncols * 8 + 12.0 * g->n_active + 8.0 * n, 0);
  *n_active_out = g->n_active;
  return 0;
}

// ---------------------------------------------------------------------------------------
// rulebook tables
// ---------------------------------------------------------------------------------------
struct Filter3 { int size[3]; int stride[3]; int out_size[3]; };

// T[k*n+s] = row of the site at coords(s)+delta_k, or -1 (SubmanifoldConvolutionRules.h:13-45)
__global__ void k_sub_table(const int32_t *__restrict__ coords, long long n, Filter3 f, int K,
                            const uint64_t *__restrict__ hk, const int32_t *__restrict__ hv,
                            uint32_t mask, int32_t *__restrict__ T) {
  pdl_sync();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * K) return;
  const int k = (int)(idx / n);
  const long long s = idx - (long long)k * n;
  const int4 c = reinterpret_cast<const int4 *>(coords)[s];
  const int iz = k % f.size[2], iy = (k / f.size[2]) % f.size[1], ix = k / (f.size[2] * f.size[1]);
  const int x = c.x + ix - f.size[0] / 2, y = c.y + iy - f.size[1] / 2, z = c.z + iz - f.size[2] / 2;
  int r = -1;
  if (x == c.x && y == c.y && z == c.z) r = (int)s;
  else if (coord_ok(x, y, z)) r = hash_find(hk, hv, mask, pack_key(x, y, z, c.w));
  T[idx] = r;
}

// meta[k] = first pair of offset k, meta[K] = total pairs
__global__ void k_pair_offsets(const int32_t *__restrict__ pos, long long n, int K,
                               int32_t *__restrict__ meta) {
  pdl_sync();
  int k = threadIdx.x;
  if (k <= K) meta[k] = pos[(long long)k * n];
}

__global__ void k_emit_pairs(const int32_t *__restrict__ T, const int32_t *__restrict__ pos,
                             long long n, long long total, int32_t *__restrict__ pairs) {
  pdl_sync();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int t = T[idx];
  if (t < 0) return;
  const int p = pos[idx];
  reinterpret_cast<int2 *>(pairs)[p] = make_int2(t, (int)(idx % n));
}

// T_in[k*n_in + in] = out for every pair (each in row occurs at most once per offset)
__global__ void k_scatter_t_in(const int32_t *__restrict__ pairs, const int32_t *__restrict__ poff,
                               int K, long long n_in, int32_t *__restrict__ t_in) {
  pdl_sync();
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= poff[K]) return;
  int k = 0;
  while (k + 1 < K && poff[k + 1] <= p) ++k;
  const int2 pr = reinterpret_cast<const int2 *>(pairs)[p];
  t_in[(long long)k * n_in + pr.x] = pr.y;
}

// ---------------------------------------------------------------------------------------
// tile books
// ---------------------------------------------------------------------------------------
// mask[r] = set of offsets at which row r has a partner.  The sort key is the mask with its bits re-ordered by
// how many pairs each offset has in this rulebook - the MOST frequent offset in bit 0, the rarest in the top bit -
// so that rows are grouped by their rare offsets first and a tile's union mask stays small.  Measured on the
// benchmark building (tile rows / pairs of the 3^3 submanifold books): scale 0 1.286 -> 1.208, scale 1 1.382 ->
// 1.251, scale 2 1.192 -> 1.130 against the mask value itself as key.  pair_off (device, K + 1 running pair
// counts) may be null: the key is then the mask.
__global__ void k_row_masks(const int32_t *__restrict__ T, long long n, int K,
                            uint32_t *__restrict__ mask, uint32_t *__restrict__ key,
                            int32_t *__restrict__ idx, int key_shift, const int32_t *__restrict__ pair_off) {
  pdl_sync();
  __shared__ int s_cnt[MAX_K], s_rank[MAX_K];
  if (pair_off) {
    if (threadIdx.x < K) s_cnt[threadIdx.x] = pair_off[threadIdx.x + 1] - pair_off[threadIdx.x];
    __syncthreads();
    if (threadIdx.x < K) {
      int rk = 0;
      const int c = s_cnt[threadIdx.x];
      for (int j = 0; j < K; ++j) rk += (s_cnt[j] > c || (s_cnt[j] == c && j < (int)threadIdx.x)) ? 1 : 0;
      s_rank[threadIdx.x] = rk;
    }
    __syncthreads();
  }
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  uint32_t m = 0, km = 0;
  for (int k = 0; k < K; ++k)
    if (T[(long long)k * n + r] >= 0) {
      m |= (1u << k);
      km |= (1u << (pair_off ? s_rank[k] : k));
    }
  mask[r] = m;
  key[r] = km >> key_shift;
  idx[r] = (int)r;
}

// one block per tile: pad the permutation, OR the row masks
__global__ void __launch_bounds__(TILE_M)
k_tile_masks(const uint32_t *__restrict__ row_mask, const int32_t *__restrict__ sorted_idx,
             long long n, int32_t *__restrict__ perm, uint32_t *__restrict__ tile_mask,
             int32_t *__restrict__ tile_pop) {
  pdl_sync();
  __shared__ uint32_t wm[TILE_M / 32];
  const long long slot = (long long)blockIdx.x * TILE_M + threadIdx.x;
  uint32_t m = 0;
  int row = -1;
  if (slot < n) { row = sorted_idx[slot]; m = row_mask[row]; }
  perm[slot] = row;
  m = __reduce_or_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t u = 0;
#pragma unroll
    for (int i = 0; i < TILE_M / 32; ++i) u |= wm[i];
    tile_mask[blockIdx.x] = u;
    tile_pop[blockIdx.x] = __popc(u);
  }
}

// tiles ordered by descending number of active offsets (counting sort, one block): the persistent gather-GEMM hands
// its work items out in this order - longest first, so the last wave consists of the shortest items.  The same block
// also writes tile_off = exclusive scan of the per-tile offset counts (n_tiles + 1 entries; a few thousand tiles).
__global__ void __launch_bounds__(1024)
k_tile_order(const int32_t *__restrict__ tile_pop, const uint32_t *__restrict__ tile_mask, int n_tiles,
             int32_t *__restrict__ order, int32_t *__restrict__ n_entries, int32_t *__restrict__ tile_off) {
  pdl_sync();
  __shared__ int cnt[40], base[40];
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x < 40) cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) atomicAdd(&cnt[min(tile_pop[t], 39)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, entries = 0;                     // (a tile has at most MAX_K = 32 active offsets)
    for (int p = 39; p >= 0; --p) { base[p] = run; run += cnt[p]; entries += p * cnt[p]; }
    *n_entries = entries;                         // the entry total the host reads back (= tile_off[n_tiles])
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) order[4 * atomicAdd(&base[min(tile_pop[t], 39)], 1)] = t;
  // exclusive scan, 1024 tiles per round
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t0 = 0; t0 <= n_tiles; t0 += 1024) {
    const int t = t0 + threadIdx.x;
    const int v = t < n_tiles ? tile_pop[t] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const int w = wsum[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += u;
      }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const int c = carry;
    if (t <= n_tiles) tile_off[t] = c + wsum[wid] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + wsum[wid] + inc;
    __syncthreads();
  }
  // the work-item records: {tile, mask, first entry, 0} (one 16-byte load per work item in the gather-GEMM's loader)
  __syncthreads();
  for (int p = threadIdx.x; p < n_tiles; p += blockDim.x) {
    const int t = order[4 * p];
    order[4 * p + 1] = (int32_t)tile_mask[t];
    order[4 * p + 2] = tile_off[t];
    order[4 * p + 3] = 0;
  }
}

// ---------------------------------------------------------------------------------------
// The whole of tilebook_phase1 in ONE block for books of at most SMALL_ROWS rows (the five smallest scales of the
// backbone: their books went through ~14 launches of 3 - 5 us each - masks, three radix passes of histogram / scan /
// scatter, tile masks, tile order - i.e. launch latency, not work).  Same results as the separate kernels: the stable
// radix sort by key equals a sort by (key, row), done here as a bitonic sort of 64-bit words in shared memory.
// ---------------------------------------------------------------------------------------
constexpr int SMALL_ROWS = 8192;          // (16384 - the backbone's scale 4 - was measured: one block is then SLOWER than the
                                          // parallel kernels, build 3.86 -> 4.46 ms)
constexpr int SMALL_SMEM = SMALL_ROWS * 8 + SMALL_ROWS * 4;       // sort words + masks

__global__ void __launch_bounds__(1024)
k_tilebook_small(const int32_t *__restrict__ T, int n, int K, int key_shift, const int32_t *__restrict__ pair_off,
                 int grouping, int n_tiles, int32_t *__restrict__ perm, uint32_t *__restrict__ tile_mask,
                 int32_t *__restrict__ tile_off, int32_t *__restrict__ order, int32_t *__restrict__ n_entries) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t small_smem[];
  unsigned long long *sk = reinterpret_cast<unsigned long long *>(small_smem);      // [P] (key << 32 | row)
  uint32_t *sm = reinterpret_cast<uint32_t *>(small_smem + (size_t)SMALL_ROWS * 8);   // [n] mask of row r
  __shared__ int s_cnt[MAX_K], s_rank[MAX_K];
  __shared__ int cnt[40], base[40];
  __shared__ int s_pop[SMALL_ROWS / TILE_M + 1], s_off[SMALL_ROWS / TILE_M + 2];
  const int tid = threadIdx.x;
  if (pair_off) {
    if (tid < K) s_cnt[tid] = pair_off[tid + 1] - pair_off[tid];
    __syncthreads();
    if (tid < K) {
      int rk = 0;
      const int c = s_cnt[tid];
      for (int j = 0; j < K; ++j) rk += (s_cnt[j] > c || (s_cnt[j] == c && j < tid)) ? 1 : 0;
      s_rank[tid] = rk;
    }
  }
  if (tid < 40) cnt[tid] = 0;
  __syncthreads();
  int P = 1;
  while (P < n) P <<= 1;
  for (int r = tid; r < P; r += 1024) {
    uint32_t m = 0, km = 0;
    if (r < n) {
      for (int k = 0; k < K; ++k)
        if (T[(long long)k * n + r] >= 0) {
          m |= (1u << k);
          km |= (1u << (pair_off ? s_rank[k] : k));
        }
      sm[r] = m;
      sk[r] = ((unsigned long long)(km >> key_shift) << 32) | (unsigned)r;
    } else {
      sk[r] = ~0ull;                                   // padding sorts to the end
    }
  }
  __syncthreads();
  if (grouping) {
    for (int k2 = 2; k2 <= P; k2 <<= 1)
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < P; i += 1024) {
          const int l = i ^ j;
          if (l > i) {
            const unsigned long long a = sk[i], b = sk[l];
            const bool up = (i & k2) == 0;
            if ((a > b) == up) { sk[i] = b; sk[l] = a; }
          }
        }
        __syncthreads();
      }
  }
  // permutation (padded with -1), per-tile union masks and offset counts: one warp per tile
  const int lane = tid & 31, wid = tid >> 5;
  for (int t = wid; t < n_tiles; t += 32) {
    uint32_t u = 0;
#pragma unroll
    for (int q = 0; q < TILE_M / 32; ++q) {
      const int slot = t * TILE_M + q * 32 + lane;
      int row = -1;
      if (slot < n) { row = (int)(unsigned)sk[slot]; u |= sm[row]; }
      perm[slot] = row;
    }
    u = __reduce_or_sync(0xffffffffu, u);
    if (lane == 0) {
      tile_mask[t] = u;
      s_pop[t] = __popc(u);
      atomicAdd(&cnt[min(__popc(u), 39)], 1);
    }
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0, entries = 0;
    for (int p = 39; p >= 0; --p) { base[p] = run; run += cnt[p]; entries += p * cnt[p]; }
    *n_entries = entries;
    int acc = 0;                                       // (at most 64 tiles)
    for (int t = 0; t < n_tiles; ++t) { s_off[t] = acc; acc += s_pop[t]; }
    s_off[n_tiles] = acc;
  }
  __syncthreads();
  for (int t = tid; t <= n_tiles; t += 1024) tile_off[t] = s_off[t];
  for (int t = tid; t < n_tiles; t += 1024) {
    const int p = atomicAdd(&base[min(s_pop[t], 39)], 1);
    order[4 * p] = t;
    order[4 * p + 1] = (int32_t)tile_mask[t];
    order[4 * p + 2] = s_off[t];
    order[4 * p + 3] = 0;
  }
}

__global__ void __launch_bounds__(TILE_M)
k_fill_entries(const int32_t *__restrict__ T, long long n, const int32_t *__restrict__ perm,
               const uint32_t *__restrict__ tile_mask, const int32_t *__restrict__ tile_off,
               int32_t *__restrict__ entries) {
  pdl_sync();
  const int row = perm[(long long)blockIdx.x * TILE_M + threadIdx.x];
  uint32_t m = tile_mask[blockIdx.x];
  long long e = tile_off[blockIdx.x];
  while (m) {
    const int k = __ffs(m) - 1;
    m &= m - 1;
    entries[e * TILE_M + threadIdx.x] = row >= 0 ? T[(long long)k * n + row] : -1;
    ++e;
  }
}

// Phase 1 (no host knowledge needed): masks, grouping, per-tile offsets. meta_slot receives
// the entry total.  Phase 2 after the read-back: allocate + fill the entry lists.
static int tilebook_phase1(TileBook &tb, const int32_t *T, int K, int64_t n_rows,
                           int64_t n_partner, int32_t *meta_slot, cudaStream_t s, const int32_t *pair_off = nullptr) {
  tb.K = K;
  tb.n_rows = n_rows;
  tb.n_partner = n_partner;
  tb.n_tiles = cdiv(n_rows, TILE_M);
  if (tb.n_tiles == 0) {
    SCN_CUDA(cudaMemsetAsync(meta_slot, 0, 4, s));
    return 0;
  }
  static const int sort_bits_small = getenv("SCN_B200_SORT_BITS") ? atoi(getenv("SCN_B200_SORT_BITS")) : 27;
  static const bool small_path = !(getenv("SCN_B200_SMALL_BOOKS") && atoi(getenv("SCN_B200_SMALL_BOOKS")) == 0);
  if (small_path && n_rows <= SMALL_ROWS) {
    static bool attr_set = false;
    if (!attr_set) {
      SCN_CUDA(cudaFuncSetAttribute((const void *)k_tilebook_small, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM));
      attr_set = true;
    }
    const int kb = K < sort_bits_small ? K : sort_bits_small;
    SCN_TRY(dev_alloc_t(&tb.perm, (size_t)tb.n_tiles * TILE_M, s));
    SCN_TRY(dev_alloc_t(&tb.tile_mask, (size_t)tb.n_tiles, s));
    SCN_TRY(dev_alloc_t(&tb.tile_off, (size_t)tb.n_tiles + 1, s));
    SCN_TRY(dev_alloc_t(&tb.order, (size_t)tb.n_tiles * 4, s));
    SCN_LAUNCH(k_tilebook_small, 1, 1024, (size_t)SMALL_SMEM, s, T, (int)n_rows, K, K - kb, pair_off,
               (g_tile_grouping && K > 1 && n_rows > 4 * TILE_M) ? 1 : 0, tb.n_tiles, tb.perm, tb.tile_mask, tb.tile_off,
               tb.order, meta_slot);
    SCN_LAUNCHED();
    return 0;
  }
  uint32_t *mask = nullptr, *key = nullptr;
  int32_t *idx = nullptr, *pop = nullptr;
  SCN_TRY(dev_alloc_t(&mask, (size_t)n_rows, s));
  SCN_TRY(dev_alloc_t(&key, (size_t)n_rows, s));
  SCN_TRY(dev_alloc_t(&idx, (size_t)n_rows, s));
  // sort key = the top `sort_bits` offsets of the mask (default: all 27 = three 9-bit radix passes; 18 saves a pass but measured 1.4-1.8x more tile rows): rows that
  // agree on them end up adjacent; the low offsets are left unsorted inside such a group
  static const int sort_bits = getenv("SCN_B200_SORT_BITS") ? atoi(getenv("SCN_B200_SORT_BITS")) : 27;
  const int kbits = K < sort_bits ? K : sort_bits;
  SCN_LAUNCH(k_row_masks, cdiv(n_rows, 256), 256, 0, s, T, n_rows, K, mask, key, idx, K - kbits, pair_off);
  SCN_LAUNCHED();
  // (books of at most 4 tiles keep the natural row order: nothing to group, and the sort is ~15 launches)
  if (g_tile_grouping && K > 1 && n_rows > 4 * TILE_M) SCN_TRY(radix_sort_pairs(key, idx, n_rows, kbits, s, false));
  SCN_TRY(dev_alloc_t(&tb.perm, (size_t)tb.n_tiles * TILE_M, s));
  SCN_TRY(dev_alloc_t(&tb.tile_mask, (size_t)tb.n_tiles, s));
  SCN_TRY(dev_alloc_t(&tb.tile_off, (size_t)tb.n_tiles + 1, s));
  SCN_TRY(dev_alloc_t(&pop, (size_t)tb.n_tiles + 1, s));
  SCN_LAUNCH(k_tile_masks, tb.n_tiles, TILE_M, 0, s, mask, idx, n_rows, tb.perm, tb.tile_mask, pop);
  SCN_LAUNCHED();
  SCN_TRY(dev_alloc_t(&tb.order, (size_t)tb.n_tiles * 4, s));
  SCN_LAUNCH(k_tile_order, 1, 1024, 0, s, pop, tb.tile_mask, tb.n_tiles, tb.order, meta_slot, tb.tile_off);
  SCN_LAUNCHED();
  dev_free(mask, s);
  dev_free(key, s);
  dev_free(idx, s);
  dev_free(pop, s);
  return 0;
}

static int tilebook_phase2(TileBook &tb, const int32_t *T, int64_t n_entries, cudaStream_t s) {
  tb.n_entries = n_entries;
  SCN_TRY(dev_alloc_t(&tb.entries, (size_t)n_entries * TILE_M, s));
  if (tb.n_tiles > 0 && n_entries > 0) {
    SCN_LAUNCH(k_fill_entries, tb.n_tiles, TILE_M, 0, s, T, tb.n_rows, tb.perm, tb.tile_mask, tb.tile_off,
                                                 tb.entries);
    SCN_LAUNCHED();
  }
  tb.built = true;
  return 0;
}

// the chain of tile books over a table T [K, n_rows]: one book per group of <= MAX_K offsets (phase 1 of each;
// meta_slot + g receives the entry total of book g)
static int tilebook_chain_phase1(TileBook &head, const int32_t *T, int K, int64_t n_rows, int64_t n_partner,
                                 int32_t *meta_slot, cudaStream_t s, const int32_t *pair_off = nullptr) {
  TileBook *tb = &head;
  for (int k0 = 0, g = 0; k0 < K; k0 += MAX_K, ++g) {
    if (g > 0) {
      tb->next = new TileBook();
      tb = tb->next;
    }
    tb->k_base = k0;
    SCN_TRY(tilebook_phase1(*tb, T + (long long)k0 * n_rows, std::min(MAX_K, K - k0), n_rows, n_partner, meta_slot + g, s,
                            pair_off ? pair_off + k0 : nullptr));
  }
  return 0;
}
static int tilebook_chain_phase2(TileBook &head, const int32_t *T, const int32_t *entry_totals, int64_t n_pairs,
                                 cudaStream_t s) {
  int g = 0;
  for (TileBook *tb = &head; tb; tb = tb->next, ++g) {
    SCN_TRY(tilebook_phase2(*tb, T + (long long)tb->k_base * tb->n_rows, entry_totals[g], s));
    tb->n_pairs = n_pairs;       // (algorithmic-bytes accounting: the whole rulebook, counted on the first book only)
  }
  return 0;
}
static int n_books(int K) { return (K + MAX_K - 1) / MAX_K; }

// pairs + tb_out from t_out with a single read-back
static int finish_rulebook(RuleBook *rb, cudaStream_t s) {
  const int K = rb->K, G = n_books(K);
  const int64_t n = rb->n_out;
  const long long total = (long long)K * n;
  SCN_CHECK(total < (1LL << 31), "rulebook table too large (%lld entries)", total);
  int32_t *pos = nullptr, *meta = nullptr;
  SCN_TRY(dev_alloc_t(&pos, (size_t)total + 1, s));
  SCN_TRY(dev_alloc_t(&meta, (size_t)K + 8 + 2 * G, s));
  SCN_TRY(exclusive_scan_flags_i32(rb->t_out, pos, total, s));   // pos[i] = present partners before table entry i
  SCN_LAUNCH(k_pair_offsets, 1, MAX_KT + 32, 0, s, pos, n, K, meta);
  SCN_LAUNCHED();
  SCN_TRY(tilebook_chain_phase1(rb->tb_out, rb->t_out, K, rb->n_out, rb->n_in, meta + K + 1, s, meta));
  // strided rulebooks also carry the in-stationary lists (conv dX, deconv forward); building them
  // here shares this read-back, so the backward pass never synchronises
  const bool both = rb->kind == 1 && rb->t_in != nullptr;
  if (both) SCN_TRY(tilebook_chain_phase1(rb->tb_in, rb->t_in, K, rb->n_in, rb->n_out, meta + K + 1 + G, s, meta));
  int64_t *hs = host_scratch((size_t)(K + 8 + 2 * G) / 2 + 8);
  int32_t *h32 = (int32_t *)hs;
  SCN_CUDA(cudaMemcpyAsync(h32, meta, (size_t)(K + 1 + 2 * G) * 4, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));  // documented read-back: pair counts + entry totals
  for (int k = 0; k <= K; ++k) rb->pair_off[k] = h32[k];
  for (int k = 0; k < K; ++k) rb->counts[k] = h32[k + 1] - h32[k];
  rb->total_pairs = h32[K];
  SCN_TRY(dev_alloc_t(&rb->pairs, (size_t)rb->total_pairs * 2, s));
  if (total > 0) {
    SCN_LAUNCH(k_emit_pairs, cdiv(total, 256), 256, 0, s, rb->t_out, pos, n, total, rb->pairs);
    SCN_LAUNCHED();
  }
  SCN_TRY(tilebook_chain_phase2(rb->tb_out, rb->t_out, h32 + K + 1, rb->total_pairs, s));
  if (both) SCN_TRY(tilebook_chain_phase2(rb->tb_in, rb->t_in, h32 + K + 1 + G, rb->total_pairs, s));
  SCN_TRY(ensure_dw_work(rb, s));
  dev_free(pos, s);
  dev_free(meta, s);
  return 0;
}

static int build_t_in(RuleBook *rb, cudaStream_t s) {
  if (rb->t_in) return 0;
  const long long total = (long long)rb->K * rb->n_in;
  SCN_TRY(dev_alloc_t(&rb->t_in, (size_t)total, s));
  SCN_CUDA(cudaMemsetAsync(rb->t_in, 0xFF, (size_t)(total ? total : 1) * 4, s));
  if (rb->total_pairs > 0) {
    int32_t *poff = nullptr;
    SCN_TRY(dev_alloc_t(&poff, (size_t)rb->K + 1, s));
    int32_t h[MAX_KT + 1];
    for (int k = 0; k <= rb->K; ++k) h[k] = (int32_t)rb->pair_off[k];
    // pageable source: the copy is staged before return, so the stack array may die
    SCN_CUDA(cudaMemcpyAsync(poff, h, (size_t)(rb->K + 1) * 4, cudaMemcpyHostToDevice, s));
    SCN_LAUNCH(k_scatter_t_in, cdiv(rb->total_pairs, 256), 256, 0, s, rb->pairs, poff, rb->K, rb->n_in,
                                                             rb->t_in);
    SCN_LAUNCHED();
    dev_free(poff, s);
  }
  return 0;
}

int ensure_tilebook(RuleBook *rb, bool stationary_out, cudaStream_t s) {
  TileBook &tb = stationary_out ? rb->tb_out : rb->tb_in;
  if (tb.built) return 0;
  if (rb->identity) {
    tb.identity = true;
    tb.K = 1;
    tb.n_rows = rb->n_out;
    tb.n_partner = rb->n_in;
    tb.n_tiles = cdiv(rb->n_out, TILE_M);
    tb.n_pairs = rb->n_out;
    tb.built = true;
    return 0;
  }
  SCN_CHECK(!stationary_out, "internal: tb_out must be built with the rulebook");
  SCN_TRY(build_t_in(rb, s));
  const int G = n_books(rb->K);
  int32_t *meta = nullptr;
  SCN_TRY(dev_alloc_t(&meta, (size_t)G + 4, s));
  SCN_TRY(tilebook_chain_phase1(tb, rb->t_in, rb->K, rb->n_in, rb->n_out, meta, s));
  int32_t *h32 = (int32_t *)host_scratch((size_t)G / 2 + 8);
  SCN_CUDA(cudaMemcpyAsync(h32, meta, (size_t)G * 4, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));  // documented read-back (first backward use only)
  SCN_TRY(tilebook_chain_phase2(tb, rb->t_in, h32, rb->total_pairs, s));
  dev_free(meta, s);
  return 0;
}

int ensure_dw_work(RuleBook *rb, cudaStream_t s) {
  if (rb->dw_work || rb->total_pairs == 0) return 0;
  // ~2 waves of work items: every item costs one Cin x Cout partial (written + re-read by the reduce)
  long long chunk = (rb->total_pairs + 2LL * num_sms() - 1) / (2LL * num_sms());
  chunk = (chunk + 63) / 64 * 64;
  chunk = std::min<long long>(std::max<long long>(chunk, 512), 16384);
  // every offset rounds its item count up: grow the chunk until the total fits 2 full waves (the kernel
  // holds one CTA per SM; 2 x 148 + 12 items ran as three waves, the last one almost empty)
  for (;;) {
    long long items = 0;
    for (int k = 0; k < rb->K; ++k) items += (rb->counts[k] + chunk - 1) / chunk;
    if (items <= 2LL * num_sms() || chunk >= 16384) break;
    chunk += 64;
  }
  std::vector<DwWork> &w = rb->dw_host;
  w.clear();
  for (int k = 0; k < rb->K; ++k) {
    int slot = 0;
    for (long long st = 0; st < rb->counts[k]; st += chunk) {
      DwWork d;
      d.k = k;
      d.start = (int32_t)(rb->pair_off[k] + st);
      d.len = (int32_t)std::min<long long>(chunk, rb->counts[k] - st);
      d.slot = slot++;                       // rewritten below
      w.push_back(d);
    }
  }
  // slot = index of the item's partial in offset-major order (what k_dw_reduce sums); the LAUNCH order
  // interleaves the offsets by position inside their pair lists: pairs are sorted by output row, so the
  // items that run together then read the same neighbourhood of X and dY rows and find them in L2
  // (offset-major order streamed both matrices from HBM once per offset: 2.7x the algorithmic bytes)
  for (size_t i = 0; i < w.size(); ++i) w[i].slot = (int32_t)i;
  {
    std::vector<std::pair<double, int>> key(w.size());
    for (size_t i = 0; i < w.size(); ++i) {
      const DwWork &d = w[i];
      const double cnt = (double)std::max<int64_t>(rb->counts[d.k], 1);
      key[i] = {((double)(d.start - rb->pair_off[d.k]) + 0.5 * d.len) / cnt, (int)i};
    }
    std::stable_sort(key.begin(), key.end(), [](const std::pair<double, int> &a, const std::pair<double, int> &b) { return a.first < b.first; });
    std::vector<DwWork> r(w.size());
    for (size_t i = 0; i < w.size(); ++i) r[i] = w[key[i].second];
    w.swap(r);
  }
  rb->n_dw_work = (int)w.size();
  rb->dw_chunk = (int)chunk;
  SCN_TRY(dev_alloc_t(&rb->dw_work, w.size(), s));
  // pageable source: staged by the driver before the call returns; rb->dw_host outlives it anyway
  SCN_CUDA(cudaMemcpyAsync(rb->dw_work, w.data(), w.size() * sizeof(DwWork),
                           cudaMemcpyHostToDevice, s));
  return 0;
}

// ---------------------------------------------------------------------------------------
// submanifold rulebook
// ---------------------------------------------------------------------------------------
static RuleBook *find_rulebook(scn_metadata *m, int kind, const int64_t *in_ss,
                               const int64_t *filter, const int64_t *stride) {
  for (RuleBook *rb : m->rulebooks)
    if (rb->kind == kind && same3(rb->in_ss, in_ss) && same3(rb->filter, filter) &&
        (kind == 0 || same3(rb->stride, stride)))
      return rb;
  return nullptr;
}

int get_submanifold_rulebook(scn_metadata *m, const int64_t *ss, const int64_t *filter,
                             cudaStream_t s, RuleBook **out) {
  m->last_stream = s;
  if (RuleBook *rb = find_rulebook(m, 0, ss, filter, nullptr)) { *out = rb; return 0; }
  Grid *g = find_grid(m, ss);
  SCN_CHECK(g, "no active sites at spatial size [%lld,%lld,%lld]", (long long)ss[0],
            (long long)ss[1], (long long)ss[2]);
  const int64_t K64 = filter[0] * filter[1] * filter[2];
  SCN_CHECK(filter[0] > 0 && filter[1] > 0 && filter[2] > 0 && K64 <= MAX_KT,
            "filter volume %lld not in 1..%d", (long long)K64, MAX_KT);
  RuleBook *rb = new RuleBook();
  rb->kind = 0;
  memcpy(rb->in_ss, ss, 24); memcpy(rb->out_ss, ss, 24); memcpy(rb->filter, filter, 24);
  rb->stride[0] = rb->stride[1] = rb->stride[2] = 1;
  rb->K = (int)K64;
  rb->n_in = rb->n_out = g->n_active;
  memset(rb->counts, 0, sizeof(rb->counts));
  memset(rb->pair_off, 0, sizeof(rb->pair_off));
  m->rulebooks.push_back(rb);
  BuildGuard guard{m, rb, nullptr, s};
  if (rb->K == 1) {
    // 1x1x1: pairs are (i,i); no hashing, no lists (the reference still probes, same result)
    rb->identity = true;
    rb->counts[0] = g->n_active;
    rb->pair_off[1] = g->n_active;
    rb->total_pairs = g->n_active;
    SCN_TRY(ensure_tilebook(rb, true, s));
    guard.ok = true;
    *out = rb;
    return 0;
  }
  const long long total = (long long)rb->K * g->n_active;
  SCN_CHECK(total < (1LL << 31), "submanifold table too large");
  prof_begin(PROF_RULES, s);
  SCN_TRY(dev_alloc_t(&rb->t_out, (size_t)total, s));
  if (total > 0) {
    Filter3 f;
    for (int d = 0; d < 3; ++d) { f.size[d] = (int)filter[d]; f.stride[d] = 1; f.out_size[d] = (int)ss[d]; }
    SCN_LAUNCH(k_sub_table, cdiv(total, 256), 256, 0, s, g->coords, g->n_active, f, rb->K, g->hkeys,
                                                 g->hvals, g->hcap - 1, rb->t_out);
    SCN_LAUNCHED();
  }
  SCN_TRY(finish_rulebook(rb, s));
  // algorithmic bytes: K probes x 12 B + one 12 B entry per site
  prof_end(PROF_RULES, s, (double)g->n_active * (12.0 + 12.0 * rb->K), 0);
  guard.ok = true;
  *out = rb;
  return 0;
}

// ---------------------------------------------------------------------------------------
// strided convolution rulebook + output grid (ConvolutionRules.h:12-105)
// ---------------------------------------------------------------------------------------
struct Region { int lb[3], cnt[3]; };
__device__ __forceinline__ Region out_region(const int4 c, const Filter3 &f) {
  Region r;
  const int in[3] = {c.x, c.y, c.z};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    // RectangularRegions.h:111-119 (C++ truncating division, computed in 64 bit)
    long long lb = ((long long)in[d] - f.size[d] + f.stride[d]) / f.stride[d];
    if (lb < 0) lb = 0;
    long long ub = in[d] / f.stride[d];
    if (ub > f.out_size[d] - 1) ub = f.out_size[d] - 1;
    r.lb[d] = (int)lb;
    r.cnt[d] = (int)(ub - lb + 1);
  }
  return r;
}

// jj enumerates the (at most R = Rx*Ry*Rz) output cells of input i, last dim fastest
__device__ __forceinline__ bool region_cell(const Region &r, int jj, const int R3[3], int j[3]) {
  const int jz = jj % R3[2], jy = (jj / R3[2]) % R3[1], jx = jj / (R3[2] * R3[1]);
  if (jx >= r.cnt[0] || jy >= r.cnt[1] || jz >= r.cnt[2]) return false;
  j[0] = r.lb[0] + jx; j[1] = r.lb[1] + jy; j[2] = r.lb[2] + jz;
  return true;
}

struct R3s { int v[3]; };

__global__ void k_conv_insert(const int32_t *__restrict__ coords, long long n_in, Filter3 f,
                              R3s R3, int R, uint64_t *hk, int32_t *hv, uint32_t mask,
                              int32_t *__restrict__ pslot) {
  pdl_sync();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_in * R) return;
  const long long i = idx / R;
  const int jj = (int)(idx - i * R);
  const int4 c = reinterpret_cast<const int4 *>(coords)[i];
  const Region r = out_region(c, f);
  int j[3];
  if (!region_cell(r, jj, R3.v, j)) { pslot[idx] = -1; return; }
  const uint32_t slot = hash_insert(hk, mask, pack_key(j[0], j[1], j[2], c.w));
  pslot[idx] = (int32_t)slot;
  atomicMin(&hv[slot], (int)idx);
}

__global__ void k_conv_flag(const int32_t *__restrict__ pslot, const int32_t *__restrict__ hv,
                            int32_t *__restrict__ flag, long long total) {
  pdl_sync();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int sl = pslot[idx];
  flag[idx] = (sl >= 0 && hv[sl] == (int)idx) ? 1 : 0;
}

__global__ void k_conv_assign(const int32_t *__restrict__ coords, long long n_in, Filter3 f,
                              R3s R3, int R, const int32_t *__restrict__ rank,
                              const int32_t *__restrict__ pslot, int32_t *hv,
                              int32_t *__restrict__ out_coords) {
  pdl_sync();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_in * R) return;
  const int r0 = rank[idx];
  if (rank[idx + 1] == r0) return;
  const long long i = idx / R;
  const int jj = (int)(idx - i * R);
  const int4 c = reinterpret_cast<const int4 *>(coords)[i];
  const Region r = out_region(c, f);
  int j[3];
  region_cell(r, jj, R3.v, j);
  reinterpret_cast<int4 *>(out_coords)[r0] = make_int4(j[0], j[1], j[2], c.w);
  hv[pslot[idx]] = r0;
}

// relabel rows by a permutation new_of_old (batch-contiguity fix-up)
__global__ void k_relabel_table(int32_t *hv, uint32_t cap, const uint64_t *__restrict__ hk,
                                const int32_t *__restrict__ new_of_old) {
  pdl_sync();
  uint32_t sl = blockIdx.x * blockDim.x + threadIdx.x;
  if (sl < cap && hk[sl] != EMPTY_KEY) hv[sl] = new_of_old[hv[sl]];
}
__global__ void k_batch_keys(const int32_t *__restrict__ coords, long long n,
                             uint32_t *__restrict__ key, int32_t *__restrict__ idx) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  key[i] = (uint32_t)coords[i * 4 + 3];
  idx[i] = (int)i;
}
__global__ void k_permute_coords(const int32_t *__restrict__ src, const int32_t *__restrict__ old_of_new,
                                 int32_t *__restrict__ dst, int32_t *__restrict__ new_of_old,
                                 long long n) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = old_of_new[i];
  reinterpret_cast<int4 *>(dst)[i] = reinterpret_cast<const int4 *>(src)[o];
  new_of_old[o] = (int)i;
}

__global__ void k_conv_tables(const int32_t *__restrict__ coords, long long n_in, long long n_out,
                              Filter3 f, R3s R3, int R, const int32_t *__restrict__ pslot,
                              const int32_t *__restrict__ hv, int32_t *__restrict__ t_out,
                              int32_t *__restrict__ t_in) {
  pdl_sync();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_in * R) return;
  const int sl = pslot[idx];
  if (sl < 0) return;
  const long long i = idx / R;
  const int jj = (int)(idx - i * R);
  const int4 c = reinterpret_cast<const int4 *>(coords)[i];
  const Region r = out_region(c, f);
  int j[3];
  region_cell(r, jj, R3.v, j);
  // offset of the input inside the output's window (RectangularRegions.h:31-38)
  const int ox = c.x - j[0] * f.stride[0], oy = c.y - j[1] * f.stride[1], oz = c.z - j[2] * f.stride[2];
  const int k = (ox * f.size[1] + oy) * f.size[2] + oz;
  const int o = hv[sl];
  t_out[(long long)k * n_out + o] = (int)i;
  t_in[(long long)k * n_in + i] = o;
}

int get_conv_rulebook(scn_metadata *m, const int64_t *in_ss, const int64_t *out_ss,
                      const int64_t *filter, const int64_t *stride, cudaStream_t s,
                      RuleBook **out) {
  m->last_stream = s;
  if (RuleBook *rb = find_rulebook(m, 1, in_ss, filter, stride)) { *out = rb; return 0; }
  Grid *gi = find_grid(m, in_ss);
  SCN_CHECK(gi, "no active sites at spatial size [%lld,%lld,%lld]", (long long)in_ss[0],
            (long long)in_ss[1], (long long)in_ss[2]);
  const int64_t K64 = filter[0] * filter[1] * filter[2];
  SCN_CHECK(K64 >= 1 && K64 <= MAX_KT, "filter volume %lld not in 1..%d", (long long)K64, MAX_KT);
  Filter3 f;
  R3s R3;
  int R = 1;
  for (int d = 0; d < 3; ++d) {
    SCN_CHECK(filter[d] > 0 && stride[d] > 0 && out_ss[d] > 0, "bad filter/stride/output size");
    f.size[d] = (int)filter[d]; f.stride[d] = (int)stride[d]; f.out_size[d] = (int)out_ss[d];
    R3.v[d] = (int)((filter[d] + stride[d] - 1) / stride[d]);
    R *= R3.v[d];
  }
  const int64_t n_in = gi->n_active;
  const long long total = (long long)n_in * R;
  SCN_CHECK(total < (1LL << 30), "strided rulebook too large");
  SCN_CHECK(!same3(in_ss, out_ss), "convolution input and output spatial sizes coincide");
  // (re)create the output grid (ConvolutionRules.h:66-70 clears it)
  for (size_t i = 0; i < m->grids.size(); ++i)
    if (same3(m->grids[i]->ss, out_ss) && m->grids[i] != gi) {
      grid_free(m->grids[i], s);
      m->grids.erase(m->grids.begin() + i);
      break;
    }
  Grid *go = new Grid();
  memcpy(go->ss, out_ss, 24);
  m->grids.push_back(go);
  RuleBook *rb = new RuleBook();
  rb->kind = 1;
  memcpy(rb->in_ss, in_ss, 24); memcpy(rb->out_ss, out_ss, 24);
  memcpy(rb->filter, filter, 24); memcpy(rb->stride, stride, 24);
  rb->K = (int)K64;
  rb->n_in = n_in;
  memset(rb->counts, 0, sizeof(rb->counts));
  memset(rb->pair_off, 0, sizeof(rb->pair_off));
  m->rulebooks.push_back(rb);
  BuildGuard guard{m, rb, go, s};

  prof_begin(PROF_RULES, s);
  SCN_TRY(grid_alloc_table(go, total, 0x7F, s));
  int32_t *pslot = nullptr, *rank = nullptr;
  SCN_TRY(dev_alloc_t(&pslot, (size_t)total, s));
  SCN_TRY(dev_alloc_t(&rank, (size_t)total + 1, s));
  int64_t n_out = 0;
  if (total > 0) {
    SCN_LAUNCH(k_conv_insert, cdiv(total, 256), 256, 0, s, gi->coords, n_in, f, R3, R, go->hkeys,
                                                   go->hvals, go->hcap - 1, pslot);
    SCN_LAUNCHED();
    SCN_LAUNCH(k_conv_flag, cdiv(total, 256), 256, 0, s, pslot, go->hvals, rank, total);
    SCN_LAUNCHED();
    SCN_TRY(exclusive_scan_i32(rank, rank, total, s));
    int32_t *h32 = (int32_t *)host_scratch(16);
    SCN_CUDA(cudaMemcpyAsync(h32, rank + total, 4, cudaMemcpyDeviceToHost, s));
    SCN_CUDA(cudaStreamSynchronize(s));  // documented read-back: nActive of the new scale
    n_out = h32[0];
  }
  go->n_active = n_out;
  rb->n_out = n_out;
  SCN_TRY(dev_alloc_t(&go->coords, (size_t)n_out * 4, s));
  if (total > 0) {
    SCN_LAUNCH(k_conv_assign, cdiv(total, 256), 256, 0, s, gi->coords, n_in, f, R3, R, rank, pslot,
                                                   go->hvals, go->coords);
    SCN_LAUNCHED();
  }
  if (!gi->batch_sorted && n_out > 1) {
    // input rows interleave samples: stable-sort the new sites by sample index so that this
    // scale is batch-contiguous ascending, as the reference's per-sample loop guarantees
    uint32_t *bk = nullptr;
    int32_t *old_of_new = nullptr, *new_of_old = nullptr, *c2 = nullptr;
    SCN_TRY(dev_alloc_t(&bk, (size_t)n_out, s));
    SCN_TRY(dev_alloc_t(&old_of_new, (size_t)n_out, s));
    SCN_TRY(dev_alloc_t(&new_of_old, (size_t)n_out, s));
    SCN_TRY(dev_alloc_t(&c2, (size_t)n_out * 4, s));
    SCN_LAUNCH(k_batch_keys, cdiv(n_out, 256), 256, 0, s, go->coords, n_out, bk, old_of_new);
    SCN_LAUNCHED();
    SCN_TRY(radix_sort_pairs(bk, old_of_new, n_out, 16, s));
    SCN_LAUNCH(k_permute_coords, cdiv(n_out, 256), 256, 0, s, go->coords, old_of_new, c2, new_of_old, n_out);
    SCN_LAUNCHED();
    SCN_LAUNCH(k_relabel_table, cdiv(go->hcap, 256), 256, 0, s, go->hvals, go->hcap, go->hkeys, new_of_old);
    SCN_LAUNCHED();
    dev_free(go->coords, s);
    go->coords = c2;
    dev_free(bk, s);
    dev_free(old_of_new, s);
    dev_free(new_of_old, s);
  }
  go->batch_sorted = true;
  const long long t_out_n = (long long)rb->K * n_out, t_in_n = (long long)rb->K * n_in;
  SCN_TRY(dev_alloc_t(&rb->t_out, (size_t)t_out_n, s));
  SCN_TRY(dev_alloc_t(&rb->t_in, (size_t)t_in_n, s));
  SCN_CUDA(cudaMemsetAsync(rb->t_out, 0xFF, (size_t)(t_out_n ? t_out_n : 1) * 4, s));
  SCN_CUDA(cudaMemsetAsync(rb->t_in, 0xFF, (size_t)(t_in_n ? t_in_n : 1) * 4, s));
  if (total > 0) {
    SCN_LAUNCH(k_conv_tables, cdiv(total, 256), 256, 0, s, gi->coords, n_in, n_out, f, R3, R, pslot,
                                                   go->hvals, rb->t_out, rb->t_in);
    SCN_LAUNCHED();
  }
  dev_free(pslot, s);
  dev_free(rank, s);
  SCN_TRY(finish_rulebook(rb, s));
  prof_end(PROF_RULES, s, (double)n_in * (12.0 + 12.0 * R) + 12.0 * n_out, 0);
  guard.ok = true;
  *out = rb;
  return 0;
}

}  // namespace scn

// =========================================================================================
// C ABI
// =========================================================================================
using namespace scn;

extern "C" {

int scn_set_tile_grouping(int enabled) {
  g_tile_grouping = enabled != 0;
  return 0;
}

int scn_metadata_create(int dimension, scn_metadata_t **out) {
  SCN_CHECK(out, "null output pointer");
  SCN_CHECK(dimension == 3, "only Metadata_3 (dimension 3) is implemented, got %d", dimension);
  *out = new scn_metadata();
  (*out)->dim = dimension;
  return 0;
}

void scn_metadata_destroy(scn_metadata_t *m) {
  if (!m) return;
  metadata_clear(m, m->last_stream);
  delete m;
}

int scn_metadata_clear(scn_metadata_t *m, void *stream) {
  SCN_CHECK(m, "null metadata");
  metadata_clear(m, (cudaStream_t)stream);
  return 0;
}

int scn_get_nactive(scn_metadata_t *m, const int64_t *ss, int64_t *n_active) {
  SCN_CHECK(m && ss && n_active, "null argument");
  Grid *g = find_grid(m, ss);
  *n_active = g ? g->n_active : -1;
  return 0;
}

int scn_get_batch_size(scn_metadata_t *m, int64_t *batch_size) {
  SCN_CHECK(m && batch_size, "null argument");
  *batch_size = m->batch_size;
  return 0;
}

int scn_get_spatial_locations_device(scn_metadata_t *m, const int64_t *ss, int32_t *out_dev,
                                     void *stream) {
  SCN_CHECK(m && ss, "null argument");
  Grid *g = find_grid(m, ss);
  SCN_CHECK(g, "no active sites at that spatial size");
  if (g->n_active)
    SCN_CUDA(cudaMemcpyAsync(out_dev, g->coords, (size_t)g->n_active * 16,
                             cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int scn_get_spatial_locations(scn_metadata_t *m, const int64_t *ss, int64_t *out_host,
                              void *stream) {
  SCN_CHECK(m && ss, "null argument");
  Grid *g = find_grid(m, ss);
  SCN_CHECK(g, "no active sites at that spatial size");
  if (g->n_active == 0) return 0;
  std::vector<int32_t> tmp((size_t)g->n_active * 4);
  SCN_CUDA(cudaMemcpyAsync(tmp.data(), g->coords, tmp.size() * 4, cudaMemcpyDeviceToHost,
                           (cudaStream_t)stream));
  SCN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  for (size_t i = 0; i < tmp.size(); ++i) out_host[i] = tmp[i];
  return 0;
}

int scn_input_layer_prepare(scn_metadata_t *m, const int64_t *spatial_size, const int64_t *coords,
                            int64_t n_points, int n_cols, int coords_on_device,
                            int64_t batch_size, int mode, void *stream, int64_t *n_active) {
  SCN_CHECK(m && spatial_size && n_active && (coords || n_points == 0), "null argument");
  return input_layer_prepare(m, spatial_size, coords, n_points, n_cols, coords_on_device,
                             batch_size, mode, (cudaStream_t)stream, n_active);
}

int scn_build_plan(scn_metadata_t *m, const int64_t *spatial_size, const int64_t *coords,
                   int64_t n_points, int n_cols, int coords_on_device, int64_t batch_size, int mode,
                   const int64_t *plan, int n_ops, void *stream, int64_t *n_active) {
  SCN_CHECK(m && spatial_size && n_active && (coords || n_points == 0) && (plan || n_ops == 0), "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  SCN_TRY(input_layer_prepare(m, spatial_size, coords, n_points, n_cols, coords_on_device, batch_size, mode, s,
                              n_active));
  for (int i = 0; i < n_ops; ++i) {
    const int64_t *op = plan + (size_t)i * 13;
    RuleBook *rb = nullptr;
    if (op[0] == 0) SCN_TRY(get_submanifold_rulebook(m, op + 1, op + 7, s, &rb));
    else SCN_TRY(get_conv_rulebook(m, op + 1, op + 4, op + 7, op + 10, s, &rb));
  }
  return 0;
}

int scn_input_rulebook_header(scn_metadata_t *m, int64_t header[4], void *stream) {
  SCN_CHECK(m && m->input.built, "input layer not prepared");
  InputRules &ir = m->input;
  if (ir.max_active < 0) {
    int32_t *h32 = (int32_t *)host_scratch(16);
    SCN_CUDA(cudaMemcpyAsync(h32, ir.stat + 4, 4, cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    SCN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    ir.max_active = std::max(1, h32[0]);
    if (ir.mode == 1 || ir.mode == 2) ir.max_active = 1;
  }
  header[0] = ir.mode;
  header[1] = ir.n_points ? ir.max_active : 0;
  header[2] = ir.n_points;
  header[3] = ir.n_active;
  return 0;
}

int scn_input_rulebook_copy(scn_metadata_t *m, int32_t *out_host, void *stream) {
  int64_t hd[4];
  SCN_TRY(scn_input_rulebook_header(m, hd, stream));
  InputRules &ir = m->input;
  if (ir.mode == 0 || ir.n_active == 0) return 0;
  std::vector<int32_t> off((size_t)ir.n_active + 1), mem((size_t)ir.n_points);
  cudaStream_t s = (cudaStream_t)stream;
  SCN_CUDA(cudaMemcpyAsync(off.data(), ir.csr_off, off.size() * 4, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaMemcpyAsync(mem.data(), ir.members, mem.size() * 4, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));
  const int64_t w = 1 + hd[1];
  memset(out_host, 0, (size_t)(ir.n_active * w) * 4);
  for (int64_t r = 0; r < ir.n_active; ++r) {
    int32_t *row = out_host + r * w;
    const int b = off[r], e = off[r + 1];
    if (ir.mode == 1) { row[0] = 1; row[1] = mem[b]; }          // IOLayersRules.h:100-105 front()
    else if (ir.mode == 2) { row[0] = 1; row[1] = mem[e - 1]; }  // :106-111 back()
    else { row[0] = e - b; for (int i = b; i < e; ++i) row[1 + i - b] = mem[i]; }
  }
  return 0;
}

int scn_submanifold_rulebook_prepare(scn_metadata_t *m, const int64_t *ss, const int64_t *filter,
                                     void *stream, int64_t *counts_host) {
  SCN_CHECK(m && ss && filter, "null argument");
  RuleBook *rb = nullptr;
  SCN_TRY(get_submanifold_rulebook(m, ss, filter, (cudaStream_t)stream, &rb));
  if (counts_host)
    for (int k = 0; k < rb->K; ++k) counts_host[k] = rb->counts[k];
  return 0;
}

int scn_conv_rulebook_prepare(scn_metadata_t *m, const int64_t *in_ss, const int64_t *out_ss,
                              const int64_t *filter, const int64_t *stride, void *stream,
                              int64_t *n_out_active, int64_t *counts_host) {
  SCN_CHECK(m && in_ss && out_ss && filter && stride, "null argument");
  RuleBook *rb = nullptr;
  SCN_TRY(get_conv_rulebook(m, in_ss, out_ss, filter, stride, (cudaStream_t)stream, &rb));
  if (n_out_active) *n_out_active = rb->n_out;
  if (counts_host)
    for (int k = 0; k < rb->K; ++k) counts_host[k] = rb->counts[k];
  return 0;
}

static int copy_pairs(RuleBook *rb, int64_t offset, int32_t *pairs_host, cudaStream_t s) {
  SCN_CHECK(offset >= 0 && offset < rb->K, "offset %lld outside filter volume %d",
            (long long)offset, rb->K);
  const int64_t c = rb->counts[offset];
  if (c == 0) return 0;
  if (rb->identity) {
    for (int64_t i = 0; i < c; ++i) { pairs_host[2 * i] = (int32_t)i; pairs_host[2 * i + 1] = (int32_t)i; }
    return 0;
  }
  SCN_CUDA(cudaMemcpyAsync(pairs_host, rb->pairs + 2 * rb->pair_off[offset], (size_t)c * 8,
                           cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int scn_submanifold_rulebook_copy(scn_metadata_t *m, const int64_t *ss, const int64_t *filter,
                                  int64_t offset, int32_t *pairs_host, void *stream) {
  RuleBook *rb = nullptr;
  SCN_TRY(get_submanifold_rulebook(m, ss, filter, (cudaStream_t)stream, &rb));
  return copy_pairs(rb, offset, pairs_host, (cudaStream_t)stream);
}

int scn_conv_rulebook_copy(scn_metadata_t *m, const int64_t *in_ss, const int64_t *filter,
                           const int64_t *stride, int64_t offset, int32_t *pairs_host,
                           void *stream) {
  RuleBook *rb = find_rulebook(m, 1, in_ss, filter, stride);
  SCN_CHECK(rb, "convolution rulebook has not been prepared");
  return copy_pairs(rb, offset, pairs_host, (cudaStream_t)stream);
}

int scn_sparse_to_dense_rules_copy(scn_metadata_t *m, const int64_t *ss, int32_t *rules_host,
                                   int32_t *sample_host, void *stream) {
  Grid *g = find_grid(m, ss);
  SCN_CHECK(g, "no active sites at that spatial size");
  if (g->n_active == 0) return 0;
  std::vector<int32_t> c((size_t)g->n_active * 4);
  SCN_CUDA(cudaMemcpyAsync(c.data(), g->coords, c.size() * 4, cudaMemcpyDeviceToHost,
                           (cudaStream_t)stream));
  SCN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  for (int64_t r = 0; r < g->n_active; ++r) {
    rules_host[2 * r] = (int32_t)r;
    rules_host[2 * r + 1] = (int32_t)((c[4 * r] * ss[1] + c[4 * r + 1]) * ss[2] + c[4 * r + 2]);
    sample_host[r] = c[4 * r + 3];
  }
  return 0;
}

int scn_rulebook_stats(scn_metadata_t *m, int kind, const int64_t *in_ss, const int64_t *filter,
                       const int64_t *stride, int64_t stats[3]) {
  SCN_CHECK(m && in_ss && filter && stats, "null argument");
  RuleBook *rb = find_rulebook(m, kind == 0 ? 0 : 1, in_ss, filter, stride);
  SCN_CHECK(rb, "rulebook not prepared");
  const bool out_side = (kind == 0 || kind == 1 || kind == 4);
  TileBook &tb = out_side ? rb->tb_out : rb->tb_in;
  SCN_CHECK(tb.built, "tile book not built yet");
  stats[0] = rb->total_pairs;
  stats[1] = 0;
  for (const TileBook *b = &tb; b; b = b->next)
    stats[1] += b->identity ? (int64_t)b->n_tiles * TILE_M : b->n_entries * TILE_M;
  stats[2] = tb.n_tiles;
  return 0;
}

}  // extern "C"
