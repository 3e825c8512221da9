// metadata.cuh - device-resident Metadata<3>: per-scale hash grids, rulebooks, tile books.
// Mirrors the role of SparseConvNet/sparseconvnet/SCN/Metadata/Metadata.h:44-163, but every
// grid, rulebook and gather list lives in HBM; the host object only owns pointers and counts.
#pragma once
#include "common.cuh"
#include <vector>

namespace scn {

constexpr int TILE_M = 128;       // output rows per gather-GEMM tile
constexpr int MAX_K = 32;         // offsets per tile book (tile masks are 32-bit); larger filters chain tile books
constexpr int MAX_KT = 512;       // filter volume limit of a rulebook (8^3; the reference has none)
constexpr uint64_t EMPTY_KEY = ~0ULL;

#ifdef __CUDACC__
// ---- hash grid: open addressing, linear probing, key = batch<<48 | x<<32 | y<<16 | z ------------------------
__device__ __forceinline__ uint64_t pack_key(int x, int y, int z, int b) {
  return ((uint64_t)(uint32_t)b << 48) | ((uint64_t)(uint32_t)x << 32) |
         ((uint64_t)(uint32_t)y << 16) | (uint64_t)(uint32_t)z;
}
__device__ __forceinline__ bool coord_ok(int x, int y, int z) {
  return ((unsigned)x | (unsigned)y | (unsigned)z) < 65536u;
}
__device__ __forceinline__ uint32_t hash_insert(uint64_t *hk, uint32_t mask, uint64_t key) {
  uint32_t slot = (uint32_t)mix64(key) & mask;
  while (true) {
    unsigned long long prev =
        atomicCAS((unsigned long long *)&hk[slot], (unsigned long long)EMPTY_KEY,
                  (unsigned long long)key);
    if (prev == EMPTY_KEY || prev == key) return slot;
    slot = (slot + 1) & mask;
  }
}
__device__ __forceinline__ int hash_find(const uint64_t *__restrict__ hk,
                                         const int32_t *__restrict__ hv, uint32_t mask,
                                         uint64_t key) {
  uint32_t slot = (uint32_t)mix64(key) & mask;
  while (true) {
    const uint64_t k = hk[slot];
    if (k == key) return hv[slot];
    if (k == EMPTY_KEY) return -1;
    slot = (slot + 1) & mask;
  }
}
#endif

struct Grid {                     // one spatial scale (Metadata.h:28-34 SparseGrids)
  int64_t ss[3] = {0, 0, 0};
  int64_t n_active = 0;
  int32_t *coords = nullptr;      // [n_active,4] x,y,z,batch
  uint64_t *hkeys = nullptr;      // open-addressing table: packed key
  int32_t *hvals = nullptr;       //                        site row
  uint32_t hcap = 0;              // power of two
  bool batch_sorted = true;       // rows are batch-contiguous ascending
};

// Output-stationary gather lists: tiles of TILE_M "stationary" rows; for every kernel offset
// active in the tile, TILE_M partner rows (-1 = none).
struct TileBook {
  bool built = false;
  bool identity = false;          // 1x1x1 filter: partner(row) = row, no lists kept
  int K = 0;                      // offsets covered by THIS book (<= MAX_K)
  int k_base = 0;                 // first filter offset of this book
  TileBook *next = nullptr;       // filters of more than MAX_K offsets: one book per group of <= MAX_K offsets (the
                                  // gather-GEMM runs once per book and sums; FPN_Net's [1,1,64] z-collapse, 4^3, 5^3)
  int64_t n_rows = 0;             // stationary rows
  int64_t n_partner = 0;          // rows of the gathered side
  int n_tiles = 0;
  int64_t n_entries = 0;
  int64_t n_pairs = 0;            // (in,out) pairs behind the lists (algorithmic-bytes accounting)
  int32_t *perm = nullptr;        // [n_tiles*TILE_M] stationary row per slot (-1 pad)
  uint32_t *tile_mask = nullptr;  // [n_tiles] union of active offsets
  int32_t *tile_off = nullptr;    // [n_tiles+1] first entry of the tile
  int32_t *order = nullptr;       // [n_tiles][4] work-item order of the gather-GEMM: tiles by descending number of active
                                  // offsets, each with its mask and first entry {tile, mask, tile_off[tile], 0}
  int32_t *entries = nullptr;     // [n_entries*TILE_M] partner rows
};

struct DwWork { int32_t k, start, len, slot; };  // one partial-dW work item

struct RuleBook {                 // Metadata.h:35 RuleBook + its derived gather lists
  int kind = 0;                   // 0 submanifold, 1 strided convolution
  int64_t in_ss[3], out_ss[3], filter[3], stride[3];
  int K = 0;
  int64_t n_in = 0, n_out = 0;
  bool identity = false;
  int64_t counts[MAX_KT];         // pairs per offset
  int64_t pair_off[MAX_KT + 1];   // host copy of the prefix
  int64_t total_pairs = 0;
  int32_t *pairs = nullptr;       // [total_pairs,2] (in,out), offsets contiguous, sorted by out
  int32_t *t_out = nullptr;       // [K,n_out] in-row feeding out-row at offset k (-1 none)
  int32_t *t_in = nullptr;        // [K,n_in]  out-row fed by in-row at offset k (-1 none)
  TileBook tb_out;                // stationary = out rows, gathers in rows
  TileBook tb_in;                 // stationary = in rows, gathers out rows
  DwWork *dw_work = nullptr;      // device work list for the weight gradient
  std::vector<DwWork> dw_host;    // its host image (kept alive for the asynchronous upload)
  int n_dw_work = 0;
  int dw_chunk = 0;
};

struct InputRules {               // IOLayersRules.h:10-15, kept as CSR instead of a padded table
  bool built = false;
  int mode = 0;
  int64_t n_points = 0, n_active = 0, max_active = 0;
  int32_t *point_row = nullptr;   // [n_points] site of each point
  int32_t *csr_off = nullptr;     // [n_active+1]
  int32_t *members = nullptr;     // [n_points] point ids grouped by site, ascending
  int32_t *stat = nullptr;        // device ints: [0] error [1] max batch [2] unsorted [4] maxActive
};

}  // namespace scn

struct scn_metadata {
  int dim = 3;
  std::vector<scn::Grid *> grids;
  std::vector<scn::RuleBook *> rulebooks;
  scn::InputRules input;
  int64_t batch_size = 0;
  cudaStream_t last_stream = 0;
};

namespace scn {
Grid *find_grid(scn_metadata *m, const int64_t *ss);
int get_submanifold_rulebook(scn_metadata *m, const int64_t *ss, const int64_t *filter,
                             cudaStream_t s, RuleBook **out);
int get_conv_rulebook(scn_metadata *m, const int64_t *in_ss, const int64_t *out_ss,
                      const int64_t *filter, const int64_t *stride, cudaStream_t s,
                      RuleBook **out);
// make sure tb_out / tb_in exists (tb_in builds t_in lazily for submanifold rulebooks)
int ensure_tilebook(RuleBook *rb, bool stationary_out, cudaStream_t s);
int ensure_dw_work(RuleBook *rb, cudaStream_t s);
}  // namespace scn
