// Device-side data layout of the B200-native DCGrid block pool and its lookup helpers.
//
// Reference structures being replaced: struct DCGrid (src/dcgrid/dcgrid.h:5-71), the index
// macros / Morton hash / lookups of src/dcgrid/dcgrid_utils.cuh.
//
// Differences that are layout only (results are identical, tests compare them entry by entry):
//  * 32-bit ids everywhere (cell id < 64*M < 2^32) instead of size_t: halves index traffic.
//  * (position, level) packed in one int4 per slot.
//  * The per-level open-addressing hash (hashKey/hashVal, dcgrid_utils.cuh:99-233) is replaced
//    by a dense "block-coordinate -> slot" map per sparse level (the coarse-to-block index
//    map): one load per level instead of a probe sequence, updated incrementally on
//    move/refine instead of memset + refill of the whole table
//    (fluid_simulation_dcgrid.cu:326-331).  4 B per potential block of a sparse level: 8 MiB
//    + 1 MiB at 512^3, sized for 180 GB of HBM.
//  * Face descriptors (6 neighbour slots + 6 closed-form codes = 48 B per block) derived from, and
//    verified entry by entry against, the 6^3 apron map, so stencil kernels read 48 B of indices
//    per block instead of the scattered 6^3 x 8 B map (dcgrid_stencil.cuh).
#pragma once
#include "common.cuh"

namespace dcg {

constexpr int kBW = 4, kBV = 64, kAW = 6, kAA = 36, kAV = 216, kSV = 8;  // dcgrid.h:51-67
constexpr uint32_t kNone = 0xFFFFFFFFu;                                   // DCGrid::notFound
constexpr int kMaxLevels = 16;
constexpr uint8_t kFlagMoved = 1, kFlagRefined = 2;                        // dcgrid.h:45-49
constexpr int kFree = 0xFF;                                               // blockLevels == 0xFF

// Plain-old-data view of the pool passed by value to kernels.
struct Pool {
  uint32_t M;
  int levels, sparse_levels;
  uint32_t offsets[kMaxLevels];     // levelOffsets
  uint32_t max_blocks[kMaxLevels];  // maxNumBlocksLevel
  uint32_t loads[kMaxLevels];       // blockLoads: active blocks of level l = slots [offsets[l], offsets[l]+loads[l])
  int4 *posl;                       // (x, y, z, level) per slot; level 0xFF = free slot
  uint8_t *flags;                   // blockFlags
  uint32_t *parent;                 // parentIndices: 8*parentSlot + subblock
  uint32_t *child;                  // childIndices[8*slot + subblock]
  uint32_t *apron;                  // cellIndices[216*slot + 36*X + 6*Y + Z]
  uint32_t *fd;                     // [12*slot]: 6 face-neighbour slots + 6 face codes (dcgrid_stencil.cuh)
  uint32_t *face;                   // [96*slot + 16*f + 4*a + b], f = -x,+x,-y,+y,-z,+z; irregular blocks only
  uint32_t *map[kMaxLevels];        // dense block-coordinate -> slot map of each sparse level
  uint32_t *fmap;                   // finest-block map (k_dc_build_fmap): level-0 block coordinate -> level << 28 | slot
};

// Work partition of the slab-decomposed solver (dcgrid.cu, "sharding"): the tiles (kTile consecutive slots) one
// rank processes in a launch, as runs of consecutive tiles.  Single GPU: one run covering everything.
constexpr int kTile = 16, kMaxRuns = 48;
struct TileRuns {
  int n;
  uint32_t first[kMaxRuns];    // first tile of run k
  uint32_t pre[kMaxRuns + 1];  // tiles in runs 0..k-1; pre[n] = total
};
__host__ __device__ __forceinline__ uint32_t run_total(const TileRuns &R) { return R.pre[R.n]; }
__host__ __device__ __forceinline__ uint32_t run_tile(const TileRuns &R, uint32_t j) {
  int k = 0;
  while (k + 1 < R.n && j >= R.pre[k + 1]) k++;
  return R.first[k] + (j - R.pre[k]);
}

// in-block cell bits: (sx<<5 | sy<<4 | sz<<3 | cx<<2 | cy<<1 | cz), coordinate = 2*s + c
// (dcgrid_utils.cuh:50-81)
__host__ __device__ __forceinline__ int cell_x(uint32_t c) { return (int)((((c >> 5) & 1) << 1) | ((c >> 2) & 1)); }
__host__ __device__ __forceinline__ int cell_y(uint32_t c) { return (int)((((c >> 4) & 1) << 1) | ((c >> 1) & 1)); }
__host__ __device__ __forceinline__ int cell_z(uint32_t c) { return (int)((((c >> 3) & 1) << 1) | (c & 1)); }
// SPREAD(data, offset), dcgrid_utils.cuh:63
__host__ __device__ __forceinline__ uint32_t spread(int d, int o) { return (uint32_t)(((d & 1) | ((d << 2) & 8)) << o); }
__host__ __device__ __forceinline__ uint32_t cell_bits(int x, int y, int z) { return spread(x, 2) + spread(y, 1) + spread(z, 0); }
// USE_APRON_INDEX, dcgrid_utils.cuh:50-59
__host__ __device__ __forceinline__ int apron_of(uint32_t c) { return kAA * (1 + cell_x(c)) + kAW * (1 + cell_y(c)) + (1 + cell_z(c)); }

// blocks per axis of a level
__device__ __forceinline__ int3 level_dims(const KParams &P, int level) {
  const int extent = kBW << level;
  return make_int3(idiv_up(P.gx, extent), idiv_up(P.gy, extent), idiv_up(P.gz, extent));
}

// getBlockIndex for ordered (fully allocated) levels, dcgrid_utils.cuh:172-183.
// (x,y,z) = level-`level` cell coordinates.
__device__ __forceinline__ uint32_t ordered_index(const Pool &T, const KParams &P, int x, int y, int z, int level) {
  const int3 r = level_dims(P, level);
  const int px = x / kBW, py = y / kBW, pz = z / kBW;
  // the reference compares int against size_t: negative coordinates convert to huge values
  if ((unsigned)px >= (unsigned)r.x || (unsigned)py >= (unsigned)r.y || (unsigned)pz >= (unsigned)r.z || x < 0 || y < 0 || z < 0)
    return kNone;
  return T.offsets[level] + ((uint32_t)px * r.y + py) * r.z + pz;
}

__device__ __forceinline__ uint32_t map_lookup(const Pool &T, const KParams &P, int x, int y, int z, int level) {
  const int3 r = level_dims(P, level);
  if (x < 0 || y < 0 || z < 0) return kNone;
  const int px = x / kBW, py = y / kBW, pz = z / kBW;
  if (px >= r.x || py >= r.y || pz >= r.z) return kNone;
  return T.map[level][((size_t)px * r.y + py) * r.z + pz];
}
__device__ __forceinline__ size_t map_slot(const KParams &P, int x, int y, int z, int level) {
  const int3 r = level_dims(P, level);
  return ((size_t)(x / kBW) * r.y + (y / kBW)) * r.z + (z / kBW);
}

// getBlockIndex, dcgrid_utils.cuh:169-199
__device__ __forceinline__ uint32_t block_index(const Pool &T, const KParams &P, int x, int y, int z, int level) {
  if (level >= T.sparse_levels) return ordered_index(T, P, x, y, z, level);
  return map_lookup(T, P, x, y, z, level);
}

// getBlockIndexDeep, dcgrid_utils.cuh:201-233: finest existing block of level >= `level`
// covering the level-`level` cell (x,y,z); updates `level`.
__device__ __forceinline__ uint32_t block_index_deep(const Pool &T, const KParams &P, int x, int y, int z, int &level) {
  for (; level < T.sparse_levels; level++, x /= 2, y /= 2, z /= 2) {
    const uint32_t b = map_lookup(T, P, x, y, z, level);
    if (b != kNone) return b;
  }
  // here level >= sparse_levels.  (For level > sparse_levels the reference shifts by a negative
  // count; that path is unreachable — only sparse-level blocks ever move.)
  return ordered_index(T, P, x, y, z, level);
}

}  // namespace dcg
