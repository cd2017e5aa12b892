// Hand-written sm_100a kernels of the DCGrid solve.  One logical block = 4^3 cells = 64 threads
// (two warps); a CTA carries kBPC logical blocks.  Each kernel cites the reference kernel whose
// results it reproduces bit for bit (compiled with -fmad=false, reference expression order).
#pragma once
#include <cfloat>

#include "dcgrid_layout.cuh"
#include "dcgrid_stencil.cuh"

namespace dcg {

constexpr int kBPC = 4;              // logical blocks per CTA
constexpr int kCTA = kBPC * kBV;     // 256 threads

// ======================================================================================
// structure
// ======================================================================================

// k_dcgrid_init_apron_indices, dcgrid_structure.cu:6-28: interior = own cells, rim = notFound
__global__ void __launch_bounds__(256) k_dc_init_apron(Pool T) {
  const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (size_t)T.M * kAV) return;
  const uint32_t b = (uint32_t)(t / kAV);
  const int ai = (int)(t % kAV);
  const int i = ai / kAA, j = (ai / kAW) % kAW, k = ai % kAW;
  const bool interior = i >= 1 && i <= kBW && j >= 1 && j <= kBW && k >= 1 && k <= kBW;
  T.apron[t] = interior ? b * kBV + cell_bits(i - 1, j - 1, k - 1) : kNone;
}

// k_dcgrid_activate_level, dcgrid_structure.cu:104-178.  One 64-thread group per block of an
// ordered level (the reference uses one thread per block with 216-iteration loops).
__global__ void __launch_bounds__(64) k_dc_activate_level(Pool T, KParams P, int level, float4 *vw0, float4 *vw1, float *q0,
                                                          float *q1, float *fl) {
  const int3 r = level_dims(P, level);
  const uint32_t lin = blockIdx.x;
  const int bz = lin % r.z, by = (lin / r.z) % r.y, bx = lin / (r.z * r.y);
  const int scale = 1 << level;
  const int px = kBW * bx, py = kBW * by, pz = kBW * bz;
  if (px * scale >= P.gx || py * scale >= P.gy || pz * scale >= P.gz) return;
  const uint32_t b = block_index(T, P, px, py, pz, level);
  if (b >= T.offsets[level] + T.max_blocks[level]) return;
  const int t = threadIdx.x;
  if (t == 0) {
    T.posl[b] = make_int4(px, py, pz, level);
    if (level < T.levels - 1)
      T.parent[b] = 8 * block_index(T, P, px / 2, py / 2, pz / 2, level + 1) + (px % (2 * kBW)) + (py % (2 * kBW)) / 2 +
                    (pz % (2 * kBW)) / 4;
  }
  if (t < 8 && level - 1 >= T.sparse_levels)
    T.child[8 * b + t] =
        block_index(T, P, 2 * px + kBW * ((t / 4) % 2), 2 * py + kBW * ((t / 2) % 2), 2 * pz + kBW * (t % 2), level - 1);
  const int i0 = px > 0 ? 0 : 1, i1 = kAW - ((px + kBW) * scale >= P.gx ? 1 : 0);
  const int j0 = py > 0 ? 0 : 1, j1 = kAW - ((py + kBW) * scale >= P.gy ? 1 : 0);
  const int k0 = pz > 0 ? 0 : 1, k1 = kAW - ((pz + kBW) * scale >= P.gz ? 1 : 0);
  for (int ai = t; ai < kAV; ai += 64) {
    const int i = ai / kAA, j = (ai / kAW) % kAW, k = ai % kAW;
    uint32_t e = kNone;
    if (i >= i0 && i < i1 && j >= j0 && j < j1 && k >= k0 && k < k1) {
      const uint32_t nb = block_index(T, P, px + i - 1, py + j - 1, pz + k - 1, level);
      // the neighbour's interior apron entry is its own cell (never rewritten after init)
      if (nb != kNone) e = nb * kBV + cell_bits((i - 1 + kBW) % kBW, (j - 1 + kBW) % kBW, (k - 1 + kBW) % kBW);
    }
    if (e == kNone) {  // outside the domain: nearest interior cell of this block (:157-167)
      const int ic = min(max(i, 1), kBW), jc = min(max(j, 1), kBW), kc = min(max(k, 1), kBW);
      e = b * kBV + cell_bits(ic - 1, jc - 1, kc - 1);
    }
    T.apron[(size_t)b * kAV + ai] = e;
  }
  {  // zero fields, set fluidity (:169-177)
    const uint32_t c = b * kBV + t;
    const float f = cell_fluidity(P, px + cell_x(t), py + cell_y(t), pz + cell_z(t), scale);
    vw0[c] = make_float4(0.f, 0.f, 0.f, f);
    vw1[c] = make_float4(0.f, 0.f, 0.f, f);
    q0[c] = 0.f;
    q1[c] = 0.f;
    fl[c] = f;
  }
}

// one bit per pool slot: "this block moved or was refined in this adaptTopology()" — 64 KiB at M = 524,288,
// cache-resident, so that the refresh below does not gather a 4-byte flag per rim entry
__global__ void __launch_bounds__(256) k_dc_flag_bits(const uint32_t *__restrict__ flags, uint32_t M, uint32_t *__restrict__ bits) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  const unsigned w = __ballot_sync(0xFFFFFFFFu, b < M && flags[b] != 0);
  if ((threadIdx.x & 31) == 0 && b < M) bits[b >> 5] = w;
}

// k_dcgrid_refresh_apron_indices, dcgrid_structure.cu:30-92 (reference: <<<M, 216>>>).  One warp per
// block walks its 6^3 map in 7 strides of 32 and touches rim entries only; a block whose flag bit and
// whose neighbours' flag bits are all clear leaves after the bitmap tests.
__device__ __forceinline__ bool flag_bit(const uint32_t *__restrict__ bits, uint32_t b) { return (bits[b >> 5] >> (b & 31)) & 1u; }
// fapron / perm: the field-order mirror of the apron map (nullptr while field order = slot order); every entry
// rewritten here is rewritten there too, translated through perm, so a topology change costs no full re-mirror.
__global__ void __launch_bounds__(256) k_dc_refresh_apron(Pool T, KParams P, const uint32_t *__restrict__ flags,
                                                          const uint32_t *__restrict__ flag_bits, uint32_t *__restrict__ fapron,
                                                          const uint32_t *__restrict__ perm) {
  const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= T.M) return;
  const int4 pl = T.posl[b];
  if (pl.w == kFree) return;
  const int level = pl.w;
  const bool self_moved = flag_bit(flag_bits, b) && (flags[b] & kFlagMoved) != 0;
  if (!self_moved) {
    // Most blocks have no touched neighbour: two coalesced 16-byte loads per lane cover the whole 864-byte map, and a
    // block none of whose entries points into a flagged block leaves here (the entry-by-entry walk below cost 336 us per
    // topology change at 512^3, profiles/README.md r2v).  Conservative: interior entries are tested too.
    const uint4 *m4 = reinterpret_cast<const uint4 *>(T.apron + (size_t)b * kAV);  // 216 entries = 54 uint4, 16-byte aligned (864 = 54 * 16)
    const int lane = threadIdx.x & 31;
    bool hit = false;
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int v = lane + 32 * r;
      if (v < kAV / 4) {
        const uint4 e = m4[v];
        auto touched = [&](uint32_t id) { return id / kBV < T.M && flag_bit(flag_bits, id / kBV); };
        hit = hit || touched(e.x) || touched(e.y) || touched(e.z) || touched(e.w);
      }
    }
    if (!__any_sync(0xFFFFFFFFu, hit)) return;
  }
  for (int ai = threadIdx.x & 31; ai < kAV; ai += 32) {
    const int i = ai / kAA, j = (ai / kAW) % kAW, k = ai % kAW;
    if (i % (kAW - 1) != 0 && j % (kAW - 1) != 0 && k % (kAW - 1) != 0) continue;
    uint32_t *entry = &T.apron[(size_t)b * kAV + ai];
    uint32_t nb = kNone;
    const int nx = pl.x + i - 1, ny = pl.y + j - 1, nz = pl.z + k - 1;
    if (self_moved) {
      const int scale = 1 << level;
      if (nx < 0 || ny < 0 || nz < 0 || nx * scale >= P.gx || ny * scale >= P.gy || nz * scale >= P.gz) {
        // :54-60 feeds APRON coordinates (1..4) to SPREAD where cell coordinates (0..3) are meant
        // (SURVEY App. B-5).  Reproduced on purpose: parity with the reference's Neumann ghosts.
        const uint32_t e = b * kBV + spread(min(max(i, 1), kBW), 2) + spread(min(max(j, 1), kBW), 1) + spread(min(max(k, 1), kBW), 0);
        *entry = e;
        if (fapron) fapron[(size_t)perm[b] * kAV + ai] = perm[b] * kBV + (e & 63u);
        continue;
      }
      int nl = level;
      nb = block_index_deep(T, P, nx, ny, nz, nl);
    } else {
      const uint32_t old = *entry;
      const uint32_t prev = old / kBV;
      if (!flag_bit(flag_bits, prev)) continue;  // neighbour untouched: entry stays
      const uint32_t cf = flags[prev];
      if (cf & kFlagMoved) {
        int nl = level;
        nb = block_index_deep(T, P, nx, ny, nz, nl);
      } else if ((cf & kFlagRefined) && T.posl[prev].w > level) {
        nb = T.child[old / kSV];
      }
    }
    if (nb == kNone) continue;
    const int4 np = T.posl[nb];
    const int s = 1 << (np.w - level);
    const uint32_t cb = cell_bits(nx / s - np.x, ny / s - np.y, nz / s - np.z);
    *entry = nb * kBV + cb;
    if (fapron) fapron[(size_t)perm[b] * kAV + ai] = perm[nb] * kBV + cb;
  }
}

// accumulate<T>, dcgrid_structure.cu:188-222: parent cell = .125 * sequential sum of the 8 cells of a
// child subblock.  One thread per subblock of `level`.
// skip_childless: the producing kernel already restricted the blocks that have no children (fused
// restriction in k_dc_advect_pipe / k_dc_apply_pressure4 / k_dc_divergence4); only blocks holding restricted
// cells of their own (= with children) still have to be pushed up.
__device__ __forceinline__ bool block_has_children(const Pool &T, uint32_t b) {
  const uint4 *c = reinterpret_cast<const uint4 *>(T.child + (size_t)b * kSV);
  const uint4 c0 = c[0], c1 = c[1];
  return (c0.x & c0.y & c0.z & c0.w & c1.x & c1.y & c1.z & c1.w) != kNone;
}
__global__ void __launch_bounds__(256) k_dc_accumulate_velocity(Pool T, int level, float4 *__restrict__ vw, int skip_childless) {
  pdl_enter();
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  if (t >= 8 * T.loads[level]) return;
  const uint32_t sb = 8 * T.offsets[level] + t, b = sb / 8;
  const uint32_t ps = T.parent[b];
  if (ps == kNone) return;
  if (skip_childless && !block_has_children(T, b)) return;
  const float4 *c = vw + (size_t)kSV * sb;
  float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
  for (int i = 0; i < kSV; i++) {
    const float4 v = c[i];
    ax += v.x; ay += v.y; az += v.z;
  }
  float *dst = reinterpret_cast<float *>(vw + ((size_t)kSV * ps + (sb % 8)));
  dst[0] = ax * .125f; dst[1] = ay * .125f; dst[2] = az * .125f;  // .w (fluidity) untouched
}
__global__ void __launch_bounds__(256) k_dc_accumulate_scalar(Pool T, int level, float *__restrict__ ch, int skip_childless) {
  pdl_enter();
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  if (t >= 8 * T.loads[level]) return;
  const uint32_t sb = 8 * T.offsets[level] + t, b = sb / 8;
  const uint32_t ps = T.parent[b];
  if (ps == kNone) return;
  if (skip_childless && !block_has_children(T, b)) return;
  const float4 lo = *reinterpret_cast<const float4 *>(ch + (size_t)kSV * sb);
  const float4 hi = *reinterpret_cast<const float4 *>(ch + (size_t)kSV * sb + 4);
  float a = 0.f;
  a += lo.x; a += lo.y; a += lo.z; a += lo.w; a += hi.x; a += hi.y; a += hi.z; a += hi.w;
  ch[(size_t)kSV * ps + (sb % 8)] = a * .125f;
}

// Blocks that still need a restriction pass when the producing kernel restricted the childless ones itself:
// the blocks WITH children (and a parent), listed per level at plist[offsets[level] ...) whenever the topology
// changes, so that a pass costs 8 threads per listed block instead of a child-link probe per pool slot
// (level 1 at 512^3: 30 k of 243 k blocks).  List order is arbitrary: every subblock is independent.
__global__ void __launch_bounds__(256) k_dc_list_parents(Pool T, const uint8_t *__restrict__ owner, uint32_t unit, int rank,
                                                         uint32_t *__restrict__ plist, uint32_t *__restrict__ pcount) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= T.M) return;
  if (owner && owner[b / unit] != rank) return;
  const int level = T.posl[b].w;
  if (level == kFree || T.parent[b] == kNone || !block_has_children(T, b)) return;
  // one atomic per warp and level (the lanes of a warp are consecutive slots: almost always one level)
  const unsigned peers = __match_any_sync(__activemask(), level);
  const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(&pcount[level], (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  plist[T.offsets[level] + base + __popc(peers & ((1u << lane) - 1u))] = b;
}
// one pass over the listed blocks of a level; kV / kS: the packed velocity and / or one scalar channel (after the fused
// advection the density and the speculative velocity are restricted together)
template <bool kV, bool kS>
__global__ void __launch_bounds__(256) k_dc_accumulate_list(Pool T, const uint32_t *__restrict__ list, uint32_t n, float4 *__restrict__ vw, float *__restrict__ ch) {
  pdl_enter();
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  if (t >= 8 * n) return;
  const uint32_t b = list[t >> 3], sb = 8 * b + (t & 7u);
  const uint32_t ps = T.parent[b];
  if (kV) {
    const float4 *c = vw + (size_t)kSV * sb;
    float4 v[kSV];
#pragma unroll
    for (int i = 0; i < kSV; i++) v[i] = c[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
    for (int i = 0; i < kSV; i++) { ax += v[i].x; ay += v[i].y; az += v[i].z; }
    float *dst = reinterpret_cast<float *>(vw + ((size_t)kSV * ps + (sb % 8)));
    dst[0] = ax * .125f; dst[1] = ay * .125f; dst[2] = az * .125f;
  }
  if (kS) {
    const float4 lo = *reinterpret_cast<const float4 *>(ch + (size_t)kSV * sb);
    const float4 hi = *reinterpret_cast<const float4 *>(ch + (size_t)kSV * sb + 4);
    float a = 0.f;
    a += lo.x; a += lo.y; a += lo.z; a += lo.w; a += hi.x; a += hi.y; a += hi.z; a += hi.w;
    ch[(size_t)kSV * ps + (sb % 8)] = a * .125f;
  }
}

// ---- restriction of the blocks WITH children in one launch: a counter-driven walk up the block tree ----------------
// The list passes above need one launch (sharded: one barrier) per level because a block with children can only be
// restricted after its children have been.  Here every block carries `expect` = the number of its children that have
// children themselves (k_dc_tree_expect, rebuilt with the topology) and a completion counter: a group of 8 lanes
// (one per subblock, as in the list passes: same sums, same order) starts at a block whose children are all
// childless — the producing kernel has already restricted those — pushes it into its parent, counts itself at the
// parent, and the group that completes the parent's count carries on with the parent.  Values written by other SMs
// in the same launch are read through L2 (__ldcg).  The walk stops above `top_level` (sharded: the ranks own whole
// subtrees below that level; the levels above it are the single-CTA-cluster pass k_dc_accumulate_coarse).
__global__ void __launch_bounds__(256) k_dc_tree_expect(Pool T, uint8_t *__restrict__ expect) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= T.M) return;
  const int level = T.posl[b].w;
  uint32_t n = 0;
  if (level != kFree) {
    for (int s = 0; s < kSV; s++) {
      const uint32_t cb = T.child[(size_t)b * kSV + s];
      if (cb != kNone && block_has_children(T, cb)) n++;
    }
  }
  expect[b] = (uint8_t)(n | ((uint32_t)(level & 15) << 4));
}
// start list: blocks of level <= top_level with children, a parent and no child that has children; world > 1: only
// those whose ancestor of level top_level (or the block itself, if the chain ends below) lies in this rank's share
// of the domain along `axis` (equal slabs of space: the small levels' slots all sit in one ownership unit, so unit
// ownership would hand every subtree to one rank; a subtree is walked by ONE rank because its counters are per rank)
__global__ void __launch_bounds__(256) k_dc_tree_starts(Pool T, KParams P, int top_level, int world, int axis, int rank,
                                                        const uint8_t *__restrict__ expect, uint32_t *__restrict__ starts, uint32_t *__restrict__ count) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= T.M) return;
  const int level = T.posl[b].w;
  if (level == kFree || level > top_level || (expect[b] & 15u) != 0 || T.parent[b] == kNone || !block_has_children(T, b)) return;
  if (world > 1) {
    uint32_t a = b;
    for (int l = level; l < top_level; l++) {
      const uint32_t ps = T.parent[a];
      if (ps == kNone) break;
      a = ps >> 3;
    }
    const int4 pa = T.posl[a];
    const long long pos = (long long)(axis == 0 ? pa.x : (axis == 1 ? pa.y : pa.z)) << pa.w;  // level-0 cell coordinate of the root's origin
    const int extent = axis == 0 ? P.gx : (axis == 1 ? P.gy : P.gz);
    const int owner = (int)min((long long)(world - 1), pos * world / extent);
    if (owner != rank) return;
  }
  starts[atomicAdd(count, 1u)] = b;
}
template <bool kV, bool kS>
__global__ void __launch_bounds__(256) k_dc_restrict_tree(Pool T, const uint32_t *__restrict__ starts, uint32_t nstarts, const uint8_t *__restrict__ expect,
                                                          uint32_t *cnt, int top_level, float4 *vw, float *ch) {
  pdl_enter();
  const uint32_t gid = blockIdx.x * 256 + threadIdx.x;
  const uint32_t task = gid >> 3, s = gid & 7u;
  if (task >= nstarts) return;  // whole groups leave together
  const unsigned lane0 = threadIdx.x & 24u, gmask = 0xFFu << lane0;
  uint32_t b = starts[task];
  for (;;) {
    const uint32_t ps = T.parent[b];
    if (ps == kNone) break;
    const size_t sb = (size_t)kSV * b + s;
    if (kV) {
      const float4 *c = vw + kSV * sb;
      float4 v[kSV];
#pragma unroll
      for (int i = 0; i < kSV; i++) v[i] = __ldcg(c + i);
      float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
      for (int i = 0; i < kSV; i++) { ax += v[i].x; ay += v[i].y; az += v[i].z; }
      float *dst = reinterpret_cast<float *>(vw + ((size_t)kSV * ps + s));
      dst[0] = ax * .125f; dst[1] = ay * .125f; dst[2] = az * .125f;  // .w (fluidity) untouched
    }
    if (kS) {
      const float4 lo = __ldcg(reinterpret_cast<const float4 *>(ch + kSV * sb)), hi = __ldcg(reinterpret_cast<const float4 *>(ch + kSV * sb + 4));
      float a = 0.f;
      a += lo.x; a += lo.y; a += lo.z; a += lo.w; a += hi.x; a += hi.y; a += hi.z; a += hi.w;
      ch[(size_t)kSV * ps + s] = a * .125f;
    }
    const uint32_t pb = ps >> 3;
    const uint32_t e = expect[pb];
    if ((int)(e >> 4) > top_level) break;
    __threadfence();  // this lane's stores, before the group is counted at the parent
    __syncwarp(gmask);
    uint32_t old = 0;
    if (s == 0) old = atomicAdd(cnt + pb, 1u);
    old = __shfl_sync(gmask, old, lane0);
    if (old + 1u != (e & 15u)) break;  // another group completes the parent
    if (s == 0) cnt[pb] = 0;           // ready for the next launch
    __threadfence();                   // the other children's stores, before this group reads them
    b = pb;
  }
}

// ======================================================================================
// adaptation (dcgrid_adaptation.cu)
// ======================================================================================

// Per-level summary of one score pass, reduced on the device so that the host can decide without
// copying the score arrays which levels can possibly move or refine anything (dcgrid.cu, move_blocks).
struct ScoreSummary {
  int max_ss[kMaxLevels];        // float bits of the largest non-negative subblock score of the level, -1 = none
  uint32_t min_bs[kMaxLevels];   // float bits of the smallest non-negative block score of the level, ~0 = none
  uint32_t n_refine[kMaxLevels]; // subblocks of the level with score > 1e-4 (refineSubblocks candidates)
};
__device__ __forceinline__ bool warp_uniform(int v) {
  const int v0 = __shfl_sync(0xFFFFFFFFu, v, 0);
  return __all_sync(0xFFFFFFFFu, v == v0);
}

// k_dcgrid_calc_subblock_scores, :10-40.  finer_full bit l = "level l cannot be refined further"
// ((l==0 && loads[0]==full[0]) || (l>0 && loads[l-1]==full[l-1]), :19-21).
// vort != nullptr: the flow-driven score (extension, dcg_ext_params.score_mode == 1) — the lines the reference carries
// commented out, :36-39, with calcCellScore (:6-8) = |vorticity|, stored in .w of vort (field order: perm maps slots)
__global__ void __launch_bounds__(256) k_dc_subblock_scores(Pool T, KParams P, uint32_t finer_full, float *__restrict__ sub_scores,
                                                            ScoreSummary *__restrict__ sum, const float4 *__restrict__ vort,
                                                            const uint32_t *__restrict__ perm) {
  const uint32_t sb = blockIdx.x * 256 + threadIdx.x;
  int level = kFree;
  float score = -FLT_MAX;
  if (sb < 8 * T.M) {
    const int4 pl = T.posl[sb / 8];
    level = pl.w;
    if (!(T.child[sb] != kNone || pl.w == kFree || ((finer_full >> pl.w) & 1)) && vort) {
      const float4 *w = vort + ((size_t)perm[sb / 8] * kBV + (size_t)kSV * (sb & 7u));
      score = 0.f;
#pragma unroll
      for (int i = 0; i < kSV; i++) score += w[i].w;
    } else if (!(T.child[sb] != kNone || pl.w == kFree || ((finer_full >> pl.w) & 1))) {
      const float s = (float)(1 << pl.w);
      const float px = s * ((float)pl.x + 2.f * (float)((sb >> 2) & 1) + 1.f);
      const float py = s * ((float)pl.y + 2.f * (float)((sb >> 1) & 1) + 1.f);
      const float pz = s * ((float)pl.z + 2.f * (float)(sb & 1) + 1.f);
      const float ex = px - .5f * (float)P.gx, ey = py - .45f * (float)P.gy, ez = pz - .5f * (float)P.gz;
      const float d = sqrtf(ex * ex + ey * ey + ez * ez);
      score = d < .2f * (float)P.gx ? 0.f : (float)P.gx / d;
    }
    sub_scores[sb] = score;
  }
  // summary: warp-shuffle reduction when the warp sits inside one level, then per CTA in shared memory: one global
  // atomic per CTA and level (131 k warps hammering two addresses per level cost 180 us at 512^3, profiles/README.md r2v)
  __shared__ int s_max[kMaxLevels];
  __shared__ uint32_t s_cnt[kMaxLevels];
  if (threadIdx.x < kMaxLevels) { s_max[threadIdx.x] = -1; s_cnt[threadIdx.x] = 0; }
  __syncthreads();
  int mx = score >= 0.f ? __float_as_int(score) : -1;
  uint32_t cnt = score > 1e-4f ? 1u : 0u;
  if (warp_uniform(level)) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
      cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && level != kFree) {
      if (mx >= 0) atomicMax(&s_max[level], mx);
      if (cnt) atomicAdd(&s_cnt[level], cnt);
    }
  } else if (level != kFree) {
    if (mx >= 0) atomicMax(&s_max[level], mx);
    if (cnt) atomicAdd(&s_cnt[level], cnt);
  }
  __syncthreads();
  if (threadIdx.x < kMaxLevels) {
    if (s_max[threadIdx.x] >= 0) atomicMax(&sum->max_ss[threadIdx.x], s_max[threadIdx.x]);
    if (s_cnt[threadIdx.x]) atomicAdd(&sum->n_refine[threadIdx.x], s_cnt[threadIdx.x]);
  }
}

// k_dcgrid_accumulate_subblock_scores, :42-63.  Sums NINE floats s[0..8]: the 9th
// is subblock 0 of the next pool slot.  sub_scores has 8*M+1 entries, the last one = -FLT_MAX.
__global__ void __launch_bounds__(256) k_dc_block_scores(Pool T, uint32_t finer_full, const float *__restrict__ sub_scores,
                                                         float *__restrict__ block_scores, ScoreSummary *__restrict__ sum) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  float out = -FLT_MAX;
  int level = kFree;
  if (b < T.M) {
    level = T.posl[b].w;
    if (level != kFree && !((finer_full >> level) & 1)) {
      const float *s = sub_scores + 8 * (size_t)b;
      if (s[0] > 0.f && s[1] > 0.f && s[2] > 0.f && s[3] > 0.f && s[4] > 0.f && s[5] > 0.f && s[6] > 0.f && s[7] > 0.f)
        out = .125f * (s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7] + s[8]);
    }
    block_scores[b] = out;
  }
  uint32_t mn = out >= 0.f ? __float_as_uint(out) : 0xFFFFFFFFu;
  if (warp_uniform(level)) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
    if ((threadIdx.x & 31) == 0 && level != kFree && mn != 0xFFFFFFFFu) atomicMin(&sum->min_bs[level], mn);
  } else if (level != kFree && mn != 0xFFFFFFFFu) {
    atomicMin(&sum->min_bs[level], mn);
  }
}

// k_dcgrid_move_blocks, :65-90, split in two so that the rank-order (sequential) semantics hold in
// parallel: the list is level-ascending, a block only reads the position of a block one level up,
// and that one is rewritten later in the sequence => every thread must see OLD positions.
__global__ void __launch_bounds__(256) k_dc_move_prepare(Pool T, const uint32_t *__restrict__ blocks, const uint32_t *__restrict__ dests,
                                                         uint32_t n, int4 *__restrict__ new_posl) {
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  if (t >= n) return;
  const uint32_t nps = dests[t];
  const int4 pp = T.posl[nps / 8];
  new_posl[t] = make_int4(2 * pp.x + kBW * (int)((nps / 4) % 2), 2 * pp.y + kBW * (int)((nps / 2) % 2),
                          2 * pp.z + kBW * (int)(nps % 2), T.posl[blocks[t]].w);
}
__global__ void __launch_bounds__(256) k_dc_move_commit(Pool T, KParams P, const uint32_t *__restrict__ blocks,
                                                        const uint32_t *__restrict__ dests, uint32_t n, const int4 *__restrict__ new_posl,
                                                        uint32_t *__restrict__ flags) {
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  if (t >= n) return;
  const uint32_t b = blocks[t], nps = dests[t];
  const int4 old = T.posl[b];
  T.child[T.parent[b]] = kNone;
  T.child[nps] = b;
  T.parent[b] = nps;
  if (old.w < T.sparse_levels) T.map[old.w][map_slot(P, old.x, old.y, old.z, old.w)] = kNone;
  T.posl[b] = new_posl[t];
  atomicOr(&flags[b], (uint32_t)kFlagMoved);
  atomicOr(&flags[nps / 8], (uint32_t)kFlagRefined);
}
// after every old position has been cleared
__global__ void __launch_bounds__(256) k_dc_map_insert(Pool T, KParams P, const uint32_t *__restrict__ blocks, uint32_t n) {
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  if (t >= n) return;
  const uint32_t b = blocks[t];
  const int4 pl = T.posl[b];
  if (pl.w < T.sparse_levels) T.map[pl.w][map_slot(P, pl.x, pl.y, pl.z, pl.w)] = b;
}

// k_dcgrid_refine_subblocks + insertBlock (:145-178, dcgrid_utils.cuh:99-140).  The reference
// allocates pool slots with a racy atomicAdd; here rank r of the level-l group takes
// freeBlockIndices[offset + load_before + (r - group_start)] — exactly what the atomics produce when
// threads run in rank order.
struct RefineGroups {
  uint32_t start[kMaxLevels];  // first rank whose CHILD level is l
  uint32_t base[kMaxLevels];   // levelOffsets[l] + blockLoads[l] before this call
};
__global__ void __launch_bounds__(256) k_dc_refine(Pool T, KParams P, const uint32_t *__restrict__ subblocks, uint32_t n,
                                                   RefineGroups G, const uint32_t *__restrict__ free_idx, uint32_t *__restrict__ flags,
                                                   uint32_t *__restrict__ touched, uint32_t num_touched, uint32_t *__restrict__ errors) {
  const uint32_t rank = blockIdx.x * 256 + threadIdx.x;
  if (rank >= n) return;
  const uint32_t sb = subblocks[rank];
  const uint32_t pb = sb / 8;
  const int4 pp = T.posl[pb];
  const int cl = pp.w - 1;
  const int cx = pp.x * 2 + (int)((sb / 4) % 2) * kBW, cy = pp.y * 2 + (int)((sb / 2) % 2) * kBW, cz = pp.z * 2 + (int)(sb % 2) * kBW;
  const size_t ms = map_slot(P, cx, cy, cz, cl);
  if (T.map[cl][ms] != kNone) {  // "Block already exists" (dcgrid_utils.cuh:109-114): never expected
    atomicAdd(errors, 1u);
    return;
  }
  const uint32_t cb = free_idx[G.base[cl] + (rank - G.start[cl])];
  T.map[cl][ms] = cb;
  T.posl[cb] = make_int4(cx, cy, cz, cl);
  T.parent[cb] = sb;
  T.child[sb] = cb;
  atomicOr(&flags[pb], (uint32_t)kFlagRefined);
  atomicOr(&flags[cb], (uint32_t)kFlagMoved);
  touched[num_touched + rank] = cb;
}

// k_dcgrid_propagate_values, :92-143: new/moved block <- 27/9/3/1 interpolation of the parent's cells
// T = the pool in field order, touched = reference slots, perm = reference slot -> field slot
__global__ void __launch_bounds__(64) k_dc_propagate(Pool T, KParams P, const uint32_t *__restrict__ touched, const uint32_t *__restrict__ perm,
                                                     int level, float4 *__restrict__ vw, float *__restrict__ q, float *__restrict__ fl) {
  const uint32_t b = perm[touched[blockIdx.x]];
  const int4 pl = T.posl[b];
  if (pl.w != level) return;
  const uint32_t ps = T.parent[b];
  if (ps == kNone) return;
  const uint32_t *pa = T.apron + (size_t)(ps / 8) * kAV;
  const uint32_t t = threadIdx.x;
  const uint32_t c = b * kBV + t;
  const int x = pl.x | cell_x(t), y = pl.y | cell_y(t), z = pl.z | cell_z(t);
  const int idx = kAA * (1 + (x / 2) % kBW) + kAW * (1 + (y / 2) % kBW) + (1 + (z / 2) % kBW);
  const int i = x % 2 ? kAA : -kAA, j = y % 2 ? kAW : -kAW, k = z % 2 ? 1 : -1;
  const uint32_t i000 = pa[idx], i001 = pa[idx + k], i010 = pa[idx + j], i100 = pa[idx + i];
  const uint32_t i011 = pa[idx + j + k], i101 = pa[idx + i + k], i110 = pa[idx + i + j], i111 = pa[idx + i + j + k];
  q[c] = ((27.f / 64.f) * q[i000] + (9.f / 64.f) * (q[i001] + q[i010] + q[i100]) + (3.f / 64.f) * (q[i011] + q[i101] + q[i110]) +
          (1.f / 64.f) * q[i111]);
  const float4 v000 = vw[i000], v001 = vw[i001], v010 = vw[i010], v100 = vw[i100];
  const float4 v011 = vw[i011], v101 = vw[i101], v110 = vw[i110], v111 = vw[i111];
  float4 o;
  o.x = ((27.f / 64.f) * v000.x + (9.f / 64.f) * (v001.x + v010.x + v100.x) + (3.f / 64.f) * (v011.x + v101.x + v110.x) + (1.f / 64.f) * v111.x);
  o.y = ((27.f / 64.f) * v000.y + (9.f / 64.f) * (v001.y + v010.y + v100.y) + (3.f / 64.f) * (v011.y + v101.y + v110.y) + (1.f / 64.f) * v111.y);
  o.z = ((27.f / 64.f) * v000.z + (9.f / 64.f) * (v001.z + v010.z + v100.z) + (3.f / 64.f) * (v011.z + v101.z + v110.z) + (1.f / 64.f) * v111.z);
  o.w = cell_fluidity(P, x, y, z, 1 << level);
  vw[c] = o;
  fl[c] = o.w;
}

// ======================================================================================
// fluid (dcgrid_fluid.cu)
// ======================================================================================

// INIT_SAMPLE, dcgrid_fluid.cu:7-72
struct DSample {
  uint32_t id[8];
  int x0, y0, z0, scale;
  float fx, fy, fz;
};
// true iff all 8 corners of the sample lie inside the domain, i.e. bc_kind() == 0 for each of them
__device__ __forceinline__ bool sample_inside(const KParams &P, const DSample &s) {
  return s.x0 >= 0 && s.y0 >= 0 && s.z0 >= 0 && (s.x0 + 1) * s.scale < P.gx && (s.y0 + 1) * s.scale < P.gy && (s.z0 + 1) * s.scale < P.gz;
}
// Second half of INIT_SAMPLE: sample set-up inside the resolved block `bp` (level, position) whose 6^3
// apron map starts at `apron` (global memory, or the CTA's staged copy in shared memory).
__device__ __forceinline__ DSample d_sample_in(const uint32_t *apron, const int4 bp, float px, float py, float pz) {
  DSample s;
  s.scale = 1 << bp.w;
  const float inv = 1.f / (float)s.scale;
  const float x = px * inv - .5f, y = py * inv - .5f, z = pz * inv - .5f;
  const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  s.fx = x - xf; s.fy = y - yf; s.fz = z - zf;
  s.x0 = (int)xf; s.y0 = (int)yf; s.z0 = (int)zf;
  const int i = min(max(s.x0 + 1 - bp.x, 0), kAW - 2), j = min(max(s.y0 + 1 - bp.y, 0), kAW - 2), k = min(max(s.z0 + 1 - bp.z, 0), kAW - 2);
  const uint32_t *a = apron + kAA * i + kAW * j + k;
  s.id[0] = a[0]; s.id[1] = a[1]; s.id[2] = a[kAW]; s.id[3] = a[kAW + 1];
  s.id[4] = a[kAA]; s.id[5] = a[kAA + 1]; s.id[6] = a[kAA + kAW]; s.id[7] = a[kAA + kAW + 1];
  return s;
}

// Gather set-up for one cell.  The reference resolves every sample with getBlockIndexDeep + two more
// dependent loads (block position, 8 apron ids).  Most backtraced samples stay inside the block they
// started from, so each CTA stages its blocks' apron maps and child links in shared memory (loads
// that do not depend on the velocity) and only samples that leave the block — or land in one of its
// refined subblocks — walk the level maps.  Both paths produce the ids of the reference's lookup:
// a sample inside an unrefined subblock of its own leaf block resolves to that block
// (dcgrid_utils.cuh:201-233 finds the finest covering block; a finer one would be a descendant of
// that subblock).
__device__ __forceinline__ DSample d_sample(const Pool &T, const KParams &P, const uint32_t *own_apron, const uint32_t *own_child,
                                            const int4 pl, float px, float py, float pz) {
  const int ix = min(max((int)floorf(px), 0), P.gx - 1), iy = min(max((int)floorf(py), 0), P.gy - 1), iz = min(max((int)floorf(pz), 0), P.gz - 1);
  const int lx = (ix >> pl.w) - pl.x, ly = (iy >> pl.w) - pl.y, lz = (iz >> pl.w) - pl.z;
  if ((unsigned)lx < (unsigned)kBW && (unsigned)ly < (unsigned)kBW && (unsigned)lz < (unsigned)kBW &&
      own_child[((lx >> 1) << 2) | ((ly >> 1) << 1) | (lz >> 1)] == kNone)
    return d_sample_in(own_apron, pl, px, py, pz);
  int level = 0;
  const uint32_t b = block_index_deep(T, P, ix, iy, iz, level);
  return d_sample_in(T.apron + (size_t)b * kAV, T.posl[b], px, py, pz);
}

// stage the apron maps + child links of the CTA's kBPC blocks
__device__ __forceinline__ void stage_apron(const Pool &T, uint32_t b, bool active, uint32_t g, uint32_t t, uint32_t (*sa)[kAV], uint32_t (*sc)[kSV]) {
  if (active) {
    const uint32_t *ap = T.apron + (size_t)b * kAV;
    for (int i = t; i < kAV; i += kBV) sa[g][i] = ap[i];
    if (t < kSV) sc[g][t] = T.child[(size_t)b * kSV + t];
  }
  __syncthreads();
}

// k_dcgrid_advect_velocity, dcgrid_fluid.cu:74-91,112-127.  Writes the other ping-pong buffer (no
// whole-pool D2D memcpy, fluid_simulation_dcgrid.cu:265-266).
__global__ void __launch_bounds__(kCTA, 5) k_dc_advect_velocity(Pool T, KParams P, const float4 *__restrict__ vin, float4 *__restrict__ vout) {
  __shared__ uint32_t sa[kBPC][kAV];
  __shared__ uint32_t sc[kBPC][kSV];
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t b = blockIdx.x * kBPC + g;
  const bool in_pool = b < T.M;
  const uint32_t c = b * kBV + t;
  int4 pl = make_int4(0, 0, 0, kFree);
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  if (in_pool) {  // position, velocity and apron map are fetched by independent loads (free slots: zeroed data, unused)
    pl = T.posl[b];
    me = vin[c];
  }
  stage_apron(T, b, in_pool, g, t, sa, sc);
  if (pl.w == kFree) return;
  float3 out = make_float3(0.f, 0.f, 0.f);
  if (sc[g][t >> 3] == kNone) {
    const float scale = (float)(1 << pl.w);
    const float alpha = P.dt * P.rdx;
    const float bx = ((float)(pl.x | cell_x(t)) + .5f) * scale - me.x * alpha;
    const float by = ((float)(pl.y | cell_y(t)) + .5f) * scale - me.y * alpha;
    const float bz = ((float)(pl.z | cell_z(t)) + .5f) * scale - me.z * alpha;
    const DSample s = d_sample(T, P, sa[g], sc[g], pl, bx, by, bz);
    float4 cv[8];
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      cv[k] = vin[s.id[k]];
      f[k] = cv[k].w;
    }
    const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
    if (!(W.acc < 1e-6f)) {
      float vx[8], vy[8], vz[8];
      // boundary conditions only substitute values of corners OUTSIDE the domain (sim_utils.cu:24-39):
      // one test for the whole 2x2x2 footprint skips them for interior samples
      const bool inside = sample_inside(P, s);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float3 v = make_float3(cv[k].x, cv[k].y, cv[k].z);
        if (!inside) v = velocity_bc(P, v, s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1), s.z0 + (k & 1), s.scale);
        vx[k] = v.x; vy[k] = v.y; vz[k] = v.z;
      }
      out = make_float3(blend8(vx, W.w), blend8(vy, W.w), blend8(vz, W.w));
    }
  }
  vout[c] = make_float4(out.x, out.y, out.z, me.w);
}

// k_dcgrid_advect_density, dcgrid_fluid.cu:93-110,129-144
__global__ void __launch_bounds__(kCTA) k_dc_advect_density(Pool T, KParams P, const float4 *__restrict__ vw, const float *__restrict__ fl,
                                                            const float *__restrict__ qin, float *__restrict__ qout) {
  __shared__ uint32_t sa[kBPC][kAV];
  __shared__ uint32_t sc[kBPC][kSV];
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t b = blockIdx.x * kBPC + g;
  const bool in_pool = b < T.M;
  const uint32_t c = b * kBV + t;
  int4 pl = make_int4(0, 0, 0, kFree);
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  if (in_pool) {  // position, velocity and apron map are fetched by independent loads (free slots: zeroed data, unused)
    pl = T.posl[b];
    me = vw[c];
  }
  stage_apron(T, b, in_pool, g, t, sa, sc);
  if (pl.w == kFree) return;
  float out = 0.f;
  if (sc[g][t >> 3] == kNone) {
    const float scale = (float)(1 << pl.w);
    const float alpha = P.dt * P.rdx;
    const float bx = ((float)(pl.x | cell_x(t)) + .5f) * scale - me.x * alpha;
    const float by = ((float)(pl.y | cell_y(t)) + .5f) * scale - me.y * alpha;
    const float bz = ((float)(pl.z | cell_z(t)) + .5f) * scale - me.z * alpha;
    const DSample s = d_sample(T, P, sa[g], sc[g], pl, bx, by, bz);
    float qv[8], f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      f[k] = fl[s.id[k]];
      qv[k] = qin[s.id[k]];
    }
    const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
    if (!(W.acc < 1e-6f)) {
      if (!sample_inside(P, s)) {
#pragma unroll
        for (int k = 0; k < 8; k++)
          qv[k] = density_bc(P, qv[k], s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1), s.z0 + (k & 1), s.scale);
      }
      out = blend8(qv, W.w);
    }
  }
  qout[c] = out;
}

// k_dcgrid_prolongate, dcgrid_multigrid_solver.cu:43-76.  One thread per quad (1x2x2 cells): the four cells
// share their parent cell, so the 4 x 8 coarse reads of the reference collapse to 2x3x3 = 18.
__device__ __forceinline__ float prolong_one(const float (*c)[3][3], int j, int k) {
  // c[di][dj+1][dk+1], di = 0 (parent cell) / 1 (its x neighbour); j,k = -1/+1 toward the nearer neighbour
  const float p000 = c[0][1][1], p001 = c[0][1][1 + k], p010 = c[0][1 + j][1], p100 = c[1][1][1];
  const float p011 = c[0][1 + j][1 + k], p101 = c[1][1][1 + k], p110 = c[1][1 + j][1], p111 = c[1][1 + j][1 + k];
  return (27.f * p000 + 9.f * (p001 + p010 + p100) + 3.f * (p011 + p101 + p110) + p111) / 64.f;
}
__global__ void __launch_bounds__(kCTA4) k_dc_prolongate4(Pool T, int level, float *__restrict__ p) {
  pdl_enter();
  const uint32_t g = threadIdx.x >> 4;
  const int t = threadIdx.x & 15;
  const uint32_t li = blockIdx.x * kB4 + g;
  if (li >= T.loads[level]) return;
  const uint32_t b = T.offsets[level] + li;
  const uint32_t ps = T.parent[b];
  if (ps == kNone) return;
  const uint32_t *pa = T.apron + (size_t)(ps / 8) * kAV;
  int X, Y0, Z0;
  quad_coords(t, X, Y0, Z0);
  // a child block covers subblock (ps & 7) of its parent: parent cell = 2*subblock bit + (child cell >> 1)
  const int PX = (int)((ps >> 2) & 1u) * 2 + (X >> 1), PY = (int)((ps >> 1) & 1u) * 2 + (Y0 >> 1), PZ = (int)(ps & 1u) * 2 + (Z0 >> 1);
  const int idx = kAA * (1 + PX) + kAW * (1 + PY) + (1 + PZ);
  const int i = (X & 1) ? kAA : -kAA;
  float c[2][3][3];
#pragma unroll
  for (int di = 0; di < 2; di++)
#pragma unroll
    for (int dj = -1; dj <= 1; dj++)
#pragma unroll
      for (int dk = -1; dk <= 1; dk++) c[di][dj + 1][dk + 1] = p[pa[idx + di * i + dj * kAW + dk]];
  float4 o;
  o.x = prolong_one(c, -1, -1);
  o.y = prolong_one(c, -1, 1);
  o.z = prolong_one(c, 1, -1);
  o.w = prolong_one(c, 1, 1);
  *reinterpret_cast<float4 *>(p + (size_t)b * kBV + 4 * t) = o;
}

// Same, with the 4^3 coarse cells a child block interpolates from (its parent subblock's 2^3 cells and their
// ring) staged once per block in shared memory: 4 index + 4 value loads per thread instead of 18 + 18 — the
// gather version keeps the L1 data pipe 85 % busy for 4.5 B/cell of useful traffic (profiles/README.md r1d).
__global__ void __launch_bounds__(kCTA4) k_dc_prolongate_staged(Pool T, TileRuns R, int level, float *__restrict__ p) {
  pdl_enter();
  __shared__ float sc[kB4][64];
  const uint32_t g = threadIdx.x >> 4;
  const int t = threadIdx.x & 15;
  const uint32_t li = run_tile(R, blockIdx.x) * kB4 + g;
  const bool active = li < T.loads[level];
  const uint32_t b = T.offsets[level] + li;
  const uint32_t ps = active ? T.parent[b] : kNone;
  const bool ok = active && ps != kNone;
  if (ok) {
    const uint32_t *pa = T.apron + (size_t)(ps / 8) * kAV;
    const int o = kAA * (int)((ps >> 2) & 1u) * 2 + kAW * (int)((ps >> 1) & 1u) * 2 + (int)(ps & 1u) * 2;  // cube origin in the apron
    uint32_t id[4];
#pragma unroll
    for (int j = 0; j < 4; j++) id[j] = pa[o + kAA * j + kAW * (t >> 2) + (t & 3)];
#pragma unroll
    for (int j = 0; j < 4; j++) sc[g][16 * j + t] = p[id[j]];
  }
  __syncthreads();
  if (!ok) return;
  int X, Y0, Z0;
  quad_coords(t, X, Y0, Z0);
  const int base = 16 * (1 + (X >> 1)) + 4 * (1 + (Y0 >> 1)) + (1 + (Z0 >> 1));
  const int i = (X & 1) ? 16 : -16;
  float c[2][3][3];
#pragma unroll
  for (int di = 0; di < 2; di++)
#pragma unroll
    for (int dj = -1; dj <= 1; dj++)
#pragma unroll
      for (int dk = -1; dk <= 1; dk++) c[di][dj + 1][dk + 1] = sc[g][base + di * i + 4 * dj + dk];
  float4 o4;
  o4.x = prolong_one(c, -1, -1);
  o4.y = prolong_one(c, -1, 1);
  o4.z = prolong_one(c, 1, -1);
  o4.w = prolong_one(c, 1, 1);
  *reinterpret_cast<float4 *>(p + (size_t)b * kBV + 4 * t) = o4;
}

// Prolongation by PARENT block, persistent and software-pipelined (levels that fill the GPU, one GPU).
// The eight children of a refined block interpolate from the same 6^3 apron of that block, and the apron map of
// a block is one contiguous 864-byte range: a CTA walks the list of blocks with children (k_dc_list_parents),
// gathers the parent's 216 apron'ed pressures ONCE (the per-child kernel gathers 64 per child: 512 for eight) and
// writes all its child blocks from shared memory.  The three dependent load rounds (list entry -> apron ids ->
// pressures) of parent k + 3 / k + 2 / k + 1 are in flight while parent k is computed.  Same values, same
// prolong_one expression as k_dc_prolongate4: every block with a parent is the child of exactly one listed block.
constexpr int kPPThreads = 256;
__global__ void __launch_bounds__(kPPThreads) k_dc_prolongate_parents(Pool T, const uint32_t *__restrict__ list, uint32_t n, float *p) {
  pdl_enter();
  __shared__ float sc[2][kAV + 8];
  __shared__ uint32_t sch[2][kSV];
  const int i = threadIdx.x;
  const bool gath = i < kAV;
  const bool link = i >= 224 && i < 224 + kSV;
  const uint32_t stride = gridDim.x;
  uint32_t j = blockIdx.x;
  auto entry = [&](uint32_t jj) -> uint32_t { return jj < n ? __ldg(list + jj) : kNone; };
  // pipeline registers: P = parent slot, id = apron id, ch = child link, v = pressure
  uint32_t P1 = entry(j + stride), P2 = entry(j + 2 * stride), P3;
  uint32_t id0 = 0, id1 = 0, id2 = 0, ch0 = kNone, ch1 = kNone, ch2 = kNone;
  float v0 = 0.f, v1 = 0.f;
  {
    const uint32_t P0 = entry(j);
    if (P0 != kNone) {
      if (gath) id0 = T.apron[(size_t)P0 * kAV + i];
      if (link) ch0 = T.child[(size_t)P0 * kSV + (i - 224)];
    }
    if (P1 != kNone) {
      if (gath) id1 = T.apron[(size_t)P1 * kAV + i];
      if (link) ch1 = T.child[(size_t)P1 * kSV + (i - 224)];
    }
    if (P0 != kNone && gath) v0 = p[id0];
  }
  for (uint32_t k = 0; j < n; k++, j += stride) {
    // issue the rounds of the following parents
    if (P1 != kNone && gath) v1 = p[id1];
    if (P2 != kNone) {
      if (gath) id2 = T.apron[(size_t)P2 * kAV + i];
      if (link) ch2 = T.child[(size_t)P2 * kSV + (i - 224)];
    }
    P3 = entry(j + 3 * stride);
    // this parent
    const int buf = (int)(k & 1u);
    if (gath) sc[buf][i] = v0;
    if (link) sch[buf][i - 224] = ch0;
    __syncthreads();
    if (i < 128) {
      const int s = i >> 4, t = i & 15;
      const uint32_t cb = sch[buf][s];
      if (cb != kNone) {
        int X, Y0, Z0;
        quad_coords(t, X, Y0, Z0);
        // the child covers subblock s of the parent: parent cell = 2 * subblock bit + (child cell >> 1)
        const int base = kAA * (1 + ((s >> 2) & 1) * 2 + (X >> 1)) + kAW * (1 + ((s >> 1) & 1) * 2 + (Y0 >> 1)) + (1 + (s & 1) * 2 + (Z0 >> 1));
        const int di = (X & 1) ? kAA : -kAA;
        float c[2][3][3];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
          for (int dj = -1; dj <= 1; dj++)
#pragma unroll
            for (int dk = -1; dk <= 1; dk++) c[a][dj + 1][dk + 1] = sc[buf][base + a * di + kAW * dj + dk];
        float4 o4;
        o4.x = prolong_one(c, -1, -1);
        o4.y = prolong_one(c, -1, 1);
        o4.z = prolong_one(c, 1, -1);
        o4.w = prolong_one(c, 1, 1);
        *reinterpret_cast<float4 *>(p + (size_t)cb * kBV + 4 * t) = o4;
      }
    }
    // rotate (one barrier per parent: the buffers alternate, and nobody gets two parents ahead of the barrier)
    v0 = v1; id1 = id2; ch0 = ch1; ch1 = ch2; P1 = P2; P2 = P3;
  }
}

// k_dcgrid_debug_stats, dcgrid_structure.cu:224-251: one thread per block, sequential i,j,k order
__global__ void __launch_bounds__(256) k_dc_debug_stats(Pool T, KParams P, const float *__restrict__ p, const float *__restrict__ div,
                                                        float *__restrict__ stats) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= T.M) return;
  float sres = 0.f;
  const int level = T.posl[b].w;
  if (level != kFree) {
    const uint32_t *a = T.apron + (size_t)b * kAV;
    const int scale = 1 << level;
    const float alpha = P.rdx * P.rdx / (float)(scale * scale);
    for (int i = 1; i <= kBW; i++)
      for (int j = 1; j <= kBW; j++)
        for (int k = 1; k <= kBW; k++) {
          const int ai = kAA * i + kAW * j + k;
          const uint32_t c = a[ai];
          if (T.child[c >> 3] != kNone) continue;
          const float r = div[c] - (p[a[ai - kAA]] + p[a[ai + kAA]] + p[a[ai - kAW]] + p[a[ai + kAW]] + p[a[ai - 1]] + p[a[ai + 1]] - 6.f * p[c]) * alpha;
          sres += (float)scale * fabsf(r);
        }
  }
  stats[b] = sres;
}

// total smoke over leaf cells, volume weighted (double, fixed reduction tree)
__global__ void __launch_bounds__(256) k_dc_total_density(Pool T, TileRuns R, const float *__restrict__ q, const float *__restrict__ fl,
                                                          double *__restrict__ partial) {
  __shared__ double sh[256];
  double s = 0.0;
  const uint32_t total = run_total(R);
  for (uint32_t j = blockIdx.x; j < total; j += gridDim.x) {
    const size_t c0 = (size_t)run_tile(R, j) * kTile * kBV;
    for (uint32_t k = threadIdx.x; k < kTile * kBV; k += 256) {
      const size_t c = c0 + k;
      if (c >= (size_t)T.M * kBV) break;
      const int level = T.posl[c >> 6].w;
      if (level == kFree || T.child[c >> 3] != kNone) continue;
      const double vol = (double)(1ull << (3 * level));
      s += vol * (double)(q[c] * fl[c]);
    }
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ======================================================================================
// accessors
// ======================================================================================
__global__ void __launch_bounds__(256) k_dc_unpack_velocity(const float4 *__restrict__ vw, float *__restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 v = vw[i];
  out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
}

// dense level-0 resampling: value of the finest covering cell (sampleCoarse-style lookup,
// dcgrid_rendering.cu:6-24)
__global__ void __launch_bounds__(256) k_dc_dense_l0(Pool T, KParams P, const float *__restrict__ src, int comps, int stride,
                                                     float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t n = (size_t)P.gx * P.gy * P.gz;
  if (i >= n) return;
  const int x = (int)(i % P.gx), y = (int)((i / P.gx) % P.gy), z = (int)(i / ((size_t)P.gx * P.gy));
  int level = 0;
  const uint32_t b = block_index_deep(T, P, x, y, z, level);
  const int4 pl = T.posl[b];
  const uint32_t c = b * kBV + cell_bits((x >> level) - pl.x, (y >> level) - pl.y, (z >> level) - pl.z);
  for (int k = 0; k < comps; k++) out[comps * i + k] = src[(size_t)stride * c + k];
}

__global__ void __launch_bounds__(256) k_dc_lookup(Pool T, KParams P, const int *__restrict__ pos, size_t n, uint32_t *__restrict__ slot,
                                                   uint8_t *__restrict__ lvl) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  int level = 0;
  slot[i] = block_index_deep(T, P, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], level);
  lvl[i] = (uint8_t)level;
}

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t *p, uint32_t v, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(256) k_iota_u32(uint32_t *p, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) k_fill_posl(int4 *p, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) p[i] = make_int4(0, 0, 0, kFree);
}

// ======================================================================================
// field order: a position-sorted renumbering of the sparse levels for the hot kernels
// ======================================================================================
// The reference numbers blocks by allocation order and a block keeps its slot when moveBlocks relocates it
// (dcgrid_adaptation.cu:65-90), so after the adaptation transient slot order says nothing about position on the
// finest level: a tile of 16 consecutive slots is 16 scattered blocks, their ghost cells miss L1/L2, and a rank
// of the slab decomposition owning a slot range owns a random half of space (47 % remote face neighbours,
// profiles/README.md).  The adaptation (scores, selection, moves, refinement, apron refresh) and every accessor
// stay in the reference's numbering; the field kernels run on a MIRROR of the pool renumbered by
// perm[reference slot] = field slot, a per-level permutation that sorts the active blocks of each sparse level by
// position (x-major like the ordered levels, then Morton).  Mirror structures hold field-space ids, the eight field
// arrays are stored in field order, so the hot kernels are unchanged: they are simply handed the mirrored Pool.
// Results cannot depend on the numbering: every cell is computed by the same expression from the same values.
// slab_axis: the axis the ranks' slabs are stacked along (the longest axis of the domain: the cut surfaces are the
// smallest cross-sections; x-slabs of a 512 x 512 x 4096 scene would have 8 x larger surfaces than z-slabs)
// sort_ordered: the fully allocated ("ordered") levels are renumbered too — their slot order is x-major by
// construction (dcgrid_utils.cuh:172-183), i.e. x-slabs; with another stacking axis the mirror addresses them through
// dense maps like the sparse levels (Tf.sparse_levels = levels) so that every level is cut along the same axis
__global__ void __launch_bounds__(256) k_dc_resort_keys(Pool T, KParams P, int world, int slab_axis, int sort_ordered,
                                                        unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= T.M) return;
  int range = T.levels;  // level whose slot range holds b (slots past the last range stay where they are)
  for (int l = 0; l < T.levels; l++)
    if (b >= T.offsets[l] && b < T.offsets[l] + T.max_blocks[l]) range = l;
  const int4 pl = T.posl[b];
  unsigned long long k = 0xFFFFFF00000000ull + b;  // free slots: behind the active ones, in slot order
  if (range >= T.sparse_levels && !sort_ordered) {
    k = b;  // ordered levels are addressed arithmetically (ordered_index): identity
  } else if (pl.w != kFree) {
    const uint32_t x = (uint32_t)(pl.x << pl.w) >> 2, y = (uint32_t)(pl.y << pl.w) >> 2, z = (uint32_t)(pl.z << pl.w) >> 2;
    auto spread3 = [](uint32_t v) {
      v &= 0x3FFu;
      v = (v | (v << 16)) & 0x030000FFu;
      v = (v | (v << 8)) & 0x0300F00Fu;
      v = (v | (v << 4)) & 0x030C30C3u;
      v = (v | (v << 2)) & 0x09249249u;
      return v;
    };
    const unsigned long long morton = (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
    // several ranks: slabs of 8 blocks along the stacking axis first, Morton inside
    const uint32_t along = slab_axis == 0 ? x : (slab_axis == 1 ? y : z);
    k = world > 1 ? ((unsigned long long)(along >> 3) << 32) | morton : morton;
  }
  keys[b] = ((unsigned long long)range << 56) | k;
  vals[b] = b;
}
// sorted[i] = reference slot that takes field slot i
__global__ void __launch_bounds__(256) k_dc_perm_from_sorted(const uint32_t *__restrict__ sorted, uint32_t M, uint32_t *__restrict__ perm) {
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i < M) perm[sorted[i]] = i;
}
// moves one field from the old to the new field order (64 * comps floats per block)
template <typename V>
__global__ void __launch_bounds__(256) k_dc_permute_field(const V *__restrict__ src, V *__restrict__ dst, const uint32_t *__restrict__ perm_old,
                                                          const uint32_t *__restrict__ perm_new, size_t cells) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= cells) return;
  const uint32_t b = (uint32_t)(i >> 6), c = (uint32_t)(i & 63);
  dst[(size_t)perm_new[b] * kBV + c] = src[(size_t)perm_old[b] * kBV + c];
}
__device__ __forceinline__ uint32_t perm_cell(const uint32_t *__restrict__ perm, uint32_t id) {
  return id == kNone ? kNone : perm[id >> 6] * kBV + (id & 63u);
}
__global__ void __launch_bounds__(256) k_dc_mirror_blocks(Pool T, Pool F, const uint32_t *__restrict__ perm) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= T.M) return;
  const uint32_t fb = perm[b];
  F.posl[fb] = T.posl[b];
  const uint32_t ps = T.parent[b];
  F.parent[fb] = ps == kNone ? kNone : 8 * perm[ps >> 3] + (ps & 7u);
#pragma unroll
  for (int s = 0; s < kSV; s++) {
    const uint32_t c = T.child[(size_t)b * kSV + s];
    F.child[(size_t)fb * kSV + s] = c == kNone ? kNone : perm[c];
  }
}
__global__ void __launch_bounds__(256) k_dc_mirror_apron(Pool T, Pool F, const uint32_t *__restrict__ perm) {
  const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (size_t)T.M * kAV) return;
  const uint32_t b = (uint32_t)(t / kAV), i = (uint32_t)(t % kAV);
  F.apron[(size_t)perm[b] * kAV + i] = perm_cell(perm, T.apron[t]);
}
__global__ void __launch_bounds__(256) k_dc_mirror_map(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t n,
                                                       const uint32_t *__restrict__ perm) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = src[i];
  dst[i] = v == kNone ? kNone : perm[v];
}
// accessors: field order -> the reference's slot order
__global__ void __launch_bounds__(256) k_dc_unpermute_f32(const float *__restrict__ src, int stride, int comps, const uint32_t *__restrict__ perm,
                                                          float *__restrict__ dst, size_t cells) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= cells) return;
  const size_t f = (size_t)perm[i >> 6] * kBV + (i & 63);
  for (int k = 0; k < comps; k++) dst[comps * i + k] = src[stride * f + k];
}
__global__ void __launch_bounds__(256) k_dc_gather_u32(const float *__restrict__ src, const uint32_t *__restrict__ perm, float *__restrict__ dst,
                                                       uint32_t n) {
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

// ======================================================================================
// slab decomposition: lock-step barrier over peer memory
// ======================================================================================
// flags[r] = rank r's control block (mapped into every process); word j of it = the last epoch rank j announced
// to r.  The epoch lives in device memory and is advanced by the kernel itself, so the barrier can be replayed
// from a CUDA graph.  Bounded spin: a rank that never arrives trips *err instead of hanging the GPU.
struct BarrierPeers {
  volatile uint32_t *flags[8];
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__global__ void __launch_bounds__(32) k_dcs_barrier(BarrierPeers B, int rank, int world, uint32_t *epoch_counter, uint32_t *err) {
  pdl_enter();
  const int t = threadIdx.x;
  uint32_t epoch = 0;
  if (t == 0) epoch = ++(*epoch_counter);
  epoch = __shfl_sync(0xFFFFFFFFu, epoch, 0);
  if (t < world && t != rank) {
    __threadfence_system();    // everything this rank wrote before the barrier is visible system-wide
    B.flags[t][rank] = epoch;  // 4-byte store into the peer's control block over NVLink
    const unsigned long long t0 = global_ns();
    while ((int32_t)(B.flags[rank][t] - epoch) < 0) {
      if (global_ns() - t0 > 10000000000ull) {  // 10 s
        *err = 1;
        break;
      }
    }
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(256) k_copy_f32(const float *__restrict__ src, float *__restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace dcg
