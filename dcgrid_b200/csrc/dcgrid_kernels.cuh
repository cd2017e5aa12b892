// Hand-written sm_100a kernels of the DCGrid solve.  One logical block = 4^3 cells = 64 threads
// (two warps); a CTA carries kBPC logical blocks.  Each kernel cites the reference kernel whose
// results it reproduces bit for bit (compiled with -fmad=false, reference expression order).
#pragma once
#include <cfloat>

#include "dcgrid_layout.cuh"

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

// k_dcgrid_refresh_apron_indices, dcgrid_structure.cu:30-92.  <<<M, 216>>>; rim entries only.
__global__ void __launch_bounds__(216) k_dc_refresh_apron(Pool T, KParams P, const uint32_t *__restrict__ flags) {
  const uint32_t b = blockIdx.x;
  const int ai = threadIdx.x;
  const int4 pl = T.posl[b];
  if (pl.w == kFree) return;
  const int i = ai / kAA, j = (ai / kAW) % kAW, k = ai % kAW;
  if (i % (kAW - 1) != 0 && j % (kAW - 1) != 0 && k % (kAW - 1) != 0) return;
  const int level = pl.w;
  uint32_t *entry = &T.apron[(size_t)b * kAV + ai];
  uint32_t nb = kNone;
  const int nx = pl.x + i - 1, ny = pl.y + j - 1, nz = pl.z + k - 1;
  if (flags[b] & kFlagMoved) {
    const int scale = 1 << level;
    if (nx < 0 || ny < 0 || nz < 0 || nx * scale >= P.gx || ny * scale >= P.gy || nz * scale >= P.gz) {
      // :54-60 feeds APRON coordinates (1..4) to SPREAD where cell coordinates (0..3) are meant
      // (SURVEY App. B-5).  Reproduced on purpose: parity with the reference's Neumann ghosts.
      *entry = b * kBV + spread(min(max(i, 1), kBW), 2) + spread(min(max(j, 1), kBW), 1) + spread(min(max(k, 1), kBW), 0);
      return;
    }
    int nl = level;
    nb = block_index_deep(T, P, nx, ny, nz, nl);
  } else {
    const uint32_t old = *entry;
    const uint32_t prev = old / kBV;
    const uint32_t cf = flags[prev];
    if (cf & kFlagMoved) {
      int nl = level;
      nb = block_index_deep(T, P, nx, ny, nz, nl);
    } else if ((cf & kFlagRefined) && T.posl[prev].w > level) {
      nb = T.child[old / kSV];
    }
  }
  if (nb == kNone) return;
  const int4 np = T.posl[nb];
  const int s = 1 << (np.w - level);
  *entry = nb * kBV + cell_bits(nx / s - np.x, ny / s - np.y, nz / s - np.z);
}

// face table = the 6 x 16 rim entries that 7-point stencils read
__device__ __forceinline__ int face_apron_index(int g) {
  const int f = g >> 4, a = (g >> 2) & 3, b = g & 3;
  const int fixed = (f & 1) ? (kAW - 1) : 0;
  switch (f >> 1) {
    case 0: return kAA * fixed + kAW * (1 + a) + (1 + b);
    case 1: return kAA * (1 + a) + kAW * fixed + (1 + b);
    default: return kAA * (1 + a) + kAW * (1 + b) + fixed;
  }
}
__global__ void __launch_bounds__(256) k_dc_build_faces(Pool T) {
  const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (size_t)T.M * 96) return;
  const uint32_t b = (uint32_t)(t / 96);
  const int g = (int)(t % 96);
  if (T.posl[b].w == kFree) return;
  T.face[t] = T.apron[(size_t)b * kAV + face_apron_index(g)];
}

// accumulate<T>, dcgrid_structure.cu:188-222: parent cell = .125 * sequential sum of the 8 cells of a
// child subblock.  One thread per subblock of `level`.
__global__ void __launch_bounds__(256) k_dc_accumulate_velocity(Pool T, int level, float4 *__restrict__ vw) {
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  if (t >= 8 * T.max_blocks[level]) return;
  const uint32_t sb = 8 * T.offsets[level] + t, b = sb / 8;
  if (T.posl[b].w != level) return;
  const uint32_t ps = T.parent[b];
  if (ps == kNone) return;
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
__global__ void __launch_bounds__(256) k_dc_accumulate_scalar(Pool T, int level, float *__restrict__ ch) {
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  if (t >= 8 * T.max_blocks[level]) return;
  const uint32_t sb = 8 * T.offsets[level] + t, b = sb / 8;
  if (T.posl[b].w != level) return;
  const uint32_t ps = T.parent[b];
  if (ps == kNone) return;
  const float4 lo = *reinterpret_cast<const float4 *>(ch + (size_t)kSV * sb);
  const float4 hi = *reinterpret_cast<const float4 *>(ch + (size_t)kSV * sb + 4);
  float a = 0.f;
  a += lo.x; a += lo.y; a += lo.z; a += lo.w; a += hi.x; a += hi.y; a += hi.z; a += hi.w;
  ch[(size_t)kSV * ps + (sb % 8)] = a * .125f;
}

// ======================================================================================
// adaptation (dcgrid_adaptation.cu)
// ======================================================================================

// k_dcgrid_calc_subblock_scores, :10-40.  finer_full bit l = "level l cannot be refined further"
// ((l==0 && loads[0]==full[0]) || (l>0 && loads[l-1]==full[l-1]), :19-21).
__global__ void __launch_bounds__(256) k_dc_subblock_scores(Pool T, KParams P, uint32_t finer_full, float *__restrict__ sub_scores) {
  const uint32_t sb = blockIdx.x * 256 + threadIdx.x;
  if (sb >= 8 * T.M) return;
  const int4 pl = T.posl[sb / 8];
  if (T.child[sb] != kNone || pl.w == kFree || ((finer_full >> pl.w) & 1)) {
    sub_scores[sb] = -FLT_MAX;
    return;
  }
  const float s = (float)(1 << pl.w);
  const float px = s * ((float)pl.x + 2.f * (float)((sb >> 2) & 1) + 1.f);
  const float py = s * ((float)pl.y + 2.f * (float)((sb >> 1) & 1) + 1.f);
  const float pz = s * ((float)pl.z + 2.f * (float)(sb & 1) + 1.f);
  const float ex = px - .5f * (float)P.gx, ey = py - .45f * (float)P.gy, ez = pz - .5f * (float)P.gz;
  const float d = sqrtf(ex * ex + ey * ey + ez * ez);
  sub_scores[sb] = d < .2f * (float)P.gx ? 0.f : (float)P.gx / d;
}

// k_dcgrid_accumulate_subblock_scores, :42-63.  Sums NINE floats s[0..8]: the 9th
// is subblock 0 of the next pool slot.  sub_scores has 8*M+1 entries, the last one = -FLT_MAX.
__global__ void __launch_bounds__(256) k_dc_block_scores(Pool T, uint32_t finer_full, const float *__restrict__ sub_scores,
                                                         float *__restrict__ block_scores) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= T.M) return;
  float out = -FLT_MAX;
  const int level = T.posl[b].w;
  if (level != kFree && !((finer_full >> level) & 1)) {
    const float *s = sub_scores + 8 * (size_t)b;
    if (s[0] > 0.f && s[1] > 0.f && s[2] > 0.f && s[3] > 0.f && s[4] > 0.f && s[5] > 0.f && s[6] > 0.f && s[7] > 0.f)
      out = .125f * (s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7] + s[8]);
  }
  block_scores[b] = out;
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
__global__ void __launch_bounds__(64) k_dc_propagate(Pool T, KParams P, const uint32_t *__restrict__ touched, int level,
                                                     float4 *__restrict__ vw, float *__restrict__ q, float *__restrict__ fl) {
  const uint32_t b = touched[blockIdx.x];
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
__device__ __forceinline__ DSample d_sample(const Pool &T, const KParams &P, float px, float py, float pz) {
  DSample s;
  int ix = min(max((int)floorf(px), 0), P.gx - 1), iy = min(max((int)floorf(py), 0), P.gy - 1), iz = min(max((int)floorf(pz), 0), P.gz - 1);
  int level = 0;
  const uint32_t b = block_index_deep(T, P, ix, iy, iz, level);
  s.scale = 1 << level;
  const float inv = 1.f / (float)s.scale;
  const float x = px * inv - .5f, y = py * inv - .5f, z = pz * inv - .5f;
  const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  s.fx = x - xf; s.fy = y - yf; s.fz = z - zf;
  s.x0 = (int)xf; s.y0 = (int)yf; s.z0 = (int)zf;
  const int4 bp = T.posl[b];
  const int i = min(max(s.x0 + 1 - bp.x, 0), kAW - 2), j = min(max(s.y0 + 1 - bp.y, 0), kAW - 2), k = min(max(s.z0 + 1 - bp.z, 0), kAW - 2);
  const uint32_t *a = T.apron + (size_t)b * kAV + kAA * i + kAW * j + k;
  s.id[0] = a[0]; s.id[1] = a[1]; s.id[2] = a[kAW]; s.id[3] = a[kAW + 1];
  s.id[4] = a[kAA]; s.id[5] = a[kAA + 1]; s.id[6] = a[kAA + kAW]; s.id[7] = a[kAA + kAW + 1];
  return s;
}

// k_dcgrid_advect_velocity, dcgrid_fluid.cu:74-91,112-127.  Writes the other ping-pong buffer (no
// whole-pool D2D memcpy, fluid_simulation_dcgrid.cu:265-266).
__global__ void __launch_bounds__(kCTA) k_dc_advect_velocity(Pool T, KParams P, const float4 *__restrict__ vin, float4 *__restrict__ vout) {
  const uint32_t b = blockIdx.x * kBPC + (threadIdx.x >> 6);
  if (b >= T.M) return;
  const int4 pl = T.posl[b];
  if (pl.w == kFree) return;
  const uint32_t t = threadIdx.x & 63, c = b * kBV + t;
  const float4 me = vin[c];
  float3 out = make_float3(0.f, 0.f, 0.f);
  if (T.child[c >> 3] == kNone) {
    const float scale = (float)(1 << pl.w);
    const float alpha = P.dt * P.rdx;
    const float bx = ((float)(pl.x | cell_x(t)) + .5f) * scale - me.x * alpha;
    const float by = ((float)(pl.y | cell_y(t)) + .5f) * scale - me.y * alpha;
    const float bz = ((float)(pl.z | cell_z(t)) + .5f) * scale - me.z * alpha;
    const DSample s = d_sample(T, P, bx, by, bz);
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
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float3 v = velocity_bc(P, make_float3(cv[k].x, cv[k].y, cv[k].z), s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1),
                                     s.z0 + (k & 1), s.scale);
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
  const uint32_t b = blockIdx.x * kBPC + (threadIdx.x >> 6);
  if (b >= T.M) return;
  const int4 pl = T.posl[b];
  if (pl.w == kFree) return;
  const uint32_t t = threadIdx.x & 63, c = b * kBV + t;
  float out = 0.f;
  if (T.child[c >> 3] == kNone) {
    const float4 me = vw[c];
    const float scale = (float)(1 << pl.w);
    const float alpha = P.dt * P.rdx;
    const float bx = ((float)(pl.x | cell_x(t)) + .5f) * scale - me.x * alpha;
    const float by = ((float)(pl.y | cell_y(t)) + .5f) * scale - me.y * alpha;
    const float bz = ((float)(pl.z | cell_z(t)) + .5f) * scale - me.z * alpha;
    const DSample s = d_sample(T, P, bx, by, bz);
    float qv[8], f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      f[k] = fl[s.id[k]];
      qv[k] = qin[s.id[k]];
    }
    const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
    if (!(W.acc < 1e-6f)) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        qv[k] = density_bc(P, qv[k], s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1), s.z0 + (k & 1), s.scale);
      out = blend8(qv, W.w);
    }
  }
  qout[c] = out;
}

// Stage one block's 6^3 apron of a scalar field in shared memory: own 64 cells + 6 x 16 face ghosts
// through the face table.  `sp` = this block's 216-float tile.
__device__ __forceinline__ void stage_scalar(float *sp, const float *__restrict__ src, const uint32_t *__restrict__ face, uint32_t b, uint32_t t) {
  sp[apron_of(t)] = src[b * kBV + t];
  const uint32_t *f = face + (size_t)b * 96;
  sp[face_apron_index(t)] = src[f[t]];
  if (t < 32) sp[face_apron_index(64 + t)] = src[f[64 + t]];
}

// k_dcgrid_calc_divergence, dcgrid_fluid.cu:174-230
__global__ void __launch_bounds__(kCTA) k_dc_divergence(Pool T, KParams P, const float4 *__restrict__ vw, float *__restrict__ div,
                                                        float *__restrict__ p, float *__restrict__ tp) {
  __shared__ float4 sv[kBPC][kAV];
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t b = blockIdx.x * kBPC + g;
  int4 pl = make_int4(0, 0, 0, kFree);
  if (b < T.M) pl = T.posl[b];
  const bool active = pl.w != kFree;
  const uint32_t c = b * kBV + t;
  const int scale = active ? (1 << pl.w) : 1;
  const int x = pl.x | cell_x(t), y = pl.y | cell_y(t), z = pl.z | cell_z(t);
  if (active) {
    float4 *s = sv[g];
    s[apron_of(t)] = vw[c];
    const uint32_t *f = T.face + (size_t)b * 96;
    // ghost cells: boundary conditions are evaluated at the ghost's own position (:193-210)
    for (int gi = t; gi < 96; gi += 64) {
      const int fc = gi >> 4, a = (gi >> 2) & 3, bb = gi & 3;
      int gx, gy, gz;
      const int fixed = (fc & 1) ? kBW : -1;
      if ((fc >> 1) == 0) { gx = fixed; gy = a; gz = bb; }
      else if ((fc >> 1) == 1) { gx = a; gy = fixed; gz = bb; }
      else { gx = a; gy = bb; gz = fixed; }
      const float4 v = vw[f[gi]];
      const float3 vb = velocity_bc(P, make_float3(v.x, v.y, v.z), pl.x + gx, pl.y + gy, pl.z + gz, scale);
      s[face_apron_index(gi)] = make_float4(vb.x, vb.y, vb.z, v.w);
    }
  }
  __syncthreads();
  if (!active) return;
  p[c] = 0.f;
  tp[c] = 0.f;
  float d = 0.f;
  if (T.child[c >> 3] == kNone) {
    const float4 *s = sv[g];
    const int ai = apron_of(t);
    const float alpha = .5f * P.rdx / (float)scale;
    const float4 l = s[ai - kAA], r = s[ai + kAA], dn = s[ai - kAW], up = s[ai + kAW], bk = s[ai - 1], fr = s[ai + 1];
    d = alpha * (r.w * r.x - l.w * l.x + up.w * up.y - dn.w * dn.y + fr.w * fr.z - bk.w * bk.z);
  }
  div[c] = d;
  (void)x; (void)y; (void)z;
}

// k_dcgrid_jacobi / k_dcgrid_jacobi_inv, dcgrid_multigrid_solver.cu:5-41.  `in`/`out` = (pressure,
// t_pressure) or (t_pressure, pressure).  Ghosts of coarser neighbours are read from `in` of the
// coarse cell — for jacobi_inv that is the coarse level's t_pressure, as in the reference (:32-37).
__global__ void __launch_bounds__(kCTA) k_dc_jacobi(Pool T, KParams P, int level, const float *__restrict__ in, float *__restrict__ out,
                                                    const float *__restrict__ div) {
  __shared__ float sp[kBPC][kAV];
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t li = blockIdx.x * kBPC + g;
  const uint32_t b = T.offsets[level] + li;
  const bool active = li < T.max_blocks[level] && T.posl[b].w == level;
  if (active) stage_scalar(sp[g], in, T.face, b, t);
  __syncthreads();
  if (!active) return;
  const float alpha = (float)((1 << level) * (1 << level)) * P.dx * P.dx;
  const float *s = sp[g];
  const int ai = apron_of(t);
  const uint32_t c = b * kBV + t;
  out[c] = (s[ai - kAA] + s[ai + kAA] + s[ai - kAW] + s[ai + kAW] + s[ai - 1] + s[ai + 1] - alpha * div[c]) / 6.f;
}

// k_dcgrid_prolongate, dcgrid_multigrid_solver.cu:43-76
__global__ void __launch_bounds__(kCTA) k_dc_prolongate(Pool T, int level, float *__restrict__ p) {
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t li = blockIdx.x * kBPC + g;
  if (li >= T.max_blocks[level]) return;
  const uint32_t b = T.offsets[level] + li;
  const int4 pl = T.posl[b];
  if (pl.w != level) return;
  const uint32_t ps = T.parent[b];
  if (ps == kNone) return;
  const uint32_t *pa = T.apron + (size_t)(ps / 8) * kAV;
  const int x = pl.x | cell_x(t), y = pl.y | cell_y(t), z = pl.z | cell_z(t);
  const int idx = kAA * (1 + (x / 2) % kBW) + kAW * (1 + (y / 2) % kBW) + (1 + (z / 2) % kBW);
  const int i = x % 2 ? kAA : -kAA, j = y % 2 ? kAW : -kAW, k = z % 2 ? 1 : -1;
  const float p000 = p[pa[idx]], p001 = p[pa[idx + k]], p010 = p[pa[idx + j]], p100 = p[pa[idx + i]];
  const float p011 = p[pa[idx + j + k]], p101 = p[pa[idx + i + k]], p110 = p[pa[idx + i + j]], p111 = p[pa[idx + i + j + k]];
  p[b * kBV + t] = (27.f * p000 + 9.f * (p001 + p010 + p100) + 3.f * (p011 + p101 + p110) + p111) / 64.f;
}

// k_dcgrid_apply_pressure, dcgrid_fluid.cu:232-259
__global__ void __launch_bounds__(kCTA) k_dc_apply_pressure(Pool T, KParams P, const float *__restrict__ p, const float *__restrict__ fl,
                                                            float4 *__restrict__ vw) {
  __shared__ float sp[kBPC][kAV];
  __shared__ float sw[kBPC][kAV];
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t b = blockIdx.x * kBPC + g;
  int level = kFree;
  if (b < T.M) level = T.posl[b].w;
  const bool active = level != kFree;
  if (active) {
    stage_scalar(sp[g], p, T.face, b, t);
    stage_scalar(sw[g], fl, T.face, b, t);
  }
  __syncthreads();
  if (!active) return;
  const uint32_t c = b * kBV + t;
  if (T.child[c >> 3] != kNone) return;
  const float alpha = .5f * P.rdx / (float)(1 << level);
  const float *s = sp[g], *w = sw[g];
  const int ai = apron_of(t);
  const float pc = s[ai];
  float4 v = vw[c];
  v.x -= alpha * (w[ai + kAA] * (s[ai + kAA] - pc) + w[ai - kAA] * (pc - s[ai - kAA]));
  v.y -= alpha * (w[ai + kAW] * (s[ai + kAW] - pc) + w[ai - kAW] * (pc - s[ai - kAW]));
  v.z -= alpha * (w[ai + 1] * (s[ai + 1] - pc) + w[ai - 1] * (pc - s[ai - 1]));
  vw[c] = v;
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
__global__ void __launch_bounds__(256) k_dc_total_density(Pool T, const float *__restrict__ q, const float *__restrict__ fl,
                                                          double *__restrict__ partial) {
  __shared__ double sh[256];
  double s = 0.0;
  const size_t n = (size_t)T.M * kBV;
  for (size_t c = (size_t)blockIdx.x * 256 + threadIdx.x; c < n; c += (size_t)gridDim.x * 256) {
    const int level = T.posl[c >> 6].w;
    if (level == kFree || T.child[c >> 3] != kNone) continue;
    const double vol = (double)(1ull << (3 * level));
    s += vol * (double)(q[c] * fl[c]);
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

}  // namespace dcg
