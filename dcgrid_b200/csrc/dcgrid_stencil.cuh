// 7-point stencil kernels of the DCGrid solve (divergence, Jacobi sweeps, pressure gradient), second
// generation: 4 cells per thread (one float4 = the 1x2x2 quad that is contiguous in the reference's
// in-block bit layout), 16 threads per logical 4^3 block, 16 blocks per 256-thread CTA, neighbours
// exchanged through half-warp shuffles (no shared memory, no barrier), and NO per-cell index table on
// the hot path.
//
// Index traffic.  The reference resolves every neighbour through cellIndices[216] (8-byte ids,
// 1,728 B per block and sweep).  Every face of that 6^3 apron map is one of three
// closed forms (checked entry by entry when the map changes, k_dc_build_fdesc):
//   * all 16 ghosts lie in ONE block `nb` that is dl >= 0 levels coarser (dl = 0: the same-level
//     neighbour; nb = the block itself: the ordered-level wall clamp, dcgrid_structure.cu:157-167),
//     at cell  o + (a >> dl, b >> dl)  of that block;
//   * the moved-block wall entry of k_dcgrid_refresh_apron_indices (dcgrid_structure.cu:54-60, SURVEY
//     App. B-5): own cell ((a+1)&3, (b+1)&3) on the tangential axes;
//   * anything else marks the block irregular and its 96 face ids are read from an explicit table.
// A block therefore carries 12 words (6 neighbour slots + 6 codes) = 48 B instead of 1,728 B, and the
// ids the kernels use are, by construction, exactly the entries of the reference's apron map.
//
// Numerics: same expressions in the same order as dcgrid_kernels.cuh / the reference; -fmad=false.
#pragma once
#include "dcgrid_layout.cuh"

namespace dcg {

constexpr int kQT = 16;    // threads per logical block (4 cells each)
constexpr int kB4 = 16;    // logical blocks per CTA
constexpr int kCTA4 = kQT * kB4;
constexpr uint32_t kFdIrregular = 0x80000000u;
constexpr uint32_t kFdQuirk = 1u << 4;
// in-block cell bits of the two tangential axes of a face (the normal axis' bits cleared):
// x faces keep (y,z) = 0b011011, y faces (x,z) = 0b101101, z faces (x,y) = 0b110110
__host__ __device__ __forceinline__ uint32_t tang_mask(int axis) { return axis == 0 ? 27u : (axis == 1 ? 45u : 54u); }

// Face descriptor = (base, code).  base = id of the ghost cell at tangential (0,0) [for the quirk form:
// with its tangential bits cleared]; code = dl | quirk << 4 (| kFdIrregular on face 0).  The ghost at
// tangential (a,b) is base + bits(ta) + bits(tb) with ta = a >> dl (quirk: (a+1)&3): the sum never
// carries because the tangential origin is 0 for dl = 0, even for dl = 1, and ta = 0 for dl >= 2.
// code == 0 (same-level neighbour, the common case): ghost = base + the tangential bits of the cell
// the ghost is for, which each lane already knows.
__device__ __forceinline__ uint32_t fd_ghost(uint32_t base, uint32_t code, int axis, int a, int b) {
  const int dl = (int)(code & 15u);
  int ta, tb;
  if (code & kFdQuirk) {
    ta = (a + 1) & 3;
    tb = (b + 1) & 3;
  } else {
    ta = a >> dl;
    tb = b >> dl;
  }
  const int s1 = axis == 0 ? 1 : 2, s2 = axis == 2 ? 1 : 0;
  return base + spread(ta, s1) + spread(tb, s2);
}
// `own_tang` = tangential bits of the cell the ghost belongs to (its in-block index & tang_mask(axis))
__device__ __forceinline__ uint32_t fd_ghost_fast(uint32_t base, uint32_t code, int axis, int a, int b, uint32_t own_tang) {
  if (code == 0) return base + own_tang;
  return fd_ghost(base, code, axis, a, b);
}

// apron index of ghost g = 16*f + 4*a + b (f = -x,+x,-y,+y,-z,+z)
__device__ __forceinline__ int face_apron_index(int g) {
  const int f = g >> 4, a = (g >> 2) & 3, b = g & 3;
  const int fixed = (f & 1) ? (kAW - 1) : 0;
  switch (f >> 1) {
    case 0: return kAA * fixed + kAW * (1 + a) + (1 + b);
    case 1: return kAA * (1 + a) + kAW * fixed + (1 + b);
    default: return kAA * (1 + a) + kAW * (1 + b) + fixed;
  }
}

// Derives the 6 face descriptors of every active block from its apron map, verifying all 16 entries
// of each face against the closed form.  8 lanes per block (6 used).  Irregular blocks get their 96
// face ids copied to T.face and are counted in *irregular.
__global__ void __launch_bounds__(256) k_dc_build_fdesc(Pool T, uint32_t *__restrict__ irregular) {
  const uint32_t gid = blockIdx.x * 256 + threadIdx.x;
  const uint32_t b = gid >> 3;
  const int f = (int)(gid & 7u);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned grp_mask = 0xFFu << (lane & 24u);
  bool bad = false;
  uint32_t base = kNone, code = 0;
  uint32_t e[16];
  const bool live = b < T.M && f < 6 && T.posl[b].w != kFree;
  if (live) {
    const uint32_t *ap = T.apron + (size_t)b * kAV;
#pragma unroll
    for (int k = 0; k < 16; k++) e[k] = ap[face_apron_index(16 * f + k)];
    const int axis = f >> 1;
    const int level = T.posl[b].w;
    if (e[0] == kNone) bad = true;
    if (!bad) {
      const uint32_t nb = e[0] >> 6;
      const int nl = nb < T.M ? T.posl[nb].w : kFree;
      const int dl = nl - level;
      bool regular = nl != kFree && dl >= 0 && dl <= 15;
      if (regular) {
        base = e[0];
        code = (uint32_t)dl;
#pragma unroll
        for (int k = 0; k < 16; k++) regular = regular && e[k] == fd_ghost(base, code, axis, k >> 2, k & 3);
      }
      if (!regular) {
        // moved-block wall entry: own block, normal coordinate from e[0], tangential (a+1)&3
        base = e[0] & ~tang_mask(axis);
        code = kFdQuirk;
        bool quirk = nb == b;
#pragma unroll
        for (int k = 0; k < 16; k++) quirk = quirk && e[k] == fd_ghost(base, code, axis, k >> 2, k & 3);
        bad = !quirk;
      }
    }
  }
  const unsigned any_bad = __ballot_sync(0xFFFFFFFFu, bad) & grp_mask;
  if (!live) return;
  uint32_t *fd = T.fd + 12 * (size_t)b;
  fd[f] = base;
  fd[6 + f] = code | ((any_bad && f == 0) ? kFdIrregular : 0u);
  if (any_bad) {
#pragma unroll
    for (int k = 0; k < 16; k++) T.face[(size_t)b * 96 + 16 * f + k] = e[k];
    if (f == 0) atomicAdd(irregular, 1u);
  }
}

// level pools are consecutive slot ranges whose active blocks form a compact prefix
__device__ __forceinline__ bool slot_is_active(const Pool &T, uint32_t b) {
  if (b >= T.M) return false;
  int level = 0;
  while (level + 1 < T.levels && b >= T.offsets[level + 1]) level++;
  return b - T.offsets[level] < T.loads[level];
}

// thread t of a block owns the quad (X, Y0..Y0+1, Z0..Z0+1): t = sx<<3 | sy<<2 | sz<<1 | cx
__device__ __forceinline__ void quad_coords(int t, int &X, int &Y0, int &Z0) {
  X = ((t >> 3) << 1) | (t & 1);
  Y0 = ((t >> 2) & 1) << 1;
  Z0 = ((t >> 1) & 1) << 1;
}

// Ghost ids a quad needs itself: 4 on an x face (only quads with X = 0 or 3), 2 on its y face, 2 on its
// z face (every quad touches exactly one y and one z face of the block).
struct QuadGhosts {
  uint32_t x[4];  // cells k = 2*cy+cz, valid iff has_x
  uint32_t y[2];  // [cz], the quad's boundary row (cy = sy)
  uint32_t z[2];  // [cy], the quad's boundary column (cz = sz)
  bool has_x;
};
// w0..w2 = the block's 12 descriptor words (6 bases, 6 codes)
__device__ __forceinline__ QuadGhosts quad_ghosts_from(const Pool &T, uint32_t b, int t, const uint4 w0, const uint4 w1, const uint4 w2) {
  int X, Y0, Z0;
  quad_coords(t, X, Y0, Z0);
  const int sy = (t >> 2) & 1, sz = (t >> 1) & 1;
  const int fx = X == 3 ? 1 : 0, fy = 2 + sy, fz = 4 + sz;
  QuadGhosts q;
  q.has_x = X == 0 || X == 3;
  if (w1.z & kFdIrregular) {
    const uint32_t *ft = T.face + (size_t)b * 96;
#pragma unroll
    for (int k = 0; k < 4; k++) q.x[k] = q.has_x ? ft[16 * fx + 4 * (Y0 + (k >> 1)) + (Z0 + (k & 1))] : 0u;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      q.y[k] = ft[16 * fy + 4 * X + (Z0 + k)];
      q.z[k] = ft[16 * fz + 4 * X + (Y0 + k)];
    }
    return q;
  }
  const uint32_t nbx = fx ? w0.y : w0.x, cdx = (fx ? w1.w : w1.z) & ~kFdIrregular;
  const uint32_t nby = sy ? w0.w : w0.z, cdy = sy ? w2.y : w2.x;
  const uint32_t nbz = sz ? w1.y : w1.x, cdz = sz ? w2.w : w2.z;
  const uint32_t c = 4u * (uint32_t)t;  // in-block index of the quad's first cell
  if ((cdx | cdy | cdz) == 0) {
    // all three faces border same-level blocks (or the ordered-level wall clamp): ghost = face base + the
    // tangential bits of the cell it belongs to
#pragma unroll
    for (int k = 0; k < 4; k++) q.x[k] = nbx + ((c + k) & 27u);
#pragma unroll
    for (int k = 0; k < 2; k++) {
      q.y[k] = nby + ((c + 2 * sy + k) & 45u);
      q.z[k] = nbz + ((c + 2 * k + sz) & 54u);
    }
    return q;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) q.x[k] = q.has_x ? fd_ghost(nbx, cdx, 0, Y0 + (k >> 1), Z0 + (k & 1)) : 0u;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    q.y[k] = fd_ghost(nby, cdy, 1, X, Z0 + k);
    q.z[k] = fd_ghost(nbz, cdz, 2, X, Y0 + k);
  }
  return q;
}

__device__ __forceinline__ QuadGhosts quad_ghosts(const Pool &T, uint32_t b, int t) {
  const uint4 *p = reinterpret_cast<const uint4 *>(T.fd + 12 * (size_t)b);
  return quad_ghosts_from(T, b, t, __ldg(p), __ldg(p + 1), __ldg(p + 2));
}

// The 6-neighbourhood of a quad of one scalar field, exchanged through half-warp shuffles (no shared
// memory): in-block neighbours come from the registers of the lane that owns them, block-boundary
// neighbours from the ghost values this lane loaded.
struct QuadNbr {
  float4 xm, xp;             // x-1 / x+1 neighbours of cells k = 0..3
  float ym0, ym1, yp0, yp1;  // y-1 of (cy=0; cz=0,1), y+1 of (cy=1; cz=0,1)
  float zm0, zm1, zp0, zp1;  // z-1 of (cz=0; cy=0,1), z+1 of (cz=1; cy=0,1)
};
__device__ __forceinline__ void quad_exchange_x(QuadNbr &n, float4 own, int t, float4 gx) {
  const unsigned hm = 0xFFFFFFFFu;  // callers keep whole warps converged (inactive halves run along)
  const int lb = threadIdx.x & 16;
  const int cx = t & 1, sx = t >> 3;
  const int lxm = lb + ((cx ? t - 1 : t - 7) & 15), lxp = lb + ((cx ? t + 7 : t + 1) & 15);
  n.xm.x = __shfl_sync(hm, own.x, lxm); n.xm.y = __shfl_sync(hm, own.y, lxm);
  n.xm.z = __shfl_sync(hm, own.z, lxm); n.xm.w = __shfl_sync(hm, own.w, lxm);
  n.xp.x = __shfl_sync(hm, own.x, lxp); n.xp.y = __shfl_sync(hm, own.y, lxp);
  n.xp.z = __shfl_sync(hm, own.z, lxp); n.xp.w = __shfl_sync(hm, own.w, lxp);
  if (!cx && !sx) n.xm = gx;
  if (cx && sx) n.xp = gx;
}
__device__ __forceinline__ void quad_exchange_y(QuadNbr &n, float4 own, int t, float gy0, float gy1) {
  const unsigned hm = 0xFFFFFFFFu;  // callers keep whole warps converged (inactive halves run along)
  const int lb = threadIdx.x & 16;
  const int sy = (t >> 2) & 1;
  const int lym = lb + ((t - 4) & 15), lyp = lb + ((t + 4) & 15);
  n.ym0 = __shfl_sync(hm, own.z, lym); n.ym1 = __shfl_sync(hm, own.w, lym);
  n.yp0 = __shfl_sync(hm, own.x, lyp); n.yp1 = __shfl_sync(hm, own.y, lyp);
  if (!sy) { n.ym0 = gy0; n.ym1 = gy1; } else { n.yp0 = gy0; n.yp1 = gy1; }
}
__device__ __forceinline__ void quad_exchange_z(QuadNbr &n, float4 own, int t, float gz0, float gz1) {
  const unsigned hm = 0xFFFFFFFFu;  // callers keep whole warps converged (inactive halves run along)
  const int lb = threadIdx.x & 16;
  const int sz = (t >> 1) & 1;
  const int lzm = lb + ((t - 2) & 15), lzp = lb + ((t + 2) & 15);
  n.zm0 = __shfl_sync(hm, own.y, lzm); n.zm1 = __shfl_sync(hm, own.w, lzm);
  n.zp0 = __shfl_sync(hm, own.x, lzp); n.zp1 = __shfl_sync(hm, own.z, lzp);
  if (!sz) { n.zm0 = gz0; n.zm1 = gz1; } else { n.zp0 = gz0; n.zp1 = gz1; }
}
__device__ __forceinline__ QuadNbr quad_exchange(float4 own, int t, float4 gx, float gy0, float gy1, float gz0, float gz1) {
  QuadNbr n;
  quad_exchange_x(n, own, t, gx);
  quad_exchange_y(n, own, t, gy0, gy1);
  quad_exchange_z(n, own, t, gz0, gz1);
  return n;
}
__device__ __forceinline__ QuadNbr quad_neighbours(const float *__restrict__ src, float4 own, int t, const QuadGhosts &q) {
  float4 gx = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q.has_x) gx = make_float4(src[q.x[0]], src[q.x[1]], src[q.x[2]], src[q.x[3]]);
  const float gy0 = src[q.y[0]], gy1 = src[q.y[1]], gz0 = src[q.z[0]], gz1 = src[q.z[1]];
  return quad_exchange(own, t, gx, gy0, gy1, gz0, gz1);
}

// Ghost values of one scalar field for a quad, with vector loads where the face is a same-level neighbour
// (code 0, the common case): the 4 x-face ghosts of a quad are 16 contiguous bytes of the neighbour block,
// the 2 y-face ghosts 8 contiguous bytes, the 2 z-face ghosts elements 0 and 2 of one aligned float4 —
// 3 load instructions touching 3 cache lines instead of 8 touching up to 8 (the L1 tag stage, not DRAM,
// was the busiest unit of the scalar version: profiles/README.md).  Other faces use the scalar ids.
struct GhostVals {
  float4 gx;
  float gy0, gy1, gz0, gz1;
};
__device__ __forceinline__ GhostVals quad_ghost_values(const Pool &T, const float *__restrict__ src, uint32_t b, int t, const uint4 w0,
                                                      const uint4 w1, const uint4 w2) {
  GhostVals g;
  g.gx = make_float4(0.f, 0.f, 0.f, 0.f);
  const int sy = (t >> 2) & 1, sz = (t >> 1) & 1;
  const int X = ((t >> 3) << 1) | (t & 1);
  const bool has_x = X == 0 || X == 3;
  const int fx = X == 3 ? 1 : 0;
  const uint32_t c = 4u * (uint32_t)t;
  if (!(w1.z & kFdIrregular)) {
    const uint32_t nbx = fx ? w0.y : w0.x, cdx = fx ? w1.w : w1.z;
    const uint32_t nby = sy ? w0.w : w0.z, cdy = sy ? w2.y : w2.x;
    const uint32_t nbz = sz ? w1.y : w1.x, cdz = sz ? w2.w : w2.z;
    if (has_x) {
      if (cdx == 0) {
        g.gx = *reinterpret_cast<const float4 *>(src + nbx + (c & 27u));
      } else {
        const int Y0 = sy << 1, Z0 = sz << 1;
        g.gx.x = src[fd_ghost(nbx, cdx, 0, Y0, Z0)];
        g.gx.y = src[fd_ghost(nbx, cdx, 0, Y0, Z0 + 1)];
        g.gx.z = src[fd_ghost(nbx, cdx, 0, Y0 + 1, Z0)];
        g.gx.w = src[fd_ghost(nbx, cdx, 0, Y0 + 1, Z0 + 1)];
      }
    }
    if (cdy == 0) {
      const float2 v = *reinterpret_cast<const float2 *>(src + nby + (c & 45u));
      g.gy0 = v.x; g.gy1 = v.y;
    } else {
      g.gy0 = src[fd_ghost(nby, cdy, 1, X, sz << 1)];
      g.gy1 = src[fd_ghost(nby, cdy, 1, X, (sz << 1) + 1)];
    }
    if (cdz == 0) {
      const float4 v = *reinterpret_cast<const float4 *>(src + nbz + (c & 54u));
      g.gz0 = v.x; g.gz1 = v.z;
    } else {
      g.gz0 = src[fd_ghost(nbz, cdz, 2, X, sy << 1)];
      g.gz1 = src[fd_ghost(nbz, cdz, 2, X, (sy << 1) + 1)];
    }
    return g;
  }
  const QuadGhosts q = quad_ghosts_from(T, b, t, w0, w1, w2);
  if (q.has_x) g.gx = make_float4(src[q.x[0]], src[q.x[1]], src[q.x[2]], src[q.x[3]]);
  g.gy0 = src[q.y[0]]; g.gy1 = src[q.y[1]]; g.gz0 = src[q.z[0]]; g.gz1 = src[q.z[1]];
  return g;
}

// ---- k_dcgrid_jacobi / k_dcgrid_jacobi_inv, dcgrid_multigrid_solver.cu:5-41 -------------------------
// Active blocks of a level are the compact slot prefix [offset, offset + blockLoads): slots come
// from freeBlockIndices[offset + load] with the identity free list (fluid_simulation_dcgrid.cu:243-249,
// dcgrid_utils.cuh:118-127) and deleteBlock is never called.
// Sum order of the reference: left + right + down + up + back + front.
__global__ void __launch_bounds__(kCTA4) k_dc_jacobi4(Pool T, KParams P, TileRuns R, int level, const float *__restrict__ in,
                                                      float *__restrict__ out, const float *__restrict__ div) {
  pdl_enter();
  const uint32_t g = threadIdx.x >> 4;
  const int t = threadIdx.x & 15;
  const uint32_t li = run_tile(R, blockIdx.x) * kB4 + g;
  // the two half-warps of a warp hold two blocks; a trailing inactive half runs along (on the level's
  // first block, results discarded) so that the shuffles below stay full-warp collectives
  const bool active = li < T.loads[level];
  if (!__any_sync(0xFFFFFFFFu, active)) return;
  const uint32_t b = T.offsets[level] + (active ? li : 0u);
  const size_t c0 = (size_t)b * kBV + 4 * t;
  const float4 own = *reinterpret_cast<const float4 *>(in + c0);
  const float4 dv = __ldcs(reinterpret_cast<const float4 *>(div + c0));  // streamed: keep L2 for the ghosts
  const uint4 *fdp = reinterpret_cast<const uint4 *>(T.fd + 12 * (size_t)b);
  const GhostVals gv = quad_ghost_values(T, in, b, t, __ldg(fdp), __ldg(fdp + 1), __ldg(fdp + 2));
  const QuadNbr n = quad_exchange(own, t, gv.gx, gv.gy0, gv.gy1, gv.gz0, gv.gz1);
  const float alpha = (float)((1 << level) * (1 << level)) * P.dx * P.dx;
  float4 o;
  o.x = div6(n.xm.x + n.xp.x + n.ym0 + own.z + n.zm0 + own.y - alpha * dv.x);
  o.y = div6(n.xm.y + n.xp.y + n.ym1 + own.w + own.x + n.zp0 - alpha * dv.y);
  o.z = div6(n.xm.z + n.xp.z + own.x + n.yp0 + n.zm1 + own.w - alpha * dv.z);
  o.w = div6(n.xm.w + n.xp.w + own.y + n.yp1 + own.z + n.zp1 - alpha * dv.w);
  if (active) *reinterpret_cast<float4 *>(out + c0) = o;
}

// ---- k_dcgrid_calc_divergence, dcgrid_fluid.cu:174-230 -----------------------------------------------
// Each lane forms fluidity*velocity-component products of its quad; x products travel to the x
// neighbours, y to y, z to z.  Ghosts pass through velocityBndCond at the ghost's own-level position
// first (:193-210).
__device__ __forceinline__ float ghost_product(const KParams &P, const float4 *__restrict__ vw, uint32_t id, int axis, int gx, int gy, int gz,
                                               int scale) {
  // only one component and the fluidity of the ghost are needed: two 4-byte loads (2 L1 wavefronts per warp
  // request) instead of one 16-byte load (4)
  const float *g = reinterpret_cast<const float *>(vw + id);
  const float w = g[3];
  float c = g[axis];
  const int k = bc_kind(P, gx, gy, gz, scale);  // velocityBndCond on the component (sim_utils.cu:24-39)
  if (k == 1) c = axis == 1 ? P.vel_rate : 0.f;
  else if (k == 2) c = 0.f;
  return w * c;
}
// zero_from: the reference clears pressure and t_pressure of the whole pool here (:186-187).  In project() every
// level below the coarsest gets its pressure from the prolongation and its t_pressure from the level's first
// sweep before either is read (ghosts included: they lie in coarser levels, which are finished by then), so
// only blocks of level >= zero_from (= the coarsest level) need the stores; project_local() starts every
// level from zero and passes 0.
__global__ void __launch_bounds__(kCTA4) k_dc_divergence4(Pool T, KParams P, const float4 *__restrict__ vw, float *__restrict__ div,
                                                          float *__restrict__ p, float *__restrict__ tp, int zero_from) {
  pdl_enter();
  const uint32_t g = threadIdx.x >> 4;
  const int t = threadIdx.x & 15;
  const uint32_t b0 = blockIdx.x * kB4 + g;
  const uint32_t b = b0 < T.M ? b0 : T.M - 1;  // out-of-range half: runs along on the last slot, results discarded
  const int4 pl = T.posl[b];                   // free slots carry level 0xFF; their (zeroed) fields are loaded but unused
  const bool active = b0 < T.M && pl.w != kFree;
  const size_t c0 = (size_t)b * kBV + 4 * t;
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = vw[c0 + k];
  const uint32_t child = T.child[(size_t)b * 8 + (t >> 1)];
  if (!__any_sync(0xFFFFFFFFu, active)) return;
  const QuadGhosts q = quad_ghosts(T, b, t);
  int X, Y0, Z0;
  quad_coords(t, X, Y0, Z0);
  const int scale = 1 << (pl.w & 15);
  const int sy = (t >> 2) & 1, sz = (t >> 1) & 1;
  float4 gx = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q.has_x) {
    const int nx = pl.x + (X == 3 ? kBW : -1);
    gx.x = ghost_product(P, vw, q.x[0], 0, nx, pl.y + Y0, pl.z + Z0, scale);
    gx.y = ghost_product(P, vw, q.x[1], 0, nx, pl.y + Y0, pl.z + Z0 + 1, scale);
    gx.z = ghost_product(P, vw, q.x[2], 0, nx, pl.y + Y0 + 1, pl.z + Z0, scale);
    gx.w = ghost_product(P, vw, q.x[3], 0, nx, pl.y + Y0 + 1, pl.z + Z0 + 1, scale);
  }
  const int ny = pl.y + (sy ? kBW : -1), nz = pl.z + (sz ? kBW : -1);
  const float gy0 = ghost_product(P, vw, q.y[0], 1, pl.x + X, ny, pl.z + Z0, scale);
  const float gy1 = ghost_product(P, vw, q.y[1], 1, pl.x + X, ny, pl.z + Z0 + 1, scale);
  const float gz0 = ghost_product(P, vw, q.z[0], 2, pl.x + X, pl.y + Y0, nz, scale);
  const float gz1 = ghost_product(P, vw, q.z[1], 2, pl.x + X, pl.y + Y0 + 1, nz, scale);
  const float4 px = make_float4(v[0].w * v[0].x, v[1].w * v[1].x, v[2].w * v[2].x, v[3].w * v[3].x);
  const float4 py = make_float4(v[0].w * v[0].y, v[1].w * v[1].y, v[2].w * v[2].y, v[3].w * v[3].y);
  const float4 pz = make_float4(v[0].w * v[0].z, v[1].w * v[1].z, v[2].w * v[2].z, v[3].w * v[3].z);
  // x products to x neighbours, y to y, z to z: three exchanges, each using only its own axis
  QuadNbr nx_, ny_, nz_;
  quad_exchange_x(nx_, px, t, gx);
  quad_exchange_y(ny_, py, t, gy0, gy1);
  quad_exchange_z(nz_, pz, t, gz0, gz1);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 d = zero;
  const bool leaf = active && child == kNone;
  if (leaf) {
    const float alpha = .5f * P.rdx / (float)scale;
    // alpha * (r - l + u - d + f - b), k = 2*cy + cz
    d.x = alpha * (nx_.xp.x - nx_.xm.x + py.z - ny_.ym0 + pz.y - nz_.zm0);
    d.y = alpha * (nx_.xp.y - nx_.xm.y + py.w - ny_.ym1 + nz_.zp0 - pz.x);
    d.z = alpha * (nx_.xp.z - nx_.xm.z + ny_.yp0 - py.x + pz.w - nz_.zm1);
    d.w = alpha * (nx_.xp.w - nx_.xm.w + ny_.yp1 - py.y + nz_.zp1 - pz.z);
  }
  // Fused restriction (accumulate<float>, dcgrid_structure.cu:188-222) of blocks without children: a subblock =
  // the quads of lanes t (cx = 0: cells 0..3) and t+1 (cx = 1: cells 4..7); sequential sum, then * .125.
  const unsigned half = 0xFFFFu << (threadIdx.x & 16u);
  const bool childless = (__ballot_sync(0xFFFFFFFFu, child != kNone) & half) == 0;
  float s = 0.f;
  s += d.x; s += d.y; s += d.z; s += d.w;
  const float lo = __shfl_sync(0xFFFFFFFFu, s, (threadIdx.x & 31u) ^ 1u);
  if (!active) return;
  if (pl.w >= zero_from) {
    __stcs(reinterpret_cast<float4 *>(p + c0), zero);
    __stcs(reinterpret_cast<float4 *>(tp + c0), zero);
  }
  // quads of refined subblocks are always overwritten by their children's restriction: nothing to store
  if (leaf) __stcs(reinterpret_cast<float4 *>(div + c0), d);
  if (childless && (t & 1)) {
    const uint32_t ps = T.parent[b];
    if (ps != kNone) {
      float a = lo;
      a += d.x; a += d.y; a += d.z; a += d.w;
      div[(size_t)kSV * ps + (t >> 1)] = a * .125f;
    }
  }
}

// ---- k_dcgrid_apply_pressure, dcgrid_fluid.cu:232-259 --------------------------------------------------
__global__ void __launch_bounds__(kCTA4) k_dc_apply_pressure4(Pool T, KParams P, const float *__restrict__ p, const float *__restrict__ fl,
                                                              float4 *__restrict__ vw) {
  pdl_enter();
  const uint32_t g = threadIdx.x >> 4;
  const int t = threadIdx.x & 15;
  const uint32_t b0 = blockIdx.x * kB4 + g;
  const uint32_t b = b0 < T.M ? b0 : T.M - 1;  // out-of-range half: runs along, results discarded
  const int level = T.posl[b].w;               // 0xFF = free slot (its zeroed fields are loaded but unused)
  const bool active = b0 < T.M && level != kFree;
  if (!__any_sync(0xFFFFFFFFu, active)) return;
  const size_t c0 = (size_t)b * kBV + 4 * t;
  const float4 op = *reinterpret_cast<const float4 *>(p + c0);
  const uint32_t child = T.child[(size_t)b * 8 + (t >> 1)];
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = vw[c0 + k];
  const float4 ow = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);  // the cells' own fluidity travels with the velocity (== fl[c0..c0+3])
  const QuadGhosts q = quad_ghosts(T, b, t);
  const QuadNbr s = quad_neighbours(p, op, t, q);
  const QuadNbr w = quad_neighbours(fl, ow, t, q);
  const unsigned half = 0xFFFFu << (threadIdx.x & 16u);
  const unsigned with_child = __ballot_sync(0xFFFFFFFFu, child != kNone);  // every lane votes (no short-circuit)
  const bool childless = active && (with_child & half) == 0;
  if (!__any_sync(0xFFFFFFFFu, active && child == kNone)) return;
  const bool leaf = active && child == kNone;
  const float alpha = .5f * P.rdx / (float)(1 << (level & 15));
  // v.x -= alpha * (w_r * (p_r - pc) + w_l * (pc - p_l)), likewise y (up/down), z (front/back)
  v[0].x -= alpha * (w.xp.x * (s.xp.x - op.x) + w.xm.x * (op.x - s.xm.x));
  v[0].y -= alpha * (ow.z * (op.z - op.x) + w.ym0 * (op.x - s.ym0));
  v[0].z -= alpha * (ow.y * (op.y - op.x) + w.zm0 * (op.x - s.zm0));
  v[1].x -= alpha * (w.xp.y * (s.xp.y - op.y) + w.xm.y * (op.y - s.xm.y));
  v[1].y -= alpha * (ow.w * (op.w - op.y) + w.ym1 * (op.y - s.ym1));
  v[1].z -= alpha * (w.zp0 * (s.zp0 - op.y) + ow.x * (op.y - op.x));
  v[2].x -= alpha * (w.xp.z * (s.xp.z - op.z) + w.xm.z * (op.z - s.xm.z));
  v[2].y -= alpha * (w.yp0 * (s.yp0 - op.z) + ow.x * (op.z - op.x));
  v[2].z -= alpha * (ow.w * (op.w - op.z) + w.zm1 * (op.z - s.zm1));
  v[3].x -= alpha * (w.xp.w * (s.xp.w - op.w) + w.xm.w * (op.w - s.xm.w));
  v[3].y -= alpha * (w.yp1 * (s.yp1 - op.w) + ow.y * (op.w - op.y));
  v[3].z -= alpha * (w.zp1 * (s.zp1 - op.w) + ow.z * (op.w - op.z));
  if (leaf) {
#pragma unroll
    for (int k = 0; k < 4; k++) vw[c0 + k] = v[k];
  }
  // Fused restriction (accumulate<float3>, dcgrid_structure.cu:188-222) of blocks without children: lane t
  // (cx = 0) sums cells 0..3 of the subblock, lane t+1 (cx = 1) continues with cells 4..7 and stores.
  float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++) { ax += v[k].x; ay += v[k].y; az += v[k].z; }
  const unsigned src = (threadIdx.x & 31u) ^ 1u;
  const float lx = __shfl_sync(0xFFFFFFFFu, ax, src), ly = __shfl_sync(0xFFFFFFFFu, ay, src), lz = __shfl_sync(0xFFFFFFFFu, az, src);
  if (childless && (t & 1)) {
    const uint32_t ps = T.parent[b];
    if (ps != kNone) {
      ax = lx; ay = ly; az = lz;
#pragma unroll
      for (int k = 0; k < 4; k++) { ax += v[k].x; ay += v[k].y; az += v[k].z; }
      float *dst = reinterpret_cast<float *>(vw + ((size_t)kSV * ps + (t >> 1)));
      dst[0] = ax * .125f; dst[1] = ay * .125f; dst[2] = az * .125f;
    }
  }
}

// ---- coarse levels: the whole cascade of the small levels in ONE single-CTA launch ---------------------
// Levels with at most kCoarseBlocks blocks (<= 4 cells per thread and sweep) are launch-latency bound:
// at 512^3 levels 5..7 hold 64/8/1 blocks and take 33 launches of ~4 us each.  One CTA walks them
// coarse -> fine exactly like FluidSimulationDCGrid::project (fluid_simulation_dcgrid.cu:270-294):
// [prolongate] then `pairs` x (jacobi, jacobi_inv), with block-wide barriers between the phases.
// Plain (coherent) loads only: the data changes inside the kernel.
constexpr uint32_t kCoarseBlocks = 64;

__device__ __forceinline__ uint32_t face_neighbour_cell(const Pool &T, uint32_t b, int f, int a, int bb) {
  const uint32_t *fd = T.fd + 12 * (size_t)b;
  if (fd[6] & kFdIrregular) return T.face[(size_t)b * 96 + 16 * f + 4 * a + bb];
  return fd_ghost(fd[f], fd[6 + f], f >> 1, a, bb);
}
// value of the neighbour of cell (X,Y,Z) of block b across direction f (-x,+x,-y,+y,-z,+z); src[0] = cell `base`
__device__ __forceinline__ float neighbour_value(const Pool &T, const float *src, uint32_t base, uint32_t b, int X, int Y, int Z, int f) {
  int c[3] = {X, Y, Z};
  const int axis = f >> 1;
  c[axis] += (f & 1) ? 1 : -1;
  if ((unsigned)c[axis] < (unsigned)kBW) return src[b * kBV + cell_bits(c[0], c[1], c[2]) - base];
  const int a = axis == 0 ? Y : X, bb = axis == 2 ? Y : Z;
  return src[face_neighbour_cell(T, b, f, a, bb) - base];
}
__device__ __forceinline__ void coarse_sweep(const Pool &T, const KParams &P, int level, const float *in, float *out, const float *div,
                                             uint32_t base) {
  const float alpha = (float)((1 << level) * (1 << level)) * P.dx * P.dx;
  const uint32_t n = T.loads[level] * kBV;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t b = T.offsets[level] + (i >> 6), c = i & 63u;
    const int X = cell_x(c), Y = cell_y(c), Z = cell_z(c);
    const float l = neighbour_value(T, in, base, b, X, Y, Z, 0), r = neighbour_value(T, in, base, b, X, Y, Z, 1);
    const float dn = neighbour_value(T, in, base, b, X, Y, Z, 2), up = neighbour_value(T, in, base, b, X, Y, Z, 3);
    const float bk = neighbour_value(T, in, base, b, X, Y, Z, 4), fr = neighbour_value(T, in, base, b, X, Y, Z, 5);
    out[b * kBV + c - base] = (l + r + dn + up + bk + fr - alpha * div[b * kBV + c - base]) / 6.f;
  }
}
// kShared: pressure, t_pressure and divergence of the levels >= finest (the slot range [offsets[finest], end),
// `ncell` cells; their ghosts never leave it: a ghost lies in the same or a coarser level) live in shared memory
// for the whole cascade, together with a table of the six neighbour cells of every cell (16-bit offsets into the
// range, resolved once per launch through the face descriptors).  The single CTA is bound by instruction issue,
// not by latency: resolving neighbours on every sweep cost ~200 instructions per cell, the table makes a sweep
// 6 index + 6 value shared-memory loads.  The host falls back to the global-memory instantiation when the range
// does not fit (ncell > 65535 or 24 B * ncell > kCoarseSmemMax).
template <bool kShared>
__global__ void __launch_bounds__(1024) k_dc_coarse_cascade(Pool T, KParams P, int finest, int prolong_coarsest, int pairs_coarsest,
                                                            int pairs_level, int prolong_levels, float *gp, float *gtp, const float *gdiv,
                                                            uint32_t ncell) {
  pdl_enter();
  extern __shared__ __align__(16) float coarse_smem[];
  const uint32_t base = kShared ? T.offsets[finest] * kBV : 0u;
  float *p = gp, *tp = gtp;
  const float *div = gdiv;
  uint16_t *nb = nullptr;
  if (kShared) {
    p = coarse_smem; tp = coarse_smem + ncell;
    float *sdiv = coarse_smem + 2 * (size_t)ncell;
    nb = reinterpret_cast<uint16_t *>(coarse_smem + 3 * (size_t)ncell);
    for (uint32_t i = threadIdx.x; i < ncell; i += blockDim.x) {
      p[i] = gp[base + i]; tp[i] = gtp[base + i]; sdiv[i] = gdiv[base + i];
    }
    div = sdiv;
    for (int level = T.levels - 1; level >= finest; level--) {
      const uint32_t n = T.loads[level] * kBV;
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t b = T.offsets[level] + (i >> 6), c = i & 63u;
        const int X = cell_x(c), Y = cell_y(c), Z = cell_z(c);
        const uint32_t o = b * kBV + c - base;
#pragma unroll
        for (int f = 0; f < 6; f++) {
          int cc[3] = {X, Y, Z};
          const int axis = f >> 1;
          cc[axis] += (f & 1) ? 1 : -1;
          uint32_t id;
          if ((unsigned)cc[axis] < (unsigned)kBW) id = b * kBV + cell_bits(cc[0], cc[1], cc[2]);
          else id = face_neighbour_cell(T, b, f, axis == 0 ? Y : X, axis == 2 ? Y : Z);
          nb[6 * o + f] = (uint16_t)(id - base);
        }
      }
    }
    __syncthreads();
  }
  for (int level = T.levels - 1; level >= finest; level--) {
    const bool top = level == T.levels - 1;
    if (top ? prolong_coarsest : prolong_levels) {  // k_dcgrid_prolongate, dcgrid_multigrid_solver.cu:43-76
      const uint32_t n = T.loads[level] * kBV;
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t b = T.offsets[level] + (i >> 6), c = i & 63u;
        const uint32_t ps = T.parent[b];
        if (ps == kNone) continue;
        const uint32_t *pa = T.apron + (size_t)(ps / 8) * kAV;
        const int X = cell_x(c), Y = cell_y(c), Z = cell_z(c);
        const int idx = kAA * (1 + (int)((ps >> 2) & 1u) * 2 + (X >> 1)) + kAW * (1 + (int)((ps >> 1) & 1u) * 2 + (Y >> 1)) +
                        (1 + (int)(ps & 1u) * 2 + (Z >> 1));
        const int ii = (X & 1) ? kAA : -kAA, jj = (Y & 1) ? kAW : -kAW, kk = (Z & 1) ? 1 : -1;
        const float p000 = p[pa[idx] - base], p001 = p[pa[idx + kk] - base], p010 = p[pa[idx + jj] - base], p100 = p[pa[idx + ii] - base];
        const float p011 = p[pa[idx + jj + kk] - base], p101 = p[pa[idx + ii + kk] - base], p110 = p[pa[idx + ii + jj] - base],
                    p111 = p[pa[idx + ii + jj + kk] - base];
        p[b * kBV + c - base] = (27.f * p000 + 9.f * (p001 + p010 + p100) + 3.f * (p011 + p101 + p110) + p111) / 64.f;
      }
      __syncthreads();
    }
    const int pairs = top ? pairs_coarsest : pairs_level;
    for (int s = 0; s < 2 * pairs; s++) {
      const float *in = (s & 1) ? tp : p;
      float *out = (s & 1) ? p : tp;
      if (kShared) {
        const float alpha = (float)((1 << level) * (1 << level)) * P.dx * P.dx;
        const uint32_t o0 = T.offsets[level] * kBV - base, n = T.loads[level] * kBV;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
          const uint32_t o = o0 + i;
          const uint16_t *q = nb + 6 * o;
          out[o] = div6(in[q[0]] + in[q[1]] + in[q[2]] + in[q[3]] + in[q[4]] + in[q[5]] - alpha * div[o]);
        }
      } else {
        coarse_sweep(T, P, level, in, out, div, base);
      }
      __syncthreads();
    }
  }
  if (kShared) {
    for (uint32_t i = threadIdx.x; i < ncell; i += blockDim.x) {
      gp[base + i] = p[i]; gtp[base + i] = tp[i];
    }
  }
}

// accumulate<T> (dcgrid_structure.cu:188-222) for the small levels first..levels-2, fine -> coarse, in ONE launch:
// a thread-block cluster of kAccClusterCTAs CTAs (one subblock per thread and level at 512^3), levels separated by
// the hardware cluster barrier (release/acquire at cluster scope orders the global stores of one level before the
// loads of the next).
constexpr int kAccClusterCTAs = 8, kAccClusterThreads = 512;
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(kAccClusterCTAs, 1, 1) __launch_bounds__(kAccClusterThreads)
    k_dc_accumulate_coarse(Pool T, int first, float4 *vw, float *ch) {
  pdl_enter();
  const uint32_t rank = blockIdx.x * kAccClusterThreads + threadIdx.x, stride = kAccClusterCTAs * kAccClusterThreads;
  for (int level = first; level < T.levels - 1; level++) {
    const uint32_t n = 8 * T.loads[level];
    for (uint32_t i = rank; i < n; i += stride) {
      const uint32_t sb = 8 * T.offsets[level] + i, b = sb / 8;
      const uint32_t ps = T.parent[b];
      if (ps == kNone) continue;
      if (vw) {
        const float4 *c = vw + (size_t)kSV * sb;
        float4 v[kSV];
#pragma unroll
        for (int k = 0; k < kSV; k++) v[k] = c[k];
        float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
        for (int k = 0; k < kSV; k++) { ax += v[k].x; ay += v[k].y; az += v[k].z; }
        float *dst = reinterpret_cast<float *>(vw + ((size_t)kSV * ps + (sb % 8)));
        dst[0] = ax * .125f; dst[1] = ay * .125f; dst[2] = az * .125f;
      }
      if (ch) {
        const float4 lo = *reinterpret_cast<const float4 *>(ch + (size_t)kSV * sb), hi = *reinterpret_cast<const float4 *>(ch + (size_t)kSV * sb + 4);
        float a = 0.f;
        a += lo.x; a += lo.y; a += lo.z; a += lo.w; a += hi.x; a += hi.y; a += hi.z; a += hi.w;
        ch[(size_t)kSV * ps + (sb % 8)] = a * .125f;
      }
    }
    cluster_sync_all();
  }
}

}  // namespace dcg
