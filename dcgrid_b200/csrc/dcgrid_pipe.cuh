// Persistent, TMA-fed variants of the DCGrid stencil kernels (sm_100a).
//
// Why: the one-CTA-per-16-blocks kernels of dcgrid_stencil.cuh are latency-bound (profiles/README.md r1a:
// 48 % of DRAM peak with two dependent load rounds per CTA and a CTA lifetime of ~3 us).  The active blocks
// of a level are a compact slot prefix and a slot's 64 cells are contiguous, so the fields of 16
// consecutive blocks are ONE contiguous 4 KiB range per field and their face descriptors one 768-byte range.
// Here each CTA stays resident, walks tiles of 16 blocks, and an elected thread streams the next tiles'
// ranges into a multi-stage shared-memory ring with cp.async.bulk (UBLKCP) completing on mbarriers, so the
// bytes in flight per SM are set by the ring depth instead of by occupancy x one load round.  Ghost values
// (other blocks' cells, mostly L2 hits) are still gathered with ordinary loads through the face
// descriptors.  Arithmetic, operand order and the shuffle exchange are those of dcgrid_stencil.cuh.
#pragma once
#include "dcgrid_stencil.cuh"

namespace dcg {
namespace pipe {

__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Orders this thread's earlier generic-proxy accesses to shared memory (LD.SHARED of a ring slot) before later
// async-proxy accesses (the cp.async.bulk that refills the slot).  A CTA barrier or an mbarrier arrival alone does
// NOT order the two proxies: without this fence a refill occasionally overtook reads of the slot still in flight —
// never on the small test scenes, on ~0.1 % of the tiles of the 512^3 scene (profiles/README.md r2b: the ring Jacobi
// kernels differed from the reference at the first step; found by tests/test_golden_big_gpu.py).
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(saddr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(dst)), "l"(src),
               "r"(bytes), "r"(saddr(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(saddr(dst)),
               "l"(src), "r"(bytes), "r"(saddr(bar)), "l"(pol)
               : "memory");
}

}  // namespace pipe

// ---- Jacobi sweep (k_dcgrid_jacobi / _inv, dcgrid_multigrid_solver.cu:5-41), persistent + TMA ring --------
// CTA = 256 threads = one tile of 16 consecutive blocks per iteration (16 threads per block, one 1x2x2 quad
// each, half-warp shuffle exchange exactly as in k_dc_jacobi4).  in[], div[] and the face descriptors of a
// tile are three contiguous ranges streamed through a kJStages-deep ring by cp.async.bulk; the ghost
// values of tile i+1 are gathered into registers (vector loads, quad_ghost_values) while tile i is
// computed, so no thread ever waits for DRAM or L2 inside an iteration.
// (Two variants that kept the neighbour exchange and the ghosts in shared memory — LDS at loop-invariant
// offsets, ghosts gathered by 4-byte cp.async — were slower: they saturate the shared-memory data pipe,
// l1tex__data_pipe_lsu_wavefronts 71 %, profiles/README.md.)
constexpr int kJStages = 4;
struct alignas(128) JacobiStage {
  float p[kB4 * kBV];     // 4 KiB: in[] of the tile's 16 blocks
  float dv[kB4 * kBV];    // 4 KiB: div[]
  uint32_t fd[kB4 * 12];  // 768 B: face descriptors
};
constexpr size_t kJacobiPipeSmem = kJStages * sizeof(JacobiStage) + kJStages * sizeof(uint64_t);

// tile order: `reverse` walks the level's tiles from the last to the first, so that a sweep starts on the
// tiles the previous sweep wrote last (still L2-resident: a level-0 field is 62 MB at 512^3, L2 is 126 MB)
// R: the level-local tiles (tile k = slots offsets[level] + 16k ...) this rank sweeps, clipped to the active prefix
__global__ void __launch_bounds__(kCTA4, 4) k_dc_jacobi_pipe(Pool T, KParams P, TileRuns R, int level, const float *__restrict__ in,
                                                            float *__restrict__ out, const float *__restrict__ div, int reverse) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  JacobiStage *st = reinterpret_cast<JacobiStage *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kJStages * sizeof(JacobiStage));
  const uint32_t loads = T.loads[level], off = T.offsets[level];
  const uint32_t total = run_total(R), ntiles = 0xFFFFFFFFu;  // ntiles = "no tile"
  const uint32_t g = threadIdx.x >> 4;
  const int t = threadIdx.x & 15;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kJStages; s++) pipe::mbar_init(&full[s], 1);
    pipe::fence_barrier_init();
  }
  __syncthreads();
  const uint64_t pol_stream = pipe::policy_evict_first();
  auto tile_of = [&](uint32_t it) -> uint32_t {  // this CTA's it-th tile, ntiles = none
    const uint32_t j = blockIdx.x + it * gridDim.x;
    if (j >= total) return ntiles;
    return run_tile(R, reverse ? total - 1 - j : j);
  };
  auto issue_tile = [&](uint32_t it) {  // producer (thread 0): tile `it` of this CTA -> ring slot it % kJStages
    const uint32_t tl = tile_of(it);
    if (tl >= ntiles) return;
    const uint32_t li0 = tl * kB4;
    const uint32_t nvalid = min((uint32_t)kB4, loads - li0);
    const size_t b0 = (size_t)off + li0;
    JacobiStage &S = st[it % kJStages];
    uint64_t *bar = &full[it % kJStages];
    pipe::mbar_expect_tx(bar, nvalid * (2u * kBV * 4u + 48u));
    pipe::bulk_g2s(S.p, in + b0 * kBV, nvalid * kBV * 4u, bar);
    pipe::bulk_g2s_hint(S.dv, div + b0 * kBV, nvalid * kBV * 4u, bar, pol_stream);
    pipe::bulk_g2s(S.fd, T.fd + b0 * 12, nvalid * 48u, bar);
  };
  // ghost values of tile `it` (waits for its ring slot; an inactive trailing block holds stale descriptors)
  auto load_ghosts = [&](uint32_t it) -> GhostVals {
    GhostVals gv;
    gv.gx = make_float4(0.f, 0.f, 0.f, 0.f);
    gv.gy0 = gv.gy1 = gv.gz0 = gv.gz1 = 0.f;
    const uint32_t tl = tile_of(it);
    if (tl < ntiles) {
      pipe::mbar_wait(&full[it % kJStages], (it / kJStages) & 1u);
      const uint32_t li = tl * kB4 + g;
      if (li < loads) {
        const uint4 *fdp = reinterpret_cast<const uint4 *>(&st[it % kJStages].fd[g * 12]);
        gv = quad_ghost_values(T, in, off + li, t, fdp[0], fdp[1], fdp[2]);
      }
    }
    return gv;
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (uint32_t it = 0; it < (uint32_t)kJStages; it++) issue_tile(it);
  }
  const float alpha = (float)((1 << level) * (1 << level)) * P.dx * P.dx;
  GhostVals gv = load_ghosts(0);
  for (uint32_t it = 0;; it++) {
    const uint32_t tl = tile_of(it);
    if (tl >= ntiles) break;
    const uint32_t s = it % kJStages;
    const GhostVals gn = load_ghosts(it + 1);  // in flight while this tile is computed
    const uint32_t li = tl * kB4 + g;
    const bool active = li < loads;
    const JacobiStage &S = st[s];  // its barrier was waited for by load_ghosts(it)
    const float4 own = *reinterpret_cast<const float4 *>(&S.p[g * kBV + 4 * t]);
    const float4 dv = *reinterpret_cast<const float4 *>(&S.dv[g * kBV + 4 * t]);
    pipe::fence_proxy_async();  // the reads above, before the async-proxy refill issued after the barrier
    __syncthreads();            // every thread holds its part of ring slot s in registers: the slot can be refilled
    if (threadIdx.x == 0) issue_tile(it + kJStages);
    const QuadNbr n = quad_exchange(own, t, gv.gx, gv.gy0, gv.gy1, gv.gz0, gv.gz1);
    float4 o;
    o.x = div6(n.xm.x + n.xp.x + n.ym0 + own.z + n.zm0 + own.y - alpha * dv.x);
    o.y = div6(n.xm.y + n.xp.y + n.ym1 + own.w + own.x + n.zp0 - alpha * dv.y);
    o.z = div6(n.xm.z + n.xp.z + own.x + n.yp0 + n.zm1 + own.w - alpha * dv.z);
    o.w = div6(n.xm.w + n.xp.w + own.y + n.yp1 + own.z + n.zp1 - alpha * dv.w);
    if (active) *reinterpret_cast<float4 *>(out + ((size_t)off + li) * kBV + 4 * t) = o;
    gv = gn;
  }
}

// ---- Jacobi sweep, 8 cells per thread --------------------------------------------------------------------------
// k_dc_jacobi_pipe is bound by instruction issue as much as by DRAM (64 % of the issue slots, 287 warp
// instructions per tile iteration, profiles/README.md r1d): per 4 cells a thread pays 16 shuffles + 16 selects
// for the neighbour exchange and the whole per-thread overhead (ring bookkeeping, descriptor decode, ghost
// addresses).  Here a thread owns one 2x2x2 subblock (8 contiguous cells = two float4), 8 threads per block,
// 128 threads per 16-block tile: the x/y/z neighbours inside the subblock are the thread's own registers, only
// one 2x2 face per axis crosses lanes (12 shuffles per 8 cells), and the per-thread overhead is paid once per
// 8 cells.  Ring, barriers, ghost prefetch, arithmetic and its order are those of k_dc_jacobi_pipe.
// (Measured and rejected, profiles/README.md r2l: one range test for the eight div6 of a thread instead of eight, the
// tile loop unrolled over three named ghost sets instead of the rotating copies, one merged branch for three
// same-level faces, a single-run shortcut for the tile index — each removes instructions, ptxas then allocates 79-80
// registers instead of 94 and schedules the ghost prefetch later: 39 -> 44-47 us per level-0 sweep, and the same
// happens to THIS code under __launch_bounds__(128, 6): 79 registers, 47 us at 6 CTAs per SM, 55 us at 5.  The kernel
// is bound by how long its ghost loads stay in flight, not by instruction count; with the 79-register schedule it
// scales 62 / 52 / 45 us at 4 / 5 / 6 resident CTAs per SM.  Ghosts one tile ahead instead of two: 71 registers,
// 43 us at 6 CTAs per SM.)
constexpr int kJ8Threads = kB4 * 8;
struct Ghost12 {
  float gx[4];  // [cy*2+cz]: the subblock's x face (-x if sx = 0, +x if sx = 1)
  float gy[4];  // [cx*2+cz]
  float gz[4];  // [cx*2+cy]
};
__device__ __forceinline__ Ghost12 sub_ghost_values(const Pool &T, const float *__restrict__ src, uint32_t b, int t, const uint4 w0,
                                                    const uint4 w1, const uint4 w2) {
  Ghost12 g;
  const int sx = t >> 2, sy = (t >> 1) & 1, sz = t & 1;
  const int fx = sx, fy = 2 + sy, fz = 4 + sz;
  if (!(w1.z & kFdIrregular)) {
    const uint32_t nbx = fx ? w0.y : w0.x, cdx = fx ? w1.w : w1.z;
    const uint32_t nby = sy ? w0.w : w0.z, cdy = sy ? w2.y : w2.x;
    const uint32_t nbz = sz ? w1.y : w1.x, cdz = sz ? w2.w : w2.z;
    if (cdx == 0) {  // same-level neighbour: the facing 2x2 cells are 16 contiguous bytes
      const float4 v = *reinterpret_cast<const float4 *>(src + nbx + (uint32_t)((sy << 4) | (sz << 3)));
      g.gx[0] = v.x; g.gx[1] = v.y; g.gx[2] = v.z; g.gx[3] = v.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) g.gx[k] = src[fd_ghost(nbx, cdx, 0, 2 * sy + (k >> 1), 2 * sz + (k & 1))];
    }
    if (cdy == 0) {  // two 8-byte pairs (cx = 0, 1)
      const float *a = src + nby + (uint32_t)((sx << 5) | (sz << 3));
      const float2 v0 = *reinterpret_cast<const float2 *>(a), v1 = *reinterpret_cast<const float2 *>(a + 4);
      g.gy[0] = v0.x; g.gy[1] = v0.y; g.gy[2] = v1.x; g.gy[3] = v1.y;
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) g.gy[k] = src[fd_ghost(nby, cdy, 1, 2 * sx + (k >> 1), 2 * sz + (k & 1))];
    }
    if (cdz == 0) {
      // four 4-byte loads at stride 2 (one sector).  Not two float4 + a select on the base's low bit: the select
      // would consume the loads where they are issued and turn the two-tiles-ahead prefetch into a blocking wait
      // (it was the top long-scoreboard stall of the kernel)
      const float *a = src + nbz + (uint32_t)((sx << 5) | (sy << 4));
      g.gz[0] = a[0]; g.gz[1] = a[2]; g.gz[2] = a[4]; g.gz[3] = a[6];
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) g.gz[k] = src[fd_ghost(nbz, cdz, 2, 2 * sx + (k >> 1), 2 * sy + (k & 1))];
    }
    return g;
  }
  const uint32_t *ft = T.face + (size_t)b * 96;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    g.gx[k] = src[ft[16 * fx + 4 * (2 * sy + (k >> 1)) + (2 * sz + (k & 1))]];
    g.gy[k] = src[ft[16 * fy + 4 * (2 * sx + (k >> 1)) + (2 * sz + (k & 1))]];
    g.gz[k] = src[ft[16 * fz + 4 * (2 * sx + (k >> 1)) + (2 * sy + (k & 1))]];
  }
  return g;
}

__global__ void __launch_bounds__(kJ8Threads, 5) k_dc_jacobi_pipe8(Pool T, KParams P, TileRuns R, int level, const float *__restrict__ in,
                                                                  float *__restrict__ out, const float *__restrict__ div, int reverse) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  JacobiStage *st = reinterpret_cast<JacobiStage *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kJStages * sizeof(JacobiStage));
  const uint32_t loads = T.loads[level], off = T.offsets[level];
  const uint32_t total = run_total(R), none = 0xFFFFFFFFu;
  const uint32_t g = threadIdx.x >> 3;
  const int t = threadIdx.x & 7;
  const int sx = t >> 2, sy = (t >> 1) & 1, sz = t & 1;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kJStages; s++) pipe::mbar_init(&full[s], 1);
    pipe::fence_barrier_init();
  }
  __syncthreads();
  const uint64_t pol_stream = pipe::policy_evict_first();
  auto tile_of = [&](uint32_t it) -> uint32_t {
    const uint32_t j = blockIdx.x + it * gridDim.x;
    if (j >= total) return none;
    return run_tile(R, reverse ? total - 1 - j : j);
  };
  auto issue_tile = [&](uint32_t it) {
    const uint32_t tl = tile_of(it);
    if (tl == none) return;
    const uint32_t li0 = tl * kB4;
    const uint32_t nvalid = min((uint32_t)kB4, loads - li0);
    const size_t b0 = (size_t)off + li0;
    JacobiStage &S = st[it % kJStages];
    uint64_t *bar = &full[it % kJStages];
    pipe::mbar_expect_tx(bar, nvalid * (2u * kBV * 4u + 48u));
    pipe::bulk_g2s(S.p, in + b0 * kBV, nvalid * kBV * 4u, bar);
    pipe::bulk_g2s_hint(S.dv, div + b0 * kBV, nvalid * kBV * 4u, bar, pol_stream);
    pipe::bulk_g2s(S.fd, T.fd + b0 * 12, nvalid * 48u, bar);
  };
  auto load_ghosts = [&](uint32_t it) -> Ghost12 {
    Ghost12 gv;
#pragma unroll
    for (int k = 0; k < 4; k++) gv.gx[k] = gv.gy[k] = gv.gz[k] = 0.f;
    const uint32_t tl = tile_of(it);
    if (tl != none) {
      pipe::mbar_wait(&full[it % kJStages], (it / kJStages) & 1u);
      const uint32_t li = tl * kB4 + g;
      if (li < loads) {
        const uint4 *fdp = reinterpret_cast<const uint4 *>(&st[it % kJStages].fd[g * 12]);
        gv = sub_ghost_values(T, in, off + li, t, fdp[0], fdp[1], fdp[2]);
      }
    }
    return gv;
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (uint32_t it = 0; it < (uint32_t)kJStages; it++) issue_tile(it);
  }
  const float alpha = (float)((1 << level) * (1 << level)) * P.dx * P.dx;
  Ghost12 gv = load_ghosts(0), g1 = load_ghosts(1);
  for (uint32_t it = 0;; it++) {
    const uint32_t tl = tile_of(it);
    if (tl == none) break;
    const uint32_t s = it % kJStages;
    const Ghost12 gn = load_ghosts(it + 2);  // two tiles ahead: in flight while this tile and the next are computed
    const uint32_t li = tl * kB4 + g;
    const bool active = li < loads;
    const JacobiStage &S = st[s];
    // two float4 per thread at a 32-byte lane stride: the upper four lanes of a block start with the second one,
    // which makes both requests bank-conflict free
    float4 a0 = *reinterpret_cast<const float4 *>(&S.p[g * kBV + 8 * t + 4 * sx]);
    float4 a1 = *reinterpret_cast<const float4 *>(&S.p[g * kBV + 8 * t + 4 * (sx ^ 1)]);
    float4 d0 = *reinterpret_cast<const float4 *>(&S.dv[g * kBV + 8 * t + 4 * sx]);
    float4 d1 = *reinterpret_cast<const float4 *>(&S.dv[g * kBV + 8 * t + 4 * (sx ^ 1)]);
    pipe::fence_proxy_async();  // the reads above, before the async-proxy refill issued after the barrier
    __syncthreads();            // every thread holds its part of ring slot s in registers: the slot can be refilled
    if (threadIdx.x == 0) issue_tile(it + kJStages);
    const float4 lo = sx ? a1 : a0, hi = sx ? a0 : a1, dlo = sx ? d1 : d0, dhi = sx ? d0 : d1;
    const float own[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};  // k = cx*4 + cy*2 + cz
    const float dv[8] = {dlo.x, dlo.y, dlo.z, dlo.w, dhi.x, dhi.y, dhi.z, dhi.w};
    float rx[4], ry[4], rz[4];  // the facing 2x2 cells of the x / y / z sibling subblock
#pragma unroll
    for (int j = 0; j < 4; j++) {
      rx[j] = __shfl_xor_sync(0xFFFFFFFFu, sx ? own[j] : own[4 + j], 4);                                            // j = cy*2+cz
      ry[j] = __shfl_xor_sync(0xFFFFFFFFu, sy ? own[(j >> 1) * 4 + (j & 1)] : own[(j >> 1) * 4 + 2 + (j & 1)], 2);  // j = cx*2+cz
      rz[j] = __shfl_xor_sync(0xFFFFFFFFu, sz ? own[(j >> 1) * 4 + (j & 1) * 2] : own[(j >> 1) * 4 + (j & 1) * 2 + 1], 1);  // j = cx*2+cy
    }
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int cx = k >> 2, cy = (k >> 1) & 1, cz = k & 1;
      const int jx = cy * 2 + cz, jy = cx * 2 + cz, jz = cx * 2 + cy;
      // a cell's outward neighbour along an axis: the sibling subblock's facing cell, or the block's ghost
      const float xm = cx ? own[k - 4] : (sx ? rx[jx] : gv.gx[jx]);
      const float xp = cx ? (sx ? gv.gx[jx] : rx[jx]) : own[k + 4];
      const float ym = cy ? own[k - 2] : (sy ? ry[jy] : gv.gy[jy]);
      const float yp = cy ? (sy ? gv.gy[jy] : ry[jy]) : own[k + 2];
      const float zm = cz ? own[k - 1] : (sz ? rz[jz] : gv.gz[jz]);
      const float zp = cz ? (sz ? gv.gz[jz] : rz[jz]) : own[k + 1];
      o[k] = div6(xm + xp + ym + yp + zm + zp - alpha * dv[k]);
    }
    if (active) {
      float4 *dst = reinterpret_cast<float4 *>(out + ((size_t)off + li) * kBV + 8 * t);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    gv = g1;
    g1 = gn;
  }
}

// ---- semi-Lagrangian advection (k_dcgrid_advect_velocity / _density, dcgrid_fluid.cu:7-144), persistent ----
// The one-CTA-per-4-blocks kernels of dcgrid_kernels.cuh spend ~45 % of their stall samples staging the
// blocks' 6^3 apron maps (LDG -> STS -> barrier) before the first gather can be issued
// (profiles/README.md).  A tile's apron maps (4 x 864 B), own velocities (4 KiB), child links and
// positions are four contiguous ranges, so a resident CTA streams them through a ring of cp.async.bulk
// stages and the only latency a thread waits for is that of its own gathers.
//
// Samples that leave their block (69 % of the warps hold at least one in the developed flow) resolve the
// covering block with one load from the finest-block map (the reference walks the level maps one dependent
// load at a time, dcgrid_utils.cuh:201-233) and derive the block origin from the sample position (block
// origins are multiples of 4 cells of their level) instead of loading it.
// Ring depth 2: the consumers release a slot before their gathers, so two tiles in flight cover the copy latency, and
// every KiB of shared memory not used is L1 for the gathers (4 stages: 853 us, 2 stages: 840 us at dcgrid512).
// (Measured and rejected, profiles/README.md r2n: one private ring per warp, no producer warp, 4-warp CTAs at 80 / 95
// registers = 6 / 5 CTAs per SM: 901 / 971 us.  The kernel is bound by the L1 data pipe and its hit rate — twice as
// many blocks in flight per SM, each at half speed, only widen the gather footprint that has to stay in L1.)
constexpr int kAStages = 2;
constexpr int kAdvectThreads = kCTA + 32;  // 8 consumer warps (4 blocks) + 1 producer warp
struct alignas(128) AdvectStage {
  uint32_t apron[kBPC * kAV];  // 3456 B
  float4 me[kBPC * kBV];       // 4096 B: the cells' own velocity (+ fluidity)
  uint32_t child[kBPC * kSV];  // 128 B
  int4 posl[kBPC];             // 64 B
  uint32_t slot[kBPC];         // 16 B: the tile's pool slots (kNone = padding), written by the producer
};
constexpr size_t kAdvectPipeSmem = kAStages * sizeof(AdvectStage) + 2 * kAStages * sizeof(uint64_t);

// getBlockIndexDeep(position, 0) as ONE load.  The finest block covering a level-0 cell is the same for all cells of a
// level-0 block footprint (every block origin is a multiple of 4 cells of its level), so the walk over the level maps
// (dcgrid_utils.cuh:201-233: one dependent probe per sparse level) is tabulated per level-0 block coordinate whenever
// the topology changes: fmap[(bx * ry + by) * rz + bz] = level << 28 | slot (k_dc_build_fmap evaluates the walk
// itself, so the table cannot disagree with it).  8 MiB at 512^3.
constexpr uint32_t kFmapSlotMask = 0x0FFFFFFFu;
__global__ void __launch_bounds__(256) k_dc_build_fmap(Pool T, KParams P) {
  const int3 r = level_dims(P, 0);
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)r.x * r.y * r.z) return;
  const int bz = (int)(i % r.z), by = (int)((i / r.z) % r.y), bx = (int)(i / ((size_t)r.z * r.y));
  int level = 0;
  const uint32_t b = block_index_deep(T, P, bx * kBW, by * kBW, bz * kBW, level);
  T.fmap[i] = b == kNone ? kNone : ((uint32_t)level << 28 | b);
}
__device__ __forceinline__ uint32_t block_index_finest(const Pool &T, const KParams &P, int ix, int iy, int iz, int4 &bp) {
  const int ry = idiv_up(P.gy, kBW), rz = idiv_up(P.gz, kBW);
  const uint32_t e = __ldg(T.fmap + ((size_t)(ix >> 2) * ry + (iy >> 2)) * rz + (iz >> 2));
  const int level = (int)(e >> 28);
  bp = make_int4((ix >> level) & ~(kBW - 1), (iy >> level) & ~(kBW - 1), (iz >> level) & ~(kBW - 1), level);
  return e & kFmapSlotMask;
}

__device__ __forceinline__ DSample d_sample_pipe(const Pool &T, const KParams &P, const uint32_t *own_apron, const uint32_t *own_child,
                                                 const int4 pl, float px, float py, float pz) {
  const int ix = min(max((int)floorf(px), 0), P.gx - 1), iy = min(max((int)floorf(py), 0), P.gy - 1), iz = min(max((int)floorf(pz), 0), P.gz - 1);
  const int lx = (ix >> pl.w) - pl.x, ly = (iy >> pl.w) - pl.y, lz = (iz >> pl.w) - pl.z;
  if ((unsigned)lx < (unsigned)kBW && (unsigned)ly < (unsigned)kBW && (unsigned)lz < (unsigned)kBW &&
      own_child[((lx >> 1) << 2) | ((ly >> 1) << 1) | (lz >> 1)] == kNone)
    return d_sample_in(own_apron, pl, px, py, pz);
  int4 bp;
  const uint32_t b = block_index_finest(T, P, ix, iy, iz, bp);
  return d_sample_in(T.apron + (size_t)b * kAV, bp, px, py, pz);
}

// kMode 0: velocity (k_dcgrid_advect_velocity), 1: density (k_dcgrid_advect_density), 2: both at once.
//
// Mode 2.  advectDensity() of step n and advectVelocity() of step n+1 backtrace every cell through the SAME
// velocity field (nothing runs between them, simulation.cpp:104-111), so both resolve the same sample: the same
// covering block, the same 8 cell ids and the same fluidity-weighted weights.  The reference does that work
// twice; here the density pass also gathers the velocities at the 8 ids it already holds and writes the next
// step's advected velocity into the idle ping-pong buffer.  The host consumes it in the next advect_velocity()
// if nothing touched the state in between (dcgrid.cu, `spec_velocity`).
//
// Fused restriction (accumulate<T>, dcgrid_structure.cu:188-222) of blocks without children, as in
// k_dc_divergence4: the 8 cells of a subblock are 8 consecutive lanes; every lane sums them in cell order
// (sequentially, starting from 0.f like the reference) and lane 7 of the group stores into the parent cell.
// Cells of refined subblocks are not stored: their children's restriction owns those words (the fluidity in
// .w is stored separately, it is the one component restriction leaves alone).
__device__ __forceinline__ float subblock_sum(float v) {
  const unsigned base = threadIdx.x & 24u;
  float a = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) a += __shfl_sync(0xFFFFFFFFu, v, base + k);
  return a;
}

// Processing order.  Tiles are kBPC consecutive entries of `order` (active slots, padded with kNone), not
// consecutive slots: the host sorts the active blocks along a Morton curve of their positions (dcgrid.cu,
// rebuild_order) so that the blocks in flight at any time — and the upstream cells their samples gather — form
// a compact region that stays in L2; in slot order 2/3 of the gathered sectors came from DRAM
// (profiles/README.md r1d).  Every block's inputs are four contiguous ranges, fetched by one cp.async.bulk each.
//
// Warp specialisation.  Warp 8 is the producer (one lane: waits for a ring slot to drain, issues the copies);
// the 8 consumer warps never meet at a CTA barrier: each waits for the `full` mbarrier of its tile, resolves
// its samples, releases the slot (`empty` mbarrier, one arrival per warp) and goes on to its gathers, so the
// warps of a CTA drift apart by up to kAStages tiles and their load rounds overlap.
// kExt (extension, dcg_ext_params.sources): potential temperature and vapor ride along with the density — same sample,
// same weights, their own boundary values (ext::scalar_bc), the ambient value where a sample has no fluid weight
// (oracle advect_scalar; k_dc_ext_advect_scalar is the stand-alone kernel) — instead of two more gather passes.
struct AdvectScalars {
  const float *th_in, *qv_in;
  float *th_out, *qv_out;
  dcg_ext_params E;
};
template <int kMode, int kMinBlocks, bool kExt>
__global__ void __launch_bounds__(kAdvectThreads, kMinBlocks) k_dc_advect_pipe(Pool T, KParams P, const uint32_t *__restrict__ order, uint32_t norder,
                                                                               const float4 *__restrict__ vin, float4 *__restrict__ vout,
                                                                               const float *__restrict__ fl, const float *__restrict__ qin,
                                                                               float *__restrict__ qout, AdvectScalars X) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AdvectStage *st = reinterpret_cast<AdvectStage *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kAStages * sizeof(AdvectStage));
  uint64_t *empty = full + kAStages;
  const uint32_t ntiles = norder / kBPC;  // norder is a multiple of kBPC
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kAStages; s++) {
      pipe::mbar_init(&full[s], 1);
      pipe::mbar_init(&empty[s], kCTA / 32);
    }
    pipe::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x >= kCTA) {  // ---- producer warp ----
    if (threadIdx.x != kCTA) return;
    const uint64_t pol_stream = pipe::policy_evict_first();
    const uint4 *order4 = reinterpret_cast<const uint4 *>(order);
    uint32_t tl = blockIdx.x;
    uint4 nxt = tl < ntiles ? __ldg(order4 + tl) : make_uint4(kNone, kNone, kNone, kNone);
    for (uint32_t it = 0; tl < ntiles; it++, tl += gridDim.x) {
      const uint4 cur = nxt;
      if (tl + gridDim.x < ntiles) nxt = __ldg(order4 + tl + gridDim.x);  // in flight while this tile is issued
      const uint32_t s = it % kAStages;
      if (it >= (uint32_t)kAStages) pipe::mbar_wait(&empty[s], ((it / kAStages) - 1u) & 1u);
      AdvectStage &S = st[s];
      *reinterpret_cast<uint4 *>(S.slot) = cur;
      const uint32_t b4[kBPC] = {cur.x, cur.y, cur.z, cur.w};
      uint32_t nvalid = 0;
#pragma unroll
      for (int g = 0; g < kBPC; g++) nvalid += b4[g] != kNone ? 1u : 0u;
      pipe::mbar_expect_tx(&full[s], nvalid * (kAV * 4u + kBV * 16u + kSV * 4u + 16u));
#pragma unroll
      for (int g = 0; g < kBPC; g++) {
        const size_t b = b4[g];
        if (b4[g] == kNone) continue;
        pipe::bulk_g2s_hint(S.apron + g * kAV, T.apron + b * kAV, kAV * 4u, &full[s], pol_stream);
        pipe::bulk_g2s(S.me + g * kBV, vin + b * kBV, kBV * 16u, &full[s]);
        pipe::bulk_g2s(S.child + g * kSV, T.child + b * kSV, kSV * 4u, &full[s]);
        pipe::bulk_g2s(S.posl + g, T.posl + b, 16u, &full[s]);
      }
    }
    return;
  }
  // ---- consumer warps ----
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const float alpha = P.dt * P.rdx;
  uint32_t tl = blockIdx.x;
  for (uint32_t it = 0; tl < ntiles; it++, tl += gridDim.x) {
    const uint32_t s = it % kAStages;
    pipe::mbar_wait(&full[s], (it / kAStages) & 1u);
    const AdvectStage &S = st[s];
    const uint32_t b = S.slot[g];
    const bool live = b != kNone;  // uniform over the block's two warps
    int4 pl = make_int4(0, 0, 0, 0);
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t ps = kNone;
    bool childless = false, leaf = false;
    DSample smp;
    if (live) {
      ps = __ldg(T.parent + b);
      pl = S.posl[g];
      me = S.me[g * kBV + t];
      const uint4 c0 = *reinterpret_cast<const uint4 *>(&S.child[g * kSV]), c1 = *reinterpret_cast<const uint4 *>(&S.child[g * kSV + 4]);
      childless = (c0.x & c0.y & c0.z & c0.w & c1.x & c1.y & c1.z & c1.w) == kNone;
      leaf = S.child[g * kSV + (t >> 3)] == kNone;
      if (leaf) {
        const float scale = (float)(1 << pl.w);
        const float bx = ((float)(pl.x | cell_x(t)) + .5f) * scale - me.x * alpha;
        const float by = ((float)(pl.y | cell_y(t)) + .5f) * scale - me.y * alpha;
        const float bz = ((float)(pl.z | cell_z(t)) + .5f) * scale - me.z * alpha;
        smp = d_sample_pipe(T, P, S.apron + g * kAV, S.child + g * kSV, pl, bx, by, bz);
      }
    }
    pipe::fence_proxy_async();
    __syncwarp();  // every lane is done with ring slot s (positions, velocities, apron ids)
    if ((threadIdx.x & 31u) == 0) pipe::mbar_arrive(&empty[s]);
    if (!live) continue;
    const uint32_t c = b * kBV + t;
    float3 vo = make_float3(0.f, 0.f, 0.f);
    float qo = 0.f, to = 0.f, wo = 0.f;
    if (leaf) {
      float4 cv[8];
      float qv[8], f[8], tv[8], wv[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (kMode != 1) cv[k] = vin[smp.id[k]];
        if (kMode != 0) qv[k] = qin[smp.id[k]];
        if (kExt) {
          tv[k] = X.th_in[smp.id[k]];
          wv[k] = X.qv_in[smp.id[k]];
        }
        f[k] = kMode == 1 ? fl[smp.id[k]] : cv[k].w;
      }
      if (kExt) {  // a sample without fluid weight takes the ambient values
        to = ext::ambient_theta(X.E, ext::cell_height(P, pl.y | cell_y(t), 1 << pl.w));
        wo = X.E.ambient_vapor;
      }
      const Weights8 W = corner_weights(f, smp.fx, smp.fy, smp.fz);
      if (!(W.acc < 1e-6f)) {
        const bool inside = sample_inside(P, smp);
        if (kExt) {
          if (!inside) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const int cx = smp.x0 + ((k >> 2) & 1), cy = smp.y0 + ((k >> 1) & 1), cz = smp.z0 + (k & 1);
              tv[k] = ext::scalar_bc<1>(P, X.E, tv[k], cx, cy, cz, smp.scale);
              wv[k] = ext::scalar_bc<2>(P, X.E, wv[k], cx, cy, cz, smp.scale);
            }
          }
          to = blend8(tv, W.w);
          wo = blend8(wv, W.w);
        }
        if (kMode != 1) {
          float vx[8], vy[8], vz[8];
#pragma unroll
          for (int k = 0; k < 8; k++) {
            float3 v = make_float3(cv[k].x, cv[k].y, cv[k].z);
            if (!inside) v = velocity_bc(P, v, smp.x0 + ((k >> 2) & 1), smp.y0 + ((k >> 1) & 1), smp.z0 + (k & 1), smp.scale);
            vx[k] = v.x; vy[k] = v.y; vz[k] = v.z;
          }
          vo = make_float3(blend8(vx, W.w), blend8(vy, W.w), blend8(vz, W.w));
        }
        if (kMode != 0) {
          if (!inside) {
#pragma unroll
            for (int k = 0; k < 8; k++)
              qv[k] = density_bc(P, qv[k], smp.x0 + ((k >> 2) & 1), smp.y0 + ((k >> 1) & 1), smp.z0 + (k & 1), smp.scale);
          }
          qo = blend8(qv, W.w);
        }
      }
    }
    const bool push = childless && ps != kNone;  // uniform over the block
    if (kMode != 1) {
      if (leaf) vout[c] = make_float4(vo.x, vo.y, vo.z, me.w);
      else reinterpret_cast<float *>(vout + c)[3] = me.w;
      if (push) {
        const float ax = subblock_sum(vo.x), ay = subblock_sum(vo.y), az = subblock_sum(vo.z);
        if ((t & 7u) == 7u) {
          float *dst = reinterpret_cast<float *>(vout + ((size_t)kSV * ps + (t >> 3)));
          dst[0] = ax * .125f; dst[1] = ay * .125f; dst[2] = az * .125f;
        }
      }
    }
    if (kMode != 0) {
      if (leaf) qout[c] = qo;
      if (push) {
        const float a = subblock_sum(qo);
        if ((t & 7u) == 7u) qout[(size_t)kSV * ps + (t >> 3)] = a * .125f;
      }
    }
    if (kExt) {
      if (leaf) {
        X.th_out[c] = to;
        X.qv_out[c] = wo;
      }
      if (push) {
        const float a = subblock_sum(to), b2 = subblock_sum(wo);
        if ((t & 7u) == 7u) {
          X.th_out[(size_t)kSV * ps + (t >> 3)] = a * .125f;
          X.qv_out[(size_t)kSV * ps + (t >> 3)] = b2 * .125f;
        }
      }
    }
  }
}

// Morton key of a block for the processing order: its origin in level-0 block units (4 cells), bits interleaved
__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
// mode 0: key = slot (pool order); 1: Morton.  Free slots get key 0xFFFFFFFF and sort to the end.
// owner != nullptr: only the blocks of units owned by `rank` (owner[b / unit]) are listed
__global__ void __launch_bounds__(256) k_dc_order_keys(Pool T, int mode, const uint8_t *__restrict__ owner, uint32_t unit, int rank,
                                                       uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ count) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  const bool in_pool = b < T.M;
  const int4 pl = in_pool ? T.posl[b] : make_int4(0, 0, 0, kFree);
  const bool listed = in_pool && pl.w != kFree && (!owner || owner[b / unit] == rank);
  const unsigned votes = __ballot_sync(0xFFFFFFFFu, listed);  // one atomic per warp, not per block
  if ((threadIdx.x & 31u) == 0 && votes) atomicAdd(count, (uint32_t)__popc(votes));
  if (!in_pool) return;
  uint32_t key = 0xFFFFFFFFu;
  if (listed) {
    if (mode == 0) key = b;
    else key = (spread3((uint32_t)(pl.x << pl.w) >> 2) << 2) | (spread3((uint32_t)(pl.y << pl.w) >> 2) << 1) | spread3((uint32_t)(pl.z << pl.w) >> 2);
  }
  keys[b] = key;
  vals[b] = b;
}
__global__ void __launch_bounds__(256) k_dc_order_pad(uint32_t *__restrict__ order, uint32_t n, uint32_t padded) {
  const uint32_t i = n + blockIdx.x * 256 + threadIdx.x;
  if (i < padded) order[i] = kNone;
}

// ---- divergence and pressure gradient (k_dcgrid_calc_divergence / k_dcgrid_apply_pressure), persistent + TMA ring ----
// The one-CTA-per-tile kernels of dcgrid_stencil.cuh read a quad's four packed velocities with four LDG.128 at a
// 64-byte lane stride: every request touches 32 half-used sectors and costs 4x the L1 wavefronts of a coalesced
// one (l1tex data-pipe 78 % busy at 3.5 TB/s, profiles/README.md r1d).  Here a tile's velocities (16 KiB), face
// descriptors, child links, positions and parent links are contiguous ranges streamed by cp.async.bulk into a
// 3-stage ring by a producer warp; the consumers read their quads from shared memory in a rotated chunk order
// (lane t starts at chunk (t >> 1) & 3) that is bank-conflict free, and un-rotate in registers.  Ghosts are
// still gathered from global memory through the face descriptors; arithmetic and its order are unchanged.
constexpr int kSStages = 3;
constexpr int kStencilThreads = kCTA4 + 32;
struct alignas(128) DivStage {
  float4 vw[kB4 * kBV];      // 16 KiB
  uint32_t fd[kB4 * 12];     // 768 B
  uint32_t child[kB4 * kSV]; // 512 B
  int4 posl[kB4];            // 256 B
  uint32_t parent[kB4];      // 64 B (T.parent is padded to a multiple of kB4 entries)
};
struct alignas(128) ApplyStage {
  float4 vw[kB4 * kBV];
  float p[kB4 * kBV];        // 4 KiB
  uint32_t fd[kB4 * 12];
  uint32_t child[kB4 * kSV];
  int4 posl[kB4];
  uint32_t parent[kB4];
};
constexpr size_t kDivPipeSmem = kSStages * sizeof(DivStage) + 2 * kSStages * sizeof(uint64_t);
constexpr size_t kApplyPipeSmem = kSStages * sizeof(ApplyStage) + 2 * kSStages * sizeof(uint64_t);

// the quad's four packed velocities from a staged tile: conflict-free rotated reads, then un-rotation
__device__ __forceinline__ void load_quad_velocities(const float4 *blk, int t, float4 v[4]) {
  const int r = (t >> 1) & 3;
  float4 a[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; k++) a[k] = blk[4 * t + ((k + r) & 3)];  // a[k] = chunk (k + r) & 3
#pragma unroll
  for (int c = 0; c < 4; c++) b[c] = (r & 1) ? a[(c + 3) & 3] : a[c];
#pragma unroll
  for (int c = 0; c < 4; c++) v[c] = (r & 2) ? b[(c + 2) & 3] : b[c];
}

template <class Stage>
__device__ __forceinline__ void stencil_ring_init(uint64_t *full, uint64_t *empty) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kSStages; s++) {
      pipe::mbar_init(&full[s], 1);
      pipe::mbar_init(&empty[s], kCTA4 / 32);
    }
    pipe::fence_barrier_init();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kStencilThreads, 3) k_dc_divergence_pipe(Pool T, KParams P, TileRuns R, const float4 *__restrict__ vw,
                                                                           float *__restrict__ div, float *__restrict__ p, float *__restrict__ tp,
                                                                           int zero_from) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DivStage *st = reinterpret_cast<DivStage *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kSStages * sizeof(DivStage));
  uint64_t *empty = full + kSStages;
  const uint32_t ntiles = run_total(R);
  stencil_ring_init<DivStage>(full, empty);
  if (threadIdx.x >= kCTA4) {  // producer warp
    if (threadIdx.x != kCTA4) return;
    uint32_t j = blockIdx.x;
    for (uint32_t it = 0; j < ntiles; it++, j += gridDim.x) {
      const uint32_t s = it % kSStages;
      if (it >= (uint32_t)kSStages) pipe::mbar_wait(&empty[s], ((it / kSStages) - 1u) & 1u);
      const size_t b0 = (size_t)run_tile(R, j) * kB4;
      const uint32_t nv = min((uint32_t)kB4, T.M - (uint32_t)b0);
      DivStage &S = st[s];
      pipe::mbar_expect_tx(&full[s], nv * (kBV * 16u + 48u + kSV * 4u + 16u) + kB4 * 4u);
      pipe::bulk_g2s(S.vw, vw + b0 * kBV, nv * kBV * 16u, &full[s]);
      pipe::bulk_g2s(S.fd, T.fd + b0 * 12, nv * 48u, &full[s]);
      pipe::bulk_g2s(S.child, T.child + b0 * kSV, nv * kSV * 4u, &full[s]);
      pipe::bulk_g2s(S.posl, T.posl + b0, nv * 16u, &full[s]);
      pipe::bulk_g2s(S.parent, T.parent + b0, kB4 * 4u, &full[s]);
    }
    return;
  }
  const uint32_t g = threadIdx.x >> 4;
  const int t = threadIdx.x & 15;
  int X, Y0, Z0;
  quad_coords(t, X, Y0, Z0);
  const int sy = (t >> 2) & 1, sz = (t >> 1) & 1;
  const unsigned half = 0xFFFFu << (threadIdx.x & 16u);
  uint32_t j = blockIdx.x;
  for (uint32_t it = 0; j < ntiles; it++, j += gridDim.x) {
    const uint32_t s = it % kSStages;
    pipe::mbar_wait(&full[s], (it / kSStages) & 1u);
    const DivStage &S = st[s];
    const uint32_t b = run_tile(R, j) * kB4 + g;
    int4 pl = make_int4(0, 0, 0, kFree);
    if (b < T.M) pl = S.posl[g];
    const bool active = pl.w != kFree;
    float4 v[4];
    uint32_t child = kNone, ps = kNone;
    QuadGhosts q;
    q.has_x = false;
    if (active) {
      load_quad_velocities(S.vw + g * kBV, t, v);
      child = S.child[g * kSV + (t >> 1)];
      ps = S.parent[g];
      const uint4 *fdp = reinterpret_cast<const uint4 *>(&S.fd[g * 12]);
      q = quad_ghosts_from(T, b, t, fdp[0], fdp[1], fdp[2]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pipe::fence_proxy_async();
    __syncwarp();
    if ((threadIdx.x & 31u) == 0) pipe::mbar_arrive(&empty[s]);
    if (!__any_sync(0xFFFFFFFFu, active)) continue;
    const int scale = 1 << (pl.w & 15);
    float4 gx = make_float4(0.f, 0.f, 0.f, 0.f);
    float gy0 = 0.f, gy1 = 0.f, gz0 = 0.f, gz1 = 0.f;
    if (active) {
      if (q.has_x) {
        const int nx = pl.x + (X == 3 ? kBW : -1);
        gx.x = ghost_product(P, vw, q.x[0], 0, nx, pl.y + Y0, pl.z + Z0, scale);
        gx.y = ghost_product(P, vw, q.x[1], 0, nx, pl.y + Y0, pl.z + Z0 + 1, scale);
        gx.z = ghost_product(P, vw, q.x[2], 0, nx, pl.y + Y0 + 1, pl.z + Z0, scale);
        gx.w = ghost_product(P, vw, q.x[3], 0, nx, pl.y + Y0 + 1, pl.z + Z0 + 1, scale);
      }
      const int ny = pl.y + (sy ? kBW : -1), nz = pl.z + (sz ? kBW : -1);
      gy0 = ghost_product(P, vw, q.y[0], 1, pl.x + X, ny, pl.z + Z0, scale);
      gy1 = ghost_product(P, vw, q.y[1], 1, pl.x + X, ny, pl.z + Z0 + 1, scale);
      gz0 = ghost_product(P, vw, q.z[0], 2, pl.x + X, pl.y + Y0, nz, scale);
      gz1 = ghost_product(P, vw, q.z[1], 2, pl.x + X, pl.y + Y0 + 1, nz, scale);
    }
    const float4 px = make_float4(v[0].w * v[0].x, v[1].w * v[1].x, v[2].w * v[2].x, v[3].w * v[3].x);
    const float4 py = make_float4(v[0].w * v[0].y, v[1].w * v[1].y, v[2].w * v[2].y, v[3].w * v[3].y);
    const float4 pz = make_float4(v[0].w * v[0].z, v[1].w * v[1].z, v[2].w * v[2].z, v[3].w * v[3].z);
    QuadNbr nx_, ny_, nz_;
    quad_exchange_x(nx_, px, t, gx);
    quad_exchange_y(ny_, py, t, gy0, gy1);
    quad_exchange_z(nz_, pz, t, gz0, gz1);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 d = zero;
    const bool leaf = active && child == kNone;
    if (leaf) {
      const float alpha = .5f * P.rdx / (float)scale;
      d.x = alpha * (nx_.xp.x - nx_.xm.x + py.z - ny_.ym0 + pz.y - nz_.zm0);
      d.y = alpha * (nx_.xp.y - nx_.xm.y + py.w - ny_.ym1 + nz_.zp0 - pz.x);
      d.z = alpha * (nx_.xp.z - nx_.xm.z + ny_.yp0 - py.x + pz.w - nz_.zm1);
      d.w = alpha * (nx_.xp.w - nx_.xm.w + ny_.yp1 - py.y + nz_.zp1 - pz.z);
    }
    const bool childless = (__ballot_sync(0xFFFFFFFFu, child != kNone) & half) == 0;
    float sm = 0.f;
    sm += d.x; sm += d.y; sm += d.z; sm += d.w;
    const float lo = __shfl_sync(0xFFFFFFFFu, sm, (threadIdx.x & 31u) ^ 1u);
    if (!active) continue;
    const size_t c0 = (size_t)b * kBV + 4 * t;
    if (pl.w >= zero_from) {
      __stcs(reinterpret_cast<float4 *>(p + c0), zero);
      __stcs(reinterpret_cast<float4 *>(tp + c0), zero);
    }
    if (leaf) __stcs(reinterpret_cast<float4 *>(div + c0), d);
    if (childless && (t & 1) && ps != kNone) {
      float a = lo;
      a += d.x; a += d.y; a += d.z; a += d.w;
      div[(size_t)kSV * ps + (t >> 1)] = a * .125f;
    }
  }
}

template <int kMinBlocks>
__global__ void __launch_bounds__(kStencilThreads, kMinBlocks) k_dc_apply_pipe(Pool T, KParams P, TileRuns R, const float *__restrict__ p,
                                                                               const float *__restrict__ fl, float4 *__restrict__ vw) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ApplyStage *st = reinterpret_cast<ApplyStage *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kSStages * sizeof(ApplyStage));
  uint64_t *empty = full + kSStages;
  const uint32_t ntiles = run_total(R);
  stencil_ring_init<ApplyStage>(full, empty);
  if (threadIdx.x >= kCTA4) {  // producer warp
    if (threadIdx.x != kCTA4) return;
    uint32_t j = blockIdx.x;
    for (uint32_t it = 0; j < ntiles; it++, j += gridDim.x) {
      const uint32_t s = it % kSStages;
      if (it >= (uint32_t)kSStages) pipe::mbar_wait(&empty[s], ((it / kSStages) - 1u) & 1u);
      const size_t b0 = (size_t)run_tile(R, j) * kB4;
      const uint32_t nv = min((uint32_t)kB4, T.M - (uint32_t)b0);
      ApplyStage &S = st[s];
      pipe::mbar_expect_tx(&full[s], nv * (kBV * 16u + kBV * 4u + 48u + kSV * 4u + 16u) + kB4 * 4u);
      pipe::bulk_g2s(S.vw, vw + b0 * kBV, nv * kBV * 16u, &full[s]);
      pipe::bulk_g2s(S.p, p + b0 * kBV, nv * kBV * 4u, &full[s]);
      pipe::bulk_g2s(S.fd, T.fd + b0 * 12, nv * 48u, &full[s]);
      pipe::bulk_g2s(S.child, T.child + b0 * kSV, nv * kSV * 4u, &full[s]);
      pipe::bulk_g2s(S.posl, T.posl + b0, nv * 16u, &full[s]);
      pipe::bulk_g2s(S.parent, T.parent + b0, kB4 * 4u, &full[s]);
    }
    return;
  }
  const uint32_t g = threadIdx.x >> 4;
  const int t = threadIdx.x & 15;
  const unsigned half = 0xFFFFu << (threadIdx.x & 16u);
  uint32_t j = blockIdx.x;
  for (uint32_t it = 0; j < ntiles; it++, j += gridDim.x) {
    const uint32_t s = it % kSStages;
    pipe::mbar_wait(&full[s], (it / kSStages) & 1u);
    const ApplyStage &S = st[s];
    const uint32_t b = run_tile(R, j) * kB4 + g;
    int level = kFree;
    if (b < T.M) level = S.posl[g].w;
    const bool active = level != kFree;
    float4 v[4];
    float4 op = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t child = kNone, ps = kNone;
    QuadGhosts q;
    q.has_x = false;
    if (active) {
      load_quad_velocities(S.vw + g * kBV, t, v);
      op = *reinterpret_cast<const float4 *>(&S.p[g * kBV + 4 * t]);
      child = S.child[g * kSV + (t >> 1)];
      ps = S.parent[g];
      const uint4 *fdp = reinterpret_cast<const uint4 *>(&S.fd[g * 12]);
      q = quad_ghosts_from(T, b, t, fdp[0], fdp[1], fdp[2]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pipe::fence_proxy_async();
    __syncwarp();
    if ((threadIdx.x & 31u) == 0) pipe::mbar_arrive(&empty[s]);
    if (!__any_sync(0xFFFFFFFFu, active)) continue;
    const float4 ow = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);  // == fl[c0..c0+3]
    float4 pgx = make_float4(0.f, 0.f, 0.f, 0.f), wgx = pgx;
    float pg[4] = {0.f, 0.f, 0.f, 0.f}, wg[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
      if (q.has_x) {
        pgx = make_float4(p[q.x[0]], p[q.x[1]], p[q.x[2]], p[q.x[3]]);
        wgx = make_float4(fl[q.x[0]], fl[q.x[1]], fl[q.x[2]], fl[q.x[3]]);
      }
      pg[0] = p[q.y[0]]; pg[1] = p[q.y[1]]; pg[2] = p[q.z[0]]; pg[3] = p[q.z[1]];
      wg[0] = fl[q.y[0]]; wg[1] = fl[q.y[1]]; wg[2] = fl[q.z[0]]; wg[3] = fl[q.z[1]];
    }
    const QuadNbr sN = quad_exchange(op, t, pgx, pg[0], pg[1], pg[2], pg[3]);
    const QuadNbr w = quad_exchange(ow, t, wgx, wg[0], wg[1], wg[2], wg[3]);
    const unsigned with_child = __ballot_sync(0xFFFFFFFFu, child != kNone);
    const bool childless = active && (with_child & half) == 0;
    const bool leaf = active && child == kNone;
    const float alpha = .5f * P.rdx / (float)(1 << (level & 15));
    v[0].x -= alpha * (w.xp.x * (sN.xp.x - op.x) + w.xm.x * (op.x - sN.xm.x));
    v[0].y -= alpha * (ow.z * (op.z - op.x) + w.ym0 * (op.x - sN.ym0));
    v[0].z -= alpha * (ow.y * (op.y - op.x) + w.zm0 * (op.x - sN.zm0));
    v[1].x -= alpha * (w.xp.y * (sN.xp.y - op.y) + w.xm.y * (op.y - sN.xm.y));
    v[1].y -= alpha * (ow.w * (op.w - op.y) + w.ym1 * (op.y - sN.ym1));
    v[1].z -= alpha * (w.zp0 * (sN.zp0 - op.y) + ow.x * (op.y - op.x));
    v[2].x -= alpha * (w.xp.z * (sN.xp.z - op.z) + w.xm.z * (op.z - sN.xm.z));
    v[2].y -= alpha * (w.yp0 * (sN.yp0 - op.z) + ow.x * (op.z - op.x));
    v[2].z -= alpha * (ow.w * (op.w - op.z) + w.zm1 * (op.z - sN.zm1));
    v[3].x -= alpha * (w.xp.w * (sN.xp.w - op.w) + w.xm.w * (op.w - sN.xm.w));
    v[3].y -= alpha * (w.yp1 * (sN.yp1 - op.w) + ow.y * (op.w - op.y));
    v[3].z -= alpha * (w.zp1 * (sN.zp1 - op.w) + ow.z * (op.w - op.z));
    const size_t c0 = (size_t)b * kBV + 4 * t;
    if (leaf) {
#pragma unroll
      for (int k = 0; k < 4; k++) vw[c0 + k] = v[k];
    }
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) { ax += v[k].x; ay += v[k].y; az += v[k].z; }
    const unsigned src = (threadIdx.x & 31u) ^ 1u;
    const float lx = __shfl_sync(0xFFFFFFFFu, ax, src), ly = __shfl_sync(0xFFFFFFFFu, ay, src), lz = __shfl_sync(0xFFFFFFFFu, az, src);
    if (childless && (t & 1) && ps != kNone) {
      ax = lx; ay = ly; az = lz;
#pragma unroll
      for (int k = 0; k < 4; k++) { ax += v[k].x; ay += v[k].y; az += v[k].z; }
      float *dst = reinterpret_cast<float *>(vw + ((size_t)kSV * ps + (t >> 1)));
      dst[0] = ax * .125f; dst[1] = ay * .125f; dst[2] = az * .125f;
    }
  }
}

}  // namespace dcg
