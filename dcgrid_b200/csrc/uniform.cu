// B200-native dense-grid solver: drop-in for FluidSimulationUniform
// (reference src/uniformgrid/fluid_simulation_uniform.{h,cu}, kernels in
// src/uniformgrid/uniformgrid_{structure,fluid}.cu).
//
// HBM layout (all fp32, x fastest: idx = (z*gy+y)*gx+x, src/utils/grid_math.cuh:10):
//   vw[2]      float4[N]   (vx,vy,vz,fluidity) ping-pong — one 16-byte load per gather corner;
//                          replaces velocity float3[N] + the level-0 fluidity read + the
//                          whole-field D2D memcpy of fluid_simulation_uniform.cu:92-93
//   q[2]       float[N]    density ping-pong (no D2D memcpy, :139-140)
//   fluidity   float[pyr]  full mip pyramid (mipmapCells/mipmapIdx, grid_math.cuh:24-52);
//                          static: recomputed only when SimParams change (the reference
//                          recomputes every step, :143-147)
//   p, tp, div float[pyr]  pressure / t_pressure / divergence pyramids (explicit buffers
//                          instead of the aliased `temporary`, uniformgrid_structure.cu:7-13)
// Arithmetic keeps the reference's expression order; compiled with -fmad=false.
#include <algorithm>
#include <string>
#include <vector>

#include "shard_vmm.h"
#include "sim.h"

namespace dcg {
namespace {

constexpr int BX = 32, BY = 4, BZ = 2;  // 256 threads, x-contiguous 128-byte rows
// planes [zb, ze) of the level a launch covers: everything on one GPU, a rank's slab in the sharded solver (blockIdx.z counts from zb)
struct ZRange {
  int zb, ze;
};

struct UGrid {
  int gx, gy, gz;
  uint64_t N;
};

__device__ __forceinline__ uint64_t lidx(const KParams &P, int x, int y, int z) {
  return ((uint64_t)z * P.gy + y) * P.gx + x;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }

// k_uniform_set_solidity_ratio, uniformgrid_structure.cu:23-31
__global__ void __launch_bounds__(256) k_u_fluidity(KParams P, float *__restrict__ fluidity, uint64_t off, int level,
                                                    float4 *__restrict__ vw0, float4 *__restrict__ vw1, ZRange Z) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y, z = Z.zb + blockIdx.z * BZ + threadIdx.z;
  const int w = P.gx >> level, h = P.gy >> level, d = P.gz >> level;
  if (x >= w || y >= h || z >= min(d, Z.ze)) return;
  const int scale = 1 << level;
  const float f = cell_fluidity(P, x, y, z, scale);
  const uint64_t i = ((uint64_t)z * (P.gy / scale) + y) * (P.gx / scale) + x;
  fluidity[off + i] = f;
  if (level == 0) {  // keep the packed copies in sync
    vw0[i].w = f;
    vw1[i].w = f;
  }
}

// Common gather set-up: INIT_SAMPLE, uniformgrid_fluid.cu:7-27
struct USample {
  uint64_t id[8];
  int x0, y0, z0;
  float fx, fy, fz;
};
__device__ __forceinline__ USample u_sample(const KParams &P, float px, float py, float pz) {
  USample s;
  const float x = px - .5f, y = py - .5f, z = pz - .5f;
  const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  s.x0 = (int)xf; s.y0 = (int)yf; s.z0 = (int)zf;
  s.fx = x - xf; s.fy = y - yf; s.fz = z - zf;
  const int xa = clampi(s.x0, 0, P.gx - 1), xb = clampi(s.x0 + 1, 0, P.gx - 1);
  const int ya = clampi(s.y0, 0, P.gy - 1), yb = clampi(s.y0 + 1, 0, P.gy - 1);
  const int za = clampi(s.z0, 0, P.gz - 1), zb = clampi(s.z0 + 1, 0, P.gz - 1);
  // one full index, the other seven by adding the (0 or 1 cell) steps: the kernels that gather are bound by instruction
  // issue, not by memory (ncu: 71 % of the issue slots, 500 warp instructions per 32 cells)
  const uint64_t base = lidx(P, xa, ya, za);
  const uint64_t ox = (uint64_t)(xb - xa), oy = (uint64_t)(yb - ya) * P.gx, oz = (uint64_t)(zb - za) * P.gx * P.gy;
  s.id[0] = base; s.id[1] = base + oz;
  s.id[2] = base + oy; s.id[3] = base + oy + oz;
  s.id[4] = base + ox; s.id[5] = base + ox + oz;
  s.id[6] = base + ox + oy; s.id[7] = base + ox + oy + oz;
  return s;
}
// true iff all 8 corners lie inside the domain: the boundary conditions are the identity there (sim_utils.cu:24-55)
__device__ __forceinline__ bool u_inside(const KParams &P, const USample &s) {
  return s.x0 >= 0 && s.y0 >= 0 && s.z0 >= 0 && s.x0 + 1 < P.gx && s.y0 + 1 < P.gy && s.z0 + 1 < P.gz;
}

// k_uniform_advect_velocity, uniformgrid_fluid.cu:50-67,88-95
__global__ void __launch_bounds__(256) k_u_advect_velocity(KParams P, const float4 *__restrict__ vin, float4 *__restrict__ vout, ZRange Z) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y, z = Z.zb + blockIdx.z * BZ + threadIdx.z;
  if (x >= P.gx || y >= P.gy || z >= min(P.gz, Z.ze)) return;
  const uint64_t i = lidx(P, x, y, z);
  const float4 me = vin[i];
  const float bx = ((float)x + .5f) - me.x * P.dt * P.rdx;
  const float by = ((float)y + .5f) - me.y * P.dt * P.rdx;
  const float bz = ((float)z + .5f) - me.z * P.dt * P.rdx;
  const USample s = u_sample(P, bx, by, bz);
  float4 c[8];
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    c[k] = vin[s.id[k]];
    f[k] = c[k].w;
  }
  const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
  float3 out = make_float3(0.f, 0.f, 0.f);
  if (!(W.acc < 1e-6f)) {
    float vx[8], vy[8], vz[8];
    const bool inside = u_inside(P, s);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float3 v = make_float3(c[k].x, c[k].y, c[k].z);
      if (!inside) v = velocity_bc(P, v, s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1), s.z0 + (k & 1), 1);
      vx[k] = v.x; vy[k] = v.y; vz[k] = v.z;
    }
    out = make_float3(blend8(vx, W.w), blend8(vy, W.w), blend8(vz, W.w));
  }
  vout[i] = make_float4(out.x, out.y, out.z, me.w);
}

// k_uniform_advect_density, uniformgrid_fluid.cu:69-86,97-105
__global__ void __launch_bounds__(256) k_u_advect_density(KParams P, const float4 *__restrict__ vw, const float *__restrict__ qin,
                                                          float *__restrict__ qout, ZRange Z) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y, z = Z.zb + blockIdx.z * BZ + threadIdx.z;
  if (x >= P.gx || y >= P.gy || z >= min(P.gz, Z.ze)) return;
  const uint64_t i = lidx(P, x, y, z);
  const float4 me = vw[i];
  const float bx = ((float)x + .5f) - me.x * P.dt * P.rdx;
  const float by = ((float)y + .5f) - me.y * P.dt * P.rdx;
  const float bz = ((float)z + .5f) - me.z * P.dt * P.rdx;
  const USample s = u_sample(P, bx, by, bz);
  float q[8], f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    f[k] = vw[s.id[k]].w;
    q[k] = qin[s.id[k]];
  }
  const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
  float out = 0.f;
  if (!(W.acc < 1e-6f)) {
    if (!u_inside(P, s)) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        q[k] = density_bc(P, q[k], s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1), s.z0 + (k & 1), 1);
    }
    out = blend8(q, W.w);
  }
  qout[i] = out;
}

// advectDensity() of step n and advectVelocity() of step n+1 in one pass: both backtrace every cell through the
// same velocity field (nothing runs between them, simulation.cpp:104-111) and therefore share the sample — the 8
// clamped ids and the fluidity-weighted weights.  The advected velocity goes to the idle ping-pong buffer; the
// host uses it in the next advect_velocity() if nothing touched the state in between (`spec_velocity`).
__global__ void __launch_bounds__(256) k_u_advect_both(KParams P, const float4 *__restrict__ vin, float4 *__restrict__ vout,
                                                       const float *__restrict__ qin, float *__restrict__ qout, ZRange Z) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y, z = Z.zb + blockIdx.z * BZ + threadIdx.z;
  if (x >= P.gx || y >= P.gy || z >= min(P.gz, Z.ze)) return;
  const uint64_t i = lidx(P, x, y, z);
  const float4 me = vin[i];
  const float bx = ((float)x + .5f) - me.x * P.dt * P.rdx;
  const float by = ((float)y + .5f) - me.y * P.dt * P.rdx;
  const float bz = ((float)z + .5f) - me.z * P.dt * P.rdx;
  const USample s = u_sample(P, bx, by, bz);
  float4 c[8];
  float q[8], f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    c[k] = vin[s.id[k]];
    q[k] = qin[s.id[k]];
    f[k] = c[k].w;
  }
  const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
  float3 vo = make_float3(0.f, 0.f, 0.f);
  float qo = 0.f;
  if (!(W.acc < 1e-6f)) {
    float vx[8], vy[8], vz[8];
    const bool inside = u_inside(P, s);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float3 v = make_float3(c[k].x, c[k].y, c[k].z);
      if (!inside) {
        const int cx = s.x0 + ((k >> 2) & 1), cy = s.y0 + ((k >> 1) & 1), cz = s.z0 + (k & 1);
        v = velocity_bc(P, v, cx, cy, cz, 1);
        q[k] = density_bc(P, q[k], cx, cy, cz, 1);
      }
      vx[k] = v.x; vy[k] = v.y; vz[k] = v.z;
    }
    vo = make_float3(blend8(vx, W.w), blend8(vy, W.w), blend8(vz, W.w));
    qo = blend8(q, W.w);
  }
  vout[i] = make_float4(vo.x, vo.y, vo.z, me.w);
  qout[i] = qo;
}

// The same, marching along z: the cell's own velocity — the first of the two dependent load rounds of a gather — is
// fetched one plane ahead, so that only the corner loads are on a thread's critical path, and the index set-up is paid
// once per column.
// 4 CTAs per SM (64 registers, 16 bytes of spill) instead of the 3 ptxas picks on its own (79 registers): the kernel waits on
// its gathers, a fourth CTA's warps are worth more than the registers — 14.5 -> 13.75 ms at 1024^3; 5 CTAs (48 registers,
// 160 bytes of spill): 34 ms (profiles/README.md r4b-r4d).
__global__ void __launch_bounds__(256, 4) k_u_advect_both_zm(KParams P, const float4 *__restrict__ vin, float4 *__restrict__ vout,
                                                          const float *__restrict__ qin, float *__restrict__ qout, int zc, ZRange Z) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  const int z0 = Z.zb + blockIdx.z * zc, z1 = min(z0 + zc, min(P.gz, Z.ze));
  if (x >= P.gx || y >= P.gy) return;
  const uint64_t sz = (uint64_t)P.gx * P.gy;
  const uint64_t col = (uint64_t)y * P.gx + x;
  const float alpha = P.dt * P.rdx;  // (me.x * P.dt * P.rdx evaluates as (me.x * dt) * rdx in the reference: kept below)
  (void)alpha;
  float4 nxt = vin[col + (uint64_t)z0 * sz];
  for (int z = z0; z < z1; z++) {
    const uint64_t i = col + (uint64_t)z * sz;
    const float4 me = nxt;
    if (z + 1 < z1) nxt = vin[i + sz];
    const float bx = ((float)x + .5f) - me.x * P.dt * P.rdx;
    const float by = ((float)y + .5f) - me.y * P.dt * P.rdx;
    const float bz = ((float)z + .5f) - me.z * P.dt * P.rdx;
    const USample s = u_sample(P, bx, by, bz);
    float4 c[8];
    float q[8], f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      c[k] = vin[s.id[k]];
      q[k] = qin[s.id[k]];
      f[k] = c[k].w;
    }
    const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
    float3 vo = make_float3(0.f, 0.f, 0.f);
    float qo = 0.f;
    if (!(W.acc < 1e-6f)) {
      float vx[8], vy[8], vz[8];
      const bool inside = u_inside(P, s);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float3 v = make_float3(c[k].x, c[k].y, c[k].z);
        if (!inside) {
          const int cx = s.x0 + ((k >> 2) & 1), cy = s.y0 + ((k >> 1) & 1), cz = s.z0 + (k & 1);
          v = velocity_bc(P, v, cx, cy, cz, 1);
          q[k] = density_bc(P, q[k], cx, cy, cz, 1);
        }
        vx[k] = v.x; vy[k] = v.y; vz[k] = v.z;
      }
      vo = make_float3(blend8(vx, W.w), blend8(vy, W.w), blend8(vz, W.w));
      qo = blend8(q, W.w);
    }
    vout[i] = make_float4(vo.x, vo.y, vo.z, me.w);
    qout[i] = qo;
  }
}

// k_uniform_calc_divergence, uniformgrid_fluid.cu:107-132
// zero: bit 0 = clear p, bit 1 = clear t_p like the reference does here; project() skips the level-0 clears it provably
// never reads (the prolongation rewrites every p of the level, the first sweep every t_p, before either is read)
__global__ void __launch_bounds__(256) k_u_divergence(KParams P, const float4 *__restrict__ vw, float *__restrict__ div,
                                                      float *__restrict__ p, float *__restrict__ tp, int zero, ZRange Z) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y, z = Z.zb + blockIdx.z * BZ + threadIdx.z;
  if (x >= P.gx || y >= P.gy || z >= min(P.gz, Z.ze)) return;
  const uint64_t i = lidx(P, x, y, z);
  const uint64_t sx = 1, sy = P.gx, sz = (uint64_t)P.gx * P.gy;
  const float4 l = vw[x > 0 ? i - sx : i], r = vw[x < P.gx - 1 ? i + sx : i];
  const float4 dn = vw[y > 0 ? i - sy : i], up = vw[y < P.gy - 1 ? i + sy : i];
  const float4 b = vw[z > 0 ? i - sz : i], f = vw[z < P.gz - 1 ? i + sz : i];
  const float3 vl = velocity_bc(P, make_float3(l.x, l.y, l.z), x - 1, y, z, 1);
  const float3 vr = velocity_bc(P, make_float3(r.x, r.y, r.z), x + 1, y, z, 1);
  const float3 vd = velocity_bc(P, make_float3(dn.x, dn.y, dn.z), x, y - 1, z, 1);
  const float3 vu = velocity_bc(P, make_float3(up.x, up.y, up.z), x, y + 1, z, 1);
  const float3 vb = velocity_bc(P, make_float3(b.x, b.y, b.z), x, y, z - 1, 1);
  const float3 vf = velocity_bc(P, make_float3(f.x, f.y, f.z), x, y, z + 1, 1);
  if (zero & 1) p[i] = 0.f;
  if (zero & 2) tp[i] = 0.f;
  div[i] = .5f * P.rdx * (r.w * vr.x - l.w * vl.x + up.w * vu.y - dn.w * vd.y + f.w * vf.z - b.w * vb.z);
}

// k_uniform_restrict, uniformgrid_fluid.cu:134-160.  `off`/`coff` = pyramid offsets of this/child level.
__global__ void __launch_bounds__(256) k_u_restrict(KParams P, int level, uint64_t off, uint64_t coff, float *__restrict__ div,
                                                    float *__restrict__ p, float *__restrict__ tp, ZRange Z) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y, z = Z.zb + blockIdx.z * BZ + threadIdx.z;
  const int W = P.gx >> level, H = P.gy >> level, D = P.gz >> level;
  if (x >= W || y >= H || z >= min(D, Z.ze)) return;
  const uint64_t cw = (uint64_t)(P.gx >> (level - 1)), ch = (uint64_t)(P.gy >> (level - 1));
  const uint64_t i = off + ((uint64_t)z * H + y) * W + x;
  const float *c = div + coff + ((uint64_t)(2 * z) * ch + 2 * y) * cw + 2 * x;
  p[i] = 0.f;
  tp[i] = 0.f;
  div[i] = .125f * (c[0] + c[1] + c[cw] + c[cw + 1] + c[cw * ch] + c[cw * ch + 1] + c[cw * ch + cw] + c[cw * ch + cw + 1]);
}

// calcPressure<in,out>, uniformgrid_fluid.cu:162-192 (k_uniform_jacobi / k_uniform_jacobi_inv :194-204)
__global__ void __launch_bounds__(256) k_u_jacobi(KParams P, int level, uint64_t off, const float *__restrict__ in,
                                                  float *__restrict__ out, const float *__restrict__ div, ZRange Z) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y, z = Z.zb + blockIdx.z * BZ + threadIdx.z;
  const int w = P.gx >> level, h = P.gy >> level, d = P.gz >> level;
  if (x >= w || y >= h || z >= min(d, Z.ze)) return;
  const int scale = 1 << level;
  const float alpha = P.dx * P.dx * scale * scale;
  const uint64_t i = off + ((uint64_t)z * h + y) * w + x;
  const uint64_t sy = w, sz = (uint64_t)w * h;
  const float pl = in[x > 0 ? i - 1 : i], pr = in[x < w - 1 ? i + 1 : i];
  const float pd = in[y > 0 ? i - sy : i], pu = in[y < h - 1 ? i + sy : i];
  const float pb = in[z > 0 ? i - sz : i], pf = in[z < d - 1 ? i + sz : i];
  out[i] = (pl + pr + pd + pu + pb + pf - alpha * div[i]) / 6.f;
}

// ---- z-marching variants (sm_100a tuning; profiles/README.md r2h) --------------------------------------------------
// ncu on the one-thread-per-cell kernels above: the Jacobi sweep issues 120 warp instructions per 32 cells (64-bit index
// arithmetic, boundary selects, the generic division) and sits at 79 % of the issue slots with DRAM traffic exactly at
// its 12 B/cell; the divergence keeps the L1 data pipe 99 % busy with six 16-byte neighbour loads per cell.  Here a
// thread marches along z and keeps its column's three planes in registers, so that the z neighbours cost nothing, the x
// neighbours come from the adjacent lanes by shuffle (only the two edge lanes of a warp load), and the index arithmetic
// is paid once per thread: 4 cells per thread and plane for the scalar fields (one float4), 1 cell for the packed
// velocity.  Same expressions in the same order; x / 6.f through div6() (bit-identical, common.cuh).
constexpr int ZTX = 32, ZTY = 8;  // 256 threads: a warp covers one x row of the tile

__global__ void __launch_bounds__(ZTX * ZTY) k_u_jacobi_zm(KParams P, int level, uint64_t off, const float *__restrict__ in, float *__restrict__ out,
                                                           const float *__restrict__ div, int zc, ZRange Z) {
  const int w = P.gx >> level, h = P.gy >> level, d = P.gz >> level;
  const int x = 4 * (blockIdx.x * ZTX + threadIdx.x), y = blockIdx.y * ZTY + threadIdx.y;
  const int z0 = Z.zb + blockIdx.z * zc, z1 = min(z0 + zc, min(d, Z.ze));
  const int scale = 1 << level;
  const float alpha = P.dx * P.dx * scale * scale;
  const size_t sy = (size_t)w, sz = (size_t)w * h;
  const size_t col = off + (size_t)y * w + x;  // + z * sz
  const size_t dnrow = y > 0 ? col - sy : col, uprow = y < h - 1 ? col + sy : col;
  const unsigned lane = threadIdx.x;
  const bool left_edge = lane == 0, right_edge = lane == ZTX - 1;
  auto ld4 = [&](size_t i) { return *reinterpret_cast<const float4 *>(in + i); };
  float4 cur = ld4(col + (size_t)z0 * sz);
  float4 prev = z0 > 0 ? ld4(col + (size_t)(z0 - 1) * sz) : cur;
  for (int z = z0; z < z1; z++) {
    const size_t zo = (size_t)z * sz;
    const float4 next = z < d - 1 ? ld4(col + zo + sz) : cur;
    const float4 dn = ld4(dnrow + zo), up = ld4(uprow + zo);
    const float4 dv = *reinterpret_cast<const float4 *>(div + col + zo);
    float xl = __shfl_up_sync(0xFFFFFFFFu, cur.w, 1), xr = __shfl_down_sync(0xFFFFFFFFu, cur.x, 1);
    if (left_edge) xl = x > 0 ? in[col + zo - 1] : cur.x;
    if (right_edge) xr = x + 4 < w ? in[col + zo + 4] : cur.w;
    float4 o;
    o.x = div6(xl + cur.y + dn.x + up.x + prev.x + next.x - alpha * dv.x);
    o.y = div6(cur.x + cur.z + dn.y + up.y + prev.y + next.y - alpha * dv.y);
    o.z = div6(cur.y + cur.w + dn.z + up.z + prev.z + next.z - alpha * dv.z);
    o.w = div6(cur.z + xr + dn.w + up.w + prev.w + next.w - alpha * dv.w);
    *reinterpret_cast<float4 *>(out + col + zo) = o;
    prev = cur;
    cur = next;
  }
}

// (Measured and rejected, profiles/README.md r3c: the y neighbours' (vy, fluidity) exchanged between the warps of the tile through
// shared memory instead of two 16-byte loads per cell: 6.74 ms at 1024^3 either way — those loads hit L1 / L2.)
__global__ void __launch_bounds__(ZTX * ZTY) k_u_divergence_zm(KParams P, const float4 *__restrict__ vw, float *__restrict__ div, float *__restrict__ p,
                                                               float *__restrict__ tp, int zc, int zero, ZRange Z) {
  const int x = blockIdx.x * ZTX + threadIdx.x, y = blockIdx.y * ZTY + threadIdx.y;
  const int z0 = Z.zb + blockIdx.z * zc, z1 = min(z0 + zc, min(P.gz, Z.ze));
  const size_t sy = (size_t)P.gx, sz = (size_t)P.gx * P.gy;
  const size_t col = (size_t)y * P.gx + x;
  const size_t dnrow = y > 0 ? col - sy : col, uprow = y < P.gy - 1 ? col + sy : col;
  const unsigned lane = threadIdx.x;
  const bool left_edge = lane == 0, right_edge = lane == ZTX - 1;
  // boundary conditions only substitute values of neighbours OUTSIDE the domain (sim_utils.cu:24-39)
  const bool xy_inner = x > 0 && x < P.gx - 1 && y > 0 && y < P.gy - 1;
  float4 cur = vw[col + (size_t)z0 * sz];
  float4 prev = z0 > 0 ? vw[col + (size_t)(z0 - 1) * sz] : cur;
  for (int z = z0; z < z1; z++) {
    const size_t zo = (size_t)z * sz;
    const float4 next = z < P.gz - 1 ? vw[col + zo + sz] : cur;
    const float4 dn = vw[dnrow + zo], up = vw[uprow + zo];
    float4 l, r;
    l.x = __shfl_up_sync(0xFFFFFFFFu, cur.x, 1); l.w = __shfl_up_sync(0xFFFFFFFFu, cur.w, 1);
    r.x = __shfl_down_sync(0xFFFFFFFFu, cur.x, 1); r.w = __shfl_down_sync(0xFFFFFFFFu, cur.w, 1);
    l.y = l.z = r.y = r.z = 0.f;
    if (left_edge) l = x > 0 ? vw[col + zo - 1] : cur;
    if (right_edge) r = x < P.gx - 1 ? vw[col + zo + 1] : cur;
    float lx = l.x, rx = r.x, dy = dn.y, uy = up.y, bz = prev.z, fz = next.z;
    if (!(xy_inner && z > 0 && z < P.gz - 1)) {
      lx = velocity_bc(P, make_float3(l.x, l.y, l.z), x - 1, y, z, 1).x;
      rx = velocity_bc(P, make_float3(r.x, r.y, r.z), x + 1, y, z, 1).x;
      dy = velocity_bc(P, make_float3(dn.x, dn.y, dn.z), x, y - 1, z, 1).y;
      uy = velocity_bc(P, make_float3(up.x, up.y, up.z), x, y + 1, z, 1).y;
      bz = velocity_bc(P, make_float3(prev.x, prev.y, prev.z), x, y, z - 1, 1).z;
      fz = velocity_bc(P, make_float3(next.x, next.y, next.z), x, y, z + 1, 1).z;
    }
    const size_t i = col + zo;
    if (zero & 1) p[i] = 0.f;
    if (zero & 2) tp[i] = 0.f;
    div[i] = .5f * P.rdx * (r.w * rx - l.w * lx + up.w * uy - dn.w * dy + next.w * fz - prev.w * bz);
    prev = cur;
    cur = next;
  }
}

// Two rows per thread.  ncu of the kernel above at 1024^3 (profiles/README.md r4a): 5.92 ms, DRAM traffic at its minimum
// (22.5 GB) but only 3.8 TB/s — long-scoreboard stalls, ~124 executed instructions per 32 cells at 63 % of the issue slots,
// the L1 data pipe at 63 % (three 16-byte row loads per cell; ptxas splits the y neighbours into 4-byte loads at a 16-byte
// stride).  Here a thread owns the cells (x, 2j) and (x, 2j + 1) of a plane: each is the other's y neighbour (4 row loads per
// 2 cells instead of 6) and pointer arithmetic, chunk bounds and the loop are paid once per 2 cells.  Same expression per
// cell.  What the A/B runs say (r4b-r4d, ms per divergence stage at 1024^3; one row 6.71): the kernel lives on resident
// warps — four rows at 79 registers / 3 CTAs per SM 6.34; with the column fetched two planes ahead 7.77 (108 registers,
// 2 CTAs); two rows fetched two planes ahead 5.96 (64 registers, 4 CTAs); two rows, the plane ahead fetched in the iteration
// that uses it, 48 registers / 5 CTAs: 5.86 (shipped); capped at 40 registers / 6 CTAs (48 bytes of spill) 7.1.  Fetching two
// planes ahead changed nothing for the one-row kernel either.
//
// kRestrict (one GPU, even extents, chunks of an even number of planes): the level-1 divergence — the reference's sequential
// sum over a coarse cell's 2x2x2 children (k_uniform_restrict, uniformgrid_fluid.cu:134-160: x fastest, then y, then z) — is
// accumulated on the way: the children are the lane pair (2i, 2i + 1), the thread's two rows and two consecutive iterations.
// Saves the pass that re-reads the level-0 divergence (0.83 ms for 4.8 GB at 1024^3; 0.1-0.5 ms of it arrive in the step).
constexpr int ZRY = 2;
template <bool kRestrict>
__global__ void __launch_bounds__(ZTX * ZTY, 5) k_u_divergence_zm2(KParams P, const float4 *__restrict__ vw, float *__restrict__ div, float *__restrict__ p,
                                                                   float *__restrict__ tp, int zc, int zero, ZRange Z, uint64_t off1) {
  const int x = blockIdx.x * ZTX + threadIdx.x, y0 = (blockIdx.y * ZTY + threadIdx.y) * ZRY;
  const int z0 = Z.zb + blockIdx.z * zc, z1 = min(z0 + zc, min(P.gz, Z.ze));
  const ptrdiff_t sy = P.gx, sz = (ptrdiff_t)P.gx * P.gy;
  // y neighbours outside the thread's rows, relative to its first row (clamped at the walls like the reference's indices)
  const ptrdiff_t dno = y0 > 0 ? -sy : 0, upo = y0 + ZRY < P.gy ? ZRY * sy : (ZRY - 1) * sy;
  const bool left_edge = threadIdx.x == 0, right_edge = threadIdx.x == ZTX - 1;
  // boundary conditions only substitute values of neighbours OUTSIDE the domain (sim_utils.cu:24-39)
  const bool x_inner = x > 0 && x < P.gx - 1;
  const size_t first = (size_t)z0 * sz + (size_t)y0 * P.gx + x;
  const float4 *pc = vw + first;  // (x, y0) of plane z
  float *pd = div + first;
  float4 cur[ZRY];
  float pz[ZRY], pw[ZRY];  // plane z - 1: only (vz, fluidity) are used
#pragma unroll
  for (int k = 0; k < ZRY; k++) {
    cur[k] = __ldg(pc + k * sy);
    const float4 b = z0 > 0 ? __ldg(pc + k * sy - sz) : cur[k];
    pz[k] = b.z; pw[k] = b.w;
  }
  float carry = 0.f;  // kRestrict: the sum over the coarse cell's four children of the even plane
  for (int z = z0; z < z1; z++, pc += sz, pd += sz) {
    float4 next[ZRY];
#pragma unroll
    for (int k = 0; k < ZRY; k++) next[k] = z < P.gz - 1 ? __ldg(pc + k * sy + sz) : cur[k];
    const float4 dn = __ldg(pc + dno), up = __ldg(pc + upo);
    const bool z_inner = z > 0 && z < P.gz - 1;
    float dv[ZRY];
#pragma unroll
    for (int k = 0; k < ZRY; k++) {
      const int y = y0 + k;
      const float4 c = cur[k];
      const float4 d = k == 0 ? dn : cur[0], u = k == ZRY - 1 ? up : cur[ZRY - 1];
      float4 l, r;
      l.x = __shfl_up_sync(0xFFFFFFFFu, c.x, 1); l.w = __shfl_up_sync(0xFFFFFFFFu, c.w, 1);
      r.x = __shfl_down_sync(0xFFFFFFFFu, c.x, 1); r.w = __shfl_down_sync(0xFFFFFFFFu, c.w, 1);
      l.y = l.z = r.y = r.z = 0.f;
      if (left_edge) l = x > 0 ? pc[k * sy - 1] : c;
      if (right_edge) r = x < P.gx - 1 ? pc[k * sy + 1] : c;
      float lx = l.x, rx = r.x, dy = d.y, uy = u.y, bz = pz[k], fz = next[k].z;
      if (!(x_inner && z_inner && y > 0 && y < P.gy - 1)) {
        lx = velocity_bc(P, make_float3(l.x, l.y, l.z), x - 1, y, z, 1).x;
        rx = velocity_bc(P, make_float3(r.x, r.y, r.z), x + 1, y, z, 1).x;
        dy = velocity_bc(P, make_float3(d.x, d.y, d.z), x, y - 1, z, 1).y;
        uy = velocity_bc(P, make_float3(u.x, u.y, u.z), x, y + 1, z, 1).y;
        bz = velocity_bc(P, make_float3(0.f, 0.f, pz[k]), x, y, z - 1, 1).z;
        fz = velocity_bc(P, make_float3(next[k].x, next[k].y, next[k].z), x, y, z + 1, 1).z;
      }
      if (zero) {
        const size_t i = (size_t)(pd - div) + k * sy;
        if (zero & 1) p[i] = 0.f;
        if (zero & 2) tp[i] = 0.f;
      }
      dv[k] = .5f * P.rdx * (r.w * rx - l.w * lx + u.w * uy - d.w * dy + next[k].w * fz - pw[k] * bz);
      pd[k * sy] = dv[k];
    }
    if (kRestrict) {
      const float o0 = __shfl_down_sync(0xFFFFFFFFu, dv[0], 1), o1 = __shfl_down_sync(0xFFFFFFFFu, dv[1], 1);
      if (z & 1) {
        const float s = carry + dv[0] + o0 + dv[1] + o1;
        if (!(threadIdx.x & 1)) {
          const size_t ci = off1 + ((size_t)(z >> 1) * (P.gy >> 1) + (y0 >> 1)) * (P.gx >> 1) + (x >> 1);
          p[ci] = 0.f;
          tp[ci] = 0.f;
          div[ci] = .125f * s;
        }
      } else {
        carry = dv[0] + o0 + dv[1] + o1;
      }
    }
#pragma unroll
    for (int k = 0; k < ZRY; k++) {
      pz[k] = cur[k].z; pw[k] = cur[k].w;
      cur[k] = next[k];
    }
  }
}

// k_uniform_prolongate, uniformgrid_fluid.cu:206-237
__global__ void __launch_bounds__(256) k_u_prolongate(KParams P, int level, uint64_t off, uint64_t poff, float *__restrict__ p, ZRange Z) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y, z = Z.zb + blockIdx.z * BZ + threadIdx.z;
  const int w = P.gx >> level, h = P.gy >> level, d = P.gz >> level;
  if (x >= w || y >= h || z >= min(d, Z.ze)) return;
  const uint64_t i = off + ((uint64_t)z * h + y) * w + x;
  const int64_t pw = w / 2, ph = h / 2;
  const int64_t i000 = (int64_t)poff + ((int64_t)(z / 2) * ph + y / 2) * pw + x / 2;
  const int sx = (x == 0 || x == w - 1) ? 0 : 2 * (x % 2) - 1;
  const int sy = (y == 0 || y == h - 1) ? 0 : 2 * (y % 2) - 1;
  const int sz = (z == 0 || z == d - 1) ? 0 : 2 * (z % 2) - 1;
  const int64_t ox = sx, oy = sy * pw, oz = sz * pw * ph;
  const float p000 = p[i000], p001 = p[i000 + ox], p010 = p[i000 + oy], p011 = p[i000 + oy + ox];
  const float p100 = p[i000 + oz], p101 = p[i000 + oz + ox], p110 = p[i000 + oz + oy], p111 = p[i000 + oz + oy + ox];
  p[i] = (27.f * p000 + 9.f * (p001 + p010 + p100) + 3.f * (p011 + p101 + p110) + p111) / 64.f;
}

// z-marching prolongation: 4 fine cells (one float4 store) per thread and plane.  The one-thread-per-cell kernel above
// took 4.5 ms at 1024^3 for 4.8 GB of traffic (8 coarse loads and a page of 64-bit index arithmetic per cell); here the
// four cells x = 4k .. 4k+3 share the coarse cells 2k-1 .. 2k+2 of four coarse rows.  Same expression per cell.
__global__ void __launch_bounds__(ZTX * ZTY) k_u_prolongate_zm(KParams P, int level, uint64_t off, uint64_t poff, float *__restrict__ p, int zc, ZRange Z) {
  const int w = P.gx >> level, h = P.gy >> level, d = P.gz >> level;
  const int x = 4 * (blockIdx.x * ZTX + threadIdx.x), y = blockIdx.y * ZTY + threadIdx.y;
  const int z0 = Z.zb + blockIdx.z * zc, z1 = min(z0 + zc, min(d, Z.ze));
  const int pw = w / 2, ph = h / 2;
  const int sy = (y == 0 || y == h - 1) ? 0 : 2 * (y % 2) - 1;
  const float *c = p + poff;
  // coarse columns 2k-1 .. 2k+2; the outer two are only used by cells whose x step is not suppressed at a wall
  const int X = x / 2;
  const int xm = x == 0 ? X : X - 1, xp = x + 4 >= w ? X + 1 : X + 2;
  const size_t row0 = (size_t)(y / 2) * pw, row1 = (size_t)(y / 2 + sy) * pw;
  for (int z = z0; z < z1; z++) {
    const int sz = (z == 0 || z == d - 1) ? 0 : 2 * (z % 2) - 1;
    const size_t pl0 = (size_t)(z / 2) * pw * ph, pl1 = (size_t)(z / 2 + sz) * pw * ph;
    float v[4][4];  // [row: (y0,z0) (y1,z0) (y0,z1) (y1,z1)][coarse column xm, X, X+1, xp]
    const size_t r[4] = {pl0 + row0, pl0 + row1, pl1 + row0, pl1 + row1};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float2 mid = *reinterpret_cast<const float2 *>(c + r[k] + X);
      v[k][0] = c[r[k] + xm];
      v[k][1] = mid.x;
      v[k][2] = mid.y;
      v[k][3] = c[r[k] + xp];
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int a = 1 + (j >> 1);  // own coarse column: X for j = 0, 1; X + 1 for j = 2, 3
      // x step: -1 for even x, +1 for odd x, 0 on the two wall cells (uniformgrid_fluid.cu:218)
      const int b = (j & 1) ? a + 1 : a - 1;
      const bool wall = (x + j == 0) || (x + j == w - 1);
      const float p000 = v[0][a], p010 = v[1][a], p100 = v[2][a], p110 = v[3][a];
      const float p001 = wall ? p000 : v[0][b], p011 = wall ? p010 : v[1][b], p101 = wall ? p100 : v[2][b], p111 = wall ? p110 : v[3][b];
      o[j] = (27.f * p000 + 9.f * (p001 + p010 + p100) + 3.f * (p011 + p101 + p110) + p111) / 64.f;
    }
    *reinterpret_cast<float4 *>(p + off + ((size_t)z * h + y) * w + x) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// k_uniform_apply_pressure, uniformgrid_fluid.cu:239-260
__global__ void __launch_bounds__(256) k_u_apply_pressure(KParams P, const float *__restrict__ p, const float *__restrict__ fl,
                                                          float4 *__restrict__ vw, ZRange Z) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y, z = Z.zb + blockIdx.z * BZ + threadIdx.z;
  if (x >= P.gx || y >= P.gy || z >= min(P.gz, Z.ze)) return;
  const uint64_t i = lidx(P, x, y, z);
  const uint64_t sy = P.gx, sz = (uint64_t)P.gx * P.gy;
  const uint64_t il = x > 0 ? i - 1 : i, ir = x < P.gx - 1 ? i + 1 : i;
  const uint64_t id = y > 0 ? i - sy : i, iu = y < P.gy - 1 ? i + sy : i;
  const uint64_t ib = z > 0 ? i - sz : i, iff = z < P.gz - 1 ? i + sz : i;
  const float alpha = .5f * P.rdx;
  const float pc = p[i];
  float4 v = vw[i];
  // neighbours' fluidity comes from the level-0 part of the pyramid so that the in-place float4
  // update of vw never races with a neighbour's read
  const float wl = fl[il], wr = fl[ir], wd = fl[id], wu = fl[iu], wb = fl[ib], wf = fl[iff];
  v.x -= alpha * (wr * (p[ir] - pc) + wl * (pc - p[il]));
  v.y -= alpha * (wu * (p[iu] - pc) + wd * (pc - p[id]));
  v.z -= alpha * (wf * (p[iff] - pc) + wb * (pc - p[ib]));
  vw[i] = v;
}

// z-marching pressure gradient: the column's pressure and fluidity planes in registers, x neighbours by shuffle.
// ncu (profiles/README.md r4a): 7.60 ms at 1024^3 = 5.7 TB/s of DRAM traffic, 87 % of the measured peak for the 40 B/cell it moves.
// (Measured and rejected, r4b: the front plane's fluidity taken from `.w` of that plane's packed velocity, fetched two planes
// ahead by the thread that updates it — 36 B/cell, but 59 registers instead of 40 and 9.7 ms instead of 7.7.)
__global__ void __launch_bounds__(ZTX * ZTY) k_u_apply_zm(KParams P, const float *__restrict__ p, const float *__restrict__ fl, float4 *__restrict__ vw,
                                                          int zc, ZRange Z) {
  const int x = blockIdx.x * ZTX + threadIdx.x, y = blockIdx.y * ZTY + threadIdx.y;
  const int z0 = Z.zb + blockIdx.z * zc, z1 = min(z0 + zc, min(P.gz, Z.ze));
  const size_t sy = (size_t)P.gx, sz = (size_t)P.gx * P.gy;
  const size_t col = (size_t)y * P.gx + x;
  const size_t dnrow = y > 0 ? col - sy : col, uprow = y < P.gy - 1 ? col + sy : col;
  const bool left_edge = threadIdx.x == 0, right_edge = threadIdx.x == ZTX - 1;
  const float alpha = .5f * P.rdx;
  float pc = p[col + (size_t)z0 * sz], wc = fl[col + (size_t)z0 * sz];
  float pb = pc, wb = wc;
  if (z0 > 0) { pb = p[col + (size_t)(z0 - 1) * sz]; wb = fl[col + (size_t)(z0 - 1) * sz]; }
  for (int z = z0; z < z1; z++) {
    const size_t zo = (size_t)z * sz, i = col + zo;
    float pf = pc, wf = wc;
    if (z < P.gz - 1) { pf = p[i + sz]; wf = fl[i + sz]; }
    const float pd = p[dnrow + zo], pu = p[uprow + zo], wd = fl[dnrow + zo], wu = fl[uprow + zo];
    float4 v = vw[i];
    float pl = __shfl_up_sync(0xFFFFFFFFu, pc, 1), wl = __shfl_up_sync(0xFFFFFFFFu, wc, 1);
    float pr = __shfl_down_sync(0xFFFFFFFFu, pc, 1), wr = __shfl_down_sync(0xFFFFFFFFu, wc, 1);
    if (left_edge) { pl = x > 0 ? p[i - 1] : pc; wl = x > 0 ? fl[i - 1] : wc; }
    if (right_edge) { pr = x < P.gx - 1 ? p[i + 1] : pc; wr = x < P.gx - 1 ? fl[i + 1] : wc; }
    v.x -= alpha * (wr * (pr - pc) + wl * (pc - pl));
    v.y -= alpha * (wu * (pu - pc) + wd * (pc - pd));
    v.z -= alpha * (wf * (pf - pc) + wb * (pc - pb));
    vw[i] = v;
    pb = pc; wb = wc;
    pc = pf; wc = wf;
  }
}

// k_uniform_debug_stats, uniformgrid_structure.cu:33-43: per-256-cell bins, sequential sums
__global__ void k_u_debug_stats(const float *__restrict__ q, const float4 *__restrict__ vw, float *__restrict__ stats, uint64_t bins) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bins) return;
  float s = 0.f;
  for (uint64_t i = b * 256; i < b * 256 + 256; i++) s += q[i] * vw[i].w;
  stats[b] = s;
}

// deterministic total (double accumulation, fixed tree): dcg_total_density
__global__ void __launch_bounds__(256) k_u_total_density(const float *__restrict__ q, const float4 *__restrict__ vw, uint64_t n,
                                                         double *__restrict__ partial) {
  __shared__ double sh[256];
  double s = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256)
    s += (double)(q[i] * vw[i].w);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void k_unpack_velocity(const float4 *__restrict__ vw, float *__restrict__ out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = vw[i];
  out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
}

uint64_t pyramid_cells(uint64_t w, uint64_t h, uint64_t d) {  // grid_math.cuh:24-30
  uint64_t n = 0;
  for (uint64_t s = 1; w % s == 0 && h % s == 0 && d % s == 0; s *= 2) n += (w * h * d) / (s * s * s);
  return n;
}
uint64_t pyramid_offset(uint64_t w, uint64_t h, uint64_t d, uint64_t scale) {  // grid_math.cuh:46-48
  uint64_t off = 0;
  for (uint64_t s = 1; s < scale; s *= 2) off += (w * h * d) / (s * s * s);
  return off;
}

// ---- slab decomposition ------------------------------------------------------------------------------
// The dense solver sharded over `world` ranks by contiguous z-slabs runs THE SAME kernels: every field lives in one
// virtual address range that looks the same on every GPU (one process per GPU: stitched from physical pieces the ranks
// allocate on their own devices, shard_vmm.h; all ranks in one process: plain allocations), a launch covers the planes
// [zb, ze) of its rank (ZRange), and a neighbour plane, a backtraced gather corner (CFL ~ 23-46 cells: no fixed-width
// halo would do) or a coarse parent cell that another rank owns is read in place over NVLink.  A coarse plane belongs
// to the owner of its first fine plane.  Ranks meet at a flag barrier over peer memory before every stage.
constexpr int kMaxShardRanks = 8;
struct ShardFlags {
  volatile uint32_t *flags[kMaxShardRanks];  // flags[r][j] = the last epoch rank j announced to rank r
};
__device__ __forceinline__ unsigned long long shard_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// epoch in device memory: the launch is identical every time (graph-capturable)
__global__ void __launch_bounds__(32) k_u_barrier(ShardFlags B, int rank, int world, uint32_t *epoch_counter, uint32_t *err) {
  const int t = threadIdx.x;
  uint32_t epoch = 0;
  if (t == 0) epoch = ++(*epoch_counter);
  epoch = __shfl_sync(0xFFFFFFFFu, epoch, 0);
  if (t < world && t != rank) {
    __threadfence_system();       // everything this rank wrote before the barrier is visible system-wide
    B.flags[t][rank] = epoch;     // 4-byte store into the peer's control block over NVLink
    const unsigned long long t0 = shard_global_ns();
    while ((int32_t)(B.flags[rank][t] - epoch) < 0) {
      if (shard_global_ns() - t0 > 10000000000ull) {  // 10 s: a rank that never arrives trips *err instead of hanging the GPU
        *err = 1;
        break;
      }
    }
    __threadfence_system();
  }
}

struct UniformSim : dcg_sim {
  int gx = 0, gy = 0, gz = 0, mip_levels = 1;
  uint64_t N = 0, pyr = 0;
  float4 *vw[2] = {nullptr, nullptr};
  float *q[2] = {nullptr, nullptr};
  float *fluidity = nullptr, *p = nullptr, *tp = nullptr, *div = nullptr;
  float *scratch = nullptr;       // 3N floats: accessor staging / stats bins
  double *d_partial = nullptr;    // total-density partials
  double *h_partial = nullptr;    // pinned
  int cur_v = 0, cur_q = 0;
  bool fluidity_dirty = true;
  // k_u_advect_both: advect_density() also produces the NEXT step's advected velocity in vw[cur_v ^ 1];
  // spec_velocity = that buffer is valid (nothing has touched velocity or parameters since)
  bool fuse_advect = true, spec_velocity = false;
  std::vector<uint64_t> level_off;

  // slab decomposition: ranks [rank0, rank0 + nlocal) of `world`; world == 1: the plain single-GPU solver
  int world = 1, rank0 = 0, nlocal = 1, slab = 0;
  bool vmm = false, ready = true;  // vmm: one rank per process, fields stitched from every rank's physical pieces
  vmm::Driver drv;
  vmm::FdServer fd_server;
  static constexpr int kFields = 8;  // vw0 vw1 q0 q1 fluidity p tp div
  size_t gran = 0, field_bytes[kFields] = {0}, field_va_bytes[kFields] = {0};
  CUdeviceptr field_va[kFields] = {0}, ctrl_va = 0;
  std::vector<std::vector<CUmemGenericAllocationHandle>> pieces;  // [rank]: control block, then its non-empty field pieces in field order
  ShardFlags peers{};
  uint32_t *d_epoch = nullptr, *d_barrier_err = nullptr;
  uint64_t n_barriers = 0;

  // CUDA graph of one full step (advectVelocity, adaptTopology, project, advectDensity)
  cudaGraphExec_t step_graph[4] = {nullptr, nullptr, nullptr, nullptr};  // one per (cur_v, cur_q) start state
  uint64_t step_graph_launches = 0, step_graph_barriers = 0;

  ~UniformSim() override {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    drop_graphs();
    if (vmm) {
      fd_server.finish();
      release_vmm();
    } else {
      for (int i = 0; i < 2; i++) { cudaFree(vw[i]); cudaFree(q[i]); }
      cudaFree(fluidity); cudaFree(p); cudaFree(tp); cudaFree(div);
    }
    cudaFree(scratch); cudaFree(d_partial); cudaFree(d_epoch); cudaFree(d_barrier_err);
    if (h_partial) cudaFreeHost(h_partial);
    base_teardown();
  }
  void drop_graphs() {
    for (auto &g : step_graph) {
      if (g) cudaGraphExecDestroy(g);
      g = nullptr;
    }
  }

  void invalidate_graphs() override { drop_graphs(); }

  int construct(const dcg_sim_params *prm, int dev) override { return construct_sharded(prm, dev, 0, 1, 1); }

  // rank / wsize / nloc: this instance holds ranks [rank, rank + nloc) of a wsize-way z-slab decomposition
  // (nloc == wsize: all of them, on one device, in stream order; nloc == 1 < wsize: one rank per process)
  int construct_sharded(const dcg_sim_params *prm, int dev, int rank, int wsize, int nloc) {
    DCG_TRY(base_setup(prm, dev));
    project_coarsest_pairs = 2; project_level_pairs = 1; local_pairs = 5;  // fluid_simulation_uniform.cu:103,116,129
    gx = prm->gx; gy = prm->gy; gz = prm->gz;
    world = wsize; rank0 = rank; nlocal = nloc;
    if (gx <= 0 || gy <= 0 || gz <= 0) return fail(DCG_ERR_INVALID, "grid size must be positive");
    if (world < 1 || world > kMaxShardRanks) return fail(DCG_ERR_INVALID, "world size must be in [1, %d]", kMaxShardRanks);
    if (nlocal != 1 && nlocal != world) return fail(DCG_ERR_INVALID, "nlocal must be 1 (one rank per process) or world (all ranks in this process)");
    if (rank0 < 0 || rank0 + nlocal > world) return fail(DCG_ERR_INVALID, "rank out of range");
    if (gz % world != 0) return fail(DCG_ERR_INVALID, "gz (%d) must be a multiple of the world size (%d): slabs are whole z-planes of equal count", gz, world);
    if (((uint64_t)gx * gy) % 256 != 0 && world > 1) return fail(DCG_ERR_INVALID, "gx*gy must be a multiple of 256 (debugStats bins must not straddle slabs)");
    slab = gz / world;
    vmm = nlocal != world;
    N = (uint64_t)gx * gy * gz;
    // mipmapLevels rule, fluid_simulation_uniform.cu:8-17
    uint64_t min_dim = (gx < gy && gx < gz) ? gx : (gy < gz ? gy : gz);
    uint64_t cell = 2;
    mip_levels = 1;
    while (gx % cell == 0 && gy % cell == 0 && gz % cell == 0 && cell * 4 <= min_dim) {
      mip_levels++;
      cell *= 2;
    }
    pyr = pyramid_cells(gx, gy, gz);
    level_off.resize(mip_levels);
    for (int l = 0; l < mip_levels; l++) level_off[l] = pyramid_offset(gx, gy, gz, 1ull << l);
    for (int f = 0; f < kFields; f++) field_bytes[f] = f < 2 ? N * sizeof(float4) : (f < 4 ? N * sizeof(float) : pyr * sizeof(float));
    DCG_CUDA_TRY(cudaMalloc(&scratch, 3 * local_cells() * sizeof(float)));
    DCG_CUDA_TRY(cudaMalloc(&d_partial, 1024 * sizeof(double)));
    DCG_CUDA_TRY(cudaMallocHost(&h_partial, 1024 * sizeof(double)));
    fuse_advect = !opt.advect_no_fuse;
    if (vmm) return create_arena();  // the caller exchanges handles, then import_handles() maps the peers and resets
    for (int i = 0; i < 2; i++) {
      DCG_CUDA_TRY(cudaMalloc(&vw[i], field_bytes[i]));
      DCG_CUDA_TRY(cudaMalloc(&q[i], field_bytes[2 + i]));
    }
    DCG_CUDA_TRY(cudaMalloc(&fluidity, field_bytes[4]));
    DCG_CUDA_TRY(cudaMalloc(&p, field_bytes[5]));
    DCG_CUDA_TRY(cudaMalloc(&tp, field_bytes[6]));
    DCG_CUDA_TRY(cudaMalloc(&div, field_bytes[7]));
    return reset();
  }

  // ---- one rank per process: the fields as one virtual range per field, stitched from every rank's pieces ----
  // Who holds a byte decides where it lives, not who computes it — but a remote byte costs ~8x a local one, so the physical
  // pieces follow the ownership of the cells: every allocation granule (2 MiB) of a field's byte range belongs to the rank that
  // owns the cell in its middle (level-0 arrays: its z-slab; pyramids: the slab of whichever level the granule lies in), and a
  // rank allocates one piece per maximal run of its granules.
  struct PieceRun { size_t b0, b1; int owner; };
  std::vector<PieceRun> piece_runs[kFields];
  int owner_of_byte(int f, size_t byte) const {
    if (f < 4) {
      const uint64_t cell = std::min<uint64_t>(N - 1, byte / (f < 2 ? sizeof(float4) : sizeof(float)));
      return (int)std::min<uint64_t>(world - 1, cell / ((uint64_t)gx * gy) / slab);
    }
    const uint64_t cell = std::min<uint64_t>(pyr - 1, byte / sizeof(float));
    int l = 0;
    while (l + 1 < mip_levels && cell >= level_off[l + 1]) l++;
    if (cell >= level_off[l] + (uint64_t)(gx >> l) * (gy >> l) * (gz >> l)) return world - 1;  // levels beyond mip_levels (never touched)
    const uint64_t z = (cell - level_off[l]) / ((uint64_t)(gx >> l) * (gy >> l));
    return (int)std::min<uint64_t>(world - 1, (z << l) / slab);
  }
  void build_piece_runs() {
    for (int f = 0; f < kFields; f++) {
      piece_runs[f].clear();
      const size_t ngran = (field_bytes[f] + gran - 1) / gran;
      for (size_t g = 0; g < ngran; g++) {
        const int o = owner_of_byte(f, g * gran + gran / 2);
        if (!piece_runs[f].empty() && piece_runs[f].back().owner == o) piece_runs[f].back().b1 = (g + 1) * gran;
        else piece_runs[f].push_back({g * gran, (g + 1) * gran, o});
      }
    }
  }
  CUmemAccessDesc access_desc() const {
    CUmemAccessDesc acc;
    std::memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    return acc;
  }
  int map_piece(CUdeviceptr va, size_t bytes, CUmemGenericAllocationHandle h) {
    if (drv.memMap(va, bytes, 0, h, 0) != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemMap failed (%zu bytes)", bytes);
    CUmemAccessDesc acc = access_desc();
    if (drv.memSetAccess(va, bytes, &acc, 1) != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemSetAccess failed: peer access between the GPUs is required");
    return DCG_OK;
  }
  int create_arena() {
    if (!drv.load()) return fail(DCG_ERR_CUDA, "CUDA virtual memory management entry points are not available");
    CUmemAllocationProp prop = vmm::device_prop(device);
    if (drv.memGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return fail(DCG_ERR_CUDA, "cuMemGetAllocationGranularity failed");
    pieces.assign(world, {});
    std::vector<int> fds;
    auto create = [&](size_t bytes) -> int {
      CUmemGenericAllocationHandle h;
      if (drv.memCreate(&h, bytes, &prop, 0) != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemCreate(%zu bytes) failed", bytes);
      pieces[rank0].push_back(h);
      int fd = -1;
      if (drv.memExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemExportToShareableHandle failed");
      fds.push_back(fd);
      return DCG_OK;
    };
    auto prepare = [&]() -> int {
      DCG_TRY(create(gran));  // control block (barrier flags)
      build_piece_runs();
      for (int f = 0; f < kFields; f++)
        for (const PieceRun &r : piece_runs[f])
          if (r.owner == rank0) DCG_TRY(create(r.b1 - r.b0));
      // this rank's control block is mapped and cleared BEFORE the pieces are published: a peer may announce its
      // first barrier epoch as soon as it has imported them
      if (drv.memReserve(&ctrl_va, (size_t)world * gran, gran, 0, 0) != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemAddressReserve failed");
      DCG_TRY(map_piece(ctrl_va + (size_t)rank0 * gran, gran, pieces[rank0][0]));
      DCG_CUDA_TRY(cudaMemset(reinterpret_cast<void *>(ctrl_va + (size_t)rank0 * gran), 0, 4096));
      DCG_CUDA_TRY(cudaDeviceSynchronize());
      return DCG_OK;
    };
    const int rc = prepare();
    if (rc != DCG_OK) {
      for (int fd : fds) close(fd);
      return rc;
    }
    if (!fd_server.start(fds, world - 1)) return fail(DCG_ERR_CUDA, "cannot open the descriptor socket");
    ready = false;
    return DCG_OK;
  }
  int export_handle(void *out, uint64_t cap) override {
    if (!vmm) return fail(DCG_ERR_INVALID, "not a one-rank-per-process instance");
    if (cap < vmm::kHandleBytes) return fail(DCG_ERR_INVALID, "handle buffer too small");
    std::memcpy(out, fd_server.name, vmm::kHandleBytes);
    return DCG_OK;
  }
  int import_handles(const void *handles, int count) override {
    if (!vmm) return fail(DCG_ERR_INVALID, "all ranks are local: nothing to import");
    if (count != world) return fail(DCG_ERR_INVALID, "expected %d handles, got %d", world, count);
    DCG_CUDA_TRY(cudaSetDevice(device));
    for (int r = 0; r < world; r++) {
      if (r == rank0) continue;
      int npieces = 1;
      for (int f = 0; f < kFields; f++)
        for (const PieceRun &pr : piece_runs[f]) npieces += pr.owner == r ? 1 : 0;
      char name[vmm::kHandleBytes + 1] = {0};
      std::memcpy(name, static_cast<const char *>(handles) + (size_t)r * vmm::kHandleBytes, vmm::kHandleBytes);
      std::vector<int> fds;
      if (!vmm::fetch_fds(name, npieces, fds)) return fail(DCG_ERR_CUDA, "could not fetch the memory descriptors of rank %d", r);
      for (int fd : fds) {
        CUmemGenericAllocationHandle h;
        const CUresult rc = drv.memImport(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(fd);
        if (rc != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemImportFromShareableHandle failed for rank %d", r);
        pieces[r].push_back(h);
      }
    }
    fd_server.finish_after_serving(world - 1, 120000);
    if (fd_server.served.load() != world - 1) return fail(DCG_ERR_CUDA, "only %d of %d peers fetched this rank's memory", fd_server.served.load(), world - 1);
    for (int r = 0; r < world; r++) {
      if (r != rank0) DCG_TRY(map_piece(ctrl_va + (size_t)r * gran, gran, pieces[r][0]));
      peers.flags[r] = reinterpret_cast<volatile uint32_t *>(ctrl_va + (size_t)r * gran);
    }
    std::vector<size_t> next(world, 1);  // pieces[r][0] = control block, then the rank's runs in (field, run) order
    for (int f = 0; f < kFields; f++) {
      field_va_bytes[f] = (field_bytes[f] + gran - 1) / gran * gran;
      if (drv.memReserve(&field_va[f], field_va_bytes[f], gran, 0, 0) != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemAddressReserve failed");
      for (const PieceRun &pr : piece_runs[f]) DCG_TRY(map_piece(field_va[f] + pr.b0, pr.b1 - pr.b0, pieces[pr.owner][next[pr.owner]++]));
    }
    vw[0] = reinterpret_cast<float4 *>(field_va[0]); vw[1] = reinterpret_cast<float4 *>(field_va[1]);
    q[0] = reinterpret_cast<float *>(field_va[2]); q[1] = reinterpret_cast<float *>(field_va[3]);
    fluidity = reinterpret_cast<float *>(field_va[4]); p = reinterpret_cast<float *>(field_va[5]);
    tp = reinterpret_cast<float *>(field_va[6]); div = reinterpret_cast<float *>(field_va[7]);
    DCG_CUDA_TRY(cudaMalloc(&d_epoch, 4));
    DCG_CUDA_TRY(cudaMalloc(&d_barrier_err, 4));
    DCG_CUDA_TRY(cudaMemset(d_epoch, 0, 4));
    DCG_CUDA_TRY(cudaMemset(d_barrier_err, 0, 4));
    DCG_CUDA_TRY(cudaDeviceSynchronize());
    ready = true;
    DCG_TRY(reset());  // starts with a barrier: every rank has mapped every piece before any field is touched
    return check_barrier_error();
  }
  void release_vmm() {
    if (ctrl_va) { drv.memUnmap(ctrl_va, (size_t)world * gran); drv.memFree(ctrl_va, (size_t)world * gran); }
    for (int f = 0; f < kFields; f++)
      if (field_va[f]) { drv.memUnmap(field_va[f], field_va_bytes[f]); drv.memFree(field_va[f], field_va_bytes[f]); }
    for (auto &v : pieces)
      for (auto h : v) drv.memRelease(h);
  }
  int enter() {
    DCG_CUDA_TRY(cudaSetDevice(device));
    return ready ? DCG_OK : fail(DCG_ERR_INVALID, "sharded instance not finalized: call dcg_shard_import_handles first");
  }
  int check_barrier_error() {
    if (!vmm || !d_barrier_err) return DCG_OK;  // (before import_handles there is no barrier yet)
    uint32_t e = 0;
    DCG_CUDA_TRY(cudaMemcpyAsync(&e, d_barrier_err, 4, cudaMemcpyDeviceToHost, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    if (e) return fail(DCG_ERR_CUDA, "shard barrier timed out: a peer rank never arrived (ranks must issue identical call sequences)");
    return DCG_OK;
  }
  int health_check() override { return check_barrier_error(); }

  // ---- the planes of a rank, stages ----
  uint64_t local_cells() const { return (uint64_t)gx * gy * (uint64_t)(gz / world) * nlocal; }
  uint64_t local_first() const { return (uint64_t)gx * gy * (uint64_t)(gz / world) * rank0; }
  // first plane of mip level `level` owned by rank r: a coarse plane belongs to the owner of its first fine plane
  int zfirst(int level, int r) const { return std::min(gz >> level, (r * slab + (1 << level) - 1) >> level); }
  // every rank has finished everything it launched so far, and its writes are visible to all
  void barrier() {
    if (!vmm) return;  // all ranks share this stream: program order is the barrier
    k_u_barrier<<<1, 32, 0, stream>>>(peers, rank0, world, d_epoch, d_barrier_err);
    launches++;
    n_barriers++;
  }
  // one lock-step stage: [barrier,] launch(Z, planes) for every local rank that owns planes of `level`
  template <typename F>
  void stage(int level, F &&launch) {
    for (int lr = 0; lr < nlocal; lr++) {
      barrier();
      const ZRange Z{zfirst(level, rank0 + lr), zfirst(level, rank0 + lr + 1)};
      if (Z.ze > Z.zb) {
        launch(Z, Z.ze - Z.zb);
        launches++;
      }
    }
  }
  dim3 grid_for(int level, int planes) const { return dim3(idiv_up(gx >> level, BX), idiv_up(gy >> level, BY), idiv_up(planes, BZ)); }
  static dim3 block() { return dim3(BX, BY, BZ); }

  int on_params_changed() override {
    if (params.gx != gx || params.gy != gy || params.gz != gz)
      return fail(DCG_ERR_INVALID, "grid size is fixed at construction (the reference sizes its buffers in the ctor)");
    fluidity_dirty = true;
    spec_velocity = false;
    drop_graphs();
    return DCG_OK;
  }

  // clears this instance's planes of a field (level 0 of `bytes_per_cell`-sized cells, or the whole pyramid share)
  int clear_level0(void *base, size_t bytes_per_cell) {
    DCG_CUDA_TRY(cudaMemsetAsync(static_cast<char *>(base) + local_first() * bytes_per_cell, 0, local_cells() * bytes_per_cell, stream));
    return DCG_OK;
  }
  int clear_pyramid(float *base) {
    for (int l = 0; l < mip_levels; l++) {
      const uint64_t plane = (uint64_t)(gx >> l) * (gy >> l);
      const int z0 = zfirst(l, rank0), z1 = zfirst(l, rank0 + nlocal);
      if (z1 > z0) DCG_CUDA_TRY(cudaMemsetAsync(base + level_off[l] + (uint64_t)z0 * plane, 0, (uint64_t)(z1 - z0) * plane * sizeof(float), stream));
    }
    return DCG_OK;
  }
  int reset() override {  // fluid_simulation_uniform.cu:81-88
    DCG_TRY(enter());
    barrier();  // nobody is still reading the fields we are about to clear
    DCG_TRY(clear_pyramid(p));
    DCG_TRY(clear_pyramid(tp));
    DCG_TRY(clear_pyramid(div));
    DCG_TRY(clear_pyramid(fluidity));
    cur_v = cur_q = 0;
    spec_velocity = false;
    return init();
  }
  int init() override {  // fluid_simulation_uniform.cu:76-79: adaptTopology + k_uniform_init (zero density, velocity)
    DCG_TRY(enter());
    for (int i = 0; i < 2; i++) {
      DCG_TRY(clear_level0(vw[i], sizeof(float4)));
      DCG_TRY(clear_level0(q[i], sizeof(float)));
    }
    fluidity_dirty = true;  // the memset cleared the packed .w lanes
    spec_velocity = false;
    return adapt_topology();
  }

  int adapt_topology() override {  // fluid_simulation_uniform.cu:143-147
    DCG_TRY(enter());
    if (!fluidity_dirty) return DCG_OK;  // static field: identical values every step in the reference
    for (int l = 0; l < mip_levels; l++)
      stage(l, [&](ZRange Z, int planes) { k_u_fluidity<<<grid_for(l, planes), block(), 0, stream>>>(kp, fluidity, level_off[l], l, vw[0], vw[1], Z); });
    DCG_CUDA_TRY(cudaGetLastError());
    fluidity_dirty = false;
    return DCG_OK;
  }

  int advect_velocity() override {  // fluid_simulation_uniform.cu:90-94
    DCG_TRY(enter());
    if (spec_velocity) {
      spec_velocity = false;  // vw[cur_v ^ 1] already holds this step's advected velocity (k_u_advect_both)
    } else {
      stage(0, [&](ZRange Z, int planes) { k_u_advect_velocity<<<grid_for(0, planes), block(), 0, stream>>>(kp, vw[cur_v], vw[cur_v ^ 1], Z); });
    }
    cur_v ^= 1;
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  int advect_density() override {  // fluid_simulation_uniform.cu:137-141
    DCG_TRY(enter());
    if (fuse_advect && !opt.advect && gz >= 64) {
      const int zc = opt.advect_ctas_per_sm > 0 ? opt.advect_ctas_per_sm : 16;  // (option reused: planes per thread)
      stage(0, [&](ZRange Z, int planes) {
        const dim3 g(idiv_up(gx, 32), idiv_up(gy, 8), idiv_up(planes, zc));
        k_u_advect_both_zm<<<g, dim3(32, 8), 0, stream>>>(kp, vw[cur_v], vw[cur_v ^ 1], q[cur_q], q[cur_q ^ 1], zc, Z);
      });
      spec_velocity = true;
    } else if (fuse_advect) {
      stage(0, [&](ZRange Z, int planes) { k_u_advect_both<<<grid_for(0, planes), block(), 0, stream>>>(kp, vw[cur_v], vw[cur_v ^ 1], q[cur_q], q[cur_q ^ 1], Z); });
      spec_velocity = true;
    } else {
      stage(0, [&](ZRange Z, int planes) { k_u_advect_density<<<grid_for(0, planes), block(), 0, stream>>>(kp, vw[cur_v], q[cur_q], q[cur_q ^ 1], Z); });
    }
    cur_q ^= 1;
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  // z-marching kernels: the level's x extent in whole warps of float4 (128 cells), y in tiles of 8, 16-byte aligned rows
  bool zm_ok(int l) const {
    const int w = gx >> l, h = gy >> l;
    return !opt.stencil && w % (4 * ZTX) == 0 && h % ZTY == 0 && level_off[l] % 4 == 0;
  }
  static int zm_chunk(int d) { return d >= 256 ? 32 : (d >= 32 ? 16 : d); }
  void jacobi_sweep(int l, const float *in, float *out) {
    if (zm_ok(l)) {
      const int zc = zm_chunk(gz >> l);
      stage(l, [&](ZRange Z, int planes) {
        k_u_jacobi_zm<<<dim3((gx >> l) / (4 * ZTX), (gy >> l) / ZTY, idiv_up(planes, zc)), dim3(ZTX, ZTY), 0, stream>>>(kp, l, level_off[l], in, out, div, zc, Z);
      });
    } else {
      stage(l, [&](ZRange Z, int planes) { k_u_jacobi<<<grid_for(l, planes), block(), 0, stream>>>(kp, l, level_off[l], in, out, div, Z); });
    }
  }
  void jacobi_pair(int l) {
    jacobi_sweep(l, p, tp);
    jacobi_sweep(l, tp, p);
  }
  void launch_apply() {
    if (!opt.stencil && gx % ZTX == 0 && gy % ZTY == 0) {
      const int zc = zm_chunk(gz);
      stage(0, [&](ZRange Z, int planes) {
        k_u_apply_zm<<<dim3(gx / ZTX, gy / ZTY, idiv_up(planes, zc)), dim3(ZTX, ZTY), 0, stream>>>(kp, p, fluidity, vw[cur_v], zc, Z);
      });
    } else {
      stage(0, [&](ZRange Z, int planes) { k_u_apply_pressure<<<grid_for(0, planes), block(), 0, stream>>>(kp, p, fluidity, vw[cur_v], Z); });
    }
  }
  bool fused_restrict1 = false;  // set by launch_divergence when the level-1 divergence has been written along the way
  void launch_divergence(int zero = 3, bool restrict1 = false) {
    fused_restrict1 = false;
    if (opt.zero_all) zero = 3;
    if (!opt.stencil && gx % ZTX == 0 && gy % ZTY == 0) {
      const int zc = zm_chunk(gz);
      stage(0, [&](ZRange Z, int planes) {
        const dim3 g2(gx / ZTX, gy / (ZTY * ZRY), idiv_up(planes, zc)), b(ZTX, ZTY);
        const int e = opt.experiment;  // tests / A/B: 64 = the one-row kernel, 1024 = no fused level-1 restriction
        if (gy % (ZTY * ZRY) == 0 && !(e & 64)) {
          // the level-1 restriction rides along on one GPU (coarse cells never straddle a chunk: even chunk sizes, even extents)
          if (restrict1 && world == 1 && !(e & 1024) && gx % 2 == 0 && gz % 2 == 0 && zc % 2 == 0) {
            k_u_divergence_zm2<true><<<g2, b, 0, stream>>>(kp, vw[cur_v], div, p, tp, zc, zero, Z, level_off[1]);
            fused_restrict1 = true;
          } else {
            k_u_divergence_zm2<false><<<g2, b, 0, stream>>>(kp, vw[cur_v], div, p, tp, zc, zero, Z, 0);
          }
        } else
          k_u_divergence_zm<<<dim3(gx / ZTX, gy / ZTY, idiv_up(planes, zc)), dim3(ZTX, ZTY), 0, stream>>>(kp, vw[cur_v], div, p, tp, zc, zero, Z);
      });
    } else {
      stage(0, [&](ZRange Z, int planes) { k_u_divergence<<<grid_for(0, planes), block(), 0, stream>>>(kp, vw[cur_v], div, p, tp, zero, Z); });
    }
  }
  int project() override {  // fluid_simulation_uniform.cu:96-124
    DCG_TRY(enter());
    spec_velocity = false;
    launch_divergence(mip_levels > 1 && project_level_pairs >= 1 ? 0 : 3, mip_levels > 1);
    for (int l = fused_restrict1 ? 2 : 1; l < mip_levels; l++)
      stage(l, [&](ZRange Z, int planes) { k_u_restrict<<<grid_for(l, planes), block(), 0, stream>>>(kp, l, level_off[l], level_off[l - 1], div, p, tp, Z); });
    for (int i = 0; i < project_coarsest_pairs; i++) jacobi_pair(mip_levels - 1);
    for (int l = mip_levels - 2; l >= 0; l--) {
      if (zm_ok(l) && level_off[l + 1] % 2 == 0) {
        const int zc = zm_chunk(gz >> l);
        stage(l, [&](ZRange Z, int planes) {
          k_u_prolongate_zm<<<dim3((gx >> l) / (4 * ZTX), (gy >> l) / ZTY, idiv_up(planes, zc)), dim3(ZTX, ZTY), 0, stream>>>(kp, l, level_off[l], level_off[l + 1], p, zc, Z);
        });
      } else {
        stage(l, [&](ZRange Z, int planes) { k_u_prolongate<<<grid_for(l, planes), block(), 0, stream>>>(kp, l, level_off[l], level_off[l + 1], p, Z); });
      }
      for (int i = 0; i < project_level_pairs; i++) jacobi_pair(l);
    }
    launch_apply();
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  int project_local() override {  // fluid_simulation_uniform.cu:126-135
    DCG_TRY(enter());
    spec_velocity = false;
    launch_divergence();
    for (int i = 0; i < local_pairs; i++) jacobi_pair(0);
    launch_apply();
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }

  // One full step = src/simulation.cpp:104-111, captured once per ping-pong parity as a CUDA graph
  // (29 launches at 64^3 become one graph launch) and replayed.
  int step(int n) override {
    DCG_TRY(enter());
    DCG_TRY(adapt_topology());  // flush a pending fluidity rebuild outside the graph
    DCG_CUDA_TRY(cudaEventRecord(ev_begin, stream));
    for (int done = 0; done < n; done++) {
      if (spec_velocity != fuse_advect) {  // the graphs are captured in the fused steady state only
        DCG_TRY(dcg_sim::step(1));
        continue;
      }
      cudaGraphExec_t &ge = step_graph[cur_v * 2 + cur_q];
      if (!ge) {
        const uint64_t before = launches, barriers_before = n_barriers;
        const int sv = cur_v, sq = cur_q;
        const bool sspec = spec_velocity;
        cudaGraph_t g = nullptr;
        DCG_CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        const int rc = dcg_sim::step(1);
        const cudaError_t ce = cudaStreamEndCapture(stream, &g);
        cur_v = sv; cur_q = sq; spec_velocity = sspec;  // capture records, it does not execute
        step_graph_launches = launches - before;
        step_graph_barriers = n_barriers - barriers_before;
        launches = before;
        n_barriers = barriers_before;
        if (rc != DCG_OK) return rc;
        DCG_CUDA_TRY(ce);
        DCG_CUDA_TRY(cudaGraphInstantiate(&ge, g, 0));
        cudaGraphDestroy(g);
      }
      DCG_CUDA_TRY(cudaGraphLaunch(ge, stream));
      launches += step_graph_launches;
      n_barriers += step_graph_barriers;
      cur_v ^= 1;
      cur_q ^= 1;
    }
    DCG_CUDA_TRY(cudaEventRecord(ev_end, stream));
    step_timing_pending = true;
    return DCG_OK;
  }

  // sharded: local partial results (this instance's slabs); the host side (bench / tests) combines the ranks
  int debug_stats(float *out) override {  // fluid_simulation_uniform.cu:160-176 (host sums the bins in order)
    DCG_TRY(enter());
    const uint64_t bins = local_cells() / 256, first = local_first();
    if (bins == 0) { *out = 0.f; return DCG_OK; }
    barrier();
    k_u_debug_stats<<<(unsigned)((bins + 255) / 256), 256, 0, stream>>>(q[cur_q] + first, vw[cur_v] + first, scratch, bins);
    launches++;
    std::vector<float> h(bins);
    DCG_CUDA_TRY(cudaMemcpyAsync(h.data(), scratch, bins * sizeof(float), cudaMemcpyDeviceToHost, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    float sum = 0.f;
    for (uint64_t i = 0; i < bins; i++) sum += h[i];
    *out = sum;
    return check_barrier_error();
  }

  int total_density(double *out) override {
    DCG_TRY(enter());
    const uint64_t n = local_cells(), first = local_first();
    const int blocks = (int)std::min<uint64_t>(1024, (n + 255) / 256);
    barrier();
    k_u_total_density<<<blocks, 256, 0, stream>>>(q[cur_q] + first, vw[cur_v] + first, n, d_partial);
    launches++;
    DCG_CUDA_TRY(cudaMemcpyAsync(h_partial, d_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, stream));
    DCG_TRY(synchronize());
    double s = 0.0;
    for (int i = 0; i < blocks; i++) s += h_partial[i];
    *out = s;
    return check_barrier_error();
  }

  uint64_t num_cells() const override { return local_cells(); }  // cells held by this instance
  int num_levels() const override { return mip_levels; }

  // one launch of a single stage, `reps` times, CUDA-event timed on the instance's stream
  int bench_stage(const char *stage_name, int level, int reps, float *ms_per_launch, double *alg_bytes) override {
    DCG_TRY(enter());
    if (world > 1) return fail(DCG_ERR_UNSUPPORTED, "bench_stage: not available on sharded instances");
    const std::string st(stage_name);
    if (level < 0 || level >= mip_levels) return fail(DCG_ERR_INVALID, "bench_stage: bad level");
    const double n0 = (double)N, nl = (double)((uint64_t)(gx >> level) * (gy >> level) * (gz >> level));
    double bytes = 0;
    DCG_CUDA_TRY(cudaEventRecord(ev_begin, stream));
    for (int r = 0; r < reps; r++) {
      if (st == "jacobi") {
        jacobi_sweep(level, (r & 1) ? tp : p, (r & 1) ? p : tp);
        bytes = 12.0 * nl;
      } else if (st == "advect_velocity") { spec_velocity = false; DCG_TRY(advect_velocity()); bytes = 28.0 * n0; }
      else if (st == "advect_density") {
        const bool saved = fuse_advect;
        fuse_advect = false;
        DCG_TRY(advect_density());
        fuse_advect = saved;
        bytes = 24.0 * n0;
      } else if (st == "advect_both") { DCG_TRY(advect_density()); spec_velocity = false; bytes = 52.0 * n0; }
      else if (st == "divergence") {
        launch_divergence();
        bytes = 28.0 * n0;
      } else if (st == "apply_pressure") {
        launch_apply();
        bytes = 32.0 * n0;
      } else return fail(DCG_ERR_INVALID, "bench_stage: unknown stage %s", stage_name);
    }
    DCG_CUDA_TRY(cudaEventRecord(ev_end, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    float ms = 0.f;
    DCG_CUDA_TRY(cudaEventElapsedTime(&ms, ev_begin, ev_end));
    step_timing_pending = false;
    if (ms_per_launch) *ms_per_launch = ms / reps;
    if (alg_bytes) *alg_bytes = bytes;
    return DCG_OK;
  }

  // sharded: this instance's slab(s), ranks in order — with nlocal == world the whole field in the reference's memory
  // order (a z-slab is a contiguous index range).  Pyramid fields return their level-0 part.
  int get_field(int field, int layout, float *dst, uint64_t count) override {
    DCG_TRY(enter());
    (void)layout;  // NATIVE == DENSE_L0 on the uniform grid
    const uint64_t n = local_cells(), first = local_first();
    const uint64_t need = field == DCG_FIELD_VELOCITY ? 3 * n : n;
    if (!dst || count < need) return fail(DCG_ERR_INVALID, "get_field: destination too small (%llu < %llu)", (unsigned long long)count, (unsigned long long)need);
    const float *src = nullptr;
    barrier();
    switch (field) {
      case DCG_FIELD_DENSITY: src = q[cur_q] + first; break;
      case DCG_FIELD_VELOCITY:
        k_unpack_velocity<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(vw[cur_v] + first, scratch, n);
        launches++;
        src = scratch;
        break;
      case DCG_FIELD_FLUIDITY: src = fluidity + first; break;
      case DCG_FIELD_PRESSURE: src = p + first; break;
      case DCG_FIELD_DIVERGENCE: src = div + first; break;
      case DCG_FIELD_T_PRESSURE: src = tp + first; break;
      default: return fail(DCG_ERR_INVALID, "get_field: unknown field %d", field);
    }
    DCG_CUDA_TRY(cudaMemcpyAsync(dst, src, need * sizeof(float), cudaMemcpyDeviceToHost, stream));
    DCG_TRY(synchronize());
    return check_barrier_error();
  }

  int get_counters(uint64_t out[8]) override {
    for (int i = 0; i < 8; i++) out[i] = 0;
    out[6] = launches;
    out[7] = n_barriers;
    return DCG_OK;
  }

  // fields only, fp32, each field read once + written once per stage; the cells THIS instance owns.
  int algorithmic_bytes(double *bytes, uint64_t *active_blocks) override {
    const double per_level0 = 28.0 /*advectV*/ + 28.0 /*divergence*/ + 32.0 /*apply*/ + 24.0 /*advectQ*/;
    double b = per_level0 * (double)local_cells();
    for (int l = 0; l < mip_levels; l++) {
      const double n = (double)(zfirst(l, rank0 + nlocal) - zfirst(l, rank0)) * (double)(gx >> l) * (double)(gy >> l);
      const int pairs = (l == mip_levels - 1) ? project_coarsest_pairs : project_level_pairs;
      b += n * 12.0 * 2 * pairs;               // Jacobi sweeps
      if (l >= 1) b += n * (8 * 4.0 + 12.0);   // restrict: read 8 children, write div,p,tp
      if (l < mip_levels - 1) b += n * 4.5;    // prolongate: write p, read 1/8 coarse
    }
    if (bytes) *bytes = b;
    if (active_blocks) *active_blocks = 0;
    return DCG_OK;
  }
};

}  // namespace
}  // namespace dcg

int dcg_sim::step(int n) {
  for (int i = 0; i < n; i++) {
    DCG_TRY(advect_velocity());
    DCG_TRY(adapt_topology());
    if (ext.sources) DCG_TRY(apply_sources());  // extension: the fused source pass
    DCG_TRY(project());
    DCG_TRY(advect_density());
  }
  return DCG_OK;
}

dcg_sim *dcg_make_uniform() { return new dcg::UniformSim(); }

extern "C" {

DCG_API int dcg_create_uniform_sharded(const dcg_sim_params *params, int device, int rank, int world, int nlocal, dcg_sim **out) {
  if (!params || !out) return DCG_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return DCG_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) return DCG_ERR_INVALID;
  auto *s = new dcg::UniformSim();
  int rc = s->construct_sharded(params, device, rank, world, nlocal);
  if (rc == DCG_OK) rc = s->synchronize();
  if (rc != DCG_OK) {
    dcg_set_create_error(s->err.c_str());
    delete s;
    return rc;
  }
  *out = s;
  return DCG_OK;
}

DCG_API uint64_t dcg_shard_handle_bytes(void) { return dcg::vmm::kHandleBytes; }

DCG_API int dcg_shard_export_handle(dcg_sim *sim, void *out, uint64_t capacity) {
  if (!sim || !out) return DCG_ERR_INVALID;
  return sim->export_handle(out, capacity);
}

DCG_API int dcg_shard_import_handles(dcg_sim *sim, const void *handles, int count) {
  if (!sim || !handles) return DCG_ERR_INVALID;
  int rc = sim->import_handles(handles, count);
  if (rc == DCG_OK) rc = sim->synchronize();
  return rc;
}

}  // extern "C"
