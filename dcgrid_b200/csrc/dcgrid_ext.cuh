// Extensions beyond the reference snapshot (include/dcgrid_b200.h "extensions"): flow-driven
// refinement score, temperature / vapor scalars, the fused source pass (condensation, buoyancy, vorticity
// confinement), MacCormack advection, point sampling.  None of this exists in /root/reference; the
// specification is the CPU oracle (oracle/dcgrid_oracle.cpp, "EXTENSIONS"), and every expression below keeps the
// oracle's operand order (this TU is compiled with -fmad=false) so that the two agree bit for bit.
//
// These kernels are written for clarity, one thread per cell on the staged-apron pattern of dcgrid_kernels.cuh: the
// hot path of the benchmarked step (dcgrid_pipe.cuh) is untouched when every switch is off.
#pragma once
#include "dcgrid_kernels.cuh"

namespace dcg {
namespace ext {

__device__ __forceinline__ float cell_height(const KParams &P, int y, int scale) { return ((float)y + .5f) * (float)scale * P.dx; }
__device__ __forceinline__ float ambient_theta(const dcg_ext_params &E, float h) { return E.ambient_temperature + E.ambient_lapse * h; }

// kBC: 0 density (sim_utils.cu:41-55), 1 temperature, 2 vapor (oracle temperature_bc / vapor_bc)
template <int kBC>
__device__ __forceinline__ float scalar_bc(const KParams &P, const dcg_ext_params &E, float v, int x, int y, int z, int scale) {
  const int k = bc_kind(P, x, y, z, scale);
  if (kBC == 0) {
    if (k == 1) return P.dens_rate;
    if (k == 2) return 0.f;
  } else if (kBC == 1) {
    if (k == 1) return E.ambient_temperature + E.temperature_emission;
    if (k == 2) return ambient_theta(E, cell_height(P, y, scale));
  } else {
    if (k == 1) return E.vapor_emission;
    if (k == 2) return E.ambient_vapor;
  }
  return v;
}

// k_dcgrid_calc_vorticity, dcgrid_fluid.cu:146-172, for every cell of every active block; .w = |omega| (the
// calcCellScore of dcgrid_adaptation.cu:6-8)
__global__ void __launch_bounds__(kCTA) k_dc_ext_vorticity(Pool T, KParams P, const float4 *__restrict__ vw, float4 *__restrict__ vort) {
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t b = blockIdx.x * kBPC + g;
  if (b >= T.M) return;
  const int4 pl = T.posl[b];
  if (pl.w == kFree) return;
  const uint32_t *a = T.apron + (size_t)b * kAV;  // (staging the map in shared memory first: 351 -> 443 us, profiles/README.md r3e)
  const int ai = apron_of(t);
  const float4 l = vw[a[ai - kAA]], r = vw[a[ai + kAA]], d = vw[a[ai - kAW]], u = vw[a[ai + kAW]], bk = vw[a[ai - 1]], f = vw[a[ai + 1]];
  const int scale = 1 << pl.w;
  const float alpha = .5f * P.rdx / (float)scale;
  float4 o;
  o.x = alpha * ((u.w * u.z - d.w * d.z) - (f.w * f.y - bk.w * bk.y));
  o.y = alpha * ((f.w * f.x - bk.w * bk.x) - (r.w * r.z - l.w * l.z));
  o.z = alpha * ((r.w * r.y - l.w * l.y) - (u.w * u.x - d.w * d.x));
  o.w = sqrtf(o.x * o.x + o.y * o.y + o.z * o.z);
  vort[(size_t)b * kBV + t] = o;
}

// the fused source pass (oracle apply_sources): leaf cells only, in place; reads neighbours' |omega| only; blocks without
// children restrict what they wrote themselves
__global__ void __launch_bounds__(kCTA) k_dc_ext_sources(Pool T, KParams P, dcg_ext_params E, float4 *__restrict__ vw, float *__restrict__ q,
                                                         float *__restrict__ th, float *__restrict__ qv, const float4 *__restrict__ vort) {
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t b = blockIdx.x * kBPC + g;
  if (b >= T.M) return;
  const int4 pl = T.posl[b];
  if (pl.w == kFree) return;
  // fused restriction (accumulate<T>, dcgrid_structure.cu:188-222) of blocks without children, as in the field kernels: the 8 cells
  // of a subblock are 8 consecutive lanes, summed in cell order from 0.f; blocks WITH children are left to the list passes
  const uint4 *cl = reinterpret_cast<const uint4 *>(T.child + (size_t)b * kSV);
  const uint4 c0 = cl[0], c1 = cl[1];
  const uint32_t ps = T.parent[b];
  const bool push = (c0.x & c0.y & c0.z & c0.w & c1.x & c1.y & c1.z & c1.w) == kNone && ps != kNone;  // uniform over the block
  auto sub_sum = [&](float v) {
    const unsigned base = threadIdx.x & 24u;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) acc += __shfl_sync(0xFFFFFFFFu, v, base + k);
    return acc;
  };
  if (!push && T.child[(size_t)b * kSV + (t >> 3)] != kNone) return;  // (a block that pushes has no refined subblock)
  const size_t c = (size_t)b * kBV + t;
  const uint32_t *a = T.apron + (size_t)b * kAV;
  const int scale = 1 << pl.w;
  const float alpha = .5f * P.rdx / (float)scale;
  const int y = pl.y | cell_y(t);
  const float h = cell_height(P, y, scale);
  float tc = th[c], vc = qv[c], qc = q[c];
  const float tabs = tc - E.adiabatic_lapse * h;
  const float qs = fmaxf(0.f, E.saturation_base + E.saturation_slope * (tabs - E.ambient_temperature));
  float dq = E.condensation_rate * (vc - qs);
  dq = fmaxf(dq, -qc);
  vc = vc - dq;
  qc = qc + dq;
  tc = tc + E.latent_heat * dq;
  const float tha = ambient_theta(E, h);
  const float lift = E.buoyancy * ((tc - tha) / E.ambient_temperature) + E.vapor_buoyancy * vc - E.smoke_weight * qc;
  const int ai = apron_of(t);
  const float gx_ = alpha * (vort[a[ai + kAA]].w - vort[a[ai - kAA]].w);
  const float gy_ = alpha * (vort[a[ai + kAW]].w - vort[a[ai - kAW]].w);
  const float gz_ = alpha * (vort[a[ai + 1]].w - vort[a[ai - 1]].w);
  const float glen = sqrtf(gx_ * gx_ + gy_ * gy_ + gz_ * gz_);
  float fx = 0.f, fy = 0.f, fz = 0.f;
  if (glen > 1e-12f) {
    const float inv = 1.f / glen;
    const float nx = gx_ * inv, ny = gy_ * inv, nz = gz_ * inv;
    const float4 w = vort[c];
    const float k = E.vorticity_confinement * (P.dx * (float)scale);
    fx = k * (ny * w.z - nz * w.y);
    fy = k * (nz * w.x - nx * w.z);
    fz = k * (nx * w.y - ny * w.x);
  }
  float4 v = vw[c];
  const float gg = P.dt * v.w;
  v.x = v.x + gg * fx;
  v.y = v.y + gg * (fy + lift);
  v.z = v.z + gg * fz;
  vw[c] = v;
  th[c] = tc;
  qv[c] = vc;
  q[c] = qc;
  if (push) {
    const float ax = sub_sum(v.x), ay = sub_sum(v.y), az = sub_sum(v.z), at = sub_sum(tc), av = sub_sum(vc), aq = sub_sum(qc);
    if ((t & 7u) == 7u) {
      const size_t pc = (size_t)kSV * ps + (t >> 3);
      float *dst = reinterpret_cast<float *>(vw + pc);
      dst[0] = ax * .125f; dst[1] = ay * .125f; dst[2] = az * .125f;  // .w (fluidity) untouched
      th[pc] = at * .125f;
      qv[pc] = av * .125f;
      q[pc] = aq * .125f;
    }
  }
}

// initial temperature / vapor of every active block (oracle activate_level): the ambient profile
__global__ void __launch_bounds__(kCTA) k_dc_ext_init_scalars(Pool T, KParams P, dcg_ext_params E, float *__restrict__ th0, float *__restrict__ th1,
                                                              float *__restrict__ qv0, float *__restrict__ qv1) {
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t b = blockIdx.x * kBPC + g;
  if (b >= T.M) return;
  const int4 pl = T.posl[b];
  if (pl.w == kFree) return;
  const size_t c = (size_t)b * kBV + t;
  const float v = ambient_theta(E, cell_height(P, pl.y | cell_y(t), 1 << pl.w));
  th0[c] = v; th1[c] = v;
  qv0[c] = E.ambient_vapor; qv1[c] = E.ambient_vapor;
}

// k_dcgrid_propagate_values (dcgrid_adaptation.cu:92-143) for the two extension scalars
__global__ void __launch_bounds__(64) k_dc_ext_propagate(Pool T, const uint32_t *__restrict__ touched, const uint32_t *__restrict__ perm, int level,
                                                         float *__restrict__ th, float *__restrict__ qv) {
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
  th[c] = ((27.f / 64.f) * th[i000] + (9.f / 64.f) * (th[i001] + th[i010] + th[i100]) + (3.f / 64.f) * (th[i011] + th[i101] + th[i110]) +
           (1.f / 64.f) * th[i111]);
  qv[c] = ((27.f / 64.f) * qv[i000] + (9.f / 64.f) * (qv[i001] + qv[i010] + qv[i100]) + (3.f / 64.f) * (qv[i011] + qv[i101] + qv[i110]) +
           (1.f / 64.f) * qv[i111]);
}

// position of cell t of the block at `pl`, moved along its own velocity: sign -1 = the reference's backtrace
// (dcgrid_fluid.cu:119-123), +1 = the reversed trajectory of the MacCormack correction
__device__ __forceinline__ void trace(const KParams &P, const int4 pl, uint32_t t, const float4 me, bool forward, float &bx, float &by, float &bz) {
  const float scale = (float)(1 << pl.w);
  const float alpha = P.dt * P.rdx;
  const float fx = (float)(pl.x | cell_x(t)), fy = (float)(pl.y | cell_y(t)), fz = (float)(pl.z | cell_z(t));
  if (!forward) {
    bx = (fx + .5f) * scale - me.x * alpha;
    by = (fy + .5f) * scale - me.y * alpha;
    bz = (fz + .5f) * scale - me.z * alpha;
  } else {
    bx = (fx + .5f) * scale + me.x * alpha;
    by = (fy + .5f) * scale + me.y * alpha;
    bz = (fz + .5f) * scale + me.z * alpha;
  }
}

// Scalar advection with a selectable boundary rule.  kMC = false: out = semi-Lagrangian gather of phi (the reference's
// k_dcgrid_advect_density with the scalar's own boundary values).  kMC = true: the MacCormack correction (oracle
// maccormack()): hat = SL(phi) already restricted; out = clamp(hat + .5 (phi - SL_reversed(hat)), corners of the
// forward sample), hat where either sample has no fluid weight.
template <int kBC, bool kMC>
__global__ void __launch_bounds__(kCTA) k_dc_ext_advect_scalar(Pool T, KParams P, dcg_ext_params E, const float4 *__restrict__ vw,
                                                               const float *__restrict__ fl, const float *__restrict__ phi,
                                                               const float *__restrict__ hat, float *__restrict__ out) {
  __shared__ uint32_t sa[kBPC][kAV];
  __shared__ uint32_t sc[kBPC][kSV];
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t b = blockIdx.x * kBPC + g;
  const bool in_pool = b < T.M;
  const uint32_t c = b * kBV + t;
  int4 pl = make_int4(0, 0, 0, kFree);
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  if (in_pool) {
    pl = T.posl[b];
    me = vw[c];
  }
  stage_apron(T, b, in_pool, g, t, sa, sc);
  if (pl.w == kFree) return;
  float o = 0.f;
  if (sc[g][t >> 3] == kNone) {
    // a sample without fluid weight (inside a solid): the reference's 0 for the density, the ambient value for the
    // extension scalars
    if (kBC == 1) o = ambient_theta(E, cell_height(P, pl.y | cell_y(t), 1 << pl.w));
    if (kBC == 2) o = E.ambient_vapor;
    float bx, by, bz;
    trace(P, pl, t, me, false, bx, by, bz);
    const DSample s = d_sample(T, P, sa[g], sc[g], pl, bx, by, bz);
    float v8[8], f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      f[k] = fl[s.id[k]];
      v8[k] = phi[s.id[k]];
    }
    const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
    if (!(W.acc < 1e-6f)) {
#pragma unroll
      for (int k = 0; k < 8; k++) v8[k] = scalar_bc<kBC>(P, E, v8[k], s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1), s.z0 + (k & 1), s.scale);
      if (!kMC) {
        o = blend8(v8, W.w);
      } else {
        float mn = v8[0], mx = v8[0];
#pragma unroll
        for (int k = 1; k < 8; k++) {
          mn = fminf(mn, v8[k]);
          mx = fmaxf(mx, v8[k]);
        }
        o = hat[c];
        float fx, fy, fz;
        trace(P, pl, t, me, true, fx, fy, fz);
        const DSample sf = d_sample(T, P, sa[g], sc[g], pl, fx, fy, fz);
        float h8[8], ff[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          ff[k] = fl[sf.id[k]];
          h8[k] = hat[sf.id[k]];
        }
        const Weights8 Wf = corner_weights(ff, sf.fx, sf.fy, sf.fz);
        if (!(Wf.acc < 1e-6f)) {
#pragma unroll
          for (int k = 0; k < 8; k++)
            h8[k] = scalar_bc<kBC>(P, E, h8[k], sf.x0 + ((k >> 2) & 1), sf.y0 + ((k >> 1) & 1), sf.z0 + (k & 1), sf.scale);
          const float back = blend8(h8, Wf.w);
          const float r = o + .5f * (phi[c] - back);
          o = fminf(fmaxf(r, mn), mx);
        }
      }
    }
  }
  out[c] = o;
}

// MacCormack correction of the velocity (same scheme per component; trajectories along the pre-advection velocity
// `vin`, whose .w is the fluidity)
__global__ void __launch_bounds__(kCTA) k_dc_ext_maccormack_velocity(Pool T, KParams P, const float4 *__restrict__ vin, const float4 *__restrict__ hat,
                                                                     float4 *__restrict__ out) {
  __shared__ uint32_t sa[kBPC][kAV];
  __shared__ uint32_t sc[kBPC][kSV];
  const uint32_t g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const uint32_t b = blockIdx.x * kBPC + g;
  const bool in_pool = b < T.M;
  const uint32_t c = b * kBV + t;
  int4 pl = make_int4(0, 0, 0, kFree);
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  if (in_pool) {
    pl = T.posl[b];
    me = vin[c];
  }
  stage_apron(T, b, in_pool, g, t, sa, sc);
  if (pl.w == kFree) return;
  float3 o = make_float3(0.f, 0.f, 0.f);
  if (sc[g][t >> 3] == kNone) {
    float bx, by, bz;
    trace(P, pl, t, me, false, bx, by, bz);
    const DSample s = d_sample(T, P, sa[g], sc[g], pl, bx, by, bz);
    float f[8], vx[8], vy[8], vz[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float4 cv = vin[s.id[k]];
      f[k] = cv.w;
      vx[k] = cv.x; vy[k] = cv.y; vz[k] = cv.z;
    }
    const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
    if (!(W.acc < 1e-6f)) {
      float3 mn, mx;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float3 v = velocity_bc(P, make_float3(vx[k], vy[k], vz[k]), s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1), s.z0 + (k & 1), s.scale);
        if (k == 0) {
          mn = v; mx = v;
        } else {
          mn.x = fminf(mn.x, v.x); mn.y = fminf(mn.y, v.y); mn.z = fminf(mn.z, v.z);
          mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z);
        }
      }
      const float4 hc = hat[c];
      o = make_float3(hc.x, hc.y, hc.z);
      float fx, fy, fz;
      trace(P, pl, t, me, true, fx, fy, fz);
      const DSample sf = d_sample(T, P, sa[g], sc[g], pl, fx, fy, fz);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float4 cv = hat[sf.id[k]];
        f[k] = cv.w;
        vx[k] = cv.x; vy[k] = cv.y; vz[k] = cv.z;
      }
      const Weights8 Wf = corner_weights(f, sf.fx, sf.fy, sf.fz);
      if (!(Wf.acc < 1e-6f)) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const float3 v = velocity_bc(P, make_float3(vx[k], vy[k], vz[k]), sf.x0 + ((k >> 2) & 1), sf.y0 + ((k >> 1) & 1), sf.z0 + (k & 1), sf.scale);
          vx[k] = v.x; vy[k] = v.y; vz[k] = v.z;
        }
        const float backx = blend8(vx, Wf.w), backy = blend8(vy, Wf.w), backz = blend8(vz, Wf.w);
        const float rx = hc.x + .5f * (me.x - backx), ry = hc.y + .5f * (me.y - backy), rz = hc.z + .5f * (me.z - backz);
        o.x = fminf(fmaxf(rx, mn.x), mx.x);
        o.y = fminf(fmaxf(ry, mn.y), mx.y);
        o.z = fminf(fmaxf(rz, mn.z), mx.z);
      }
    }
  }
  out[c] = make_float4(o.x, o.y, o.z, me.w);
}

// sampleCoarse / samplePrecise (dcgrid_rendering.cu:6-58, interpolate() of raymarching.cuh:26-40) at n positions
__global__ void __launch_bounds__(256) k_dc_ext_sample(Pool T, KParams P, const float *__restrict__ src, int comps, int stride, int mode,
                                                       const float *__restrict__ xyz, size_t n, float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
  const int ix = min(max((int)floorf(px), 0), P.gx - 1), iy = min(max((int)floorf(py), 0), P.gy - 1), iz = min(max((int)floorf(pz), 0), P.gz - 1);
  int level = 0;
  const uint32_t b = block_index_deep(T, P, ix, iy, iz, level);
  const int4 pl = T.posl[b];
  if (mode == 0) {
    const uint32_t c = b * kBV + cell_bits((ix >> level) % kBW, (iy >> level) % kBW, (iz >> level) % kBW);
    for (int k = 0; k < comps; k++) out[comps * i + k] = src[(size_t)stride * c + k];
    return;
  }
  const float inv = 1.f / (float)(1 << level);
  const float x = px * inv - .5f, y = py * inv - .5f, z = pz * inv - .5f;
  const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  const float dx = x - xf, dy = y - yf, dz = z - zf;
  const int ai = min(max((int)xf + 1 - pl.x, 0), kAW - 2), aj = min(max((int)yf + 1 - pl.y, 0), kAW - 2), ak = min(max((int)zf + 1 - pl.z, 0), kAW - 2);
  const uint32_t *a = T.apron + (size_t)b * kAV + kAA * ai + kAW * aj + ak;
  const uint32_t id[8] = {a[0], a[1], a[kAW], a[kAW + 1], a[kAA], a[kAA + 1], a[kAA + kAW], a[kAA + kAW + 1]};
  for (int k = 0; k < comps; k++) {
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = src[(size_t)stride * id[q] + k];
    const float dxi = 1.f - dx;
    const float c00 = v[0] * dxi + v[4] * dx, c01 = v[1] * dxi + v[5] * dx, c10 = v[2] * dxi + v[6] * dx, c11 = v[3] * dxi + v[7] * dx;
    const float dyi = 1.f - dy;
    const float c0 = c00 * dyi + c10 * dy, c1 = c01 * dyi + c11 * dy;
    out[comps * i + k] = c0 * (1.f - dz) + c1 * dz;
  }
}

// state load: packed velocity + fluidity from the two dumped arrays; dense level maps from the block positions
__global__ void __launch_bounds__(256) k_dc_ext_pack_vw(const float *__restrict__ v3, const float *__restrict__ fl, float4 *__restrict__ vw0,
                                                        float4 *__restrict__ vw1, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 v = make_float4(v3[3 * i], v3[3 * i + 1], v3[3 * i + 2], fl[i]);
  vw0[i] = v;
  vw1[i] = v;
}
__global__ void __launch_bounds__(256) k_dc_ext_rebuild_maps(Pool T, KParams P) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= T.M) return;
  const int4 pl = T.posl[b];
  if (pl.w == kFree || pl.w >= T.sparse_levels) return;
  T.map[pl.w][map_slot(P, pl.x, pl.y, pl.z, pl.w)] = b;
}

}  // namespace ext
}  // namespace dcg

// ======================================================================================================================
// Device-side selection of adaptTopology (north_star (1): warp-ballot and prefix-sum compaction; SURVEY App. B-8 stage 2)
// ======================================================================================================================
namespace dcg {
namespace sel {

constexpr int kSelCta = 1024;  // one element per thread: a CTA covers 1024 consecutive ids, a warp 32

// Order-preserving compaction of the ids i in [first, first + n) with score[i] > thresh — the candidate scan of
// refineSubblocks (fluid_simulation_dcgrid.cu:455-463: `if (subblockScores[i] > 1e-4f) list[n++] = i`) — in three
// passes: per-CTA counts from warp ballots, an exclusive scan of the counts, and the scatter at
// CTA offset + warp prefix + popc(ballot below the lane).  The output is in ascending id order, exactly the sequence the
// reference's host loop produces.
__device__ __forceinline__ uint32_t cta_rank(bool pred, uint32_t &cta_total) {
  __shared__ uint32_t warp_sum[kSelCta / 32];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned votes = __ballot_sync(0xFFFFFFFFu, pred);
  if (lane == 0) warp_sum[warp] = __popc(votes);
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kSelCta / 32; w++) {
    const uint32_t v = warp_sum[w];
    before += w < (int)warp ? v : 0u;
    total += v;
  }
  cta_total = total;
  return before + __popc(votes & ((1u << lane) - 1u));
}
__global__ void __launch_bounds__(kSelCta) k_sel_count(const float *__restrict__ score, uint32_t first, uint32_t n, float thresh,
                                                       uint32_t *__restrict__ cta_counts) {
  const uint32_t i = blockIdx.x * kSelCta + threadIdx.x;
  uint32_t total;
  cta_rank(i < n && score[first + i] > thresh, total);
  if (threadIdx.x == 0) cta_counts[blockIdx.x] = total;
}
// in-place exclusive scan of `n` counts by one CTA (n <= a few 10^4: chunks of 1024 with a running carry); total -> *sum
__global__ void __launch_bounds__(kSelCta) k_sel_scan(uint32_t *__restrict__ counts, uint32_t n, uint32_t *__restrict__ sum) {
  __shared__ uint32_t part[kSelCta / 32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n; base += kSelCta) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n ? counts[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    if (lane == 31) part[warp] = incl;
    __syncthreads();
    uint32_t wbefore = 0, chunk = 0;
#pragma unroll
    for (int w = 0; w < kSelCta / 32; w++) {
      const uint32_t pv = part[w];
      wbefore += w < (int)warp ? pv : 0u;
      chunk += pv;
    }
    const uint32_t carry = carry_s;
    if (i < n) counts[i] = carry + wbefore + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + chunk;
    __syncthreads();
  }
  if (threadIdx.x == 0) *sum = carry_s;
}
__global__ void __launch_bounds__(kSelCta) k_sel_scatter(const float *__restrict__ score, uint32_t first, uint32_t n, float thresh,
                                                         const uint32_t *__restrict__ cta_offsets, uint32_t *__restrict__ out) {
  const uint32_t i = blockIdx.x * kSelCta + threadIdx.x;
  const bool pred = i < n && score[first + i] > thresh;
  uint32_t total;
  const uint32_t r = cta_rank(pred, total);
  if (pred) out[cta_offsets[blockIdx.x] + r] = first + i;
}

// Sort keys of the total-order selection (dcg_ext_params.selection == 1).  Non-negative floats order like their bit
// patterns: blocks ascending by (score, slot) -> key = bits << 32 | slot; destinations descending by score, ascending by
// id -> key = ~bits << 32 | id.  Negative scores (unmovable / unusable) get the all-ones key and sort behind everything.
// counts[0] += number of valid keys (one atomic per warp).
template <bool kDescending>
__global__ void __launch_bounds__(256) k_sel_keys(const float *__restrict__ score, uint32_t first, uint32_t n, unsigned long long *__restrict__ keys,
                                                  uint32_t *__restrict__ count) {
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  bool valid = false;
  if (i < n) {
    const float s = score[first + i];
    valid = s >= 0.f;
    const uint32_t bits = __float_as_uint(s);
    keys[i] = valid ? ((unsigned long long)(kDescending ? ~bits : bits) << 32) | (unsigned long long)(first + i) : ~0ull;
  }
  const unsigned votes = __ballot_sync(0xFFFFFFFFu, valid);
  if ((threadIdx.x & 31u) == 0 && votes) atomicAdd(count, (uint32_t)__popc(votes));
}
// sc[0] = valid blocks, sc[1] = valid destinations, -> sc[2] = K = min(l, sc[0], sc[1]), sc[3] = matches (starts at K)
__global__ void k_sel_match_init(uint32_t *__restrict__ sc, uint32_t l) {
  const uint32_t K = min(l, min(sc[0], sc[1]));
  sc[2] = K;
  sc[3] = K;
}
// the greedy rule (:410-418) on the two sorted lists: matches = first m with !(bs[mc[m]] < ss[dc[m]]) — block scores
// ascend and destination scores descend along the lists, so the condition is monotone and the first failure is a minimum
__global__ void __launch_bounds__(256) k_sel_match(const float *__restrict__ bs, const float *__restrict__ ss, const unsigned long long *__restrict__ mc,
                                                   const unsigned long long *__restrict__ dc, uint32_t *__restrict__ sc) {
  const uint32_t m = blockIdx.x * 256 + threadIdx.x;
  if (m >= sc[2]) return;
  if (!(bs[(uint32_t)mc[m]] < ss[(uint32_t)dc[m]])) atomicMin(&sc[3], m);
}
// the matched pairs -> lists; the parents that receive a block are protected for the next level's selection
__global__ void __launch_bounds__(256) k_sel_apply(float *__restrict__ bs, const unsigned long long *__restrict__ mc, const unsigned long long *__restrict__ dc,
                                                   const uint32_t *__restrict__ sc, uint32_t *__restrict__ to_move, uint32_t *__restrict__ dest) {
  const uint32_t m = blockIdx.x * 256 + threadIdx.x;
  if (m >= sc[3]) return;
  const uint32_t b = (uint32_t)mc[m], d = (uint32_t)dc[m];
  bs[d >> 3] = -FLT_MAX;
  to_move[m] = b;
  dest[m] = d;
}

}  // namespace sel
}  // namespace dcg
