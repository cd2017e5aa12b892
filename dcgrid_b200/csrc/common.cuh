// Shared device/host helpers of the B200-native solver.
//
// Numerics contract: this translation unit is compiled with -fmad=false, and every
// expression below keeps the reference's operand order, so results are bit-identical
// to the reference CUDA built with -fmad=false and to the strict-IEEE CPU oracle
// (oracle/dcgrid_oracle.cpp).  Do not reorder floating-point expressions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dcgrid_b200.h"

namespace dcg {

// The subset of SimParams (reference src/data/sim_params.h:14-47) the solve reads,
// passed BY VALUE to every kernel: the reference keeps it in a process-global
// __constant__ symbol (src/utils/sim_utils.cu:6-9); per-launch arguments make it
// per-instance and multi-GPU safe.
struct KParams {
  float dt, dx, rdx;
  float vel_rate, dens_rate, emit_radius;
  int gx, gy, gz;
  int solids;
  // extension (dcg_ext_params.terrain): height-field terrain instead of the sphere; 0 = the reference's scene
  int terrain;
  float terrain_height, terrain_wavelength;
};

inline KParams make_kparams(const dcg_sim_params &p, const dcg_ext_params &e = dcg_ext_params{}) {
  KParams k;
  k.dt = p.dt; k.dx = p.dx; k.rdx = p.rdx;
  k.vel_rate = p.velocity_emission_rate;
  k.dens_rate = p.density_emission_rate;
  k.emit_radius = p.emission_radius;
  k.gx = p.gx; k.gy = p.gy; k.gz = p.gz;
  k.solids = p.enable_additional_solids ? 1 : 0;
  k.terrain = e.terrain ? 1 : 0;
  k.terrain_height = e.terrain_height;
  k.terrain_wavelength = e.terrain_wavelength;
  return k;
}

// ---- solids: src/utils/sim_utils.cu:11-22, src/sdf.cuh:8-20 -----------------------
__device__ __forceinline__ float cell_fluidity(const KParams &P, int x, int y, int z, int scale) {
  if (!P.solids) return 1.f;
  const float fs = (float)scale;
  const float px = ((float)x + .5f) * fs, py = ((float)y + .5f) * fs, pz = ((float)z + .5f) * fs;
  float d;
  if (P.terrain) {  // extension, specified by oracle/dcgrid_oracle.cpp terrain_sdf()
    float ux = px / P.terrain_wavelength;
    ux = ux - floorf(ux);
    const float wx = 2.f * ux - 1.f;
    const float hx = 1.f - wx * wx;
    float uz = pz / P.terrain_wavelength;
    uz = uz - floorf(uz);
    const float wz = 2.f * uz - 1.f;
    const float hz = 1.f - wz * wz;
    d = py - P.terrain_height * hx * hz;
  } else {
    const float ex = px - (float)P.gx * .5f, ey = py - (float)P.gy * .45f, ez = pz - (float)P.gz * .5f;
    d = sqrtf(ex * ex + ey * ey + ez * ez) - 2000.f * P.rdx;
  }
  const float overlap = fmaxf(0.f, fminf(.5f - d / (fs * 1.73205f), 1.f));
  return 1.f - overlap;
}

// ---- boundary conditions: src/utils/sim_utils.cu:24-55 ----------------------------
// 0 = interior (pass-through), 1 = inlet disc under the floor, 2 = any other outside cell
__device__ __forceinline__ int bc_kind(const KParams &P, int x, int y, int z, int scale) {
  if (y < 0) {
    const float a = (float)(x * scale) - .5f * (float)P.gx;
    const float b = (float)(z * scale) - .5f * (float)P.gz;
    if (sqrtf(a * a + b * b) < P.emit_radius * P.rdx) return 1;
    return 2;
  }
  if (x < 0 || z < 0 || x * scale >= P.gx || y * scale >= P.gy || z * scale >= P.gz) return 2;
  return 0;
}
__device__ __forceinline__ float3 velocity_bc(const KParams &P, float3 v, int x, int y, int z, int scale) {
  const int k = bc_kind(P, x, y, z, scale);
  if (k == 1) return make_float3(0.f, P.vel_rate, 0.f);
  if (k == 2) return make_float3(0.f, 0.f, 0.f);
  return v;
}
__device__ __forceinline__ float density_bc(const KParams &P, float q, int x, int y, int z, int scale) {
  const int k = bc_kind(P, x, y, z, scale);
  if (k == 1) return P.dens_rate;
  if (k == 2) return 0.f;
  return q;
}

// ---- fluidity-weighted, renormalised trilinear weights ------------------------------
// uniformgrid_fluid.cu:28-48 / dcgrid_fluid.cu:48-72.  f[] = fluidity of the 8 corners in
// the reference's 000,001,010,011,100,101,110,111 order (x is the slowest bit).
struct Weights8 {
  float w[8];
  float acc;
};
__device__ __forceinline__ Weights8 corner_weights(const float f[8], float dx, float dy, float dz) {
  const float Dx = 1.f - dx, Dy = 1.f - dy, Dz = 1.f - dz;
  Weights8 c;
  c.w[0] = f[0] * Dx * Dy * Dz;
  c.w[1] = f[1] * Dx * Dy * dz;
  c.w[2] = f[2] * Dx * dy * Dz;
  c.w[3] = f[3] * Dx * dy * dz;
  c.w[4] = f[4] * dx * Dy * Dz;
  c.w[5] = f[5] * dx * Dy * dz;
  c.w[6] = f[6] * dx * dy * Dz;
  c.w[7] = f[7] * dx * dy * dz;
  c.acc = c.w[0] + c.w[1] + c.w[2] + c.w[3] + c.w[4] + c.w[5] + c.w[6] + c.w[7];
  const float inv = 1.f / c.acc;
#pragma unroll
  for (int i = 0; i < 8; i++) c.w[i] *= inv;
  return c;
}
__device__ __forceinline__ float blend8(const float q[8], const float w[8]) {
  return q[0] * w[0] + q[1] * w[1] + q[2] * w[2] + q[3] * w[3] + q[4] * w[4] + q[5] * w[5] + q[6] * w[6] +
         q[7] * w[7];
}

// x / 6.f, correctly rounded, without the generic division routine: q = RN(x*r), rem = x - 6q (exact, one
// FMA), q' = RN(q + rem*r) with r = RN(1/6).  Checked exhaustively over all 2^32 inputs against x / 6.f
// (tests/tools/div6_exhaustive.c): identical for every |x| >= 2^-125; smaller magnitudes (denormal
// quotients) take the IEEE division.  The explicit fmaf is a single-rounding FMA regardless of -fmad.
__device__ __forceinline__ float div6(float x) {
  const float r = 0x1.555556p-3f;
  if (fabsf(x) < 0x1p-120f) return x / 6.f;
  const float q = x * r;
  return fmaf(fmaf(-6.f, q, x), r, q);
}

// ---- programmatic dependent launch (sm_90+) ------------------------------------------------
// Every kernel of a step starts with pdl_enter(): when the host launched it with
// cudaLaunchAttributeProgrammaticStreamSerialization its CTAs may become resident while the previous
// kernel of the stream is still draining (launch latency and CTA ramp-up overlap the predecessor's tail);
// griddepcontrol.wait then blocks until that kernel has completed and its writes are visible, so the data
// dependence is exactly that of plain stream order.  launch_dependents lets the NEXT kernel do the same
// with this one.  Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_trigger();
}

// ---- misc ------------------------------------------------------------------------
__host__ __device__ __forceinline__ int idiv_up(int a, int b) { return (a + b - 1) / b; }

}  // namespace dcg
