// dcgrid_oracle.cpp — CPU restatement of the reference's per-timestep fluid solve.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product path (dcgrid_b200/csrc) never links or
// calls it and has no CPU fallback.
//
// What it is: every CUDA kernel of the reference's solver path rewritten as a loop
// in ascending global-thread-id order ("rank-order serialisation": a legal execution
// of the CUDA code that fixes the order of its atomics), plus the host orchestration
// of FluidSimulationUniform / FluidSimulationDCGrid.  Arithmetic is strict IEEE
// binary32 in the reference's expression order; build with
//   g++ -O2 -ffp-contract=off -fno-fast-math [-fopenmp]
// so that it is bit-comparable to the reference CUDA built with -fmad=false.
//
// Pinning: the reference ships no tests or golden vectors, so this
// restatement is pinned against outputs of the reference's own CUDA kernels run on
// a B200 by oracle/ref_harness (built into oracle/_ref/), see tests/golden/README.md.
//
// Each function cites the reference file:line it follows (paths relative to the
// reference root).
#include "../include/dcgrid_b200.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <vector>

namespace {

using u64 = uint64_t;
constexpr u64 kNone = UINT64_MAX;      // DCGrid::notFound, dcgrid.h:69
constexpr uint32_t kHashEmpty = UINT32_MAX;  // DCGrid::hashEmpty, dcgrid.h:70

struct V3 { float x, y, z; };

inline int iclamp(int v, int lo, int hi) { return std::max(lo, std::min(v, hi)); }
inline int idiv_up(int a, int b) { return (a % b != 0) ? (a / b + 1) : (a / b); }  // grid_math.cuh:5-7

// ---------------------------------------------------------------------------
// Boundary conditions and solids: src/utils/sim_utils.cu:11-55, src/sdf.cuh:8-20
// ---------------------------------------------------------------------------
inline float sphere_sdf(const dcg_sim_params &P, float px, float py, float pz) {
  const float cx = P.gx * .5f, cy = P.gy * .45f, cz = P.gz * .5f;  // sdf.cuh:9-11
  const float r = 2000.f * P.rdx;                                  // sdf.cuh:12
  const float dx = px - cx, dy = py - cy, dz = pz - cz;
  return sqrtf(dx * dx + dy * dy + dz * dz) - r;                   // sdf.cuh:17
}

inline float cell_fluidity(const dcg_sim_params &P, int x, int y, int z, int scale) {
  if (!P.enable_additional_solids) return 1.f;                     // sim_utils.cu:15-16
  const float sqrt3 = 1.73205f;
  const float px = ((float)x + .5f) * (float)scale;
  const float py = ((float)y + .5f) * (float)scale;
  const float pz = ((float)z + .5f) * (float)scale;
  const float d = sphere_sdf(P, px, py, pz);
  const float overlap = fmaxf(0.f, fminf(.5f - d / ((float)scale * sqrt3), 1.f));  // sim_utils.cu:19
  return 1.f - overlap;
}

// ---------------------------------------------------------------------------
// EXTENSIONS (include/dcgrid_b200.h "extensions").  Nothing below this banner exists in the
// reference snapshot: THIS FILE IS THE SPECIFICATION, parity against the reference is unpinned.  Only
// + - * / sqrtf floorf fminf fmaxf, in the order written, so the CUDA kernels can match bit for bit.
// ---------------------------------------------------------------------------
// terrain height field replacing sceneSDF (src/sdf.cuh:20): parabolic hills of period `terrain_wavelength` in x and
// z (valleys along the lines x, z = k * wavelength, so the inlet disc in the domain centre stays open when the
// wavelength divides half the domain); signed distance = height above the surface (vertical, conservative)
inline float terrain_sdf(const dcg_ext_params &E, float px, float py, float pz) {
  float ux = px / E.terrain_wavelength;
  ux = ux - floorf(ux);
  const float wx = 2.f * ux - 1.f;
  const float hx = 1.f - wx * wx;
  float uz = pz / E.terrain_wavelength;
  uz = uz - floorf(uz);
  const float wz = 2.f * uz - 1.f;
  const float hz = 1.f - wz * wz;
  return py - E.terrain_height * hx * hz;
}
inline float cell_fluidity(const dcg_sim_params &P, const dcg_ext_params &E, int x, int y, int z, int scale) {
  if (!E.terrain) return cell_fluidity(P, x, y, z, scale);
  if (!P.enable_additional_solids) return 1.f;
  const float sqrt3 = 1.73205f;
  const float px = ((float)x + .5f) * (float)scale;
  const float py = ((float)y + .5f) * (float)scale;
  const float pz = ((float)z + .5f) * (float)scale;
  const float d = terrain_sdf(E, px, py, pz);
  const float overlap = fmaxf(0.f, fminf(.5f - d / ((float)scale * sqrt3), 1.f));
  return 1.f - overlap;
}
// height of a cell centre above the floor, world units; y in cells of size `scale`
inline float cell_height(const dcg_sim_params &P, int y, int scale) { return ((float)y + .5f) * (float)scale * P.dx; }
inline float ambient_theta(const dcg_ext_params &E, float h) { return E.ambient_temperature + E.ambient_lapse * h; }

inline bool in_inlet(const dcg_sim_params &P, int x, int y, int z, int scale) {
  if (!(y < 0)) return false;                                      // sim_utils.cu:28,44
  const float a = (float)(x * scale) - .5f * (float)P.gx;
  const float b = (float)(z * scale) - .5f * (float)P.gz;
  return sqrtf(a * a + b * b) < P.emission_radius * P.rdx;
}
inline bool out_of_domain(const dcg_sim_params &P, int x, int y, int z, int scale) {
  return x < 0 || y < 0 || z < 0 || x * scale >= P.gx || y * scale >= P.gy || z * scale >= P.gz;
}
inline V3 velocity_bc(const dcg_sim_params &P, V3 v, int x, int y, int z, int scale) {  // sim_utils.cu:24-39
  if (in_inlet(P, x, y, z, scale)) return V3{0.f, P.velocity_emission_rate, 0.f};
  if (out_of_domain(P, x, y, z, scale)) return V3{0.f, 0.f, 0.f};
  return v;
}
inline float density_bc(const dcg_sim_params &P, float q, int x, int y, int z, int scale) {  // sim_utils.cu:41-55
  if (in_inlet(P, x, y, z, scale)) return P.density_emission_rate;
  if (out_of_domain(P, x, y, z, scale)) return 0.f;
  return q;
}

// extension scalars: inlet -> emission value, any other outside cell -> the ambient profile, else pass-through
inline float temperature_bc(const dcg_sim_params &P, const dcg_ext_params &E, float t, int x, int y, int z, int scale) {
  if (in_inlet(P, x, y, z, scale)) return E.ambient_temperature + E.temperature_emission;
  if (out_of_domain(P, x, y, z, scale)) return ambient_theta(E, cell_height(P, y, scale));
  return t;
}
inline float vapor_bc(const dcg_sim_params &P, const dcg_ext_params &E, float v, int x, int y, int z, int scale) {
  if (in_inlet(P, x, y, z, scale)) return E.vapor_emission;
  if (out_of_domain(P, x, y, z, scale)) return E.ambient_vapor;
  return v;
}

// grid_math.cuh:24-30 (3-D mipmapCells) and :42-52 (3-D mipmapIdx offset part)
u64 pyramid_cells(u64 w, u64 h, u64 d) {
  u64 n = 0;
  for (u64 s = 1; w % s == 0 && h % s == 0 && d % s == 0; s *= 2) n += (w * h * d) / (s * s * s);
  return n;
}
u64 pyramid_offset(u64 w, u64 h, u64 d, u64 scale) {
  u64 off = 0;
  for (u64 s = 1; s < scale; s *= 2) off += (w * h * d) / (s * s * s);
  return off;
}

// Trilinear weights shared by both gathers (uniformgrid_fluid.cu:28-48, dcgrid_fluid.cu:48-72):
// w = fluidity * three 1-D factors, renormalised by 1/sum.
struct Corner8 { float w[8]; float acc; };
inline Corner8 corner_weights(const float f[8], float dx, float dy, float dz) {
  const float Dx = 1.f - dx, Dy = 1.f - dy, Dz = 1.f - dz;
  Corner8 c;
  c.w[0] = f[0] * Dx * Dy * Dz;  // 000
  c.w[1] = f[1] * Dx * Dy * dz;  // 001
  c.w[2] = f[2] * Dx * dy * Dz;  // 010
  c.w[3] = f[3] * Dx * dy * dz;  // 011
  c.w[4] = f[4] * dx * Dy * Dz;  // 100
  c.w[5] = f[5] * dx * Dy * dz;  // 101
  c.w[6] = f[6] * dx * dy * Dz;  // 110
  c.w[7] = f[7] * dx * dy * dz;  // 111
  c.acc = c.w[0] + c.w[1] + c.w[2] + c.w[3] + c.w[4] + c.w[5] + c.w[6] + c.w[7];
  const float inv = 1.f / c.acc;
  for (int i = 0; i < 8; i++) c.w[i] *= inv;
  return c;
}
inline float blend8(const float q[8], const float w[8]) {
  return q[0] * w[0] + q[1] * w[1] + q[2] * w[2] + q[3] * w[3] + q[4] * w[4] + q[5] * w[5] +
         q[6] * w[6] + q[7] * w[7];
}

}  // namespace

// ===========================================================================
// Common base
// ===========================================================================
struct orc_sim {
  dcg_sim_params P{};
  dcg_ext_params E{};  // extensions, all zero = the reference snapshot
  bool is_dcgrid = false;
  virtual void apply_sources() {}
  int coarse_pairs = 0, level_pairs = 0, local_pairs = 0;
  virtual ~orc_sim() {}
  virtual void reset() = 0;
  virtual void init() = 0;
  virtual void adapt_topology() = 0;
  virtual void advect_velocity() = 0;
  virtual void project() = 0;
  virtual void project_local() = 0;
  virtual void advect_density() = 0;
  virtual float debug_stats() = 0;
  virtual u64 num_cells() const = 0;
  virtual int get_field(int field, float *dst) = 0;
};

// ===========================================================================
// Uniform grid: src/uniformgrid/*
// ===========================================================================
struct UniformOracle : orc_sim {
  int gx, gy, gz;
  u64 N, pyr;
  int mip_levels;
  std::vector<float> density, velocity /*3N*/, fluidity /*pyr*/, pressure /*pyr*/, temp /*3N*/;
  // aliases of temp (uniformgrid_structure.cu:7-13)
  float *t_velocity() { return temp.data(); }
  float *t_pressure() { return temp.data(); }
  float *divergence() { return temp.data() + pyr; }
  float *t_density() { return temp.data(); }

  explicit UniformOracle(const dcg_sim_params &p) {
    P = p;
    coarse_pairs = 2; level_pairs = 1; local_pairs = 5;  // fluid_simulation_uniform.cu:103,116,129
    gx = p.gx; gy = p.gy; gz = p.gz;
    N = (u64)gx * gy * gz;
    // fluid_simulation_uniform.cu:8-17
    u64 min_dim = (gx < gy && gx < gz) ? gx : (gy < gz ? gy : gz);
    u64 cell = 2;
    mip_levels = 1;
    while (gx % cell == 0 && gy % cell == 0 && gz % cell == 0 && cell * 4 <= min_dim) {
      mip_levels++;
      cell *= 2;
    }
    pyr = pyramid_cells(gx, gy, gz);
    density.assign(N, 0.f);
    velocity.assign(3 * N, 0.f);
    fluidity.assign(pyr, 0.f);
    pressure.assign(pyr, 0.f);
    temp.assign(3 * N, 0.f);
    reset();
  }

  u64 num_cells() const override { return N; }
  u64 lidx(int x, int y, int z) const { return ((u64)z * gy + y) * gx + x; }  // grid_math.cuh:10

  void reset() override {  // fluid_simulation_uniform.cu:81-88
    std::fill(density.begin(), density.end(), 0.f);
    std::fill(velocity.begin(), velocity.end(), 0.f);
    std::fill(pressure.begin(), pressure.end(), 0.f);
    std::fill(fluidity.begin(), fluidity.end(), 0.f);
    init();
  }
  void init() override {  // fluid_simulation_uniform.cu:76-79 + uniformgrid_structure.cu:15-21
    adapt_topology();
    std::fill(density.begin(), density.end(), 0.f);
    std::fill(velocity.begin(), velocity.end(), 0.f);
  }
  void adapt_topology() override {  // fluid_simulation_uniform.cu:143-147, uniformgrid_structure.cu:23-31
    for (int l = 0; l < mip_levels; l++) {
      const int scale = 1 << l, w = gx >> l, h = gy >> l, d = gz >> l;
      const u64 off = pyramid_offset(gx, gy, gz, scale);
#pragma omp parallel for schedule(static)
      for (int z = 0; z < d; z++)
        for (int y = 0; y < h; y++)
          for (int x = 0; x < w; x++)
            fluidity[off + ((u64)z * (gy / scale) + y) * (gx / scale) + x] = cell_fluidity(P, x, y, z, scale);
    }
  }

  // INIT_SAMPLE, uniformgrid_fluid.cu:7-48
  struct Sample { u64 id[8]; int x0, y0, z0; Corner8 c; };
  Sample sample(float px, float py, float pz) const {
    Sample s;
    const float x = px - .5f, y = py - .5f, z = pz - .5f;
    const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
    s.x0 = (int)xf; s.y0 = (int)yf; s.z0 = (int)zf;
    const int xc[2] = {iclamp(s.x0, 0, gx - 1), iclamp(s.x0 + 1, 0, gx - 1)};
    const int yc[2] = {iclamp(s.y0, 0, gy - 1), iclamp(s.y0 + 1, 0, gy - 1)};
    const int zc[2] = {iclamp(s.z0, 0, gz - 1), iclamp(s.z0 + 1, 0, gz - 1)};
    float f[8];
    for (int c = 0; c < 8; c++) {  // corner bits: (x<<2)|(y<<1)|z == the reference's i{x}{y}{z}
      s.id[c] = lidx(xc[(c >> 2) & 1], yc[(c >> 1) & 1], zc[c & 1]);
      f[c] = fluidity[s.id[c]];
    }
    s.c = corner_weights(f, x - xf, y - yf, z - zf);
    return s;
  }

  void advect_velocity() override {  // fluid_simulation_uniform.cu:90-94, uniformgrid_fluid.cu:50-67,88-95
    float *tv = t_velocity();
#pragma omp parallel for schedule(static)
    for (int z = 0; z < gz; z++)
      for (int y = 0; y < gy; y++)
        for (int x = 0; x < gx; x++) {
          const u64 i = lidx(x, y, z);
          const float bx = ((float)x + .5f) - velocity[3 * i + 0] * P.dt * P.rdx;
          const float by = ((float)y + .5f) - velocity[3 * i + 1] * P.dt * P.rdx;
          const float bz = ((float)z + .5f) - velocity[3 * i + 2] * P.dt * P.rdx;
          const Sample s = sample(bx, by, bz);
          V3 out{0.f, 0.f, 0.f};
          if (!(s.c.acc < 1e-6f)) {
            float vx[8], vy[8], vz[8];
            for (int c = 0; c < 8; c++) {
              const V3 v = velocity_bc(P, V3{velocity[3 * s.id[c]], velocity[3 * s.id[c] + 1], velocity[3 * s.id[c] + 2]},
                                       s.x0 + ((c >> 2) & 1), s.y0 + ((c >> 1) & 1), s.z0 + (c & 1), 1);
              vx[c] = v.x; vy[c] = v.y; vz[c] = v.z;
            }
            out = V3{blend8(vx, s.c.w), blend8(vy, s.c.w), blend8(vz, s.c.w)};
          }
          tv[3 * i] = out.x; tv[3 * i + 1] = out.y; tv[3 * i + 2] = out.z;
        }
    std::memcpy(velocity.data(), tv, 3 * N * sizeof(float));
  }

  void advect_density() override {  // fluid_simulation_uniform.cu:137-141, uniformgrid_fluid.cu:69-86,97-105
    float *tq = t_density();
#pragma omp parallel for schedule(static)
    for (int z = 0; z < gz; z++)
      for (int y = 0; y < gy; y++)
        for (int x = 0; x < gx; x++) {
          const u64 i = lidx(x, y, z);
          const float bx = ((float)x + .5f) - velocity[3 * i + 0] * P.dt * P.rdx;
          const float by = ((float)y + .5f) - velocity[3 * i + 1] * P.dt * P.rdx;
          const float bz = ((float)z + .5f) - velocity[3 * i + 2] * P.dt * P.rdx;
          const Sample s = sample(bx, by, bz);
          float out = 0.f;
          if (!(s.c.acc < 1e-6f)) {
            float q[8];
            for (int c = 0; c < 8; c++)
              q[c] = density_bc(P, density[s.id[c]], s.x0 + ((c >> 2) & 1), s.y0 + ((c >> 1) & 1), s.z0 + (c & 1), 1);
            out = blend8(q, s.c.w);
          }
          tq[i] = out;
        }
    std::memcpy(density.data(), tq, N * sizeof(float));
  }

  void calc_divergence() {  // uniformgrid_fluid.cu:107-132
    float *div = divergence(), *tp = t_pressure();
    // NB: t_pressure[idx] and divergence[idx] are distinct addresses (offset = pyr), and the
    // kernel's zero stores are overwritten/independent per cell, so a per-cell loop is exact.
#pragma omp parallel for schedule(static)
    for (int z = 0; z < gz; z++)
      for (int y = 0; y < gy; y++)
        for (int x = 0; x < gx; x++) {
          const u64 i = lidx(x, y, z);
          const u64 il = x > 0 ? i - 1 : i, ir = x < gx - 1 ? i + 1 : i;
          const u64 id = y > 0 ? i - gx : i, iu = y < gy - 1 ? i + gx : i;
          const u64 ib = z > 0 ? i - (u64)gx * gy : i, iff = z < gz - 1 ? i + (u64)gx * gy : i;
          pressure[i] = 0.f;
          tp[i] = 0.f;
          auto vel = [&](u64 j, int X, int Y, int Z) {
            return velocity_bc(P, V3{velocity[3 * j], velocity[3 * j + 1], velocity[3 * j + 2]}, X, Y, Z, 1);
          };
          const V3 vl = vel(il, x - 1, y, z), vr = vel(ir, x + 1, y, z);
          const V3 vd = vel(id, x, y - 1, z), vu = vel(iu, x, y + 1, z);
          const V3 vb = vel(ib, x, y, z - 1), vf = vel(iff, x, y, z + 1);
          div[i] = .5f * P.rdx *
                   (fluidity[ir] * vr.x - fluidity[il] * vl.x + fluidity[iu] * vu.y - fluidity[id] * vd.y +
                    fluidity[iff] * vf.z - fluidity[ib] * vb.z);
        }
  }

  void restrict_level(int l) {  // uniformgrid_fluid.cu:134-160
    const int scale = 1 << l, cs = scale / 2;
    const int w = gx / cs, h = gy / cs;  // child-level dims
    const u64 off = pyramid_offset(gx, gy, gz, scale), coff = pyramid_offset(gx, gy, gz, cs);
    float *div = divergence(), *tp = t_pressure();
    const int W = gx >> l, H = gy >> l, D = gz >> l;
    for (int z = 0; z < D; z++)
      for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
          const u64 i = off + ((u64)z * (gy / scale) + y) * (gx / scale) + x;
          const u64 c0 = coff + ((u64)(2 * z) * (gy / cs) + 2 * y) * (gx / cs) + 2 * x;
          const u64 i001 = c0 + 1, i010 = c0 + w, i011 = i010 + 1;
          const u64 i100 = c0 + (u64)w * h, i101 = i100 + 1, i110 = i100 + w, i111 = i110 + 1;
          pressure[i] = 0.f;
          tp[i] = 0.f;
          div[i] = .125f * (div[c0] + div[i001] + div[i010] + div[i011] + div[i100] + div[i101] + div[i110] + div[i111]);
        }
  }

  void jacobi(int l, const float *in, float *out) {  // calcPressure, uniformgrid_fluid.cu:162-192
    const int scale = 1 << l;
    const int w = gx / scale, h = gy / scale, d = gz / scale;
    const u64 off = pyramid_offset(gx, gy, gz, scale);
    const float alpha = P.dx * P.dx * scale * scale;
    const float *div = divergence();
    const int W = gx >> l, H = gy >> l, D = gz >> l;
#pragma omp parallel for schedule(static)
    for (int z = 0; z < D; z++)
      for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
          const u64 i = off + ((u64)z * h + y) * w + x;
          const u64 il = x > 0 ? i - 1 : i, ir = x < w - 1 ? i + 1 : i;
          const u64 id = y > 0 ? i - w : i, iu = y < h - 1 ? i + w : i;
          const u64 ib = z > 0 ? i - (u64)w * h : i, iff = z < d - 1 ? i + (u64)w * h : i;
          out[i] = (in[il] + in[ir] + in[id] + in[iu] + in[ib] + in[iff] - alpha * div[i]) / 6.f;
        }
  }
  void jacobi_pair(int l) {
    jacobi(l, pressure.data(), t_pressure());  // k_uniform_jacobi, :194-198
    jacobi(l, t_pressure(), pressure.data());  // k_uniform_jacobi_inv, :200-204
  }

  void prolongate(int l) {  // uniformgrid_fluid.cu:206-237
    const int scale = 1 << l;
    const int w = gx / scale, h = gy / scale, d = gz / scale;
    const u64 off = pyramid_offset(gx, gy, gz, scale), poff = pyramid_offset(gx, gy, gz, 2 * scale);
    float *p = pressure.data();
    for (int z = 0; z < d; z++)
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
          const u64 i = off + ((u64)z * h + y) * w + x;
          const u64 i000 = poff + ((u64)(z / 2) * (gy / (2 * scale)) + y / 2) * (gx / (2 * scale)) + x / 2;
          const int sx = (x == 0 || x == w - 1) ? 0 : 2 * (x % 2) - 1;
          const int sy = (y == 0 || y == h - 1) ? 0 : 2 * (y % 2) - 1;
          const int sz = (z == 0 || z == d - 1) ? 0 : 2 * (z % 2) - 1;
          // size_t + int arithmetic of the reference (wraps for negative steps) == signed offsets
          const int64_t ox = sx, oy = (int64_t)(sy * w / 2), oz = (int64_t)sz * (w / 2) * (h / 2);
          const u64 i001 = i000 + ox, i010 = i000 + oy, i011 = i010 + ox;
          const u64 i100 = i000 + oz, i101 = i100 + ox, i110 = i100 + (int64_t)sy * (w / 2), i111 = i110 + ox;
          p[i] = (27.f * p[i000] + 9.f * (p[i001] + p[i010] + p[i100]) + 3.f * (p[i011] + p[i101] + p[i110]) + p[i111]) / 64.f;
        }
  }

  void apply_pressure() {  // uniformgrid_fluid.cu:239-260
    const float alpha = .5f * P.rdx;
    const float *p = pressure.data();
#pragma omp parallel for schedule(static)
    for (int z = 0; z < gz; z++)
      for (int y = 0; y < gy; y++)
        for (int x = 0; x < gx; x++) {
          const u64 i = lidx(x, y, z);
          const u64 il = x > 0 ? i - 1 : i, ir = x < gx - 1 ? i + 1 : i;
          const u64 id = y > 0 ? i - gx : i, iu = y < gy - 1 ? i + gx : i;
          const u64 ib = z > 0 ? i - (u64)gx * gy : i, iff = z < gz - 1 ? i + (u64)gx * gy : i;
          const float pc = p[i];
          velocity[3 * i + 0] -= alpha * (fluidity[ir] * (p[ir] - pc) + fluidity[il] * (pc - p[il]));
          velocity[3 * i + 1] -= alpha * (fluidity[iu] * (p[iu] - pc) + fluidity[id] * (pc - p[id]));
          velocity[3 * i + 2] -= alpha * (fluidity[iff] * (p[iff] - pc) + fluidity[ib] * (pc - p[ib]));
        }
  }

  void project() override {  // fluid_simulation_uniform.cu:96-124
    calc_divergence();
    for (int l = 1; l < mip_levels; l++) restrict_level(l);
    for (int i = 0; i < coarse_pairs; i++) jacobi_pair(mip_levels - 1);
    for (int l = mip_levels - 2; l >= 0; l--) {
      prolongate(l);
      for (int i = 0; i < level_pairs; i++) jacobi_pair(l);
    }
    apply_pressure();
  }
  void project_local() override {  // fluid_simulation_uniform.cu:126-135
    calc_divergence();
    for (int i = 0; i < local_pairs; i++) jacobi_pair(0);
    apply_pressure();
  }

  float debug_stats() override {  // uniformgrid_structure.cu:33-43, fluid_simulation_uniform.cu:160-176
    const u64 bins = N / 256;
    float sum = 0.f;
    for (u64 b = 0; b < bins; b++) {
      float s = 0.f;
      for (u64 i = b * 256; i < b * 256 + 256; i++) s += density[i] * fluidity[i];
      sum += s;
    }
    return sum;
  }

  int get_field(int field, float *dst) override {
    switch (field) {
      case DCG_FIELD_DENSITY: std::memcpy(dst, density.data(), N * 4); return 0;
      case DCG_FIELD_VELOCITY: std::memcpy(dst, velocity.data(), 3 * N * 4); return 0;
      case DCG_FIELD_FLUIDITY: std::memcpy(dst, fluidity.data(), N * 4); return 0;
      case DCG_FIELD_PRESSURE: std::memcpy(dst, pressure.data(), N * 4); return 0;
      case DCG_FIELD_DIVERGENCE: std::memcpy(dst, divergence(), N * 4); return 0;
      case DCG_FIELD_T_PRESSURE: std::memcpy(dst, t_pressure(), N * 4); return 0;
    }
    return 1;
  }
};

// ===========================================================================
// DCGrid: src/dcgrid/*
// ===========================================================================
struct DCGridOracle : orc_sim {
  enum { BW = 4, BV = 64, AW = 6, AA = 36, AV = 216, SV = 8 };  // dcgrid.h:51-67
  enum : uint8_t { kClean = 0, kMoved = 1, kRefined = 2 };      // dcgrid.h:45-49

  int gx, gy, gz, levels = 0, sparse_levels = 0;
  u64 M = 0, num_cells_ = 0;
  std::vector<u64> max_blocks, full_blocks, loads, offsets, hash_size, hash_off, move_limit;
  std::vector<uint32_t> hash_key;
  std::vector<u64> hash_val;
  std::vector<int32_t> pos;  // 3*M
  std::vector<uint8_t> lvl, flags;
  std::vector<u64> apron, free_idx, parent, child;
  std::vector<float> density, velocity /*3*64M*/, fluidity, temp /*3*64M*/;
  // extensions: potential temperature, vapor, vorticity (3 floats / cell), MacCormack scratch
  std::vector<float> temperature, vapor, vort, mc_hat, mc_out;
  std::vector<float> block_scores, sub_scores;
  std::vector<u64> to_move, dest, touched;
  bool pool_error = false;
  // counters
  u64 n_adapt = 0, n_changed = 0, n_moved = 0, n_refined = 0, n_failed = 0;

  // aliases of temp (dcgrid_structure.cu:94-102)
  float *t_velocity() { return temp.data(); }
  float *divergence() { return temp.data(); }
  float *pressure() { return temp.data() + num_cells_; }
  float *t_pressure() { return temp.data() + 2 * num_cells_; }
  float *t_density() { return temp.data(); }

  DCGridOracle(const dcg_sim_params &p, u64 max_num_blocks) {
    P = p;
    is_dcgrid = true;
    level_pairs = 5; coarse_pairs = 5; local_pairs = 10;  // fluid_simulation_dcgrid.cu:274,283,301
    gx = p.gx; gy = p.gy; gz = p.gz;
    M = max_num_blocks;
    // fluid_simulation_dcgrid.cu:12-22
    const int min_dim = (gx < gy && gx < gz) ? gx : (gy < gz ? gy : gz);
    int cell = 2;
    levels = 1;
    while (gx % cell == 0 && gy % cell == 0 && gz % cell == 0 && cell * BW <= min_dim) {
      levels++;
      cell *= 2;
    }
    max_blocks.assign(levels, 0); full_blocks.assign(levels, 0); loads.assign(levels, 0);
    offsets.assign(levels, 0); move_limit.assign(levels, 0);
    for (int l = 0, cs = 1; l < levels; l++, cs *= 2)  // :29-32
      full_blocks[l] = (u64)idiv_up(gx, cs * BW) * idiv_up(gy, cs * BW) * idiv_up(gz, cs * BW);
    if (M < full_blocks[levels - 1]) { pool_error = true; return; }  // :35-39
    max_blocks[levels - 1] = full_blocks[levels - 1];
    u64 left = M - max_blocks[levels - 1];
    for (int l = levels - 2; l >= 0; l--) {  // :44-48
      max_blocks[l] = std::min(left / (u64)(l + 1), full_blocks[l]);
      left -= max_blocks[l];
    }
    if (max_blocks[0] == 0) { pool_error = true; return; }  // :50-53
    for (int l = 1; l < levels; l++) offsets[l] = offsets[l - 1] + max_blocks[l - 1];
    sparse_levels = 0;
    for (int l = 0; l < levels; l++)
      if (max_blocks[l] < full_blocks[l]) sparse_levels = l + 1;  // :66-69
    num_cells_ = M * BV;
    hash_size.assign(levels, 0); hash_off.assign(levels, 0);
    for (int l = 0; l < levels; l++) hash_size[l] = 4 * max_blocks[l];  // :101-102
    for (int l = 1; l < levels; l++) hash_off[l] = hash_off[l - 1] + hash_size[l - 1];
    hash_key.assign(4 * M, kHashEmpty);
    hash_val.assign(4 * M, kNone);
    pos.assign(3 * M, 0); lvl.assign(M, 0xFF); flags.assign(M, 0);
    apron.assign(M * AV, 0); free_idx.assign(M, 0); parent.assign(M, kNone); child.assign(8 * M, kNone);
    density.assign(num_cells_, 0.f); velocity.assign(3 * num_cells_, 0.f);
    fluidity.assign(num_cells_, 0.f); temp.assign(3 * num_cells_, 0.f);
    temperature.assign(num_cells_, 0.f); vapor.assign(num_cells_, 0.f); vort.assign(3 * num_cells_, 0.f);
    block_scores.assign(M, 0.f);
    sub_scores.assign(8 * M + 1, -FLT_MAX);  // +1: the reference reads one float past the end for slot M-1 (App. B-1)
    to_move.assign(M, 0); dest.assign(8 * M, 0); touched.assign(M, 0);
    reset();
  }

  u64 num_cells() const override { return num_cells_; }

  // ---- dcgrid_utils.cuh -------------------------------------------------
  static uint32_t bit_spread3(uint32_t d) {  // :83-90
    uint32_t r = 0;
    for (uint32_t mask = 1u; mask; mask <<= 3, d <<= 2) r |= d & mask;
    return r;
  }
  static uint32_t grid_hash(int x, int y, int z) {  // :92-96
    return (bit_spread3((uint32_t)(x / BW)) << 2) | (bit_spread3((uint32_t)(y / BW)) << 1) | bit_spread3((uint32_t)(z / BW));
  }
  static u64 spread(int d, int o) { return (u64)(((d & 1) | ((d << 2) & 8)) << o); }  // SPREAD, :63

  u64 ordered_index(int x, int y, int z, int level) const {  // :172-183 / :222-232
    const int extent = BW << level;
    const u64 rx = idiv_up(gx, extent), ry = idiv_up(gy, extent), rz = idiv_up(gz, extent);
    const int px = x / BW, py = y / BW, pz = z / BW;
    // int-vs-size_t comparison of the reference: a negative px converts to a huge value
    return ((u64)(int64_t)px >= rx || (u64)(int64_t)py >= ry || (u64)(int64_t)pz >= rz)
               ? kNone
               : offsets[level] + ((u64)px * ry + py) * rz + pz;
  }
  u64 hash_find(uint32_t key, int level) const {  // probe loop of :186-198
    if (hash_size[level] == 0) return kNone;
    const u64 o = hash_off[level];
    u64 slot = key % hash_size[level];
    const u64 slot0 = slot;
    do {
      if (hash_key[o + slot] == key) return hash_val[o + slot];
      if (hash_key[o + slot] == kHashEmpty) return kNone;
      slot = (slot + 127) % hash_size[level];
    } while (slot != slot0);
    return kNone;
  }
  u64 block_index(int x, int y, int z, int level) const {  // getBlockIndex, :169-199
    if (level >= sparse_levels) return ordered_index(x, y, z, level);
    return hash_find(grid_hash(x, y, z), level);
  }
  // getBlockIndexDeep, :201-233. (x,y,z) in level-`level` cells; returns the finest existing block of
  // level >= `level` that covers it and updates `level`.
  u64 block_index_deep(int x, int y, int z, int &level) const {
    if (level < sparse_levels) {
      uint32_t key = grid_hash(x, y, z);
      for (; level < sparse_levels; level++, key >>= 3, x /= 2, y /= 2, z /= 2) {
        const u64 r = hash_find(key, level);
        if (r != kNone) return r;
      }
    }
    const int sh = sparse_levels - level;  // 0 after the loop; >0 never reached with level>sparse (App. B-11)
    if (sh > 0) { x >>= sh; y >>= sh; z >>= sh; }
    level = sparse_levels;
    return ordered_index(x, y, z, level);
  }
  // insertBlock, :99-140 (sequential: CAS always sees the current table)
  u64 insert_block(int x, int y, int z, int level) {
    const uint32_t key = grid_hash(x, y, z);
    const u64 o = hash_off[level];
    if (hash_size[level] == 0) return kNone;
    u64 slot = key % hash_size[level];
    const u64 slot0 = slot;
    do {
      const uint32_t prev = hash_key[o + slot];
      if (prev == kHashEmpty) hash_key[o + slot] = key;
      if (prev == key) return hash_val[o + slot];
      if (prev == kHashEmpty) {
        const u64 allocated = loads[level]++;
        if (allocated >= max_blocks[level]) {
          loads[level]--;
          n_failed++;
          return kNone;
        }
        const u64 b = free_idx[offsets[level] + allocated];
        hash_val[o + slot] = b;
        pos[3 * b] = x; pos[3 * b + 1] = y; pos[3 * b + 2] = z;
        lvl[b] = (uint8_t)level;
        return b;
      }
      slot = (slot + 127) % hash_size[level];
    } while (slot != slot0);
    return kNone;
  }

  // ---- dcgrid_structure.cu -------------------------------------------------
  void init_apron_indices() {  // :6-28
    for (u64 b = 0; b < M; b++) {
      u64 *a = &apron[b * AV];
      for (int i = 0; i < AV; i++) a[i] = kNone;
      for (int c = 0; c < BV; c++) {
        const int X = 1 + (((c >> 5) & 1) << 1 | ((c >> 2) & 1));
        const int Y = 1 + (((c >> 4) & 1) << 1 | ((c >> 1) & 1));
        const int Z = 1 + (((c >> 3) & 1) << 1 | (c & 1));
        a[AA * X + AW * Y + Z] = b * BV + c;
      }
    }
  }

  void activate_level(int level) {  // :104-178
    const int scale = 1 << level;
    const int extent = BW << level;
    const int rx = idiv_up(gx, extent), ry = idiv_up(gy, extent), rz = idiv_up(gz, extent);
    for (int bz = 0; bz < rz; bz++)
      for (int by = 0; by < ry; by++)
        for (int bx = 0; bx < rx; bx++) {
          const int px = BW * bx, py = BW * by, pz = BW * bz;
          if (px * scale >= gx || py * scale >= gy || pz * scale >= gz) continue;
          const u64 b = block_index(px, py, pz, level);
          if (b >= offsets[level] + max_blocks[level]) continue;
          pos[3 * b] = px; pos[3 * b + 1] = py; pos[3 * b + 2] = pz;
          lvl[b] = (uint8_t)level;
          if (level < levels - 1)
            parent[b] = 8 * block_index(px / 2, py / 2, pz / 2, level + 1) + (px % (2 * BW)) + (py % (2 * BW)) / 2 +
                        (pz % (2 * BW)) / 4;
          if (level - 1 >= sparse_levels)
            for (int c = 0; c < 8; c++)
              child[8 * b + c] = block_index(2 * px + BW * ((c / 4) % 2), 2 * py + BW * ((c / 2) % 2), 2 * pz + BW * (c % 2), level - 1);
          u64 *a = &apron[b * AV];
          const int i0 = px > 0 ? 0 : 1, i1 = AW - ((px + BW) * scale >= gx ? 1 : 0);
          const int j0 = py > 0 ? 0 : 1, j1 = AW - ((py + BW) * scale >= gy ? 1 : 0);
          const int k0 = pz > 0 ? 0 : 1, k1 = AW - ((pz + BW) * scale >= gz ? 1 : 0);
          for (int i = i0; i < i1; i++)
            for (int j = j0; j < j1; j++)
              for (int k = k0; k < k1; k++) {
                const u64 nb = block_index(px + i - 1, py + j - 1, pz + k - 1, level);
                const u64 *na = &apron[nb * AV];
                a[AA * i + AW * j + k] =
                    na[AA * (1 + ((i - 1 + BW) % BW)) + AW * (1 + ((j - 1 + BW) % BW)) + (1 + ((k - 1 + BW) % BW))];
              }
          for (int i = 0; i < AW; i++)
            for (int j = 0; j < AW; j++)
              for (int k = 0; k < AW; k++)
                if (a[AA * i + AW * j + k] == kNone)
                  a[AA * i + AW * j + k] = a[AA * iclamp(i, 1, BW) + AW * iclamp(j, 1, BW) + iclamp(k, 1, BW)];
          for (int i = 1; i <= BW; i++)
            for (int j = 1; j <= BW; j++)
              for (int k = 1; k <= BW; k++) {
                const u64 c = a[AA * i + AW * j + k];
                density[c] = 0.f;
                velocity[3 * c] = velocity[3 * c + 1] = velocity[3 * c + 2] = 0.f;
                fluidity[c] = cell_fluidity(P, E, px + i - 1, py + j - 1, pz + k - 1, scale);
                temperature[c] = ambient_theta(E, cell_height(P, py + j - 1, scale));  // extension: the ambient profile
                vapor[c] = E.ambient_vapor;
              }
        }
  }

  void refresh_apron_indices() {  // :30-92
    for (u64 b = 0; b < M; b++) {
      if (lvl[b] == 0xFF) continue;
      const int level = lvl[b];
      const int p0x = pos[3 * b], p0y = pos[3 * b + 1], p0z = pos[3 * b + 2];
      for (int ai = 0; ai < AV; ai++) {
        const int i = (ai / AA) % AW, j = (ai / AW) % AW, k = ai % AW;
        if (i % (AW - 1) != 0 && j % (AW - 1) != 0 && k % (AW - 1) != 0) continue;  // interior
        u64 &entry = apron[b * AV + ai];
        u64 nb = kNone;
        if (flags[b] & kMoved) {
          const int nx = p0x + i - 1, ny = p0y + j - 1, nz = p0z + k - 1;
          const int scale = 1 << level;
          if (nx < 0 || ny < 0 || nz < 0 || nx * scale >= gx || ny * scale >= gy || nz * scale >= gz) {
            // :54-60 — apron coords where cell coords are meant; replicated
            entry = b * BV + spread(iclamp(i, 1, BW), 2) + spread(iclamp(j, 1, BW), 1) + spread(iclamp(k, 1, BW), 0);
            continue;
          }
          int nl = level;
          nb = block_index_deep(nx, ny, nz, nl);
        } else {
          const u64 prev = entry / BV;
          const uint8_t cf = flags[prev];
          if (cf & kMoved) {
            int nl = level;
            nb = block_index_deep(p0x + i - 1, p0y + j - 1, p0z + k - 1, nl);
          } else if ((cf & kRefined) && lvl[prev] > level) {
            nb = child[entry / SV];
          }
        }
        if (nb == kNone) continue;
        const int s = 1 << (lvl[nb] - level);
        const int x = (p0x + i - 1) / s - pos[3 * nb];
        const int y = (p0y + j - 1) / s - pos[3 * nb + 1];
        const int z = (p0z + k - 1) / s - pos[3 * nb + 2];
        entry = nb * BV + spread(x, 2) + spread(y, 1) + spread(z, 0);
      }
    }
  }

  // accumulate<T>, :188-205, with USE_SUBBLOCK_INDEX_LEVEL (dcgrid_utils.cuh:20-29)
  void accumulate(float *ch, int comps, int level) {
    const u64 first = 8 * offsets[level], count = 8 * max_blocks[level];
#pragma omp parallel for schedule(static)
    for (u64 t = 0; t < count; t++) {
      const u64 sb = first + t, b = sb / 8;
      if (lvl[b] != level) continue;
      const u64 ps = parent[b];
      if (ps == kNone) continue;
      const u64 pc = SV * ps + (sb % 8);
      for (int k = 0; k < comps; k++) {
        float acc = 0.f;
        for (int i = 0; i < SV; i++) acc += ch[comps * (SV * sb + i) + k];
        ch[comps * pc + k] = acc * .125f;
      }
    }
  }
  void accumulate_velocity() { for (int l = 0; l < levels - 1; l++) accumulate(velocity.data(), 3, l); }  // fluid_simulation_dcgrid.cu:496-501
  void accumulate_density() { for (int l = 0; l < levels - 1; l++) accumulate(density.data(), 1, l); }    // :503-508
  void accumulate_divergence() { for (int l = 0; l < levels - 1; l++) accumulate(divergence(), 1, l); }  // :510-515

  // ---- dcgrid_fluid.cu ---------------------------------------------------
  static int cell_x(u64 c) { return (int)((((c >> 5) & 1) << 1) | ((c >> 2) & 1)); }
  static int cell_y(u64 c) { return (int)((((c >> 4) & 1) << 1) | ((c >> 1) & 1)); }
  static int cell_z(u64 c) { return (int)((((c >> 3) & 1) << 1) | (c & 1)); }
  static int apron_of(u64 c) { return AA * (1 + cell_x(c)) + AW * (1 + cell_y(c)) + (1 + cell_z(c)); }  // USE_APRON_INDEX

  struct Sample { u64 id[8]; int x0, y0, z0, scale; Corner8 c; };
  Sample sample(float px, float py, float pz) const {  // INIT_SAMPLE, dcgrid_fluid.cu:7-72
    Sample s;
    int ix = iclamp((int)floorf(px), 0, gx - 1), iy = iclamp((int)floorf(py), 0, gy - 1), iz = iclamp((int)floorf(pz), 0, gz - 1);
    int level = 0;
    const u64 b = block_index_deep(ix, iy, iz, level);
    s.scale = 1 << level;
    const float inv = 1.f / s.scale;
    const float x = px * inv - .5f, y = py * inv - .5f, z = pz * inv - .5f;
    const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
    s.x0 = (int)xf; s.y0 = (int)yf; s.z0 = (int)zf;
    const int i = iclamp(s.x0 + 1 - pos[3 * b], 0, AW - 2);
    const int j = iclamp(s.y0 + 1 - pos[3 * b + 1], 0, AW - 2);
    const int k = iclamp(s.z0 + 1 - pos[3 * b + 2], 0, AW - 2);
    const u64 *a = &apron[AV * b + AA * i + AW * j + k];
    float f[8];
    for (int c = 0; c < 8; c++) {
      s.id[c] = a[AA * ((c >> 2) & 1) + AW * ((c >> 1) & 1) + (c & 1)];
      f[c] = fluidity[s.id[c]];
    }
    s.c = corner_weights(f, x - xf, y - yf, z - zf);
    return s;
  }

  // backtraced (sign = -1) or forward-traced (sign = +1) position of cell c of block b along the cell's own velocity
  void trace(u64 b, u64 c, float sign, float &bx, float &by, float &bz) const {
    const float scale = (float)(1 << lvl[b]);
    const float alpha = P.dt * P.rdx;
    const float fx = (float)(pos[3 * b] | cell_x(c)), fy = (float)(pos[3 * b + 1] | cell_y(c)), fz = (float)(pos[3 * b + 2] | cell_z(c));
    if (sign < 0.f) {  // the reference's expression, dcgrid_fluid.cu:119-123
      bx = (fx + .5f) * scale - velocity[3 * c] * alpha;
      by = (fy + .5f) * scale - velocity[3 * c + 1] * alpha;
      bz = (fz + .5f) * scale - velocity[3 * c + 2] * alpha;
    } else {
      bx = (fx + .5f) * scale + velocity[3 * c] * alpha;
      by = (fy + .5f) * scale + velocity[3 * c + 1] * alpha;
      bz = (fz + .5f) * scale + velocity[3 * c + 2] * alpha;
    }
  }
  // semi-Lagrangian gather of a `comps`-component field through the cell's own velocity (k_dcgrid_advect_velocity /
  // _density, dcgrid_fluid.cu:74-144): out[c] for every cell of every active block, 0 for non-leaf cells.
  // bc(values[comps], x, y, z, scale) substitutes boundary values per corner; fb(values, y, scale) is what a leaf cell
  // gets when its sample has no fluid weight (the reference: 0; the extension scalars: their ambient value).
  template <class BC, class FB>
  void gather_sl(const float *phi, int comps, float *out, BC bc, FB fb) const {
#pragma omp parallel for schedule(dynamic, 64)
    for (u64 b = 0; b < M; b++) {
      if (lvl[b] == 0xFF) continue;
      for (u64 c = b * BV; c < (b + 1) * BV; c++) {
        float o[3] = {0.f, 0.f, 0.f};
        if (child[c >> 3] == kNone) {
          fb(o, pos[3 * b + 1] | cell_y(c), 1 << lvl[b]);
          float bx, by, bz;
          trace(b, c, -1.f, bx, by, bz);
          const Sample s = sample(bx, by, bz);
          if (!(s.c.acc < 1e-6f)) {
            float v8[3][8];
            for (int q = 0; q < 8; q++) {
              float v[3] = {0.f, 0.f, 0.f};
              for (int k = 0; k < comps; k++) v[k] = phi[comps * s.id[q] + k];
              bc(v, s.x0 + ((q >> 2) & 1), s.y0 + ((q >> 1) & 1), s.z0 + (q & 1), s.scale);
              for (int k = 0; k < comps; k++) v8[k][q] = v[k];
            }
            for (int k = 0; k < comps; k++) o[k] = blend8(v8[k], s.c.w);
          }
        }
        for (int k = 0; k < comps; k++) out[comps * c + k] = o[k];
      }
    }
  }
  void accumulate_all(float *ch, int comps) { for (int l = 0; l < levels - 1; l++) accumulate(ch, comps, l); }

  // EXTENSION (dcg_ext_params.advection == 1): MacCormack on top of the semi-Lagrangian gather.
  //   hat  = SL(phi)                                   (then restricted, like every advected field)
  //   back = SL of `hat` along the REVERSED trajectory (cell centre + v dt)
  //   out  = clamp(hat + .5 (phi - back), min, max of the 8 boundary-substituted corners of the forward sample)
  // falling back to hat where either sample has no fluid weight.  Trajectories use the pre-advection velocity.
  template <class BC, class FB>
  void maccormack(float *phi, int comps, BC bc, FB fb) {
    mc_hat.resize(3 * num_cells_); mc_out.resize(3 * num_cells_);
    float *hat = mc_hat.data(), *out = mc_out.data();
    std::fill(mc_hat.begin(), mc_hat.end(), 0.f);
    gather_sl(phi, comps, hat, bc, fb);
    accumulate_all(hat, comps);
#pragma omp parallel for schedule(dynamic, 64)
    for (u64 b = 0; b < M; b++) {
      if (lvl[b] == 0xFF) continue;
      for (u64 c = b * BV; c < (b + 1) * BV; c++) {
        float o[3] = {0.f, 0.f, 0.f};
        if (child[c >> 3] == kNone) {
          fb(o, pos[3 * b + 1] | cell_y(c), 1 << lvl[b]);
          float bx, by, bz;
          trace(b, c, -1.f, bx, by, bz);
          const Sample s = sample(bx, by, bz);
          if (!(s.c.acc < 1e-6f)) {
            float mn[3], mx[3];
            for (int q = 0; q < 8; q++) {
              float v[3] = {0.f, 0.f, 0.f};
              for (int k = 0; k < comps; k++) v[k] = phi[comps * s.id[q] + k];
              bc(v, s.x0 + ((q >> 2) & 1), s.y0 + ((q >> 1) & 1), s.z0 + (q & 1), s.scale);
              for (int k = 0; k < comps; k++) {
                mn[k] = q == 0 ? v[k] : fminf(mn[k], v[k]);
                mx[k] = q == 0 ? v[k] : fmaxf(mx[k], v[k]);
              }
            }
            for (int k = 0; k < comps; k++) o[k] = hat[comps * c + k];
            float fx, fy, fz;
            trace(b, c, 1.f, fx, fy, fz);
            const Sample sf = sample(fx, fy, fz);
            if (!(sf.c.acc < 1e-6f)) {
              float v8[3][8];
              for (int q = 0; q < 8; q++) {
                float v[3] = {0.f, 0.f, 0.f};
                for (int k = 0; k < comps; k++) v[k] = hat[comps * sf.id[q] + k];
                bc(v, sf.x0 + ((q >> 2) & 1), sf.y0 + ((q >> 1) & 1), sf.z0 + (q & 1), sf.scale);
                for (int k = 0; k < comps; k++) v8[k][q] = v[k];
              }
              for (int k = 0; k < comps; k++) {
                const float back = blend8(v8[k], sf.c.w);
                const float r = hat[comps * c + k] + .5f * (phi[comps * c + k] - back);
                o[k] = fminf(fmaxf(r, mn[k]), mx[k]);
              }
            }
          }
        }
        for (int k = 0; k < comps; k++) out[comps * c + k] = o[k];
      }
    }
    for (u64 b = 0; b < M; b++)
      if (lvl[b] != 0xFF) std::memcpy(phi + comps * b * BV, out + comps * b * BV, comps * BV * sizeof(float));
    accumulate_all(phi, comps);
  }

  void advect_velocity() override {  // fluid_simulation_dcgrid.cu:263-268, dcgrid_fluid.cu:74-91,112-127
    auto bc = [this](float *v, int x, int y, int z, int scale) {
      const V3 r = velocity_bc(P, V3{v[0], v[1], v[2]}, x, y, z, scale);
      v[0] = r.x; v[1] = r.y; v[2] = r.z;
    };
    auto zero = [](float *, int, int) {};
    if (E.advection == 1) { maccormack(velocity.data(), 3, bc, zero); return; }
    float *tv = t_velocity();
    gather_sl(velocity.data(), 3, tv, bc, zero);
    std::memcpy(velocity.data(), tv, 3 * num_cells_ * sizeof(float));  // whole-pool D2D copy, :265-266
    accumulate_velocity();
  }

  template <class BC, class FB>
  void advect_scalar(std::vector<float> &phi, BC bc, FB fb) {
    if (E.advection == 1) { maccormack(phi.data(), 1, bc, fb); return; }
    float *tq = t_density();
    gather_sl(phi.data(), 1, tq, bc, fb);
    std::memcpy(phi.data(), tq, num_cells_ * sizeof(float));  // :315-316
    accumulate_all(phi.data(), 1);
  }
  void advect_density() override {  // fluid_simulation_dcgrid.cu:313-318, dcgrid_fluid.cu:93-110,129-144
    advect_scalar(density, [this](float *v, int x, int y, int z, int scale) { v[0] = density_bc(P, v[0], x, y, z, scale); }, [](float *, int, int) {});
    if (E.sources) {  // extension: temperature and vapor ride along (same trajectories, same weights); a cell whose
                      // sample has no fluid weight (inside a solid) takes the ambient value instead of the reference's 0
      advect_scalar(temperature, [this](float *v, int x, int y, int z, int scale) { v[0] = temperature_bc(P, E, v[0], x, y, z, scale); },
                    [this](float *o, int y, int scale) { o[0] = ambient_theta(E, cell_height(P, y, scale)); });
      advect_scalar(vapor, [this](float *v, int x, int y, int z, int scale) { v[0] = vapor_bc(P, E, v[0], x, y, z, scale); },
                    [this](float *o, int, int) { o[0] = E.ambient_vapor; });
    }
  }

  void calc_vorticity() {  // dcgrid_fluid.cu:146-172 — result is dead in this snapshot (overwritten by divergence),
                           // kept because it scribbles over `temporary` exactly like the reference
    // extensions that consume the vorticity keep it in a buffer of its own (the reference's aliases `temporary`)
    float *vo = (E.score_mode == 1 || E.sources) ? vort.data() : temp.data();
    for (u64 b = 0; b < M; b++) {
      if (lvl[b] == 0xFF) continue;
      const u64 *a = &apron[AV * b];
      const int scale = 1 << lvl[b];
      const float alpha = .5f * P.rdx / scale;
      for (u64 c = b * BV; c < (b + 1) * BV; c++) {
        const int ai = apron_of(c);
        const u64 il = a[ai - AA], ir = a[ai + AA], id = a[ai - AW], iu = a[ai + AW], ib = a[ai - 1], iff = a[ai + 1];
        const float wl = fluidity[il], wr = fluidity[ir], wd = fluidity[id], wu = fluidity[iu], wb = fluidity[ib], wf = fluidity[iff];
        const float *vl = &velocity[3 * il], *vr = &velocity[3 * ir], *vd = &velocity[3 * id], *vu = &velocity[3 * iu],
                    *vb = &velocity[3 * ib], *vf = &velocity[3 * iff];
        vo[3 * c] = alpha * ((wu * vu[2] - wd * vd[2]) - (wf * vf[1] - wb * vb[1]));
        vo[3 * c + 1] = alpha * ((wf * vf[0] - wb * vb[0]) - (wr * vr[2] - wl * vl[2]));
        vo[3 * c + 2] = alpha * ((wr * vr[1] - wl * vl[1]) - (wu * vu[0] - wd * vd[0]));
      }
    }
  }

  void calc_divergence() {  // dcgrid_fluid.cu:174-230
    float *div = divergence(), *p = pressure(), *tp = t_pressure();
#pragma omp parallel for schedule(dynamic, 64)
    for (u64 b = 0; b < M; b++) {
      if (lvl[b] == 0xFF) continue;
      const u64 *a = &apron[AV * b];
      const int scale = 1 << lvl[b];
      const float alpha = .5f * P.rdx / scale;
      for (u64 c = b * BV; c < (b + 1) * BV; c++) {
        p[c] = 0.f; tp[c] = 0.f; div[c] = 0.f;
        if (child[c >> 3] != kNone) continue;
        const int x = pos[3 * b] | cell_x(c), y = pos[3 * b + 1] | cell_y(c), z = pos[3 * b + 2] | cell_z(c);
        const int ai = apron_of(c);
        const u64 il = a[ai - AA], ir = a[ai + AA], id = a[ai - AW], iu = a[ai + AW], ib = a[ai - 1], iff = a[ai + 1];
        auto vel = [&](u64 j, int X, int Y, int Z) {
          return velocity_bc(P, V3{velocity[3 * j], velocity[3 * j + 1], velocity[3 * j + 2]}, X, Y, Z, scale);
        };
        // the SMEM staging of the reference applies the BC only on block faces; in-block neighbours
        // are in-domain so the BC is the identity there
        const V3 vl = vel(il, x - 1, y, z), vr = vel(ir, x + 1, y, z), vd = vel(id, x, y - 1, z), vu = vel(iu, x, y + 1, z),
                 vb = vel(ib, x, y, z - 1), vf = vel(iff, x, y, z + 1);
        div[c] = alpha * (fluidity[ir] * vr.x - fluidity[il] * vl.x + fluidity[iu] * vu.y - fluidity[id] * vd.y +
                          fluidity[iff] * vf.z - fluidity[ib] * vb.z);
      }
    }
  }

  void apply_pressure() {  // dcgrid_fluid.cu:232-259
    const float *p = pressure();
#pragma omp parallel for schedule(dynamic, 64)
    for (u64 b = 0; b < M; b++) {
      if (lvl[b] == 0xFF) continue;
      const u64 *a = &apron[AV * b];
      const int scale = 1 << lvl[b];
      const float alpha = .5f * P.rdx / scale;
      for (u64 c = b * BV; c < (b + 1) * BV; c++) {
        if (child[c >> 3] != kNone) continue;
        const int ai = apron_of(c);
        const u64 il = a[ai - AA], ir = a[ai + AA], id = a[ai - AW], iu = a[ai + AW], ib = a[ai - 1], iff = a[ai + 1];
        const float pc = p[c];
        velocity[3 * c] -= alpha * (fluidity[ir] * (p[ir] - pc) + fluidity[il] * (pc - p[il]));
        velocity[3 * c + 1] -= alpha * (fluidity[iu] * (p[iu] - pc) + fluidity[id] * (pc - p[id]));
        velocity[3 * c + 2] -= alpha * (fluidity[iff] * (p[iff] - pc) + fluidity[ib] * (pc - p[ib]));
      }
    }
  }

  // ---- dcgrid_multigrid_solver.cu ---------------------------------------------
  void jacobi(int level, const float *in, float *out) {  // :5-41 (jacobi: in=pressure; jacobi_inv: in=t_pressure)
    const float alpha = (float)((1 << level) * (1 << level)) * P.dx * P.dx;
    const float *div = divergence();
    const u64 first = offsets[level], count = max_blocks[level];
#pragma omp parallel for schedule(static)
    for (u64 t = 0; t < count; t++) {
      const u64 b = first + t;
      if (lvl[b] != level) continue;
      const u64 *a = &apron[AV * b];
      for (u64 c = b * BV; c < (b + 1) * BV; c++) {
        const int ai = apron_of(c);
        out[c] = (in[a[ai - AA]] + in[a[ai + AA]] + in[a[ai - AW]] + in[a[ai + AW]] + in[a[ai - 1]] + in[a[ai + 1]] -
                  alpha * div[c]) / 6.f;
      }
    }
  }
  void jacobi_pair(int level) {
    jacobi(level, pressure(), t_pressure());
    jacobi(level, t_pressure(), pressure());
  }
  void prolongate(int level) {  // :43-76
    float *p = pressure();
    const u64 first = offsets[level], count = max_blocks[level];
    for (u64 t = 0; t < count; t++) {
      const u64 b = first + t;
      if (lvl[b] != level) continue;
      const u64 ps = parent[b];
      if (ps == kNone) continue;
      const u64 *pa = &apron[(ps / 8) * AV];
      for (u64 c = b * BV; c < (b + 1) * BV; c++) {
        const int x = pos[3 * b] | cell_x(c), y = pos[3 * b + 1] | cell_y(c), z = pos[3 * b + 2] | cell_z(c);
        const int idx = AA * (1 + (x / 2) % BW) + AW * (1 + (y / 2) % BW) + (1 + (z / 2) % BW);
        const int i = x % 2 ? AA : -AA, j = y % 2 ? AW : -AW, k = z % 2 ? 1 : -1;
        const float p000 = p[pa[idx]], p001 = p[pa[idx + k]], p010 = p[pa[idx + j]], p100 = p[pa[idx + i]];
        const float p011 = p[pa[idx + j + k]], p101 = p[pa[idx + i + k]], p110 = p[pa[idx + i + j]], p111 = p[pa[idx + i + j + k]];
        p[c] = (27.f * p000 + 9.f * (p001 + p010 + p100) + 3.f * (p011 + p101 + p110) + p111) / 64.f;
      }
    }
  }

  void project() override {  // fluid_simulation_dcgrid.cu:270-294
    calc_divergence();
    accumulate_divergence();
    for (int i = 0; i < coarse_pairs; i++) jacobi_pair(levels - 1);
    for (int l = levels - 2; l >= 0; l--) {
      prolongate(l);
      for (int i = 0; i < level_pairs; i++) jacobi_pair(l);
    }
    apply_pressure();
    accumulate_velocity();
  }
  void project_local() override {  // :296-311
    calc_divergence();
    accumulate_divergence();
    for (int l = levels - 1; l >= 0; l--)
      for (int i = 0; i < local_pairs; i++) jacobi_pair(l);
    apply_pressure();
    accumulate_velocity();
  }

  // ---- dcgrid_adaptation.cu ----------------------------------------------
  bool finer_level_full(int level) const {  // :19-21, :52-54
    return (level == 0 && loads[0] == full_blocks[0]) || (level > 0 && loads[level - 1] == full_blocks[level - 1]);
  }
  void calc_subblock_scores() {  // :10-40
    for (u64 sb = 0; sb < 8 * M; sb++) {
      const u64 b = sb / 8;
      const uint8_t level = lvl[b];
      if (child[sb] != kNone || level == 0xFF || finer_level_full(level)) {
        sub_scores[sb] = -FLT_MAX;
        continue;
      }
      if (E.score_mode == 1) {  // EXTENSION: the commented-out lines :36-39 with calcCellScore (:6-8) = |vorticity|
        float acc = 0.f;
        for (int i = 0; i < SV; i++) {
          const float *w = &vort[3 * (SV * sb + i)];
          acc += sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        }
        sub_scores[sb] = acc;
        continue;
      }
      const float s = (float)(1 << level);
      const float px = s * ((float)pos[3 * b] + 2.f * (float)((sb >> 2) & 1) + 1.f);
      const float py = s * ((float)pos[3 * b + 1] + 2.f * (float)((sb >> 1) & 1) + 1.f);
      const float pz = s * ((float)pos[3 * b + 2] + 2.f * (float)(sb & 1) + 1.f);
      const float ex = px - .5f * P.gx, ey = py - .45f * P.gy, ez = pz - .5f * P.gz;
      const float d = sqrtf(ex * ex + ey * ey + ez * ez);
      sub_scores[sb] = d < .2f * P.gx ? 0.f : P.gx / d;
    }
  }
  void accumulate_subblock_scores() {  // :42-63 — sums NINE floats; replicated
    for (u64 b = 0; b < M; b++) {
      block_scores[b] = -FLT_MAX;
      const uint8_t level = lvl[b];
      if (level == 0xFF || finer_level_full(level)) continue;
      const float *s = &sub_scores[8 * b];
      if (s[0] > 0.f && s[1] > 0.f && s[2] > 0.f && s[3] > 0.f && s[4] > 0.f && s[5] > 0.f && s[6] > 0.f && s[7] > 0.f)
        block_scores[b] = .125f * (s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7] + s[8]);
    }
  }
  void move_blocks_kernel(u64 n) {  // :65-90
    for (u64 t = 0; t < n; t++) {
      const u64 b = to_move[t];
      child[parent[b]] = kNone;
      const u64 nps = dest[t];
      child[nps] = b;
      parent[b] = nps;
      const u64 npb = nps / 8;
      pos[3 * b] = 2 * pos[3 * npb] + BW * (int)((nps / 4) % 2);
      pos[3 * b + 1] = 2 * pos[3 * npb + 1] + BW * (int)((nps / 2) % 2);
      pos[3 * b + 2] = 2 * pos[3 * npb + 2] + BW * (int)(nps % 2);
      flags[b] |= kMoved;
      flags[npb] |= kRefined;
    }
  }
  void refill_hash_table() {  // :180-202
    for (u64 b = 0; b < M; b++) {
      if (lvl[b] == 0xFF) continue;
      const int level = lvl[b];
      if (hash_size[level] == 0) continue;
      const uint32_t key = grid_hash(pos[3 * b], pos[3 * b + 1], pos[3 * b + 2]);
      const u64 o = hash_off[level];
      u64 slot = key % hash_size[level];
      const u64 slot0 = slot;
      do {
        const uint32_t prev = hash_key[o + slot];
        if (prev == kHashEmpty) hash_key[o + slot] = key;
        if (prev == kHashEmpty || prev == key) { hash_val[o + slot] = b; break; }
        slot = (slot + 127) % hash_size[level];
      } while (slot != slot0);
    }
  }
  void refine_subblocks_kernel(u64 n, u64 num_touched) {  // :145-178
    for (u64 rank = 0; rank < n; rank++) {
      const u64 sb = dest[rank];
      if (sb == kNone) continue;
      const u64 pb = sb / 8;
      const int level = lvl[pb];
      const int cx = pos[3 * pb] * 2 + (int)((sb / 4) % 2) * BW;
      const int cy = pos[3 * pb + 1] * 2 + (int)((sb / 2) % 2) * BW;
      const int cz = pos[3 * pb + 2] * 2 + (int)(sb % 2) * BW;
      const u64 cb = insert_block(cx, cy, cz, level - 1);
      if (cb == kNone) continue;
      parent[cb] = sb;
      child[sb] = cb;
      flags[pb] |= kRefined;
      flags[cb] |= kMoved;
      touched[num_touched + rank] = cb;
    }
  }
  void propagate_values(u64 num_touched, int level) {  // :92-143
    const int scale = 1 << level;
    for (u64 t = 0; t < num_touched; t++) {
      const u64 b = touched[t];
      if (lvl[b] != level) continue;
      const u64 ps = parent[b];
      if (ps == kNone) continue;
      const u64 *pa = &apron[(ps / 8) * AV];
      for (u64 c = b * BV; c < (b + 1) * BV; c++) {
        const int x = pos[3 * b] | cell_x(c), y = pos[3 * b + 1] | cell_y(c), z = pos[3 * b + 2] | cell_z(c);
        const int idx = AA * (1 + (x / 2) % BW) + AW * (1 + (y / 2) % BW) + (1 + (z / 2) % BW);
        const int i = x % 2 ? AA : -AA, j = y % 2 ? AW : -AW, k = z % 2 ? 1 : -1;
        const u64 i000 = pa[idx], i001 = pa[idx + k], i010 = pa[idx + j], i100 = pa[idx + i];
        const u64 i011 = pa[idx + j + k], i101 = pa[idx + i + k], i110 = pa[idx + i + j], i111 = pa[idx + i + j + k];
        density[c] = ((27.f / 64.f) * density[i000] + (9.f / 64.f) * (density[i001] + density[i010] + density[i100]) +
                      (3.f / 64.f) * (density[i011] + density[i101] + density[i110]) + (1.f / 64.f) * density[i111]);
        for (int q = 0; q < 3; q++) {
          const float *v = velocity.data() + q;
          velocity[3 * c + q] = ((27.f / 64.f) * v[3 * i000] + (9.f / 64.f) * (v[3 * i001] + v[3 * i010] + v[3 * i100]) +
                                 (3.f / 64.f) * (v[3 * i011] + v[3 * i101] + v[3 * i110]) + (1.f / 64.f) * v[3 * i111]);
        }
        fluidity[c] = cell_fluidity(P, E, x, y, z, scale);
        for (std::vector<float> *f : {&temperature, &vapor}) {  // extension scalars: interpolated like the density
          const float *a = f->data();
          (*f)[c] = ((27.f / 64.f) * a[i000] + (9.f / 64.f) * (a[i001] + a[i010] + a[i100]) + (3.f / 64.f) * (a[i011] + a[i101] + a[i110]) +
                     (1.f / 64.f) * a[i111]);
        }
      }
    }
  }

  // EXTENSION (dcg_ext_params.sources): condensation, buoyancy and vorticity confinement in one pass over the
  // leaf cells, between adaptTopology and project; then the touched fields are restricted.
  void apply_sources() override {
    if (!E.sources) return;
    calc_vorticity();  // on the topology the projection will see
    auto wlen = [this](u64 i) { const float *w = &vort[3 * i]; return sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]); };
#pragma omp parallel for schedule(dynamic, 64)
    for (u64 b = 0; b < M; b++) {
      if (lvl[b] == 0xFF) continue;
      const u64 *a = &apron[AV * b];
      const int scale = 1 << lvl[b];
      const float alpha = .5f * P.rdx / scale;
      for (u64 c = b * BV; c < (b + 1) * BV; c++) {
        if (child[c >> 3] != kNone) continue;
        const int y = pos[3 * b + 1] | cell_y(c);
        const float h = cell_height(P, y, scale);
        float th = temperature[c], qv = vapor[c], qc = density[c];
        // condensation (dq > 0) / evaporation (dq < 0, limited by the condensed water present)
        const float tabs = th - E.adiabatic_lapse * h;
        const float qs = fmaxf(0.f, E.saturation_base + E.saturation_slope * (tabs - E.ambient_temperature));
        float dq = E.condensation_rate * (qv - qs);
        dq = fmaxf(dq, -qc);
        qv = qv - dq;
        qc = qc + dq;
        th = th + E.latent_heat * dq;
        // buoyancy
        const float tha = ambient_theta(E, h);
        const float lift = E.buoyancy * ((th - tha) / E.ambient_temperature) + E.vapor_buoyancy * qv - E.smoke_weight * qc;
        // vorticity confinement: eps * cell size * (N x omega), N = grad|omega| / |grad|omega||
        const int ai = apron_of(c);
        const float gx_ = alpha * (wlen(a[ai + AA]) - wlen(a[ai - AA]));
        const float gy_ = alpha * (wlen(a[ai + AW]) - wlen(a[ai - AW]));
        const float gz_ = alpha * (wlen(a[ai + 1]) - wlen(a[ai - 1]));
        const float glen = sqrtf(gx_ * gx_ + gy_ * gy_ + gz_ * gz_);
        float fx = 0.f, fy = 0.f, fz = 0.f;
        if (glen > 1e-12f) {
          const float inv = 1.f / glen;
          const float nx = gx_ * inv, ny = gy_ * inv, nz = gz_ * inv;
          const float *w = &vort[3 * c];
          const float k = E.vorticity_confinement * (P.dx * (float)scale);
          fx = k * (ny * w[2] - nz * w[1]);
          fy = k * (nz * w[0] - nx * w[2]);
          fz = k * (nx * w[1] - ny * w[0]);
        }
        const float g = P.dt * fluidity[c];
        velocity[3 * c] = velocity[3 * c] + g * fx;
        velocity[3 * c + 1] = velocity[3 * c + 1] + g * (fy + lift);
        velocity[3 * c + 2] = velocity[3 * c + 2] + g * fz;
        temperature[c] = th;
        vapor[c] = qv;
        density[c] = qc;
      }
    }
    accumulate_velocity();
    accumulate_density();
    accumulate_all(temperature.data(), 1);
    accumulate_all(vapor.data(), 1);
  }

  // sampleCoarse / samplePrecise, dcgrid_rendering.cu:6-58 + interpolate(), raymarching.cuh:26-40 (clampPos: the
  // integer position clamped to the domain)
  void sample_field(const float *f, int comps, int mode, const float *xyz, u64 n, float *out) const {
    for (u64 i = 0; i < n; i++) {
      const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
      const int ix = iclamp((int)floorf(px), 0, gx - 1), iy = iclamp((int)floorf(py), 0, gy - 1), iz = iclamp((int)floorf(pz), 0, gz - 1);
      int level = 0;
      const u64 b = block_index_deep(ix, iy, iz, level);
      if (mode == 0) {
        const u64 c = BV * b + spread((ix >> level) % BW, 2) + spread((iy >> level) % BW, 1) + spread((iz >> level) % BW, 0);
        for (int k = 0; k < comps; k++) out[comps * i + k] = f[comps * c + k];
        continue;
      }
      const float inv = 1.f / (float)(1 << level);
      const float x = px * inv - .5f, y = py * inv - .5f, z = pz * inv - .5f;
      const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
      const float dx = x - xf, dy = y - yf, dz = z - zf;
      // (the reference does not clamp the apron offset; positions within the domain keep it inside 0..4)
      const int ai = iclamp((int)xf + 1 - pos[3 * b], 0, AW - 2), aj = iclamp((int)yf + 1 - pos[3 * b + 1], 0, AW - 2),
                ak = iclamp((int)zf + 1 - pos[3 * b + 2], 0, AW - 2);
      const u64 *a = &apron[AV * b + AA * ai + AW * aj + ak];
      const u64 id[8] = {a[0], a[1], a[AW], a[AW + 1], a[AA], a[AA + 1], a[AA + AW], a[AA + AW + 1]};
      for (int k = 0; k < comps; k++) {
        auto v = [&](int q) { return f[comps * id[q] + k]; };
        const float dxi = 1.f - dx;
        const float c00 = v(0) * dxi + v(4) * dx, c01 = v(1) * dxi + v(5) * dx, c10 = v(2) * dxi + v(6) * dx, c11 = v(3) * dxi + v(7) * dx;
        const float dyi = 1.f - dy;
        const float c0 = c00 * dyi + c10 * dy, c1 = c01 * dyi + c11 * dy;
        out[comps * i + k] = c0 * (1.f - dz) + c1 * dz;
      }
    }
  }
  const float *field_ptr(int field, int &comps) {
    comps = 1;
    switch (field) {
      case DCG_FIELD_DENSITY: return density.data();
      case DCG_FIELD_VELOCITY: comps = 3; return velocity.data();
      case DCG_FIELD_FLUIDITY: return fluidity.data();
      case DCG_FIELD_PRESSURE: return pressure();
      case DCG_FIELD_DIVERGENCE: return divergence();
      case DCG_FIELD_T_PRESSURE: return t_pressure();
      case DCG_FIELD_TEMPERATURE: return temperature.data();
      case DCG_FIELD_VAPOR: return vapor.data();
      case DCG_FIELD_VORTICITY: comps = 3; return vort.data();
    }
    return nullptr;
  }

  // ---- host orchestration: fluid_simulation_dcgrid.cu ---------------------------
  void move_blocks(u64 &num_touched) {  // :348-437
    auto block_order = [this](u64 a, u64 b) { return block_scores[a] < 0.f ? false : block_scores[a] < block_scores[b]; };
    auto sub_order = [this](u64 a, u64 b) { return sub_scores[a] > sub_scores[b]; };
    calc_subblock_scores();
    accumulate_subblock_scores();
    std::fill(flags.begin(), flags.end(), 0);
    u64 n_move = 0;
    for (int level = 0; level < levels - 1; level++) {
      const u64 d0 = max_blocks[level], d1 = 8 * max_blocks[level + 1];
      u64 l = std::min({d0, d1, loads[level], full_blocks[level] - loads[level]});
      if (move_limit[level] > 0) l = std::min(l, move_limit[level]);
      if (l == 0) continue;
      if (E.selection == 1) {
        // EXTENSION (dcg_ext_params.selection == 1): the same greedy rule on a TOTAL order — movable blocks by (score
        // ascending, slot ascending), destinations by (score descending, id ascending), negative scores excluded —
        // so that the outcome does not depend on how a sort library treats ties and a non-strict-weak comparator
        std::vector<u64> mcv, dcv;
        for (u64 b = offsets[level]; b < offsets[level] + d0; b++)
          if (block_scores[b] >= 0.f) mcv.push_back(b);
        std::stable_sort(mcv.begin(), mcv.end(), [this](u64 a, u64 b) { return block_scores[a] < block_scores[b]; });
        for (u64 sb = 8 * offsets[level + 1]; sb < 8 * offsets[level + 1] + d1; sb++)
          if (sub_scores[sb] >= 0.f) dcv.push_back(sb);
        std::stable_sort(dcv.begin(), dcv.end(), [this](u64 a, u64 b) { return sub_scores[a] > sub_scores[b]; });
        const u64 K = std::min<u64>({l, mcv.size(), dcv.size()});
        u64 matches = 0;
        while (matches < K && block_scores[mcv[matches]] < sub_scores[dcv[matches]]) matches++;
        for (u64 i = 0; i < matches; i++) {
          block_scores[dcv[i] / 8] = -FLT_MAX;  // the parents that receive a block stay where they are in this pass
          to_move[n_move + i] = mcv[i];
          dest[n_move + i] = dcv[i];
        }
        move_limit[level] = (u64)(matches * 1.2f);
        n_move += matches;
        continue;
      }
      u64 *mc = to_move.data() + n_move;
      std::iota(mc, mc + d0, offsets[level]);
      if (d0 <= l)
        std::sort(mc, mc + d0, block_order);
      else {
        std::nth_element(mc, mc + l, mc + d0, block_order);
        std::sort(mc, mc + l, block_order);
      }
      u64 *dc = dest.data() + n_move;
      std::iota(dc, dc + d1, 8 * offsets[level + 1]);
      if (d1 <= l)
        std::sort(dc, dc + d1, sub_order);
      else {
        std::nth_element(dc, dc + l, dc + d1, sub_order);
        std::sort(dc, dc + l, sub_order);
      }
      u64 matches = 0;
      while (matches < l && block_scores[mc[matches]] >= 0.f && sub_scores[dc[matches]] >= 0.f &&
             block_scores[mc[matches]] < sub_scores[dc[matches]]) {
        matches++;
        block_scores[dc[matches] / 8] = -FLT_MAX;  // :417 — indexes the NEXT candidate (App. B-3); replicated
      }
      move_limit[level] = (u64)(matches * 1.2f);
      n_move += matches;
    }
    if (n_move > 0) {
      move_blocks_kernel(n_move);
      std::copy(to_move.begin(), to_move.begin() + n_move, touched.begin());
      num_touched += n_move;
      n_moved += n_move;
    }
  }

  void refine_subblocks(u64 &num_touched) {  // :439-483
    calc_subblock_scores();
    u64 n_ref = 0;
    for (int level = 1; level < levels; level++) {
      const u64 limit = std::min(max_blocks[level - 1] - loads[level - 1], 8 * loads[level] - loads[level - 1]);
      if (limit == 0) continue;
      const u64 start = 8 * offsets[level], end = start + 8 * max_blocks[level];
      u64 *di = dest.data() + n_ref;
      u64 n = 0;
      for (u64 i = start; i < end; i++)
        if (sub_scores[i] > 1e-4f) di[n++] = i;
      if (n > limit) {
        if (E.selection == 1) std::copy(di + (n - limit), di + n, di);  // EXTENSION: the largest ids, in ascending order
        else std::nth_element(di, di + limit, di + n, std::greater<u64>{});
      }
      n_ref += std::min(n, limit);
    }
    if (n_ref > 0) {
      refine_subblocks_kernel(n_ref, num_touched);
      num_touched += n_ref;
      n_refined += n_ref;
    }
  }

  void adapt_topology() override {  // :320-346
    n_adapt++;
    calc_vorticity();
    u64 num_touched = 0;
    move_blocks(num_touched);
    if (num_touched > 0) {
      std::fill(hash_key.begin(), hash_key.end(), kHashEmpty);
      std::fill(hash_val.begin(), hash_val.end(), kNone);
      refill_hash_table();
    }
    refine_subblocks(num_touched);
    if (num_touched > 0) {
      n_changed++;
      refresh_apron_indices();
      for (int l = levels - 2; l >= 0; l--) propagate_values(num_touched, l);
    }
  }

  void init() override {  // :190-210
    init_apron_indices();
    for (int l = sparse_levels; l < levels; l++) activate_level(l);
    for (int i = 0; i < 5; i++) adapt_topology();
  }

  void reset() override {  // :212-261
    std::fill(hash_key.begin(), hash_key.end(), kHashEmpty);
    std::fill(hash_val.begin(), hash_val.end(), kNone);
    std::fill(pos.begin(), pos.end(), 0);
    std::fill(lvl.begin(), lvl.end(), 0xFF);
    std::fill(flags.begin(), flags.end(), 0);
    std::fill(parent.begin(), parent.end(), kNone);
    std::fill(child.begin(), child.end(), kNone);
    std::fill(apron.begin(), apron.end(), 0);
    std::fill(density.begin(), density.end(), 0.f);
    std::fill(velocity.begin(), velocity.end(), 0.f);
    std::fill(fluidity.begin(), fluidity.end(), 0.f);
    std::fill(temp.begin(), temp.end(), 0.f);
    std::fill(temperature.begin(), temperature.end(), 0.f);
    std::fill(vapor.begin(), vapor.end(), 0.f);
    std::fill(vort.begin(), vort.end(), 0.f);
    for (int l = 0; l < levels; l++) {
      loads[l] = (max_blocks[l] == full_blocks[l]) ? max_blocks[l] : 0;
      move_limit[l] = 0;
    }
    std::iota(free_idx.begin(), free_idx.end(), (u64)0);
    init();
  }

  float debug_stats() override {  // dcgrid_structure.cu:224-251, fluid_simulation_dcgrid.cu:517-528
    const float *p = pressure(), *div = divergence();
    float sum = 0.f;
    for (u64 b = 0; b < M; b++) {
      float s = 0.f;
      if (lvl[b] != 0xFF) {
        const u64 *a = &apron[AV * b];
        const int scale = 1 << lvl[b];
        const float alpha = P.rdx * P.rdx / (scale * scale);
        for (int i = 1; i <= BW; i++)
          for (int j = 1; j <= BW; j++)
            for (int k = 1; k <= BW; k++) {
              const int ai = AA * i + AW * j + k;
              const u64 c = a[ai];
              if (child[c >> 3] != kNone) continue;
              const float r = div[c] - (p[a[ai - AA]] + p[a[ai + AA]] + p[a[ai - AW]] + p[a[ai + AW]] + p[a[ai - 1]] +
                                        p[a[ai + 1]] - 6.f * p[c]) * alpha;
              s += scale * fabsf(r);
            }
      }
      sum += s;
    }
    return sum;
  }

  int get_field(int field, float *dst) override {
    const u64 n = num_cells_;
    switch (field) {
      case DCG_FIELD_DENSITY: std::memcpy(dst, density.data(), n * 4); return 0;
      case DCG_FIELD_VELOCITY: std::memcpy(dst, velocity.data(), 3 * n * 4); return 0;
      case DCG_FIELD_FLUIDITY: std::memcpy(dst, fluidity.data(), n * 4); return 0;
      case DCG_FIELD_PRESSURE: std::memcpy(dst, pressure(), n * 4); return 0;
      case DCG_FIELD_DIVERGENCE: std::memcpy(dst, divergence(), n * 4); return 0;
      case DCG_FIELD_T_PRESSURE: std::memcpy(dst, t_pressure(), n * 4); return 0;
      case DCG_FIELD_TEMPERATURE: std::memcpy(dst, temperature.data(), n * 4); return 0;
      case DCG_FIELD_VAPOR: std::memcpy(dst, vapor.data(), n * 4); return 0;
      case DCG_FIELD_VORTICITY: std::memcpy(dst, vort.data(), 3 * n * 4); return 0;
    }
    return 1;
  }
};

// ===========================================================================
// C API (loaded with ctypes from tests/ and bench.py)
// ===========================================================================
extern "C" {
#define ORC_API __attribute__((visibility("default")))

ORC_API orc_sim *orc_create_uniform(const dcg_sim_params *p) { return new UniformOracle(*p); }
ORC_API orc_sim *orc_create_dcgrid(const dcg_sim_params *p, uint64_t max_blocks) {
  DCGridOracle *g = new DCGridOracle(*p, max_blocks);
  if (g->pool_error) { delete g; return nullptr; }
  return g;
}
ORC_API void orc_destroy(orc_sim *s) { delete s; }
ORC_API void orc_set_params(orc_sim *s, const dcg_sim_params *p) { s->P = *p; }
ORC_API void orc_set_jacobi_schedule(orc_sim *s, int coarse, int level, int local) {
  s->coarse_pairs = coarse; s->level_pairs = level; s->local_pairs = local;
}
ORC_API void orc_init(orc_sim *s) { s->init(); }
ORC_API void orc_reset(orc_sim *s) { s->reset(); }
ORC_API void orc_adapt_topology(orc_sim *s) { s->adapt_topology(); }
ORC_API void orc_advect_velocity(orc_sim *s) { s->advect_velocity(); }
ORC_API void orc_project(orc_sim *s) { s->project(); }
ORC_API void orc_project_local(orc_sim *s) { s->project_local(); }
ORC_API void orc_advect_density(orc_sim *s) { s->advect_density(); }
ORC_API void orc_step(orc_sim *s, int n) {  // src/simulation.cpp:104-111
  for (int i = 0; i < n; i++) {
    s->advect_velocity();
    s->adapt_topology();
    s->apply_sources();  // extension; no-op unless dcg_ext_params.sources
    s->project();
    s->advect_density();
  }
}
ORC_API void orc_set_ext_params(orc_sim *s, const dcg_ext_params *e) { s->E = *e; }
ORC_API void orc_apply_sources(orc_sim *s) { s->apply_sources(); }
ORC_API int orc_sample_field(orc_sim *s, int field, int mode, const float *positions, uint64_t n, float *out) {
  if (!s->is_dcgrid) return 1;
  DCGridOracle *g = static_cast<DCGridOracle *>(s);
  int comps = 1;
  const float *f = g->field_ptr(field, comps);
  if (!f) return 1;
  g->sample_field(f, comps, mode, positions, n, out);
  return 0;
}
ORC_API float orc_debug_stats(orc_sim *s) { return s->debug_stats(); }
ORC_API uint64_t orc_num_cells(orc_sim *s) { return s->num_cells(); }
ORC_API int orc_get_field(orc_sim *s, int field, float *dst) { return s->get_field(field, dst); }
ORC_API int orc_num_levels(orc_sim *s) {
  return s->is_dcgrid ? static_cast<DCGridOracle *>(s)->levels : static_cast<UniformOracle *>(s)->mip_levels;
}
ORC_API int orc_sparse_levels(orc_sim *s) { return s->is_dcgrid ? static_cast<DCGridOracle *>(s)->sparse_levels : 0; }
ORC_API int orc_get_level_table(orc_sim *s, uint64_t *mx, uint64_t *full, uint64_t *loads, uint64_t *offs) {
  if (!s->is_dcgrid) return 1;
  DCGridOracle *g = static_cast<DCGridOracle *>(s);
  for (int l = 0; l < g->levels; l++) {
    if (mx) mx[l] = g->max_blocks[l];
    if (full) full[l] = g->full_blocks[l];
    if (loads) loads[l] = g->loads[l];
    if (offs) offs[l] = g->offsets[l];
  }
  return 0;
}
ORC_API int orc_get_topology(orc_sim *s, int32_t *positions, uint8_t *levels, uint64_t *parent, uint64_t *children,
                             uint64_t *apron) {
  if (!s->is_dcgrid) return 1;
  DCGridOracle *g = static_cast<DCGridOracle *>(s);
  if (positions) std::memcpy(positions, g->pos.data(), g->pos.size() * 4);
  if (levels) std::memcpy(levels, g->lvl.data(), g->lvl.size());
  if (parent) std::memcpy(parent, g->parent.data(), g->parent.size() * 8);
  if (children) std::memcpy(children, g->child.data(), g->child.size() * 8);
  if (apron) std::memcpy(apron, g->apron.data(), g->apron.size() * 8);
  return 0;
}
ORC_API int orc_lookup_blocks(orc_sim *s, const int32_t *positions, uint64_t n, uint64_t *out_slot, uint8_t *out_level) {
  if (!s->is_dcgrid) return 1;
  DCGridOracle *g = static_cast<DCGridOracle *>(s);
  for (uint64_t i = 0; i < n; i++) {
    int level = 0;
    out_slot[i] = g->block_index_deep(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2], level);
    out_level[i] = (uint8_t)level;
  }
  return 0;
}
ORC_API int orc_get_counters(orc_sim *s, uint64_t out[8]) {
  std::memset(out, 0, 8 * sizeof(uint64_t));
  if (!s->is_dcgrid) return 0;
  DCGridOracle *g = static_cast<DCGridOracle *>(s);
  out[0] = g->n_adapt; out[1] = g->n_changed; out[2] = g->n_moved; out[3] = g->n_refined; out[5] = g->n_failed;
  return 0;
}
ORC_API int orc_get_move_limits(orc_sim *s, uint64_t *out) {
  if (!s->is_dcgrid) return 1;
  DCGridOracle *g = static_cast<DCGridOracle *>(s);
  for (int l = 0; l < g->levels; l++) out[l] = g->move_limit[l];
  return 0;
}
}  // extern "C"
