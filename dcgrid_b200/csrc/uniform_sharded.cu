// Slab-decomposed dense-grid solver: the uniform solve of uniform.cu sharded over `world` ranks by
// contiguous z-slabs (memory order is x fastest, then y, then z — src/utils/grid_math.cuh:10 — so a
// z-slab is one contiguous range and both halo planes are contiguous).  The reference is single-GPU
// (src/main.cpp:46-48); the parity target of the N-rank run is the 1-GPU result, bit for bit: every
// cell is computed by the same expression in the same order, only the owner of the memory differs.
//
// B200-first exchange: NO halo buffers and no pack/unpack.  Every rank's slab arena is mapped into
// every other rank (cudaIpc over NVLink/NVSwitch, or plain pointers when several ranks share one
// device/process), and the stencil / gather kernels read neighbour planes THROUGH the peer mapping:
//   * Jacobi / divergence / gradient: the z-1 / z+1 plane of a slab boundary is a remote plane read
//     over NVLink inside the sweep itself (gx*gy*4 B per face per sweep);
//   * semi-Lagrangian gathers: CFL is ~23-46 cells at 512^3-1024^3 (v*dt/dx), so fixed-width halos
//     cannot work; a backtraced corner is fetched from whichever rank owns its plane;
//   * mip levels are sharded by the owner of their first fine plane, so restriction / prolongation
//     also resolve through the same accessor.
// Ranks run in lock-step: before every kernel that reads what another rank wrote, a flag barrier over
// peer memory (one 4-byte store per peer + a bounded spin, ~us over NVLink) replaces a collective.
#include <algorithm>
#include <string>
#include <vector>

#include "sim.h"

namespace dcg {
namespace {

constexpr int kMaxRanks = 8, kMaxMip = 16;
constexpr int SBX = 32, SBY = 4, SBZ = 2;

struct ShardView {
  int world, slab, levels;
  uint64_t lvl_off[kMaxMip];  // local offset (cells) of mip level l inside the per-rank pyramid arrays
  float4 *vw[2][kMaxRanks];
  float *q[2][kMaxRanks];
  float *fl[kMaxRanks], *p[kMaxRanks], *tp[kMaxRanks], *dv[kMaxRanks];
};

// first plane of mip level `level` owned by rank r: a coarse plane belongs to the owner of its first fine plane
__host__ __device__ __forceinline__ int first_plane(int slab, int level, int r) { return (r * slab + (1 << level) - 1) >> level; }
__device__ __forceinline__ int plane_owner(const ShardView &S, int level, int z) { return min((z << level) / S.slab, S.world - 1); }

// level-0 cell (x,y,z) -> (rank, local index)
__device__ __forceinline__ uint64_t l0_index(const ShardView &S, const KParams &P, int x, int y, int z, int &r) {
  r = min(z / S.slab, S.world - 1);
  return ((uint64_t)(z - r * S.slab) * P.gy + y) * P.gx + x;
}
// pyramid cell of level `level`
__device__ __forceinline__ uint64_t pyr_index(const ShardView &S, const KParams &P, int level, int x, int y, int z, int &r) {
  r = plane_owner(S, level, z);
  const int w = P.gx >> level, h = P.gy >> level;
  return S.lvl_off[level] + ((uint64_t)(z - first_plane(S.slab, level, r)) * h + y) * w + x;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }

// local thread -> (x, y, global z) of a level; false when outside this rank's planes
__device__ __forceinline__ bool my_cell(const ShardView &S, const KParams &P, int rank, int level, int &x, int &y, int &z) {
  x = blockIdx.x * SBX + threadIdx.x;
  y = blockIdx.y * SBY + threadIdx.y;
  const int zl = blockIdx.z * SBZ + threadIdx.z;
  const int z0 = first_plane(S.slab, level, rank), z1 = first_plane(S.slab, level, rank + 1);
  z = z0 + zl;
  return x < (P.gx >> level) && y < (P.gy >> level) && z < z1;
}

// ---- kernels: the bodies are those of uniform.cu with every access routed through the accessors ----

// k_uniform_set_solidity_ratio, uniformgrid_structure.cu:23-31
__global__ void __launch_bounds__(256) k_s_fluidity(ShardView S, KParams P, int rank, int level) {
  int x, y, z;
  if (!my_cell(S, P, rank, level, x, y, z)) return;
  const float f = cell_fluidity(P, x, y, z, 1 << level);
  int r;
  S.fl[rank][pyr_index(S, P, level, x, y, z, r)] = f;
  if (level == 0) {
    const uint64_t i = l0_index(S, P, x, y, z, r);
    S.vw[0][rank][i].w = f;
    S.vw[1][rank][i].w = f;
  }
}

struct SSample {
  uint64_t id[8];
  int rk[8];
  int x0, y0, z0;
  float fx, fy, fz;
};
// INIT_SAMPLE, uniformgrid_fluid.cu:7-27
__device__ __forceinline__ SSample s_sample(const ShardView &S, const KParams &P, float px, float py, float pz) {
  SSample s;
  const float x = px - .5f, y = py - .5f, z = pz - .5f;
  const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  s.x0 = (int)xf; s.y0 = (int)yf; s.z0 = (int)zf;
  s.fx = x - xf; s.fy = y - yf; s.fz = z - zf;
  const int xa = clampi(s.x0, 0, P.gx - 1), xb = clampi(s.x0 + 1, 0, P.gx - 1);
  const int ya = clampi(s.y0, 0, P.gy - 1), yb = clampi(s.y0 + 1, 0, P.gy - 1);
  const int za = clampi(s.z0, 0, P.gz - 1), zb = clampi(s.z0 + 1, 0, P.gz - 1);
  s.id[0] = l0_index(S, P, xa, ya, za, s.rk[0]); s.id[1] = l0_index(S, P, xa, ya, zb, s.rk[1]);
  s.id[2] = l0_index(S, P, xa, yb, za, s.rk[2]); s.id[3] = l0_index(S, P, xa, yb, zb, s.rk[3]);
  s.id[4] = l0_index(S, P, xb, ya, za, s.rk[4]); s.id[5] = l0_index(S, P, xb, ya, zb, s.rk[5]);
  s.id[6] = l0_index(S, P, xb, yb, za, s.rk[6]); s.id[7] = l0_index(S, P, xb, yb, zb, s.rk[7]);
  return s;
}

// k_uniform_advect_velocity, uniformgrid_fluid.cu:50-67,88-95
__global__ void __launch_bounds__(256) k_s_advect_velocity(ShardView S, KParams P, int rank, int cur) {
  int x, y, z;
  if (!my_cell(S, P, rank, 0, x, y, z)) return;
  int r;
  const uint64_t i = l0_index(S, P, x, y, z, r);
  const float4 me = S.vw[cur][rank][i];
  const float bx = ((float)x + .5f) - me.x * P.dt * P.rdx;
  const float by = ((float)y + .5f) - me.y * P.dt * P.rdx;
  const float bz = ((float)z + .5f) - me.z * P.dt * P.rdx;
  const SSample s = s_sample(S, P, bx, by, bz);
  float4 c[8];
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    c[k] = S.vw[cur][s.rk[k]][s.id[k]];  // possibly a peer's slab, over NVLink
    f[k] = c[k].w;
  }
  const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
  float3 out = make_float3(0.f, 0.f, 0.f);
  if (!(W.acc < 1e-6f)) {
    float vx[8], vy[8], vz[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float3 v = velocity_bc(P, make_float3(c[k].x, c[k].y, c[k].z), s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1), s.z0 + (k & 1), 1);
      vx[k] = v.x; vy[k] = v.y; vz[k] = v.z;
    }
    out = make_float3(blend8(vx, W.w), blend8(vy, W.w), blend8(vz, W.w));
  }
  S.vw[cur ^ 1][rank][i] = make_float4(out.x, out.y, out.z, me.w);
}

// k_uniform_advect_density, uniformgrid_fluid.cu:69-86,97-105
__global__ void __launch_bounds__(256) k_s_advect_density(ShardView S, KParams P, int rank, int cur_v, int cur_q) {
  int x, y, z;
  if (!my_cell(S, P, rank, 0, x, y, z)) return;
  int r;
  const uint64_t i = l0_index(S, P, x, y, z, r);
  const float4 me = S.vw[cur_v][rank][i];
  const float bx = ((float)x + .5f) - me.x * P.dt * P.rdx;
  const float by = ((float)y + .5f) - me.y * P.dt * P.rdx;
  const float bz = ((float)z + .5f) - me.z * P.dt * P.rdx;
  const SSample s = s_sample(S, P, bx, by, bz);
  float q[8], f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    f[k] = S.vw[cur_v][s.rk[k]][s.id[k]].w;
    q[k] = S.q[cur_q][s.rk[k]][s.id[k]];
  }
  const Weights8 W = corner_weights(f, s.fx, s.fy, s.fz);
  float out = 0.f;
  if (!(W.acc < 1e-6f)) {
#pragma unroll
    for (int k = 0; k < 8; k++) q[k] = density_bc(P, q[k], s.x0 + ((k >> 2) & 1), s.y0 + ((k >> 1) & 1), s.z0 + (k & 1), 1);
    out = blend8(q, W.w);
  }
  S.q[cur_q ^ 1][rank][i] = out;
}

// k_uniform_calc_divergence, uniformgrid_fluid.cu:107-132
__global__ void __launch_bounds__(256) k_s_divergence(ShardView S, KParams P, int rank, int cur) {
  int x, y, z;
  if (!my_cell(S, P, rank, 0, x, y, z)) return;
  int r, rb, rf;
  const uint64_t i = l0_index(S, P, x, y, z, r);
  const float4 *mine = S.vw[cur][rank];
  const uint64_t sy = P.gx;
  const float4 l = mine[x > 0 ? i - 1 : i], rr = mine[x < P.gx - 1 ? i + 1 : i];
  const float4 dn = mine[y > 0 ? i - sy : i], up = mine[y < P.gy - 1 ? i + sy : i];
  const uint64_t ib = l0_index(S, P, x, y, z > 0 ? z - 1 : z, rb), iff = l0_index(S, P, x, y, z < P.gz - 1 ? z + 1 : z, rf);
  const float4 b = S.vw[cur][rb][ib], f = S.vw[cur][rf][iff];  // slab faces: the neighbour rank's boundary plane
  const float3 vl = velocity_bc(P, make_float3(l.x, l.y, l.z), x - 1, y, z, 1);
  const float3 vr = velocity_bc(P, make_float3(rr.x, rr.y, rr.z), x + 1, y, z, 1);
  const float3 vd = velocity_bc(P, make_float3(dn.x, dn.y, dn.z), x, y - 1, z, 1);
  const float3 vu = velocity_bc(P, make_float3(up.x, up.y, up.z), x, y + 1, z, 1);
  const float3 vb = velocity_bc(P, make_float3(b.x, b.y, b.z), x, y, z - 1, 1);
  const float3 vf = velocity_bc(P, make_float3(f.x, f.y, f.z), x, y, z + 1, 1);
  int rp;
  const uint64_t ip = pyr_index(S, P, 0, x, y, z, rp);
  S.p[rank][ip] = 0.f;
  S.tp[rank][ip] = 0.f;
  S.dv[rank][ip] = .5f * P.rdx * (rr.w * vr.x - l.w * vl.x + up.w * vu.y - dn.w * vd.y + f.w * vf.z - b.w * vb.z);
}

// k_uniform_restrict, uniformgrid_fluid.cu:134-160
__global__ void __launch_bounds__(256) k_s_restrict(ShardView S, KParams P, int rank, int level) {
  int x, y, z;
  if (!my_cell(S, P, rank, level, x, y, z)) return;
  int r, r0, r1;
  const uint64_t i = pyr_index(S, P, level, x, y, z, r);
  const uint64_t cw = (uint64_t)(P.gx >> (level - 1));
  const uint64_t c0 = pyr_index(S, P, level - 1, 2 * x, 2 * y, 2 * z, r0), c1 = pyr_index(S, P, level - 1, 2 * x, 2 * y, 2 * z + 1, r1);
  const float *d0 = S.dv[r0] + c0, *d1 = S.dv[r1] + c1;
  S.p[rank][i] = 0.f;
  S.tp[rank][i] = 0.f;
  S.dv[rank][i] = .125f * (d0[0] + d0[1] + d0[cw] + d0[cw + 1] + d1[0] + d1[1] + d1[cw] + d1[cw + 1]);
}

// calcPressure<in,out>, uniformgrid_fluid.cu:162-192.  flip = 0: p -> tp, 1: tp -> p
__global__ void __launch_bounds__(256) k_s_jacobi(ShardView S, KParams P, int rank, int level, int flip) {
  int x, y, z;
  if (!my_cell(S, P, rank, level, x, y, z)) return;
  const int w = P.gx >> level, h = P.gy >> level, d = P.gz >> level;
  const int scale = 1 << level;
  const float alpha = P.dx * P.dx * scale * scale;
  int r, rb, rf;
  const uint64_t i = pyr_index(S, P, level, x, y, z, r);
  const float *mine = flip ? S.tp[rank] : S.p[rank];
  const uint64_t sy = w;
  const float pl = mine[x > 0 ? i - 1 : i], pr = mine[x < w - 1 ? i + 1 : i];
  const float pd = mine[y > 0 ? i - sy : i], pu = mine[y < h - 1 ? i + sy : i];
  const uint64_t ib = pyr_index(S, P, level, x, y, z > 0 ? z - 1 : z, rb), iff = pyr_index(S, P, level, x, y, z < d - 1 ? z + 1 : z, rf);
  const float pb = (flip ? S.tp[rb] : S.p[rb])[ib], pf = (flip ? S.tp[rf] : S.p[rf])[iff];  // slab faces: a peer's plane
  (flip ? S.p[rank] : S.tp[rank])[i] = (pl + pr + pd + pu + pb + pf - alpha * S.dv[rank][i]) / 6.f;
}

// k_uniform_prolongate, uniformgrid_fluid.cu:206-237
__global__ void __launch_bounds__(256) k_s_prolongate(ShardView S, KParams P, int rank, int level) {
  int x, y, z;
  if (!my_cell(S, P, rank, level, x, y, z)) return;
  const int w = P.gx >> level, h = P.gy >> level, d = P.gz >> level;
  int r;
  const uint64_t i = pyr_index(S, P, level, x, y, z, r);
  const int sx = (x == 0 || x == w - 1) ? 0 : 2 * (x % 2) - 1;
  const int sy = (y == 0 || y == h - 1) ? 0 : 2 * (y % 2) - 1;
  const int sz = (z == 0 || z == d - 1) ? 0 : 2 * (z % 2) - 1;
  const int X = x / 2, Y = y / 2, Z = z / 2;
  int r0, r1;
  const uint64_t a0 = pyr_index(S, P, level + 1, X, Y, Z, r0), a1 = pyr_index(S, P, level + 1, X, Y, Z + sz, r1);
  const int64_t ox = sx, oy = (int64_t)sy * (w / 2);
  const float *q0 = S.p[r0] + a0, *q1 = S.p[r1] + a1;
  const float p000 = q0[0], p001 = q0[ox], p010 = q0[oy], p011 = q0[oy + ox];
  const float p100 = q1[0], p101 = q1[ox], p110 = q1[oy], p111 = q1[oy + ox];
  S.p[rank][i] = (27.f * p000 + 9.f * (p001 + p010 + p100) + 3.f * (p011 + p101 + p110) + p111) / 64.f;
}

// k_uniform_apply_pressure, uniformgrid_fluid.cu:239-260
__global__ void __launch_bounds__(256) k_s_apply_pressure(ShardView S, KParams P, int rank, int cur) {
  int x, y, z;
  if (!my_cell(S, P, rank, 0, x, y, z)) return;
  int r, rb, rf;
  const uint64_t iv = l0_index(S, P, x, y, z, r);
  const uint64_t i = pyr_index(S, P, 0, x, y, z, r);
  const uint64_t sy = P.gx;
  const uint64_t il = x > 0 ? i - 1 : i, ir = x < P.gx - 1 ? i + 1 : i;
  const uint64_t id = y > 0 ? i - sy : i, iu = y < P.gy - 1 ? i + sy : i;
  const uint64_t ib = pyr_index(S, P, 0, x, y, z > 0 ? z - 1 : z, rb), iff = pyr_index(S, P, 0, x, y, z < P.gz - 1 ? z + 1 : z, rf);
  const float *p = S.p[rank], *fl = S.fl[rank];
  const float alpha = .5f * P.rdx;
  const float pc = p[i];
  float4 v = S.vw[cur][rank][iv];
  const float wl = fl[il], wr = fl[ir], wd = fl[id], wu = fl[iu], wb = S.fl[rb][ib], wf = S.fl[rf][iff];
  v.x -= alpha * (wr * (p[ir] - pc) + wl * (pc - p[il]));
  v.y -= alpha * (wu * (p[iu] - pc) + wd * (pc - p[id]));
  v.z -= alpha * (wf * (S.p[rf][iff] - pc) + wb * (pc - S.p[rb][ib]));
  S.vw[cur][rank][iv] = v;
}

// k_uniform_debug_stats (uniformgrid_structure.cu:33-43) bins of 256 consecutive cells; a slab holds whole bins
__global__ void k_s_debug_stats(const float *__restrict__ q, const float4 *__restrict__ vw, float *__restrict__ stats, uint64_t bins) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bins) return;
  float s = 0.f;
  for (uint64_t i = b * 256; i < b * 256 + 256; i++) s += q[i] * vw[i].w;
  stats[b] = s;
}
__global__ void k_s_unpack_velocity(const float4 *__restrict__ vw, float *__restrict__ out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = vw[i];
  out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
}
__global__ void __launch_bounds__(256) k_s_total_density(const float *__restrict__ q, const float4 *__restrict__ vw, uint64_t n,
                                                         double *__restrict__ partial) {
  __shared__ double sh[256];
  double s = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) s += (double)(q[i] * vw[i].w);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// Lock-step barrier over peer memory.  flags[r] points at rank r's flag words (its arena, mapped here);
// word j of rank r's flags = the last epoch rank j announced to r.  Bounded spin: a rank that never
// arrives trips *err instead of hanging the GPU.
struct BarrierPeers {
  volatile uint32_t *flags[kMaxRanks];
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__global__ void k_shard_barrier(BarrierPeers B, int rank, int world, uint32_t epoch, uint32_t *err) {
  const int t = threadIdx.x;
  if (t < world && t != rank) {
    __threadfence_system();       // everything this rank wrote before the barrier is visible system-wide
    B.flags[t][rank] = epoch;     // 4-byte store into the peer's arena over NVLink
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

struct UniformShardSim : dcg_sim {
  int gx = 0, gy = 0, gz = 0, mip_levels = 1;
  int world = 1, rank0 = 0, nlocal = 1, slab = 0;
  bool ipc = false, ready = false;
  uint64_t n_local = 0;            // level-0 cells per rank
  uint64_t pyr_local = 0;          // pyramid cells per rank
  size_t arena_bytes = 0;
  size_t off_flags = 0, off_vw[2] = {0, 0}, off_q[2] = {0, 0}, off_fl = 0, off_p = 0, off_tp = 0, off_dv = 0;
  std::vector<char *> arena;       // [nlocal] device arenas owned by this instance
  char *peer_base[kMaxRanks] = {};  // [world] arena base of every rank as mapped here
  bool peer_opened[kMaxRanks] = {};
  ShardView S{};
  BarrierPeers B{};
  uint32_t epoch = 0;
  uint32_t *d_err = nullptr;
  float *scratch = nullptr;
  double *d_partial = nullptr, *h_partial = nullptr;
  int cur_v = 0, cur_q = 0;
  bool fluidity_dirty = true;
  uint64_t n_barriers = 0;

  ~UniformShardSim() override {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (int r = 0; r < world; r++)
      if (peer_opened[r]) cudaIpcCloseMemHandle(peer_base[r]);
    for (char *a : arena) cudaFree(a);
    cudaFree(d_err); cudaFree(scratch); cudaFree(d_partial);
    if (h_partial) cudaFreeHost(h_partial);
    base_teardown();
  }

  static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

  int construct(const dcg_sim_params *, int) override { return fail(DCG_ERR_INVALID, "use construct_sharded"); }

  int construct_sharded(const dcg_sim_params *prm, int dev, int rank, int wsize, int nloc) {
    DCG_TRY(base_setup(prm, dev));
    project_coarsest_pairs = 2; project_level_pairs = 1; local_pairs = 5;  // fluid_simulation_uniform.cu:103,116,129
    gx = prm->gx; gy = prm->gy; gz = prm->gz;
    world = wsize; rank0 = rank; nlocal = nloc;
    if (gx <= 0 || gy <= 0 || gz <= 0) return fail(DCG_ERR_INVALID, "grid size must be positive");
    if (world < 1 || world > kMaxRanks) return fail(DCG_ERR_INVALID, "world size must be in [1, %d]", kMaxRanks);
    if (nlocal != 1 && nlocal != world) return fail(DCG_ERR_INVALID, "nlocal must be 1 (one rank per process) or world (all ranks in this process)");
    if (rank0 < 0 || rank0 + nlocal > world) return fail(DCG_ERR_INVALID, "rank out of range");
    if (gz % world != 0) return fail(DCG_ERR_INVALID, "gz (%d) must be a multiple of the world size (%d): slabs are whole z-planes of equal count", gz, world);
    if (((uint64_t)gx * gy) % 256 != 0 && world > 1) return fail(DCG_ERR_INVALID, "gx*gy must be a multiple of 256 (debugStats bins must not straddle slabs)");
    slab = gz / world;
    ipc = nlocal != world;
    const uint64_t min_dim = (uint64_t)std::min(gx, std::min(gy, gz));
    uint64_t cell = 2;
    mip_levels = 1;
    while (gx % cell == 0 && gy % cell == 0 && gz % cell == 0 && cell * 4 <= min_dim) {  // fluid_simulation_uniform.cu:8-17
      mip_levels++;
      cell *= 2;
    }
    if (mip_levels > kMaxMip) return fail(DCG_ERR_UNSUPPORTED, "too many mip levels");
    n_local = (uint64_t)gx * gy * slab;
    pyr_local = 0;
    S.world = world; S.slab = slab; S.levels = mip_levels;
    for (int l = 0; l < mip_levels; l++) {
      S.lvl_off[l] = pyr_local;
      const uint64_t planes = std::max<uint64_t>(1, ((uint64_t)slab + (1ull << l) - 1) >> l);  // max over ranks
      pyr_local += planes * (uint64_t)(gx >> l) * (gy >> l);
    }
    size_t o = 0;
    off_flags = o; o += align_up(kMaxRanks * sizeof(uint32_t));
    for (int i = 0; i < 2; i++) { off_vw[i] = o; o += align_up(n_local * sizeof(float4)); }
    for (int i = 0; i < 2; i++) { off_q[i] = o; o += align_up(n_local * sizeof(float)); }
    off_fl = o; o += align_up(pyr_local * 4);
    off_p = o; o += align_up(pyr_local * 4);
    off_tp = o; o += align_up(pyr_local * 4);
    off_dv = o; o += align_up(pyr_local * 4);
    arena_bytes = o;
    arena.assign(nlocal, nullptr);
    for (int lr = 0; lr < nlocal; lr++) {
      DCG_CUDA_TRY(cudaMalloc(&arena[lr], arena_bytes));
      DCG_CUDA_TRY(cudaMemsetAsync(arena[lr], 0, arena_bytes, stream));
      peer_base[rank0 + lr] = arena[lr];
    }
    DCG_CUDA_TRY(cudaMalloc(&d_err, 4));
    DCG_CUDA_TRY(cudaMemsetAsync(d_err, 0, 4, stream));
    DCG_CUDA_TRY(cudaMalloc(&scratch, 3 * n_local * sizeof(float)));
    DCG_CUDA_TRY(cudaMalloc(&d_partial, 1024 * sizeof(double)));
    DCG_CUDA_TRY(cudaMallocHost(&h_partial, 1024 * sizeof(double)));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    if (!ipc) return finalize();
    return DCG_OK;  // the caller exchanges handles, then import_handles() finalizes
  }

  int export_handle(void *out, uint64_t cap) override {
    if (cap < sizeof(cudaIpcMemHandle_t)) return fail(DCG_ERR_INVALID, "handle buffer too small");
    cudaIpcMemHandle_t h;
    DCG_CUDA_TRY(cudaIpcGetMemHandle(&h, arena[0]));
    std::memcpy(out, &h, sizeof h);
    return DCG_OK;
  }
  int import_handles(const void *handles, int count) override {
    if (!ipc) return fail(DCG_ERR_INVALID, "all ranks are local: nothing to import");
    if (count != world) return fail(DCG_ERR_INVALID, "expected %d handles, got %d", world, count);
    DCG_CUDA_TRY(cudaSetDevice(device));
    for (int r = 0; r < world; r++) {
      if (r == rank0) continue;
      cudaIpcMemHandle_t h;
      std::memcpy(&h, static_cast<const char *>(handles) + (size_t)r * sizeof h, sizeof h);
      void *ptr = nullptr;
      DCG_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      peer_base[r] = static_cast<char *>(ptr);
      peer_opened[r] = true;
    }
    return finalize();
  }

  int finalize() {
    for (int r = 0; r < world; r++) {
      char *b = peer_base[r];
      if (!b) return fail(DCG_ERR_INVALID, "rank %d arena not mapped", r);
      for (int i = 0; i < 2; i++) {
        S.vw[i][r] = reinterpret_cast<float4 *>(b + off_vw[i]);
        S.q[i][r] = reinterpret_cast<float *>(b + off_q[i]);
      }
      S.fl[r] = reinterpret_cast<float *>(b + off_fl);
      S.p[r] = reinterpret_cast<float *>(b + off_p);
      S.tp[r] = reinterpret_cast<float *>(b + off_tp);
      S.dv[r] = reinterpret_cast<float *>(b + off_dv);
      B.flags[r] = reinterpret_cast<volatile uint32_t *>(b + off_flags);
    }
    ready = true;
    return reset();
  }

  int need_ready() { return ready ? DCG_OK : fail(DCG_ERR_INVALID, "sharded instance not finalized: call dcg_shard_import_handles first"); }

  // every rank has finished everything it launched so far, and its writes are visible to all
  void barrier(int lr) {
    if (!ipc) return;  // all ranks share this stream: program order is the barrier
    epoch++;
    k_shard_barrier<<<1, 32, 0, stream>>>(B, rank0 + lr, world, epoch, d_err);
    launches++;
    n_barriers++;
  }

  dim3 grid_for(int level, int r) const {
    const int planes = first_plane(slab, level, r + 1) - first_plane(slab, level, r);
    return dim3(idiv_up(gx >> level, SBX), idiv_up(gy >> level, SBY), std::max(1, idiv_up(planes, SBZ)));
  }
  static dim3 block() { return dim3(SBX, SBY, SBZ); }
  bool has_planes(int level, int r) const { return first_plane(slab, level, r + 1) > first_plane(slab, level, r); }

  // one lock-step stage: [barrier,] kernel on every local rank that owns planes of `level`
  template <typename F>
  void stage(int level, F &&launch) {
    for (int lr = 0; lr < nlocal; lr++) {
      barrier(lr);
      if (has_planes(level, rank0 + lr)) {
        launch(rank0 + lr, grid_for(level, rank0 + lr));
        launches++;
      }
    }
  }

  int on_params_changed() override {
    if (params.gx != gx || params.gy != gy || params.gz != gz) return fail(DCG_ERR_INVALID, "grid size is fixed at construction");
    fluidity_dirty = true;
    return DCG_OK;
  }
  int reset() override {  // fluid_simulation_uniform.cu:81-88
    DCG_TRY(need_ready());
    DCG_CUDA_TRY(cudaSetDevice(device));
    for (int lr = 0; lr < nlocal; lr++) {
      barrier(lr);  // nobody is still reading the fields we are about to clear
      DCG_CUDA_TRY(cudaMemsetAsync(arena[lr] + off_vw[0], 0, arena_bytes - off_vw[0], stream));
    }
    cur_v = cur_q = 0;
    fluidity_dirty = true;
    return adapt_topology();
  }
  int init() override { return reset(); }
  int adapt_topology() override {  // fluid_simulation_uniform.cu:143-147 (static field: recomputed only when SimParams change)
    DCG_TRY(need_ready());
    if (!fluidity_dirty) return DCG_OK;
    for (int l = 0; l < mip_levels; l++) stage(l, [&](int r, dim3 g) { k_s_fluidity<<<g, block(), 0, stream>>>(S, kp, r, l); });
    fluidity_dirty = false;
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  int advect_velocity() override {  // fluid_simulation_uniform.cu:90-94
    DCG_TRY(need_ready());
    const int cv = cur_v;
    stage(0, [&](int r, dim3 g) { k_s_advect_velocity<<<g, block(), 0, stream>>>(S, kp, r, cv); });
    cur_v ^= 1;
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  int advect_density() override {  // fluid_simulation_uniform.cu:137-141
    DCG_TRY(need_ready());
    const int cv = cur_v, cq = cur_q;
    stage(0, [&](int r, dim3 g) { k_s_advect_density<<<g, block(), 0, stream>>>(S, kp, r, cv, cq); });
    cur_q ^= 1;
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  void jacobi_pair(int l) {
    stage(l, [&](int r, dim3 g) { k_s_jacobi<<<g, block(), 0, stream>>>(S, kp, r, l, 0); });
    stage(l, [&](int r, dim3 g) { k_s_jacobi<<<g, block(), 0, stream>>>(S, kp, r, l, 1); });
  }
  int project() override {  // fluid_simulation_uniform.cu:96-124
    DCG_TRY(need_ready());
    const int cv = cur_v;
    stage(0, [&](int r, dim3 g) { k_s_divergence<<<g, block(), 0, stream>>>(S, kp, r, cv); });
    for (int l = 1; l < mip_levels; l++) stage(l, [&](int r, dim3 g) { k_s_restrict<<<g, block(), 0, stream>>>(S, kp, r, l); });
    for (int i = 0; i < project_coarsest_pairs; i++) jacobi_pair(mip_levels - 1);
    for (int l = mip_levels - 2; l >= 0; l--) {
      stage(l, [&](int r, dim3 g) { k_s_prolongate<<<g, block(), 0, stream>>>(S, kp, r, l); });
      for (int i = 0; i < project_level_pairs; i++) jacobi_pair(l);
    }
    stage(0, [&](int r, dim3 g) { k_s_apply_pressure<<<g, block(), 0, stream>>>(S, kp, r, cv); });
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  int project_local() override {  // fluid_simulation_uniform.cu:126-135
    DCG_TRY(need_ready());
    const int cv = cur_v;
    stage(0, [&](int r, dim3 g) { k_s_divergence<<<g, block(), 0, stream>>>(S, kp, r, cv); });
    for (int i = 0; i < local_pairs; i++) jacobi_pair(0);
    stage(0, [&](int r, dim3 g) { k_s_apply_pressure<<<g, block(), 0, stream>>>(S, kp, r, cv); });
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  int step(int n) override {
    DCG_TRY(need_ready());
    DCG_CUDA_TRY(cudaSetDevice(device));
    DCG_CUDA_TRY(cudaEventRecord(ev_begin, stream));
    DCG_TRY(dcg_sim::step(n));
    DCG_CUDA_TRY(cudaEventRecord(ev_end, stream));
    step_timing_pending = true;
    return DCG_OK;
  }

  int check_barrier_error() {
    uint32_t e = 0;
    DCG_CUDA_TRY(cudaMemcpyAsync(&e, d_err, 4, cudaMemcpyDeviceToHost, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    if (e) return fail(DCG_ERR_CUDA, "shard barrier timed out: a peer rank never arrived (ranks must issue identical call sequences)");
    return DCG_OK;
  }

  // local partial results; the host side (bench / tests) combines ranks (sum / allreduce)
  int debug_stats(float *out) override {  // fluid_simulation_uniform.cu:160-176: bins of 256 cells, summed in order
    DCG_TRY(need_ready());
    DCG_CUDA_TRY(cudaSetDevice(device));
    float sum = 0.f;
    for (int lr = 0; lr < nlocal; lr++) {
      const uint64_t bins = n_local / 256;
      if (bins == 0) continue;
      const float *q = reinterpret_cast<float *>(arena[lr] + off_q[cur_q]);
      const float4 *v = reinterpret_cast<float4 *>(arena[lr] + off_vw[cur_v]);
      k_s_debug_stats<<<(unsigned)((bins + 255) / 256), 256, 0, stream>>>(q, v, scratch, bins);
      launches++;
      std::vector<float> h(bins);
      DCG_CUDA_TRY(cudaMemcpyAsync(h.data(), scratch, bins * sizeof(float), cudaMemcpyDeviceToHost, stream));
      DCG_CUDA_TRY(cudaStreamSynchronize(stream));
      for (uint64_t i = 0; i < bins; i++) sum += h[i];  // ranks in order == the reference's single ordered sum
    }
    *out = sum;
    return check_barrier_error();
  }
  int total_density(double *out) override {
    DCG_TRY(need_ready());
    DCG_CUDA_TRY(cudaSetDevice(device));
    double s = 0.0;
    for (int lr = 0; lr < nlocal; lr++) {
      const int blocks = (int)std::min<uint64_t>(1024, (n_local + 255) / 256);
      const float *q = reinterpret_cast<float *>(arena[lr] + off_q[cur_q]);
      const float4 *v = reinterpret_cast<float4 *>(arena[lr] + off_vw[cur_v]);
      k_s_total_density<<<blocks, 256, 0, stream>>>(q, v, n_local, d_partial);
      launches++;
      DCG_CUDA_TRY(cudaMemcpyAsync(h_partial, d_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, stream));
      DCG_TRY(synchronize());
      for (int i = 0; i < blocks; i++) s += h_partial[i];
    }
    *out = s;
    return check_barrier_error();
  }

  uint64_t num_cells() const override { return n_local * nlocal; }  // cells held by this instance
  int num_levels() const override { return mip_levels; }

  // Local slab(s), ranks in order: with nlocal == world this is the whole field in the reference's memory
  // order (a z-slab is a contiguous index range).  Pyramid fields return their level-0 part.
  int get_field(int field, int layout, float *dst, uint64_t count) override {
    (void)layout;
    DCG_TRY(need_ready());
    DCG_CUDA_TRY(cudaSetDevice(device));
    const uint64_t per = field == DCG_FIELD_VELOCITY ? 3 * n_local : n_local;
    if (!dst || count < per * nlocal) return fail(DCG_ERR_INVALID, "get_field: destination too small");
    for (int lr = 0; lr < nlocal; lr++) {
      const float *src = nullptr;
      char *a = arena[lr];
      switch (field) {
        case DCG_FIELD_DENSITY: src = reinterpret_cast<float *>(a + off_q[cur_q]); break;
        case DCG_FIELD_VELOCITY:
          k_s_unpack_velocity<<<(unsigned)((n_local + 255) / 256), 256, 0, stream>>>(reinterpret_cast<float4 *>(a + off_vw[cur_v]), scratch, n_local);
          launches++;
          src = scratch;
          break;
        case DCG_FIELD_FLUIDITY: src = reinterpret_cast<float *>(a + off_fl); break;
        case DCG_FIELD_PRESSURE: src = reinterpret_cast<float *>(a + off_p); break;
        case DCG_FIELD_DIVERGENCE: src = reinterpret_cast<float *>(a + off_dv); break;
        case DCG_FIELD_T_PRESSURE: src = reinterpret_cast<float *>(a + off_tp); break;
        default: return fail(DCG_ERR_INVALID, "get_field: unknown field %d", field);
      }
      DCG_CUDA_TRY(cudaMemcpyAsync(dst + (size_t)lr * per, src, per * sizeof(float), cudaMemcpyDeviceToHost, stream));
      DCG_TRY(synchronize());
    }
    return check_barrier_error();
  }

  int get_counters(uint64_t out[8]) override {
    for (int i = 0; i < 8; i++) out[i] = 0;
    out[6] = launches;
    out[7] = n_barriers;
    return DCG_OK;
  }

  // SURVEY.md §8(d) per-cell figures x the cells THIS instance owns
  int algorithmic_bytes(double *bytes, uint64_t *active_blocks) override {
    double b = 0.0;
    for (int lr = 0; lr < nlocal; lr++) {
      const int r = rank0 + lr;
      b += (28.0 + 28.0 + 32.0 + 24.0) * (double)n_local;
      for (int l = 0; l < mip_levels; l++) {
        const double n = (double)(first_plane(slab, l, r + 1) - first_plane(slab, l, r)) * (double)(gx >> l) * (double)(gy >> l);
        const int pairs = (l == mip_levels - 1) ? project_coarsest_pairs : project_level_pairs;
        b += n * 12.0 * 2 * pairs;
        if (l >= 1) b += n * (8 * 4.0 + 12.0);
        if (l < mip_levels - 1) b += n * 4.5;
      }
    }
    if (bytes) *bytes = b;
    if (active_blocks) *active_blocks = 0;
    return DCG_OK;
  }
  int bench_stage(const char *, int, int, float *, double *) override { return fail(DCG_ERR_UNSUPPORTED, "bench_stage: not available on sharded instances"); }
};

}  // namespace
}  // namespace dcg

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
  auto *s = new dcg::UniformShardSim();
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

DCG_API uint64_t dcg_shard_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

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
