// Headless driver for the REFERENCE's own CUDA solver (test infrastructure).
//
// The reference cannot run out of the box: its window/UI library is proprietary and
// absent (readme.md:3), but only main.cpp / simulation.* need it.  This driver does
// what Simulation::updateSimulation does (src/simulation.cpp:93-116) — upload
// SimParams, then advectVelocity / adaptTopology / project / advectDensity — on the
// unmodified reference classes, and dumps their device state so that
//   (1) the CPU restatement in oracle/dcgrid_oracle.cpp can be pinned against it, and
//   (2) the product's CUDA path can be compared with it on the same B200.
// It is compiled from the reference sources where they lie (see oracle/Makefile);
// no reference source is copied into this repository.
//
// usage: ref_harness key=value ...
//   grid=uniform|dcgrid  d=64 [gx= gy= gz=]  M=4096  solids=0|1  steps=10
//   schedule=project|local|jacobi<N>   (jacobi<N>: uniform only, N pairs on level 0,
//                                       launched from here: "50 Jacobi" == jacobi25)
//   out=<file>           binary dump (see tests/_refio.py for the container format)
//   dump_reset=1         also dump the state right after construction/reset
//   trace=1              per-step FNV-1a digest of density+velocity in the JSON line
//   reps=1               repeat the whole run (determinism check, prints one line each; rep r>0
//                        dumps to <out>.rep<r>)
//   serialize=1          DCGrid only: run the reference's one racy kernel
//                        (k_dcgrid_refine_subblocks: atomicAdd slot allocation,
//                        dcgrid_utils.cuh:118-127) one rank per launch, in rank order.  All
//                        kernels are still the reference's own; only the launch shape of that
//                        kernel changes, which makes the run reproducible.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <helper_math.h>

#define private public
#define protected public
#include "dcgrid/fluid_simulation_dcgrid.h"
#include "uniformgrid/fluid_simulation_uniform.h"
#undef private
#undef protected
#include "data/sim_params.h"
#include "utils/sim_utils.h"

#include "dcgrid/dcgrid.h"

static void die(const char *msg);

// The reference's DCGrid solver with adaptTopology() re-issued so that slot allocation happens in
// rank order.  Host logic follows fluid_simulation_dcgrid.cu:320-346 (adaptTopology) and :439-483
// (refineSubblocks); every kernel launched is the reference's.  moveBlocks() is called unchanged.
struct SerializedDCGrid : public FluidSimulationDCGrid {
  SerializedDCGrid(const int3 &size, const size_t &maxNumBlocks) : FluidSimulationDCGrid(size, maxNumBlocks) {}

  void refineInRankOrder(size_t &touchedCount) {
    const size_t M = h_grid->maxNumBlocks;
    k_dcgrid_calc_subblock_scores<<<(unsigned)((M * 8 + 63) / 64), 64>>>(d_grid, d_subblockScores);
    cudaMemcpy(h_subblockScores, d_subblockScores, M * 8 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(h_blockLoads, h_grid->blockLoads, h_grid->levels * sizeof(size_t), cudaMemcpyDeviceToHost);
    size_t total = 0;
    for (int level = 1; level < h_grid->levels; level++) {
      const size_t room = h_maxNumBlocksLevel[level - 1] - h_blockLoads[level - 1];
      const size_t cap = 8 * h_blockLoads[level] - h_blockLoads[level - 1];
      const size_t limit = std::min(room, cap);
      if (limit == 0) continue;
      size_t *list = h_destSubblockIndices + total;
      size_t n = 0;
      const size_t first = 8 * h_levelOffsets[level];
      for (size_t i = first; i < first + 8 * h_maxNumBlocksLevel[level]; i++)
        if (h_subblockScores[i] > 1e-4f) list[n++] = i;
      if (n > limit) std::nth_element(list, list + limit, list + n, std::greater{});
      total += std::min(n, limit);
    }
    if (total == 0) return;
    cudaMemcpy(d_destSubblockIndices, h_destSubblockIndices, total * sizeof(size_t), cudaMemcpyHostToDevice);
    for (size_t r = 0; r < total; r++)  // one rank per launch => atomics execute in rank order
      k_dcgrid_refine_subblocks<<<1, 1>>>(d_grid, d_destSubblockIndices + r, 1, touchedCount + r, d_touchedBlockIndices);
    if (cudaDeviceSynchronize() != cudaSuccess) die("serialized refine failed");
    touchedCount += total;
  }

  void adaptTopology() override {
    k_dcgrid_calc_vorticity<<<(unsigned)h_grid->maxNumBlocks, DCGrid::blockV>>>(d_grid);
    size_t touchedCount = 0;
    moveBlocks(touchedCount);
    if (touchedCount > 0) {
      cudaMemset(h_grid->hashKey, 0xff, hashTableSize * sizeof(uint32_t));
      cudaMemset(h_grid->hashVal, 0xff, hashTableSize * sizeof(size_t));
      k_dcgrid_refill_hash_table<<<(unsigned)gridSize, (unsigned)blockSize>>>(d_grid);
      cudaDeviceSynchronize();
    }
    refineInRankOrder(touchedCount);
    if (touchedCount > 0) {
      k_dcgrid_refresh_apron_indices<<<(unsigned)h_grid->maxNumBlocks, DCGrid::apronV>>>(d_grid);
      cudaDeviceSynchronize();
      for (int level = h_grid->levels - 2; level >= 0; level--)
        k_dcgrid_propagate_values<<<(unsigned)touchedCount, DCGrid::blockV>>>(d_grid, d_touchedBlockIndices, level);
      cudaDeviceSynchronize();
    }
  }
};

static void die(const char *msg) {
  fprintf(stderr, "ref_harness: %s\n", msg);
  exit(2);
}
#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "ref_harness: %s -> %s\n", #x, cudaGetErrorString(e_));          \
      exit(3);                                                                         \
    }                                                                                  \
  } while (0)

// ---- dump container: [u32 name_len][name][u32 dtype][u64 count][payload] ----
enum { DT_F32 = 0, DT_I32 = 1, DT_U8 = 2, DT_U64 = 3 };
static const size_t kDtSize[] = {4, 4, 1, 8};
struct Dump {
  FILE *f = nullptr;
  void open(const std::string &path) {
    if (path.empty()) return;
    f = fopen(path.c_str(), "wb");
    if (!f) die("cannot open out file");
  }
  void put(const std::string &name, uint32_t dt, const void *host, uint64_t count) {
    if (!f) return;
    uint32_t nl = (uint32_t)name.size();
    fwrite(&nl, 4, 1, f);
    fwrite(name.data(), 1, nl, f);
    fwrite(&dt, 4, 1, f);
    fwrite(&count, 8, 1, f);
    fwrite(host, kDtSize[dt], count, f);
  }
  void put_dev(const std::string &name, uint32_t dt, const void *dev, uint64_t count) {
    if (!f) return;
    std::vector<unsigned char> h(count * kDtSize[dt]);
    CK(cudaMemcpy(h.data(), dev, h.size(), cudaMemcpyDeviceToHost));
    put(name, dt, h.data(), count);
  }
  void close() {
    if (f) fclose(f);
    f = nullptr;
  }
};

static uint64_t fnv1a(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) {
    h ^= b[i];
    h *= 1099511628211ull;
  }
  return h;
}

struct Args {
  std::string grid = "uniform", schedule = "project", out;
  int gx = 64, gy = 64, gz = 64, solids = 0, steps = 10, dump_reset = 0, trace = 0, reps = 1, serialize = 0;
  size_t M = 4096;
};

static Args parse(int argc, char **argv) {
  Args a;
  for (int i = 1; i < argc; i++) {
    std::string s(argv[i]);
    size_t eq = s.find('=');
    if (eq == std::string::npos) die("arguments are key=value");
    std::string k = s.substr(0, eq), v = s.substr(eq + 1);
    if (k == "grid") a.grid = v;
    else if (k == "schedule") a.schedule = v;
    else if (k == "out") a.out = v;
    else if (k == "d") a.gx = a.gy = a.gz = atoi(v.c_str());
    else if (k == "gx") a.gx = atoi(v.c_str());
    else if (k == "gy") a.gy = atoi(v.c_str());
    else if (k == "gz") a.gz = atoi(v.c_str());
    else if (k == "M") a.M = strtoull(v.c_str(), nullptr, 10);
    else if (k == "solids") a.solids = atoi(v.c_str());
    else if (k == "steps") a.steps = atoi(v.c_str());
    else if (k == "dump_reset") a.dump_reset = atoi(v.c_str());
    else if (k == "trace") a.trace = atoi(v.c_str());
    else if (k == "reps") a.reps = atoi(v.c_str());
    else if (k == "serialize") a.serialize = atoi(v.c_str());
    else die("unknown key");
  }
  return a;
}

static void dump_uniform(Dump &d, const std::string &pre, FluidSimulationUniform &s, bool with_pressure) {
  const size_t N = s.numCells;
  const size_t pyr = mipmapCells(s.size.x, s.size.y, s.size.z);
  d.put_dev(pre + "density", DT_F32, s.h_grid->density, N);
  d.put_dev(pre + "velocity", DT_F32, s.h_grid->velocity, 3 * N);
  d.put_dev(pre + "fluidity", DT_F32, s.h_grid->fluidity, N);
  if (with_pressure) {
    d.put_dev(pre + "pressure", DT_F32, s.h_grid->pressure, N);
    d.put_dev(pre + "t_pressure", DT_F32, s.h_grid->t_pressure, N);
    d.put_dev(pre + "divergence", DT_F32, s.h_grid->divergence, N);
    d.put_dev(pre + "pressure_pyramid", DT_F32, s.h_grid->pressure, pyr);
  }
}

static void dump_dcgrid_topology(Dump &d, const std::string &pre, FluidSimulationDCGrid &s) {
  const size_t M = s.h_grid->maxNumBlocks;
  const int L = s.h_grid->levels;
  d.put_dev(pre + "positions", DT_I32, s.h_grid->blockPositions, 3 * M);
  d.put_dev(pre + "levels", DT_U8, s.h_grid->blockLevels, M);
  d.put_dev(pre + "flags", DT_U8, s.h_grid->blockFlags, M);
  d.put_dev(pre + "parent", DT_U64, s.h_grid->parentIndices, M);
  d.put_dev(pre + "children", DT_U64, s.h_grid->childIndices, 8 * M);
  d.put_dev(pre + "apron", DT_U64, s.h_grid->cellIndices, 216 * M);
  d.put_dev(pre + "block_loads", DT_U64, s.h_grid->blockLoads, L);
  d.put(pre + "max_blocks", DT_U64, s.h_maxNumBlocksLevel, L);
  d.put(pre + "full_blocks", DT_U64, s.h_fullBlocksLevel, L);
  d.put(pre + "level_offsets", DT_U64, s.h_levelOffsets, L);
  d.put(pre + "move_limit", DT_U64, s.moveLimit, L);
  int32_t meta[2] = {L, s.h_grid->sparseLevels};
  d.put(pre + "meta_levels_sparse", DT_I32, meta, 2);
}

static void dump_dcgrid_fields(Dump &d, const std::string &pre, FluidSimulationDCGrid &s, bool with_pressure) {
  const size_t n = s.numCells;
  if (with_pressure) {
    d.put_dev(pre + "pressure", DT_F32, s.h_grid->pressure, n);
    d.put_dev(pre + "t_pressure", DT_F32, s.h_grid->t_pressure, n);
    d.put_dev(pre + "divergence", DT_F32, s.h_grid->divergence, n);
  } else {
    d.put_dev(pre + "density", DT_F32, s.h_grid->density, n);
    d.put_dev(pre + "velocity", DT_F32, s.h_grid->velocity, 3 * n);
    d.put_dev(pre + "fluidity", DT_F32, s.h_grid->fluidity, n);
  }
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
  Args a = parse(argc, argv);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) die("no CUDA device");
  CK(cudaSetDevice(0));

  // scenes.h:13-32 + main.cpp:17
  SimParams p = SimParams::defaultParams();
  p.gx = a.gx; p.gy = a.gy; p.gz = a.gz;
  p.dx = 10000.f / a.gx;
  p.rdx = 1.f / p.dx;
  p.enable_additional_solids = a.solids != 0;
  copySimParamsToDevice(p);  // BEFORE construction (simulation.cpp:19)

  int jacobi_pairs = -1;
  if (a.schedule.rfind("jacobi", 0) == 0) jacobi_pairs = atoi(a.schedule.c_str() + 6);
  const bool local = a.schedule == "local";

  for (int rep = 0; rep < a.reps; rep++) {
    Dump d;
    if (!a.out.empty()) d.open(rep == 0 ? a.out : a.out + ".rep" + std::to_string(rep));
    const int3 size = make_int3(a.gx, a.gy, a.gz);
    FluidSimulationUniform *u = nullptr;
    FluidSimulationDCGrid *g = nullptr;
    double t0 = now_ms();
    if (a.grid == "uniform") u = new FluidSimulationUniform(size);
    else if (a.grid == "dcgrid" && a.serialize) {
      // the constructor's own reset() still runs the base-class adaptTopology (no virtual dispatch
      // during construction); a second reset() re-initialises everything through the override
      g = new SerializedDCGrid(size, a.M);
      g->reset();
    } else if (a.grid == "dcgrid") g = new FluidSimulationDCGrid(size, a.M);
    else die("grid must be uniform or dcgrid");
    FluidSimulation *sim = u ? (FluidSimulation *)u : (FluidSimulation *)g;
    CK(cudaDeviceSynchronize());
    const double t_create = now_ms() - t0;

    if (a.dump_reset) {
      if (u) dump_uniform(d, "reset/", *u, false);
      if (g) { dump_dcgrid_topology(d, "reset/", *g); dump_dcgrid_fields(d, "reset/", *g, false); }
    }

    double t_av = 0, t_ad = 0, t_pr = 0, t_aq = 0;
    std::vector<uint64_t> digests;
    std::vector<unsigned char> hbuf;
    for (int s = 0; s < a.steps; s++) {
      const bool last = s == a.steps - 1;
      copySimParamsToDevice(p);  // simulation.cpp:94
      double t = now_ms();
      sim->advectVelocity();
      CK(cudaDeviceSynchronize());
      t_av += now_ms() - t; t = now_ms();
      sim->adaptTopology();
      CK(cudaDeviceSynchronize());
      t_ad += now_ms() - t; t = now_ms();
      if (jacobi_pairs >= 0) {
        if (!u) die("jacobi<N> schedule is uniform-only");
        // same kernels as projectLocal (fluid_simulation_uniform.cu:126-135) with N pairs
        k_uniform_calc_divergence<<<u->gridSize, u->blockSize>>>(u->d_grid);
        for (int i = 0; i < jacobi_pairs; i++) {
          k_uniform_jacobi<<<u->gridSizeLevel[0], u->blockSizeLevel[0]>>>(u->d_grid, 0);
          k_uniform_jacobi_inv<<<u->gridSizeLevel[0], u->blockSizeLevel[0]>>>(u->d_grid, 0);
        }
        k_uniform_apply_pressure<<<u->gridSize, u->blockSize>>>(u->d_grid);
      } else if (local) sim->projectLocal();
      else sim->project();
      CK(cudaDeviceSynchronize());
      t_pr += now_ms() - t;
      if (last) {
        if (u) dump_uniform(d, "final/", *u, true);
        if (g) dump_dcgrid_fields(d, "final/", *g, true);
      }
      t = now_ms();
      sim->advectDensity();
      CK(cudaDeviceSynchronize());
      t_aq += now_ms() - t;
      if (a.trace) {
        const size_t n = u ? u->numCells : g->numCells;
        hbuf.resize(n * 16);
        CK(cudaMemcpy(hbuf.data(), u ? (void *)u->h_grid->density : (void *)g->h_grid->density, n * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hbuf.data() + n * 4, u ? (void *)u->h_grid->velocity : (void *)g->h_grid->velocity, n * 12,
                      cudaMemcpyDeviceToHost));
        digests.push_back(fnv1a(hbuf.data(), n * 16));
      }
    }
    if (u) dump_uniform(d, "final/", *u, false);
    if (g) { dump_dcgrid_fields(d, "final/", *g, false); dump_dcgrid_topology(d, "final/", *g); }
    d.close();

    // digest of the final state (always)
    uint64_t dig = 0;
    {
      const size_t n = u ? u->numCells : g->numCells;
      hbuf.resize(n * 16);
      CK(cudaMemcpy(hbuf.data(), u ? (void *)u->h_grid->density : (void *)g->h_grid->density, n * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hbuf.data() + n * 4, u ? (void *)u->h_grid->velocity : (void *)g->h_grid->velocity, n * 12,
                    cudaMemcpyDeviceToHost));
      dig = fnv1a(hbuf.data(), n * 16);
    }
    const double t_step = a.steps > 0 ? (t_av + t_ad + t_pr + t_aq) / a.steps : 0.0;
    printf("{\"impl\": \"reference_cuda\", \"grid\": \"%s\", \"gx\": %d, \"gy\": %d, \"gz\": %d, \"M\": %zu, \"solids\": %d, "
           "\"schedule\": \"%s\", \"serialize\": %d, \"steps\": %d, \"rep\": %d, \"create_ms\": %.3f, \"ms_per_step\": %.4f, "
           "\"advect_velocity_ms\": %.4f, \"adapt_topology_ms\": %.4f, \"project_ms\": %.4f, \"advect_density_ms\": %.4f, "
           "\"final_digest\": \"%016llx\"",
           a.grid.c_str(), a.gx, a.gy, a.gz, a.M, a.solids, a.schedule.c_str(), a.serialize, a.steps, rep, t_create, t_step,
           a.steps ? t_av / a.steps : 0, a.steps ? t_ad / a.steps : 0, a.steps ? t_pr / a.steps : 0,
           a.steps ? t_aq / a.steps : 0, (unsigned long long)dig);
    if (a.trace) {
      printf(", \"trace\": [");
      for (size_t i = 0; i < digests.size(); i++) printf("%s\"%016llx\"", i ? ", " : "", (unsigned long long)digests[i]);
      printf("]");
    }
    printf("}\n");
    fflush(stdout);
    delete sim;
    CK(cudaDeviceSynchronize());
  }
  return 0;
}
