// adapter_demo — test infrastructure.  Drives the PRODUCT (libdcgrid_b200.so) exactly the way the
// reference's UI layer drives its solvers: through a `FluidSimulation*` (the reference's abstract
// class, compiled from /root/reference/src/fluid_simulation.cpp where it lies) and the call sequence
// of Simulation::updateSimulation (src/simulation.cpp:93-116).  The only edit a maintainer makes —
// the class named in the two `new` expressions (src/simulation.cpp:32-35) — is what this file shows.
//
// usage: adapter_demo grid=uniform|dcgrid d=64 M=2000 solids=0|1 steps=4
// prints one JSON line with an FNV-1a digest of density+velocity (native layout).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fluid_simulation_b200.h"

static uint64_t fnv(uint64_t h, const void *p, size_t n) {
  const unsigned char *b = static_cast<const unsigned char *>(p);
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char **argv) {
  std::string grid = "dcgrid";
  int d = 64, solids = 0, steps = 4;
  size_t M = 2000;
  for (int i = 1; i < argc; i++) {
    std::string a(argv[i]);
    const size_t e = a.find('=');
    if (e == std::string::npos) continue;
    const std::string k = a.substr(0, e), v = a.substr(e + 1);
    if (k == "grid") grid = v;
    else if (k == "d") d = atoi(v.c_str());
    else if (k == "M") M = strtoull(v.c_str(), nullptr, 10);
    else if (k == "solids") solids = atoi(v.c_str());
    else if (k == "steps") steps = atoi(v.c_str());
  }
  SimParams p = SimParams::defaultParams();           // src/data/sim_params.cpp:4-34
  p.gx = p.gy = p.gz = d;
  p.dx = 10000.f / d;
  p.rdx = 1.f / p.dx;                                 // src/main.cpp:17
  p.enable_additional_solids = solids != 0;
  try {
    copySimParamsToB200(p);                           // was: copySimParamsToDevice(m_params), simulation.cpp:19
    const int3 size = make_int3(p.gx, p.gy, p.gz);
    FluidSimulation *sim = nullptr;
    if (grid == "dcgrid") sim = new FluidSimulationB200DCGrid(size, M);  // was: new FluidSimulationDCGrid(size, m_maxNumBlocks)
    else sim = new FluidSimulationB200Uniform(size);                     // was: new FluidSimulationUniform(size)
    for (int s = 0; s < steps; s++) {                 // Simulation::updateSimulation, simulation.cpp:93-116
      copySimParamsToB200(p);
      sim->advectVelocity();
      sim->adaptTopology();
      sim->project();
      sim->advectDensity();
    }
    auto *b = static_cast<FluidSimulationB200 *>(sim);
    const size_t n = b->cellCount();
    std::vector<float> q(n), v(3 * n);
    b->density(q.data(), n);
    b->velocity(v.data(), 3 * n);
    uint64_t h = 1469598103934665603ull;
    h = fnv(h, q.data(), n * 4);
    h = fnv(h, v.data(), 3 * n * 4);
    std::printf("{\"grid\": \"%s\", \"d\": %d, \"M\": %zu, \"steps\": %d, \"cells\": %zu, \"digest\": \"%016llx\"}\n", grid.c_str(), d, M,
                steps, n, (unsigned long long)h);
    delete sim;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "adapter_demo: %s\n", e.what());
    return 1;
  }
  return 0;
}
