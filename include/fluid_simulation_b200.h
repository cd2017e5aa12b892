// fluid_simulation_b200.h — C++ adapter that plugs libdcgrid_b200.so into the reference tree.
//
// Drop this header next to the reference's sources (include path: the reference's `src/`) and
// link `-ldcgrid_b200`.  It subclasses the reference's abstract solver interface
//   class FluidSimulation                      (reference src/fluid_simulation.h:4-27)
// so that the two `new` expressions of the UI layer (src/simulation.cpp:32-35 and :58-61)
//   new FluidSimulationDCGrid(size, m_maxNumBlocks)  ->  new FluidSimulationB200DCGrid(size, m_maxNumBlocks)
//   new FluidSimulationUniform(size)                 ->  new FluidSimulationB200Uniform(size)
// are the only lines a maintainer edits; every virtual call of Simulation::updateSimulation
// (src/simulation.cpp:104-111) then lands in the sm_100a kernels through the C ABI of
// include/dcgrid_b200.h.  See INTEGRATION.md.
//
// SimParams: the reference keeps them in a process-global `__constant__ params` that the caller
// uploads with copySimParamsToDevice() before constructing a solver and once per frame
// (src/utils/sim_utils.cu:6-9, src/simulation.cpp:19,94).  The replacement keeps them per instance;
// call copySimParamsToB200(p) wherever the reference calls copySimParamsToDevice(p): it remembers
// the struct for the next constructor and forwards it to every live adapter instance.
//
// Errors: the reference prints and exit()s (include/cuda/helper_cuda.h:771-781); the adapter
// throws std::runtime_error carrying dcg_last_error() instead.
#pragma once
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "dcgrid_b200.h"
#include "data/sim_params.h"      // reference: struct SimParams (src/data/sim_params.h:14-47)
#include "fluid_simulation.h"     // reference: class FluidSimulation (src/fluid_simulation.h:4-27)

static_assert(sizeof(SimParams) == sizeof(dcg_sim_params),
              "dcg_sim_params must stay layout-identical to the reference's SimParams");

class FluidSimulationB200;
namespace dcg_b200_detail {
struct Registry {
  std::mutex mu;
  dcg_sim_params current{};
  bool have_params = false;
  std::vector<FluidSimulationB200 *> live;
  static Registry &get() {
    static Registry r;
    return r;
  }
};
}  // namespace dcg_b200_detail

class FluidSimulationB200 : public FluidSimulation {
public:
  ~FluidSimulationB200() override {
    auto &r = dcg_b200_detail::Registry::get();
    {
      std::lock_guard<std::mutex> lk(r.mu);
      for (size_t i = 0; i < r.live.size(); i++)
        if (r.live[i] == this) {
          r.live.erase(r.live.begin() + i);
          break;
        }
    }
    if (m_sim) dcg_destroy(m_sim);
  }

  // the nine virtuals of src/fluid_simulation.h:9-22
  void init() override { check(dcg_init(m_sim)); }
  void reset() override { check(dcg_reset(m_sim)); }
  void adaptTopology() override { check(dcg_adapt_topology(m_sim)); }
  void advectVelocity() override { check(dcg_advect_velocity(m_sim)); }
  void project() override { check(dcg_project(m_sim)); }
  void projectLocal() override { check(dcg_project_local(m_sim)); }
  void advectDensity() override { check(dcg_advect_density(m_sim)); }
  // Rendering (raymarching.cuh, *_rendering.cu) is out of scope: the surface is left untouched.
  void render(cudaSurfaceObject_t, const int2 &, const float3 &, const float3 &, const float &,
              const float3 &) override {}
  void debugStats() override {
    float v = 0.f;
    check(dcg_debug_stats(m_sim, &v));
    m_lastStat = v;
    // same text as the reference: "Smoke: %g" (fluid_simulation_uniform.cu:175) /
    // "Residual: %g" (fluid_simulation_dcgrid.cu:527)
    std::printf(dcg_is_dcgrid(m_sim) ? "Residual: %g\n" : "Smoke: %g\n", v);
  }

  // additions the reference lacks (north_star: step, field accessors)
  void step(int n = 1) { check(dcg_step(m_sim, n)); }
  void synchronize() { check(dcg_synchronize(m_sim)); }
  float lastStat() const { return m_lastStat; }
  size_t cellCount() const { return (size_t)dcg_num_cells(m_sim); }
  // field = DCG_FIELD_*, layout = DCG_LAYOUT_*; dst on the HOST
  void getField(int field, int layout, float *dst, size_t count) { check(dcg_get_field(m_sim, field, layout, dst, count)); }
  void density(float *dst, size_t count, int layout = DCG_LAYOUT_NATIVE) { getField(DCG_FIELD_DENSITY, layout, dst, count); }
  void velocity(float *dst, size_t count, int layout = DCG_LAYOUT_NATIVE) { getField(DCG_FIELD_VELOCITY, layout, dst, count); }
  void pressure(float *dst, size_t count, int layout = DCG_LAYOUT_NATIVE) { getField(DCG_FIELD_PRESSURE, layout, dst, count); }
  void fluidity(float *dst, size_t count, int layout = DCG_LAYOUT_NATIVE) { getField(DCG_FIELD_FLUIDITY, layout, dst, count); }
  // extensions beyond the reference snapshot (dcg_ext_params: flow-driven score, MacCormack, temperature / vapor +
  // the fused source pass, terrain, device-side selection); off unless switched on here
  void temperature(float *dst, size_t count, int layout = DCG_LAYOUT_NATIVE) { getField(DCG_FIELD_TEMPERATURE, layout, dst, count); }
  void vapor(float *dst, size_t count, int layout = DCG_LAYOUT_NATIVE) { getField(DCG_FIELD_VAPOR, layout, dst, count); }
  void setExtensions(const dcg_ext_params &e) { check(dcg_set_ext_params(m_sim, &e)); }
  void applySources() { check(dcg_apply_sources(m_sim)); }
  void sample(int field, bool precise, const float *positions, size_t n, float *out) { check(dcg_sample_field(m_sim, field, precise ? 1 : 0, positions, n, out)); }
  void saveState(const char *path) { check(dcg_save_state(m_sim, path)); }
  void loadState(const char *path) { check(dcg_load_state(m_sim, path)); }
  dcg_sim *handle() const { return m_sim; }

  void setParams(const dcg_sim_params &p) { check(dcg_set_params(m_sim, &p)); }

protected:
  explicit FluidSimulationB200(const int3 &size) : FluidSimulation(size) {}

  static dcg_sim_params startParams(const int3 &size) {
    auto &r = dcg_b200_detail::Registry::get();
    std::lock_guard<std::mutex> lk(r.mu);
    dcg_sim_params p;
    if (r.have_params) p = r.current;
    else dcg_default_params(&p);
    // the constructor argument wins, as in the reference (`size` sizes the buffers,
    // fluid_simulation_uniform.cu:6-8, fluid_simulation_dcgrid.cu:9-12)
    p.gx = size.x; p.gy = size.y; p.gz = size.z;
    return p;
  }
  void adopt(int rc, dcg_sim *s) {
    if (rc != DCG_OK) throw std::runtime_error(std::string("dcgrid_b200: ") + dcg_last_error(nullptr));
    m_sim = s;
    numCells = (size_t)dcg_num_cells(s);  // DCGrid: 64 * maxNumBlocks (fluid_simulation_dcgrid.cu:71)
    auto &r = dcg_b200_detail::Registry::get();
    std::lock_guard<std::mutex> lk(r.mu);
    r.live.push_back(this);
  }
  void check(int rc) const {
    if (rc != DCG_OK) throw std::runtime_error(std::string("dcgrid_b200: ") + dcg_last_error(m_sim));
  }

  dcg_sim *m_sim = nullptr;
  float m_lastStat = 0.f;
};

// drop-in for FluidSimulationUniform (src/uniformgrid/fluid_simulation_uniform.h:5-39)
class FluidSimulationB200Uniform : public FluidSimulationB200 {
public:
  explicit FluidSimulationB200Uniform(const int3 &size, int device = 0) : FluidSimulationB200(size) {
    const dcg_sim_params p = startParams(size);
    dcg_sim *s = nullptr;
    const int rc = dcg_create_uniform(&p, device, &s);  // sequenced before `s` is read
    adopt(rc, s);
  }
};

// drop-in for FluidSimulationDCGrid (src/dcgrid/fluid_simulation_dcgrid.h:5-61)
class FluidSimulationB200DCGrid : public FluidSimulationB200 {
public:
  FluidSimulationB200DCGrid(const int3 &size, const size_t &maxNumBlocks, int device = 0) : FluidSimulationB200(size) {
    const dcg_sim_params p = startParams(size);
    dcg_sim *s = nullptr;
    const int rc = dcg_create_dcgrid(&p, (uint64_t)maxNumBlocks, device, &s);
    adopt(rc, s);
  }
};

// Call wherever the reference calls copySimParamsToDevice (src/simulation.cpp:19,94).
inline void copySimParamsToB200(const SimParams &h_params) {
  auto &r = dcg_b200_detail::Registry::get();
  std::vector<FluidSimulationB200 *> live;
  {
    std::lock_guard<std::mutex> lk(r.mu);
    std::memcpy(&r.current, &h_params, sizeof r.current);
    r.have_params = true;
    live = r.live;
  }
  for (auto *s : live) {
    dcg_sim_params p = r.current;
    s->setParams(p);
  }
}
