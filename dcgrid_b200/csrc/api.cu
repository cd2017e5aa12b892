// extern "C" surface of libdcgrid_b200.so — see include/dcgrid_b200.h for the contract and the
// reference interface (file:line) each entry point replaces.
#include <cstring>
#include <mutex>
#include <string>

#include "sim.h"

namespace {
std::mutex g_err_mutex;
std::string g_create_error;

void set_create_error(const std::string &s) {
  std::lock_guard<std::mutex> lk(g_err_mutex);
  g_create_error = s;
}

int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_create_error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                     "); dcgrid_b200 has no CPU fallback");
    cudaGetLastError();
    return DCG_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    set_create_error("device ordinal out of range");
    return DCG_ERR_INVALID;
  }
  return DCG_OK;
}

int finish_create(dcg_sim *s, const dcg_sim_params *params, int device, dcg_sim **out, const dcg_options *opt = nullptr) {
  s->set_options(opt);
  int rc = s->construct(params, device);
  if (rc == DCG_OK) rc = s->synchronize();
  if (rc != DCG_OK) {
    set_create_error(s->err);
    delete s;
    return rc;
  }
  *out = s;
  return DCG_OK;
}
}  // namespace

void dcg_set_create_error(const char *msg) { set_create_error(msg ? msg : ""); }

#define NEED(sim)                 \
  do {                            \
    if (!(sim)) return DCG_ERR_INVALID; \
  } while (0)

extern "C" {

int dcg_default_params(dcg_sim_params *p) {  // src/data/sim_params.cpp:4-34 (+ rdx, src/main.cpp:17)
  if (!p) return DCG_ERR_INVALID;
  std::memset(p, 0, sizeof *p);
  p->gx = p->gy = p->gz = 128;
  p->dx = 10000.f / 128;
  p->dt = 3.f;
  p->velocity_emission_rate = 150.f;
  p->density_emission_rate = 0.002f;
  p->emission_radius = 750.f;
  p->enable_additional_solids = false;
  p->render_solids = p->render_shadows = p->render_precise = true;
  p->render_channel = 4;  // RenderChannel::Resolution
  p->aa_samples = 1.f;
  p->ambient = .3f;
  p->background_color = {0.f, 0.f, 0.f};
  p->floor_color = {178.f / 255.f, 158.f / 255.f, 135.f / 255.f};
  p->smoke_color = {.9f, .9f, .9f};
  p->scene_color = {107.f / 255.f, 163.f / 255.f, 204.f / 255.f};
  p->rdx = 1.f / p->dx;
  return DCG_OK;
}

int dcg_default_options(dcg_options *o) {
  if (!o) return DCG_ERR_INVALID;
  std::memset(o, 0, sizeof *o);
  o->struct_size = (uint32_t)sizeof *o;
  return DCG_OK;
}

int dcg_create_uniform_opt(const dcg_sim_params *params, int device, const dcg_options *opt, dcg_sim **out) {
  if (!params || !out) return DCG_ERR_INVALID;
  *out = nullptr;
  int rc = check_device(device);
  if (rc != DCG_OK) return rc;
  return finish_create(dcg_make_uniform(), params, device, out, opt);
}
int dcg_create_uniform(const dcg_sim_params *params, int device, dcg_sim **out) { return dcg_create_uniform_opt(params, device, nullptr, out); }

int dcg_create_dcgrid_opt(const dcg_sim_params *params, uint64_t max_num_blocks, int device, const dcg_options *opt, dcg_sim **out) {
  if (!params || !out) return DCG_ERR_INVALID;
  *out = nullptr;
  int rc = check_device(device);
  if (rc != DCG_OK) return rc;
  return finish_create(dcg_make_dcgrid(max_num_blocks), params, device, out, opt);
}
int dcg_create_dcgrid(const dcg_sim_params *params, uint64_t max_num_blocks, int device, dcg_sim **out) {
  return dcg_create_dcgrid_opt(params, max_num_blocks, device, nullptr, out);
}

int dcg_create_dcgrid_sharded_opt(const dcg_sim_params *params, uint64_t max_num_blocks, int device, int rank, int world, int nlocal,
                                  const dcg_options *opt, dcg_sim **out) {
  if (!params || !out) return DCG_ERR_INVALID;
  *out = nullptr;
  int rc = check_device(device);
  if (rc != DCG_OK) return rc;
  return finish_create(dcg_make_dcgrid_sharded(max_num_blocks, rank, world, nlocal), params, device, out, opt);
}
int dcg_create_dcgrid_sharded(const dcg_sim_params *params, uint64_t max_num_blocks, int device, int rank, int world, int nlocal,
                              dcg_sim **out) {
  return dcg_create_dcgrid_sharded_opt(params, max_num_blocks, device, rank, world, nlocal, nullptr, out);
}

int dcg_destroy(dcg_sim *sim) {
  NEED(sim);
  delete sim;
  return DCG_OK;
}

int dcg_set_params(dcg_sim *sim, const dcg_sim_params *params) {
  NEED(sim);
  if (!params) return DCG_ERR_INVALID;
  return sim->set_params(params);
}
int dcg_get_params(const dcg_sim *sim, dcg_sim_params *out) {
  NEED(sim);
  if (!out) return DCG_ERR_INVALID;
  *out = sim->params;
  return DCG_OK;
}

int dcg_init(dcg_sim *sim) { NEED(sim); return sim->init(); }
int dcg_reset(dcg_sim *sim) { NEED(sim); return sim->reset(); }
int dcg_adapt_topology(dcg_sim *sim) { NEED(sim); return sim->adapt_topology(); }
int dcg_advect_velocity(dcg_sim *sim) { NEED(sim); return sim->advect_velocity(); }
int dcg_project(dcg_sim *sim) { NEED(sim); return sim->project(); }
int dcg_project_local(dcg_sim *sim) { NEED(sim); return sim->project_local(); }
int dcg_advect_density(dcg_sim *sim) { NEED(sim); return sim->advect_density(); }
int dcg_render(dcg_sim *sim) {
  NEED(sim);
  return sim->fail(DCG_ERR_UNSUPPORTED, "render() is out of scope (raymarching.cuh, *_rendering.cu)");
}
int dcg_debug_stats(dcg_sim *sim, float *out) {
  NEED(sim);
  if (!out) return DCG_ERR_INVALID;
  return sim->debug_stats(out);
}

int dcg_step(dcg_sim *sim, int n) {
  NEED(sim);
  if (n < 0) return sim->fail(DCG_ERR_INVALID, "step count must be >= 0");
  return sim->step(n);
}
int dcg_synchronize(dcg_sim *sim) { NEED(sim); return sim->synchronize(); }

int dcg_set_jacobi_schedule(dcg_sim *sim, int coarsest, int level, int local) {
  NEED(sim);
  if (coarsest < 0 || level < 0 || local < 0) return sim->fail(DCG_ERR_INVALID, "pair counts must be >= 0");
  sim->project_coarsest_pairs = coarsest;
  sim->project_level_pairs = level;
  sim->local_pairs = local;
  sim->invalidate_graphs();
  return DCG_OK;
}

int dcg_default_ext_params(dcg_ext_params *e) {
  if (!e) return DCG_ERR_INVALID;
  std::memset(e, 0, sizeof *e);
  e->struct_size = (uint32_t)sizeof *e;
  // switches off; coefficients of a plausible moist plume (world units as in SimParams: dx = 10000 / gx)
  e->buoyancy = 9.81f;
  e->vapor_buoyancy = 5.9f;       // 0.61 g
  e->smoke_weight = 9.81f;
  e->ambient_temperature = 290.f;
  e->ambient_lapse = 0.003f;
  e->adiabatic_lapse = 0.0098f;
  e->vorticity_confinement = 0.05f;
  e->saturation_base = 0.012f;
  e->saturation_slope = 0.0001f;
  e->condensation_rate = 0.5f;
  e->latent_heat = 2500.f;
  e->temperature_emission = 8.f;
  e->vapor_emission = 0.02f;
  e->ambient_vapor = 0.004f;
  e->terrain_height = 24.f;
  e->terrain_wavelength = 32.f;
  return DCG_OK;
}
int dcg_set_ext_params(dcg_sim *sim, const dcg_ext_params *e) {
  NEED(sim);
  if (!e) return DCG_ERR_INVALID;
  return sim->set_ext(e);
}
int dcg_get_ext_params(const dcg_sim *sim, dcg_ext_params *out) {
  NEED(sim);
  if (!out) return DCG_ERR_INVALID;
  *out = sim->ext;
  out->struct_size = (uint32_t)sizeof *out;
  return DCG_OK;
}
int dcg_apply_sources(dcg_sim *sim) { NEED(sim); return sim->apply_sources(); }
int dcg_sample_field(dcg_sim *sim, int field, int mode, const float *positions, uint64_t n, float *out) {
  NEED(sim);
  if ((!positions || !out) && n > 0) return DCG_ERR_INVALID;
  if (mode != 0 && mode != 1) return sim->fail(DCG_ERR_INVALID, "sample_field: mode must be 0 (coarse) or 1 (precise)");
  return sim->sample_field(field, mode, positions, n, out);
}
int dcg_save_state(dcg_sim *sim, const char *path) {
  NEED(sim);
  if (!path) return DCG_ERR_INVALID;
  return sim->save_state(path);
}
int dcg_load_state(dcg_sim *sim, const char *path) {
  NEED(sim);
  if (!path) return DCG_ERR_INVALID;
  return sim->load_state(path);
}

int dcg_total_density(dcg_sim *sim, double *out) {
  NEED(sim);
  if (!out) return DCG_ERR_INVALID;
  return sim->total_density(out);
}

int dcg_is_dcgrid(const dcg_sim *sim) { return sim && sim->dcgrid ? 1 : 0; }
uint64_t dcg_num_cells(const dcg_sim *sim) { return sim ? sim->num_cells() : 0; }
uint64_t dcg_max_num_blocks(const dcg_sim *sim) { return sim ? sim->max_num_blocks() : 0; }
int dcg_num_levels(const dcg_sim *sim) { return sim ? sim->num_levels() : 0; }
int dcg_sparse_levels(const dcg_sim *sim) { return sim ? sim->sparse_levels() : 0; }

int dcg_get_field(dcg_sim *sim, int field, int layout, float *dst, uint64_t count) {
  NEED(sim);
  return sim->get_field(field, layout, dst, count);
}
int dcg_get_level_table(dcg_sim *sim, uint64_t *mx, uint64_t *full, uint64_t *loads, uint64_t *offs) {
  NEED(sim);
  return sim->get_level_table(mx, full, loads, offs);
}
int dcg_get_topology(dcg_sim *sim, int32_t *positions, uint8_t *levels, uint64_t *parent, uint64_t *children,
                     uint64_t *apron) {
  NEED(sim);
  return sim->get_topology(positions, levels, parent, children, apron);
}
int dcg_lookup_blocks(dcg_sim *sim, const int32_t *positions, uint64_t n, uint64_t *out_slot, uint8_t *out_level) {
  NEED(sim);
  if (!positions || !out_slot || !out_level) return DCG_ERR_INVALID;
  return sim->lookup_blocks(positions, n, out_slot, out_level);
}
int dcg_get_counters(dcg_sim *sim, uint64_t out[8]) {
  NEED(sim);
  if (!out) return DCG_ERR_INVALID;
  return sim->get_counters(out);
}
int dcg_get_info(dcg_sim *sim, const char *key, double *out) {
  NEED(sim);
  if (!key || !out) return DCG_ERR_INVALID;
  return sim->get_info(key, out);
}
int dcg_last_step_ms(dcg_sim *sim, float *out) {
  NEED(sim);
  if (!out) return DCG_ERR_INVALID;
  *out = sim->last_step_ms;
  return DCG_OK;
}
int dcg_algorithmic_bytes(dcg_sim *sim, double *bytes, uint64_t *active) {
  NEED(sim);
  return sim->algorithmic_bytes(bytes, active);
}

int dcg_bench_stage(dcg_sim *sim, const char *stage, int level, int reps, float *ms_per_launch, double *alg_bytes) {
  NEED(sim);
  if (!stage || reps <= 0) return sim->fail(DCG_ERR_INVALID, "bench_stage: bad arguments");
  return sim->bench_stage(stage, level, reps, ms_per_launch, alg_bytes);
}

uint64_t dcg_fnv1a64(const void *data, uint64_t bytes, uint64_t seed) {
  uint64_t h = seed ? seed : 14695981039346656037ull;
  const unsigned char *b = static_cast<const unsigned char *>(data);
  for (uint64_t i = 0; i < bytes; i++) {
    h ^= b[i];
    h *= 1099511628211ull;
  }
  return h;
}

const char *dcg_last_error(const dcg_sim *sim) {
  if (sim) return sim->err.c_str();
  std::lock_guard<std::mutex> lk(g_err_mutex);
  static thread_local std::string copy;
  copy = g_create_error;
  return copy.c_str();
}
const char *dcg_version(void) { return "dcgrid_b200 0.1 (sm_100a)"; }

}  // extern "C"
