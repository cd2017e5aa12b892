// Host-side base of one solver instance behind the C ABI (include/dcgrid_b200.h).
// Mirrors the reference's abstract class FluidSimulation (src/fluid_simulation.h:4-27):
// the same nine operations, but returning status codes instead of exit()ing.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "common.cuh"

#define DCG_CUDA_TRY(expr)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) return this->fail_cuda(e__, #expr, __FILE__, __LINE__);          \
  } while (0)

#define DCG_TRY(expr)                 \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != DCG_OK) return rc__;  \
  } while (0)

struct dcg_sim {
  dcg_sim_params params{};
  dcg_options opt{};  // creation-time options (include/dcgrid_b200.h); all-zero = defaults
  dcg_ext_params ext{};  // extensions beyond the reference snapshot; all-zero = the snapshot's behaviour
  dcg::KParams kp{};
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  bool dcgrid = false;

  // Jacobi schedule in pairs (jacobi + jacobi_inv); defaults = the reference's constants
  int project_coarsest_pairs = 0, project_level_pairs = 0, local_pairs = 0;

  uint64_t launches = 0;  // kernels launched by this instance (graph replays count their nodes)
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  float last_step_ms = 0.f;
  bool step_timing_pending = false;

  virtual ~dcg_sim() {}

  // allocation + reset(), like the reference constructors
  virtual int construct(const dcg_sim_params *p, int device) = 0;
  virtual void invalidate_graphs() {}

  // the FluidSimulation virtuals
  virtual int init() = 0;
  virtual int reset() = 0;
  virtual int adapt_topology() = 0;
  virtual int advect_velocity() = 0;
  virtual int project() = 0;
  virtual int project_local() = 0;
  virtual int advect_density() = 0;
  virtual int debug_stats(float *out) = 0;

  // additions
  virtual int step(int n);
  // extensions (include/dcgrid_b200.h, "extensions"): DCGrid instances implement them
  virtual int on_ext_changed() {
    if (ext.score_mode || ext.advection || ext.sources || ext.terrain) return fail(DCG_ERR_UNSUPPORTED, "extensions are implemented for DCGrid instances only");
    return DCG_OK;
  }
  virtual int apply_sources() { return DCG_OK; }
  virtual int sample_field(int, int, const float *, uint64_t, float *) { return fail(DCG_ERR_UNSUPPORTED, "sample_field: DCGrid instances only"); }
  virtual int save_state(const char *) { return fail(DCG_ERR_UNSUPPORTED, "save_state: DCGrid instances only"); }
  virtual int load_state(const char *) { return fail(DCG_ERR_UNSUPPORTED, "load_state: DCGrid instances only"); }
  int set_ext(const dcg_ext_params *e) {
    dcg_ext_params n{};
    const size_t bytes = e->struct_size == 0 || e->struct_size > sizeof(dcg_ext_params) ? sizeof(dcg_ext_params) : e->struct_size;
    std::memcpy(&n, e, bytes);
    n.struct_size = (uint32_t)sizeof(dcg_ext_params);
    if (n.score_mode < 0 || n.score_mode > 1 || n.advection < 0 || n.advection > 1) return fail(DCG_ERR_INVALID, "ext: score_mode and advection must be 0 or 1");
    if (n.sources && !(n.ambient_temperature > 0.f)) return fail(DCG_ERR_INVALID, "ext: sources need ambient_temperature > 0");
    if (n.terrain && !(n.terrain_wavelength > 0.f)) return fail(DCG_ERR_INVALID, "ext: terrain needs terrain_wavelength > 0");
    const dcg_ext_params saved = ext;
    const dcg::KParams saved_kp = kp;
    ext = n;
    kp = dcg::make_kparams(params, ext);
    const int rc = on_ext_changed();
    if (rc != DCG_OK) {
      ext = saved;
      kp = saved_kp;
    }
    return rc;
  }
  virtual int on_params_changed() = 0;
  virtual int total_density(double *out) = 0;
  virtual uint64_t num_cells() const = 0;
  virtual uint64_t max_num_blocks() const { return 0; }
  virtual int num_levels() const = 0;
  virtual int sparse_levels() const { return 0; }
  virtual int get_field(int field, int layout, float *dst, uint64_t count) = 0;
  virtual int get_level_table(uint64_t *, uint64_t *, uint64_t *, uint64_t *) { return fail(DCG_ERR_UNSUPPORTED, "uniform grid has no level table"); }
  virtual int get_topology(int32_t *, uint8_t *, uint64_t *, uint64_t *, uint64_t *) { return fail(DCG_ERR_UNSUPPORTED, "uniform grid has no block pool"); }
  virtual int lookup_blocks(const int32_t *, uint64_t, uint64_t *, uint8_t *) { return fail(DCG_ERR_UNSUPPORTED, "uniform grid has no block pool"); }
  virtual int get_counters(uint64_t out[8]) {
    for (int i = 0; i < 8; i++) out[i] = 0;
    out[6] = launches;
    return DCG_OK;
  }
  virtual int get_info(const char *key, double *out) {
    if (std::strcmp(key, "launches") == 0) { *out = (double)launches; return DCG_OK; }
    return fail(DCG_ERR_INVALID, "get_info: unknown key %s", key);
  }
  // multi-GPU instances, one rank per process: opaque 64-byte handles exchanged through any host channel
  virtual int export_handle(void *, uint64_t) { return fail(DCG_ERR_UNSUPPORTED, "not a sharded one-rank-per-process instance"); }
  virtual int import_handles(const void *, int) { return fail(DCG_ERR_UNSUPPORTED, "not a sharded one-rank-per-process instance"); }
  virtual int algorithmic_bytes(double *bytes, uint64_t *active_blocks) = 0;
  virtual int bench_stage(const char *stage, int level, int reps, float *ms_per_launch, double *alg_bytes) = 0;

  int set_params(const dcg_sim_params *p) {
    // the reference re-uploads SimParams every frame (src/simulation.cpp:94); identical bytes change
    // nothing here: kernels take the parameters by value and captured graphs stay valid
    if (std::memcmp(&params, p, sizeof params) == 0) return DCG_OK;
    // a rejected change (grid size is fixed at construction) must leave the instance as it was: the kernels
    // index their buffers with kp.gx/gy/gz
    const dcg_sim_params saved = params;
    const dcg::KParams saved_kp = kp;
    params = *p;
    kp = dcg::make_kparams(params, ext);
    const int rc = on_params_changed();
    if (rc != DCG_OK) {
      params = saved;
      kp = saved_kp;
    }
    return rc;
  }
  void set_options(const dcg_options *o) {
    opt = dcg_options{};
    if (!o) return;
    const size_t n = o->struct_size == 0 || o->struct_size > sizeof(dcg_options) ? sizeof(dcg_options) : o->struct_size;
    std::memcpy(&opt, o, n);
  }
  // errors a kernel can only report through device memory (sharded: a peer that never reached a barrier)
  virtual int health_check() { return DCG_OK; }
  int synchronize() {
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    if (step_timing_pending) {
      DCG_CUDA_TRY(cudaEventElapsedTime(&last_step_ms, ev_begin, ev_end));
      step_timing_pending = false;
    }
    return health_check();
  }

  int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
  int fail_cuda(cudaError_t e, const char *expr, const char *file, int line) {
    return fail(DCG_ERR_CUDA, "%s failed: %s (%s:%d)", expr, cudaGetErrorString(e), file, line);
  }
  int base_setup(const dcg_sim_params *p, int dev) {
    params = *p;
    kp = dcg::make_kparams(params, ext);
    device = dev;
    DCG_CUDA_TRY(cudaSetDevice(device));
    DCG_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    DCG_CUDA_TRY(cudaEventCreate(&ev_begin));
    DCG_CUDA_TRY(cudaEventCreate(&ev_end));
    return DCG_OK;
  }
  void base_teardown() {
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    if (stream) cudaStreamDestroy(stream);
    ev_begin = ev_end = nullptr;
    stream = nullptr;
  }
};

void dcg_set_create_error(const char *msg);  // text returned by dcg_last_error(NULL)
dcg_sim *dcg_make_uniform();
dcg_sim *dcg_make_dcgrid(uint64_t max_num_blocks);
dcg_sim *dcg_make_dcgrid_sharded(uint64_t max_num_blocks, int rank, int world, int nlocal);
