// B200-native adaptive solver: drop-in for FluidSimulationDCGrid
// (reference src/dcgrid/fluid_simulation_dcgrid.{h,cu}; kernels in dcgrid_kernels.cuh).
//
// Host orchestration follows the reference's schedule exactly (same kernels per stage, same
// per-level order), with these structural changes:
//  * fields ping-pong instead of the two whole-pool D2D memcpys (fluid_simulation_dcgrid.cu:265,315);
//  * explicit p / t_p / div buffers instead of the aliased `temporary` (dcgrid_structure.cu:94-102);
//  * the dead vorticity pass (:321; its only consumer is commented out, dcgrid_adaptation.cu:36-39)
//    is not launched;
//  * the hash table is replaced by dense per-level maps updated incrementally (dcgrid_layout.cuh);
//  * pool slots are allocated in rank order instead of by racing atomicAdds;
//  * adaptTopology() is skipped once a fixed point of the (topology, moveLimit) state is proven:
//    scores depend only on geometry in this snapshot, so a call that changes nothing and leaves
//    moveLimit unchanged will repeat forever.  Until then the reference's host
//    selection (std::nth_element / std::sort with its comparators, :348-483) runs verbatim on
//    scores computed on the device, which keeps block maps identical by construction.
//  * in the steady state a 2-step CUDA graph replaces ~270 launches.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "dcgrid_ext.cuh"
#include "dcgrid_kernels.cuh"
#include "dcgrid_pipe.cuh"
#include "shard_vmm.h"
#include "sim.h"

namespace dcg {
namespace {

inline unsigned blocks_for(size_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

struct DCGridSim : dcg_sim {
  uint64_t M64 = 0;
  uint32_t M = 0;
  int gx = 0, gy = 0, gz = 0, levels = 0, sparse = 0;
  size_t cells = 0;
  std::vector<uint64_t> max_blocks, full_blocks, loads, offsets, move_limit;
  std::vector<size_t> map_size;
  size_t fmap_size = 0;

  Pool T{};   // the reference's numbering: adaptation, accessors
  Pool Tf{};  // field order (k_dc_resort_keys): what the field kernels see once `mirrored`
  bool mirrored = false;
  uint32_t *d_perm = nullptr, *d_perm_new = nullptr;  // reference slot -> field slot
  unsigned long long *d_sort_keys64[2] = {nullptr, nullptr};
  void *d_sort_tmp64 = nullptr;
  size_t sort_tmp64_bytes = 0;
  bool use_resort = true;
  bool sort_ordered = false;  // several ranks stacked along y or z: the ordered levels are renumbered along that axis too
  int resort_every = 32;      // topology changes between re-sorts during the transient (0 = only at the fixed point)
  int changes_since_resort = 0;
  uint64_t n_resorts = 0;
  const Pool &hot() const { return mirrored ? Tf : T; }
  uint32_t *d_flags = nullptr, *d_free = nullptr, *d_touched = nullptr, *d_to_move = nullptr, *d_dest = nullptr;
  int4 *d_new_posl = nullptr;
  float *d_sub_scores = nullptr, *d_block_scores = nullptr;
  ScoreSummary *d_summary = nullptr, *h_summary = nullptr;  // per-level score extrema (device-reduced, 192 B D2H)
  uint32_t *d_flag_bits = nullptr;  // one bit per slot: flags != 0 (k_dc_flag_bits)
  uint32_t *d_counters = nullptr;  // [0] failed allocations, [1] irregular-face blocks (last build)
  uint64_t n_irregular = 0, n_host_selections = 0, n_levels_shortcut = 0, n_device_selections = 0, n_selection_fallbacks = 0;
  float4 *vw[2] = {nullptr, nullptr};
  float *q[2] = {nullptr, nullptr};
  float *fl = nullptr, *p = nullptr, *tp = nullptr, *div = nullptr;
  float *scratch = nullptr;  // 3*max(cells, gx*gy*gz) floats for accessors / stats
  size_t scratch_floats = 0;
  double *d_partial = nullptr, *h_partial = nullptr;
  int cur_v = 0, cur_q = 0;
  // extensions (dcgrid_ext.cuh): temperature / vapor ping-pong pairs, vorticity (+ |omega| in .w), MacCormack scratch;
  // allocated when an extension that needs them is switched on
  float *th[2] = {nullptr, nullptr}, *qvp[2] = {nullptr, nullptr}, *s_mc = nullptr;
  float4 *vort = nullptr, *vw_mc = nullptr;
  int cur_s = 0;
  bool ext_on() const { return ext.score_mode || ext.advection || ext.sources || ext.selection; }
  // device-side selection (dcgrid_ext.cuh, namespace sel): sort keys, per-level scalars, compaction scratch
  unsigned long long *d_sel_keys[2] = {nullptr, nullptr};
  void *d_sel_tmp = nullptr;
  size_t sel_tmp_bytes = 0, sel_max_n = 0;
  uint32_t *d_sel_sc = nullptr, *h_sel_sc = nullptr, *d_sel_cta = nullptr, *d_sel_mc = nullptr, *d_sel_dc = nullptr;

  struct Cand { float s; uint32_t id; };  // (score, id): what the host selection sorts
  std::vector<Cand> sel_mc, sel_dc;
  // pinned host mirrors for the selection
  float *h_sub_scores = nullptr, *h_block_scores = nullptr;
  uint32_t *h_to_move = nullptr, *h_dest = nullptr;

  // host wall time spent in the phases of adaptTopology() (incl. their stream synchronisations): dcg_get_info "adapt_*_ms"
  double t_move_ms = 0, t_refine_ms = 0, t_apron_ms = 0, t_layout_ms = 0, t_propagate_ms = 0;
  // early scores: the geometric scores the NEXT adaptTopology() needs depend on the topology only, so they are computed as
  // soon as this call has settled it and copied to the host on a side stream while the step's field kernels run
  bool early_ready = false, early_copy_pending = false, use_early_scores = true;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_scores = nullptr, ev_scores_copied = nullptr;
  ScoreSummary *h_summary_init = nullptr;
  uint64_t n_early_scores = 0;
  double t_sel_scores_ms = 0, t_sel_d2h_ms = 0, t_sel_host_ms = 0;  // inside moveBlocks: score kernels + summary, score slices D2H, host selection
  static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  bool timing_sync = false;  // diagnostics: synchronise between the phases so that the split above is exact
  bool steady = false;
  uint64_t n_adapt = 0, n_changed = 0, n_moved = 0, n_refined = 0, n_skipped = 0, n_failed = 0;

  // persistent TMA-ring kernels (dcgrid_pipe.cuh): resident CTAs per device, sweep direction toggle
  int sm_count = 0, jacobi_pipe_ctas = 0, advect_per_sm[3] = {0, 0, 0}, advect_ext_per_sm[3] = {0, 0, 0};
  bool use_advect_pipe = true;
  // k_dc_advect_pipe<2>: advect_density() also produces the NEXT step's advected velocity in vw[cur_v ^ 1];
  // spec_velocity = that buffer is valid, i.e. nothing has touched velocity, topology or parameters since
  bool fuse_advect = true, spec_velocity = false;
  // processing order of the advection kernels: active slots sorted along a Morton curve (k_dc_order_keys)
  uint32_t *d_order_keys[2] = {nullptr, nullptr}, *d_order_vals = nullptr;
  void *d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  int order_mode = 1;    // 0 = pool-slot order, 1 = Morton
  uint32_t *d_pcount = nullptr, *h_pcount = nullptr;  // scratch of k_dc_list_parents / k_dc_order_keys counts

  // ---- slab decomposition --------------------------------------------------------------------
  // The pool is cut into `unit`-slot pieces, each owned by one of `world` ranks (per level: contiguous, equal
  // shares of the level's slot range).  A rank runs every field kernel on the tiles it owns; cells of other
  // ranks are read / written in place: one process = all ranks on one device (nlocal == world, plain
  // cudaMalloc), or one process per GPU (nlocal == 1) with the fields in a virtual range stitched from every
  // rank's arena (shard_vmm.h), the same cell id addressing the same cell on every GPU over NVLink.  Topology
  // (block pool, maps, descriptors, adaptation incl. the host selection) is replicated: every process keeps
  // it and updates it identically.  Ranks run in lock step: a flag barrier over peer memory after each phase.
  int world = 1, rank0 = 0, nlocal = 1;
  bool vmm = false, ready = true;
  uint32_t unit = 0, nunits = 0;
  std::vector<uint8_t> unit_owner;
  uint8_t *d_unit_owner = nullptr;
  struct RankWork {
    TileRuns all{};                 // absolute tiles (16 slots from slot 0) of the rank's units
    std::vector<TileRuns> level;    // level-local tiles (16 slots from levelOffsets[l]) inside the active prefix
    uint32_t *d_order = nullptr;    // Morton-ordered active slots of the rank, padded with kNone
    uint32_t n_order = 0;
    uint32_t *d_plist = nullptr;    // blocks with children, per level at [offsets[l] ...)
    uint32_t pcount[kMaxLevels] = {0};
    uint32_t *d_tstarts = nullptr;  // k_dc_restrict_tree: the blocks its 8-lane groups start from
    uint32_t n_tstarts = 0;
  };
  // one-launch restriction (k_dc_restrict_tree): per-block expected counts | level, completion counters, top level of the walk
  uint8_t *d_texpect = nullptr;
  uint32_t *d_tcount = nullptr;
  int tree_top = 0;
  bool use_tree = true;
  bool spec_restricted = false;  // the speculative velocity (advect_both) has been restricted together with the density
  std::vector<RankWork> work;
  std::vector<char> level_single;  // the active blocks of the level all belong to one rank: its sweeps need no barrier between them
  // cross-process state (vmm)
  vmm::Driver drv;
  vmm::FdServer fd_server;
  // physical pieces: per rank one control block + one allocation per (run of consecutive owned units, field)
  struct UnitRun { uint32_t u0, u1; int owner; };
  std::vector<UnitRun> unit_runs;                        // maximal runs, ascending
  std::vector<std::vector<CUmemGenericAllocationHandle>> pieces;  // [rank][0] = control, [1 + k * kFields + f] = k-th run of the rank
  size_t gran = 0;
  static constexpr int kFields = 8;              // vw0 vw1 q0 q1 fl p tp div
  CUdeviceptr field_va[kFields] = {0}, ctrl_va = 0;
  size_t field_va_bytes[kFields] = {0};
  uint32_t *d_epoch = nullptr, *d_barrier_err = nullptr;
  uint64_t n_barriers = 0;
  bool prolong_staged = true;
  bool use_stencil_pipe = true;  // k_dc_divergence_pipe / k_dc_apply_pipe instead of the one-CTA-per-tile kernels
  int div_pipe_ctas = 0, apply_pipe_ctas = 0, apply_min_blocks = 2;
  bool coarse_in_smem = true;
  static constexpr size_t kCoarseSmemMax = 200 * 1024;
  bool skip_dead_zeroing = true;  // k_dc_divergence4: no pressure clears that project() never reads
  int experiment = 0;  // dcg_options.experiment
  int advect_min_blocks = 3;  // __launch_bounds__ variant of the advection kernels (3 or 4 CTAs per SM)
  bool use_pipe = true, snake = true;
  bool jacobi8 = true;  // k_dc_jacobi_pipe8 (8 cells per thread) instead of k_dc_jacobi_pipe (4)
  int jacobi8_ctas = 0;
  unsigned pipe_min_tiles = 0;  // levels with fewer tiles take the one-CTA-per-tile kernel
  int sweep_parity = 0;

  cudaGraphExec_t step_graph[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t step_graph_launches = 0;

  explicit DCGridSim(uint64_t m) : M64(m) { dcgrid = true; }

  ~DCGridSim() override {
    cudaSetDevice(device);
    drop_graphs();
    cudaFree(T.posl); cudaFree(T.parent); cudaFree(T.child); cudaFree(T.apron); cudaFree(T.face); cudaFree(T.fd); cudaFree(T.fmap);
    for (int l = 0; l < kMaxLevels; l++) cudaFree(T.map[l]);
    cudaFree(d_perm); cudaFree(d_perm_new); cudaFree(d_sort_keys64[0]); cudaFree(d_sort_keys64[1]); cudaFree(d_sort_tmp64);
    cudaFree(Tf.posl); cudaFree(Tf.parent); cudaFree(Tf.child); cudaFree(Tf.apron);
    for (int l = 0; l < kMaxLevels; l++) cudaFree(Tf.map[l]);
    cudaFree(d_flags); cudaFree(d_free); cudaFree(d_touched); cudaFree(d_to_move); cudaFree(d_dest); cudaFree(d_counters); cudaFree(d_flag_bits); cudaFree(d_summary);
    if (h_summary) cudaFreeHost(h_summary);
    if (h_summary_init) cudaFreeHost(h_summary_init);
    if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
    if (ev_scores) cudaEventDestroy(ev_scores);
    if (ev_scores_copied) cudaEventDestroy(ev_scores_copied);
    cudaFree(d_new_posl); cudaFree(d_sub_scores); cudaFree(d_block_scores);
    cudaFree(d_pcount);
    if (h_pcount) cudaFreeHost(h_pcount);
    for (auto &w : work) { cudaFree(w.d_order); cudaFree(w.d_plist); cudaFree(w.d_tstarts); }
    cudaFree(d_texpect); cudaFree(d_tcount);
    cudaFree(d_unit_owner); cudaFree(d_epoch); cudaFree(d_barrier_err);
    cudaFree(d_order_keys[0]); cudaFree(d_order_keys[1]); cudaFree(d_order_vals); cudaFree(d_sort_tmp);
    if (vmm) {
      cudaDeviceSynchronize();
      fd_server.finish();
      release_vmm();
    } else {
      for (int i = 0; i < 2; i++) { cudaFree(vw[i]); cudaFree(q[i]); }
      cudaFree(fl); cudaFree(p); cudaFree(tp); cudaFree(div);
    }
    for (int i = 0; i < 2; i++) { cudaFree(th[i]); cudaFree(qvp[i]); }
    cudaFree(s_mc); cudaFree(vort); cudaFree(vw_mc);
    cudaFree(sel_mc_keys);
    cudaFree(d_sel_keys[0]); cudaFree(d_sel_keys[1]); cudaFree(d_sel_tmp); cudaFree(d_sel_sc); cudaFree(d_sel_cta); cudaFree(d_sel_mc); cudaFree(d_sel_dc);
    if (h_sel_sc) cudaFreeHost(h_sel_sc);
    cudaFree(scratch); cudaFree(d_partial);
    if (h_partial) cudaFreeHost(h_partial);
    if (h_sub_scores) cudaFreeHost(h_sub_scores);
    if (h_block_scores) cudaFreeHost(h_block_scores);
    if (h_to_move) cudaFreeHost(h_to_move);
    if (h_dest) cudaFreeHost(h_dest);
    base_teardown();
  }
  void drop_graphs() {
    for (auto &g : step_graph) {
      if (g) cudaGraphExecDestroy(g);
      g = nullptr;
    }
  }
  void invalidate_graphs() override { drop_graphs(); }

  // ---- construction: fluid_simulation_dcgrid.cu:9-140 ---------------------------------------
  int construct(const dcg_sim_params *prm, int dev) override {
    DCG_TRY(base_setup(prm, dev));
    project_coarsest_pairs = 5; project_level_pairs = 5; local_pairs = 10;  // :274,283,301
    gx = prm->gx; gy = prm->gy; gz = prm->gz;
    if (gx <= 0 || gy <= 0 || gz <= 0) return fail(DCG_ERR_INVALID, "grid size must be positive");
    if (M64 == 0 || M64 * kBV >= 0xFFFFFFFFull) return fail(DCG_ERR_INVALID, "maxNumBlocks must be in [1, 2^26): cell ids are 32-bit");
    M = (uint32_t)M64;
    const int min_dim = (gx < gy && gx < gz) ? gx : (gy < gz ? gy : gz);
    int cell = 2;
    levels = 1;
    while (gx % cell == 0 && gy % cell == 0 && gz % cell == 0 && cell * kBW <= min_dim) {  // :12-22
      levels++;
      cell *= 2;
    }
    if (levels > kMaxLevels) return fail(DCG_ERR_UNSUPPORTED, "more than %d levels", kMaxLevels);
    max_blocks.assign(levels, 0); full_blocks.assign(levels, 0); loads.assign(levels, 0);
    offsets.assign(levels, 0); move_limit.assign(levels, 0);
    for (int l = 0, cs = 1; l < levels; l++, cs *= 2)  // :29-32
      full_blocks[l] = (uint64_t)idiv_up(gx, cs * kBW) * idiv_up(gy, cs * kBW) * idiv_up(gz, cs * kBW);
    if (M64 < full_blocks[levels - 1])  // :35-39 (the reference exit(1)s)
      return fail(DCG_ERR_POOL, "Too few blocks to fill lowest resolution (%llu / %llu)", (unsigned long long)M64,
                  (unsigned long long)full_blocks[levels - 1]);
    max_blocks[levels - 1] = full_blocks[levels - 1];
    uint64_t left = M64 - max_blocks[levels - 1];
    for (int l = levels - 2; l >= 0; l--) {  // :44-48
      max_blocks[l] = std::min(left / (uint64_t)(l + 1), full_blocks[l]);
      left -= max_blocks[l];
    }
    if (max_blocks[0] == 0) return fail(DCG_ERR_POOL, "Too few blocks to reach highest resolution");  // :50-53
    for (int l = 1; l < levels; l++) offsets[l] = offsets[l - 1] + max_blocks[l - 1];
    sparse = 0;
    for (int l = 0; l < levels; l++)
      if (max_blocks[l] < full_blocks[l]) sparse = l + 1;  // :66-69
    cells = (size_t)M * kBV;

    T.M = M; T.levels = levels; T.sparse_levels = sparse;
    for (int l = 0; l < levels; l++) { T.offsets[l] = (uint32_t)offsets[l]; T.max_blocks[l] = (uint32_t)max_blocks[l]; }
    DCG_CUDA_TRY(cudaMalloc(&T.posl, (size_t)M * sizeof(int4)));
    DCG_CUDA_TRY(cudaMalloc(&T.parent, ((size_t)M + kB4) * 4));  // padded: the stencil rings stream kB4 entries per tile
    experiment = opt.experiment;
    DCG_TRY(setup_sharding());
    DCG_CUDA_TRY(cudaMalloc(&d_perm, (size_t)M * 4));
    use_pdl = !opt.no_pdl;
    experiment = opt.experiment;
    use_resort = !opt.no_resort;
    if (opt.resort_every != 0) resort_every = std::max(0, opt.resort_every);  // -1: only at the fixed point
    DCG_CUDA_TRY(cudaMalloc(&d_pcount, (kMaxLevels + 2) * 4));
    DCG_CUDA_TRY(cudaMallocHost(&h_pcount, (kMaxLevels + 2) * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_order_keys[0], (size_t)M * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_order_keys[1], (size_t)M * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_order_vals, (size_t)M * 4));
    DCG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp_bytes, d_order_keys[0], d_order_keys[1], d_order_vals, work[0].d_order, (int)M, 0,
                                                 32, stream));
    DCG_CUDA_TRY(cudaMalloc(&d_sort_tmp, sort_tmp_bytes + 16));
    DCG_CUDA_TRY(cudaMalloc(&T.child, (size_t)M * 8 * 4));
    DCG_CUDA_TRY(cudaMalloc(&T.apron, (size_t)M * kAV * 4));
    DCG_CUDA_TRY(cudaMalloc(&T.face, (size_t)M * 96 * 4));
    DCG_CUDA_TRY(cudaMalloc(&T.fd, (size_t)M * 12 * 4));
    map_size.assign(levels, 0);
    for (int l = 0; l < sparse; l++) {
      map_size[l] = full_blocks[l];
      DCG_CUDA_TRY(cudaMalloc(&T.map[l], map_size[l] * 4));
    }
    if (M64 > kFmapSlotMask) return fail(DCG_ERR_UNSUPPORTED, "max_num_blocks above 2^28 - 1 (finest-block map packs level and slot into 32 bits)");
    fmap_size = (size_t)idiv_up(gx, kBW) * idiv_up(gy, kBW) * idiv_up(gz, kBW);
    DCG_CUDA_TRY(cudaMalloc(&T.fmap, fmap_size * 4));
    Tf.fmap = T.fmap;  // one table, in the numbering of whichever pool the field kernels see (hot())
    DCG_CUDA_TRY(cudaMalloc(&d_flags, (size_t)M * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_free, (size_t)M * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_touched, (size_t)M * 4 * 2));
    DCG_CUDA_TRY(cudaMalloc(&d_to_move, (size_t)M * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_dest, (size_t)M * 8 * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_flag_bits, ((size_t)M / 32 + 2) * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_counters, 2 * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_summary, sizeof(ScoreSummary)));
    DCG_CUDA_TRY(cudaMallocHost(&h_summary, sizeof(ScoreSummary)));
    DCG_CUDA_TRY(cudaMallocHost(&h_summary_init, sizeof(ScoreSummary)));
    for (int l = 0; l < kMaxLevels; l++) { h_summary_init->max_ss[l] = -1; h_summary_init->min_bs[l] = 0xFFFFFFFFu; h_summary_init->n_refine[l] = 0; }
    DCG_CUDA_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    DCG_CUDA_TRY(cudaEventCreateWithFlags(&ev_scores, cudaEventDisableTiming));
    DCG_CUDA_TRY(cudaEventCreateWithFlags(&ev_scores_copied, cudaEventDisableTiming));
    use_early_scores = !(experiment & 32);
    DCG_CUDA_TRY(cudaMalloc(&d_new_posl, (size_t)M * sizeof(int4)));
    DCG_CUDA_TRY(cudaMalloc(&d_sub_scores, ((size_t)M * 8 + 1) * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_block_scores, (size_t)M * 4));
    if (!vmm) {
      for (int i = 0; i < 2; i++) {
        DCG_CUDA_TRY(cudaMalloc(&vw[i], cells * sizeof(float4)));
        DCG_CUDA_TRY(cudaMalloc(&q[i], cells * 4));
      }
      DCG_CUDA_TRY(cudaMalloc(&fl, cells * 4));
      DCG_CUDA_TRY(cudaMalloc(&p, cells * 4));
      DCG_CUDA_TRY(cudaMalloc(&tp, cells * 4));
      DCG_CUDA_TRY(cudaMalloc(&div, cells * 4));
    } else {
      DCG_TRY(create_arena());
    }
    scratch_floats = 3 * cells;  // pool-native accessors; the dense level-0 resampling grows it on demand
    DCG_CUDA_TRY(cudaMalloc(&scratch, scratch_floats * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_partial, 1024 * sizeof(double)));
    DCG_CUDA_TRY(cudaMallocHost(&h_partial, 1024 * sizeof(double)));
    DCG_CUDA_TRY(cudaMallocHost(&h_sub_scores, ((size_t)M * 8 + 1) * 4));
    DCG_CUDA_TRY(cudaMallocHost(&h_block_scores, (size_t)M * 4));
    DCG_CUDA_TRY(cudaMallocHost(&h_to_move, (size_t)M * 4));
    DCG_CUDA_TRY(cudaMallocHost(&h_dest, (size_t)M * 8 * 4));
    {
      DCG_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
      DCG_CUDA_TRY(cudaFuncSetAttribute(k_dc_jacobi_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJacobiPipeSmem));
      int per_sm = 0;
      DCG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dc_jacobi_pipe, kCTA4, kJacobiPipeSmem));
      if (per_sm < 1) return fail(DCG_ERR_CUDA, "k_dc_jacobi_pipe does not fit on an SM");
      jacobi_pipe_ctas = per_sm * sm_count;
      use_advect_pipe = opt.advect != 1;
      order_mode = opt.advect_slot_order ? 0 : 1;
      use_stencil_pipe = prolong_staged = opt.stencil != 1;
      {
        int per_sm = 0;
        DCG_CUDA_TRY(cudaFuncSetAttribute(k_dc_divergence_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDivPipeSmem));
        DCG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dc_divergence_pipe, kStencilThreads, kDivPipeSmem));
        if (per_sm < 1) return fail(DCG_ERR_CUDA, "k_dc_divergence_pipe does not fit on an SM");
        if (opt.stencil_ctas_per_sm > 0) per_sm = std::max(1, std::min(per_sm, opt.stencil_ctas_per_sm));
        div_pipe_ctas = per_sm * sm_count;
        apply_min_blocks = opt.apply_min_blocks == 3 ? 3 : 2;
        const void *afn = apply_min_blocks == 2 ? (const void *)k_dc_apply_pipe<2> : (const void *)k_dc_apply_pipe<3>;
        DCG_CUDA_TRY(cudaFuncSetAttribute(afn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kApplyPipeSmem));
        DCG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, afn, kStencilThreads, kApplyPipeSmem));
        if (per_sm < 1) return fail(DCG_ERR_CUDA, "k_dc_apply_pipe does not fit on an SM");
        if (opt.stencil_ctas_per_sm > 0) per_sm = std::max(1, std::min(per_sm, opt.stencil_ctas_per_sm));
        apply_pipe_ctas = per_sm * sm_count;
      }
      coarse_in_smem = !opt.coarse_in_gmem;
      DCG_CUDA_TRY(cudaFuncSetAttribute(k_dc_coarse_cascade<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoarseSmemMax));
      skip_dead_zeroing = !opt.zero_all;
      fuse_advect = !opt.advect_no_fuse;
      advect_min_blocks = (opt.advect_min_blocks == 4 || opt.advect_min_blocks == 2) ? opt.advect_min_blocks : 3;
      for (int mode = 0; mode < 3; mode++) {
        const void *fn = advect_fn(mode);
        DCG_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAdvectPipeSmem));
        DCG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&advect_per_sm[mode], fn, kAdvectThreads, kAdvectPipeSmem));
        if (advect_per_sm[mode] < 1) return fail(DCG_ERR_CUDA, "k_dc_advect_pipe does not fit on an SM");
        if (opt.advect_ctas_per_sm > 0) advect_per_sm[mode] = opt.advect_ctas_per_sm;
        if (mode >= 1) {
          const void *fx = advect_fn(mode, true);
          DCG_CUDA_TRY(cudaFuncSetAttribute(fx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAdvectPipeSmem));
          DCG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&advect_ext_per_sm[mode], fx, kAdvectThreads, kAdvectPipeSmem));
          if (advect_ext_per_sm[mode] < 1) return fail(DCG_ERR_CUDA, "k_dc_advect_pipe (with scalars) does not fit on an SM");
          if (opt.advect_ctas_per_sm > 0) advect_ext_per_sm[mode] = opt.advect_ctas_per_sm;
        }
      }
      pipe_min_tiles = 2u * (unsigned)sm_count;  // refined below once the resident CTA count of the ring kernel is known
      {
        DCG_CUDA_TRY(cudaFuncSetAttribute(k_dc_jacobi_pipe8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJacobiPipeSmem));
        int per8 = 0;
        DCG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per8, k_dc_jacobi_pipe8, kJ8Threads, kJacobiPipeSmem));
        if (per8 < 1) return fail(DCG_ERR_CUDA, "k_dc_jacobi_pipe8 does not fit on an SM");
        jacobi8_ctas = per8 * sm_count;
        // a resident CTA should walk >= 4 tiles to amortise its ring set-up: at 512^3 level 2 (2,048 tiles) takes
        // 8.3 us with one CTA per tile and 10.3 us with the ring
        pipe_min_tiles = 4u * (unsigned)jacobi8_ctas;
      }
      use_pipe = opt.jacobi != 1;
      if (opt.jacobi == 2 || opt.jacobi == 4) pipe_min_tiles = 1;  // tests: exercise the ring on small levels too
      if (opt.jacobi == 3 || opt.jacobi == 4) jacobi8 = false;
      snake = !opt.no_snake;
      if (opt.jacobi_ctas_per_sm > 0) jacobi_pipe_ctas = jacobi8_ctas = opt.jacobi_ctas_per_sm * sm_count;
      if (opt.jacobi_max_ctas > 0) {
        jacobi_pipe_ctas = std::min(jacobi_pipe_ctas, opt.jacobi_max_ctas);
        jacobi8_ctas = std::min(jacobi8_ctas, opt.jacobi_max_ctas);
      }
      if (world > 1 && (!use_advect_pipe || !use_stencil_pipe || !use_pipe)) return fail(DCG_ERR_UNSUPPORTED, "the legacy kernel variants are single-GPU only");
    }
    if (vmm) return DCG_OK;  // the caller exchanges handles; import_handles() maps the peers and resets
    return reset();
  }

  // ---- sharding: ownership, arenas, barriers ------------------------------------------------------------------
  int level_of_slot(uint64_t b) const {
    for (int l = 0; l < levels; l++)
      if (b >= offsets[l] && b < offsets[l] + max_blocks[l]) return l;
    return -1;
  }
  int setup_sharding() {
    if (world < 1 || world > 8 || nlocal < 1 || rank0 < 0 || rank0 + nlocal > world) return fail(DCG_ERR_INVALID, "bad rank / world");
    if (nlocal != 1 && nlocal != world) return fail(DCG_ERR_INVALID, "nlocal must be 1 (one rank per process) or world (all ranks in one process)");
    vmm = world > 1 && nlocal == 1;
    // ownership granularity: 2 MiB of the 4-byte fields in the stitched address space, one tile otherwise
    unit = vmm ? 8192u : (uint32_t)kTile;
    if (opt.shard_unit) {
      const uint32_t u = opt.shard_unit;
      if (u % kTile || (vmm && u % 8192u)) return fail(DCG_ERR_INVALID, "dcg_options.shard_unit must be a multiple of %u", vmm ? 8192u : (uint32_t)kTile);
      unit = u;
    }
    nunits = (uint32_t)((M64 + unit - 1) / unit);
    unit_owner.assign(nunits, 0);
    for (uint32_t u = 0; u < nunits; u++) {
      const uint64_t mid = std::min<uint64_t>((uint64_t)u * unit + unit / 2, M64 - 1);
      const int l = level_of_slot(mid);
      int r = 0;
      if (l >= 0 && world > 1) r = (int)std::min<uint64_t>(world - 1, (mid - offsets[l]) * (uint64_t)world / max_blocks[l]);
      unit_owner[u] = (uint8_t)r;
    }
    DCG_CUDA_TRY(cudaMalloc(&d_unit_owner, nunits));
    DCG_CUDA_TRY(cudaMemcpyAsync(d_unit_owner, unit_owner.data(), nunits, cudaMemcpyHostToDevice, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    DCG_CUDA_TRY(cudaMalloc(&d_texpect, (size_t)M));
    DCG_CUDA_TRY(cudaMalloc(&d_tcount, (size_t)M * 4));
    DCG_CUDA_TRY(cudaMemsetAsync(d_tcount, 0, (size_t)M * 4, stream));
    // one GPU: per-level list passes (2.50 against 2.54 ms per step at dcgrid512: the walk's chain of dependent rounds is longer than
    // four short launches); sharded: the tree walk (one barrier per pass instead of one per level: 2.88 against 2.94 ms on 2 B200)
    use_tree = (world > 1) != ((experiment & 8) != 0);
    work.resize(nlocal);
    for (auto &w : work) {
      DCG_CUDA_TRY(cudaMalloc(&w.d_order, ((size_t)M + kBPC) * 4));
      DCG_CUDA_TRY(cudaMalloc(&w.d_plist, (size_t)M * 4));
      DCG_CUDA_TRY(cudaMalloc(&w.d_tstarts, (size_t)M * 4));
      w.level.assign(levels, TileRuns{});
    }
    return DCG_OK;
  }
  // runs of tiles [t0, t1) (tile t = slots base + 16 t ...) whose first slot belongs to a unit of `rank`
  mutable bool runs_overflow = false;
  TileRuns tile_runs(uint64_t base, uint32_t t0, uint32_t t1, int rank) const {
    TileRuns R{};
    R.n = 0;
    R.pre[0] = 0;
    uint32_t t = t0;
    while (t < t1) {
      const uint32_t u = (uint32_t)((base + (uint64_t)t * kTile) / unit);
      // tiles up to the end of this unit share its owner
      const uint64_t unit_end = (uint64_t)(u + 1) * unit;
      uint32_t te = (uint32_t)std::min<uint64_t>(t1, (unit_end - base + kTile - 1) / kTile);
      if (te <= t) te = t + 1;
      if (unit_owner[std::min(u, nunits - 1)] == rank) {
        if (R.n > 0 && R.first[R.n - 1] + (R.pre[R.n] - R.pre[R.n - 1]) == t) {
          R.pre[R.n] += te - t;  // extends the previous run
        } else if (R.n < kMaxRuns) {
          R.first[R.n] = t;
          R.pre[R.n + 1] = R.pre[R.n] + (te - t);
          R.n++;
        } else {
          runs_overflow = true;  // rebuild_order() turns this into an error: a dropped run would never be processed
        }
      }
      t = te;
    }
    if (R.n == 0) { R.n = 1; R.first[0] = 0; R.pre[1] = 0; }
    return R;
  }
  template <class F>
  void each_rank(F f) {
    for (int lr = 0; lr < nlocal; lr++) f(rank0 + lr, work[lr]);
  }
  bool has_rank0() const { return rank0 == 0; }

  // Launch of a step kernel.  use_pdl: programmatic dependent launch — the kernel may become resident before its
  // predecessor in the stream has drained; every such kernel starts with pdl_enter() (common.cuh), which restores
  // the full data dependence.  Captured into the step graph as programmatic edges.
  bool use_pdl = true;
  uint64_t pdl_fallbacks = 0;
  template <class... P, class... A>
  void launch_pdl(void (*k)(P...), dim3 grid, dim3 block, size_t smem, A &&...a) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = use_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k, P(std::forward<A>(a))...);
  }
  BarrierPeers peers{};
  // lock-step barrier over peer memory after a phase whose results other ranks read (no-op inside one process:
  // stream order is the barrier)
  void barrier() {
    if (!vmm) return;
    launch_pdl(k_dcs_barrier, dim3(1), dim3(32), 0, peers, rank0, world, d_epoch, d_barrier_err);
    launches++;
    n_barriers++;
  }

  int create_arena() {
    if (!drv.load()) return fail(DCG_ERR_CUDA, "CUDA virtual memory management entry points are not available");
    CUmemAllocationProp prop = vmm::device_prop(device);
    if (drv.memGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return fail(DCG_ERR_CUDA, "cuMemGetAllocationGranularity failed");
    if (((size_t)unit * kBV * 4) % gran) return fail(DCG_ERR_UNSUPPORTED, "allocation granularity %zu does not divide a unit", gran);
    unit_runs.clear();
    for (uint32_t u = 0; u < nunits;) {
      uint32_t e = u + 1;
      while (e < nunits && unit_owner[e] == unit_owner[u]) e++;
      unit_runs.push_back({u, e, (int)unit_owner[u]});
      u = e;
    }
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
    // exported descriptors are closed on every early exit (FdServer::start takes them over, also when it fails)
    auto prepare = [&]() -> int {
      DCG_TRY(create(gran));  // control block
      for (const UnitRun &r : unit_runs)
        if (r.owner == rank0)
          for (int f = 0; f < kFields; f++) DCG_TRY(create((size_t)(r.u1 - r.u0) * unit * kBV * (f < 2 ? 16 : 4)));
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
  CUmemAccessDesc access_desc() const {
    CUmemAccessDesc acc;
    std::memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    return acc;
  }
  int export_handle(void *out, uint64_t cap) override {
    if (!vmm) return fail(DCG_ERR_INVALID, "not a one-rank-per-process instance");
    if (cap < vmm::kHandleBytes) return fail(DCG_ERR_INVALID, "handle buffer too small");
    std::memcpy(out, fd_server.name, vmm::kHandleBytes);
    return DCG_OK;
  }
  // maps one whole physical piece (cuMemMap takes neither offsets nor partial sizes) and opens it to this device
  int map_piece(CUdeviceptr va, size_t bytes, CUmemGenericAllocationHandle h) {
    if (drv.memMap(va, bytes, 0, h, 0) != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemMap failed (%zu bytes)", bytes);
    CUmemAccessDesc acc = access_desc();
    if (drv.memSetAccess(va, bytes, &acc, 1) != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemSetAccess failed: peer access between the GPUs is required");
    return DCG_OK;
  }
  int import_handles(const void *handles, int count) override {
    if (!vmm) return fail(DCG_ERR_INVALID, "not a one-rank-per-process instance");
    if (count != world) return fail(DCG_ERR_INVALID, "expected %d handles, got %d", world, count);
    DCG_CUDA_TRY(cudaSetDevice(device));
    std::vector<int> runs_of(world, 0);
    for (const UnitRun &r : unit_runs) runs_of[r.owner]++;
    for (int r = 0; r < world; r++) {
      if (r == rank0) continue;
      char name[vmm::kHandleBytes + 1] = {0};
      std::memcpy(name, static_cast<const char *>(handles) + (size_t)r * vmm::kHandleBytes, vmm::kHandleBytes);
      std::vector<int> fds;
      if (!vmm::fetch_fds(name, 1 + runs_of[r] * kFields, fds)) return fail(DCG_ERR_CUDA, "could not fetch the memory descriptors of rank %d", r);
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
    // control blocks: rank r's at ctrl_va + r * gran (this rank's own was mapped at creation)
    for (int r = 0; r < world; r++) {
      if (r != rank0) DCG_TRY(map_piece(ctrl_va + (size_t)r * gran, gran, pieces[r][0]));
      peers.flags[r] = reinterpret_cast<volatile uint32_t *>(ctrl_va + (size_t)r * gran);
    }
    // fields: the k-th run of rank r, field f = pieces[r][1 + k * kFields + f], mapped at the run's place
    for (int f = 0; f < kFields; f++) {
      const size_t ub = (size_t)unit * kBV * (f < 2 ? 16 : 4);
      field_va_bytes[f] = (size_t)nunits * ub;
      if (drv.memReserve(&field_va[f], field_va_bytes[f], gran, 0, 0) != CUDA_SUCCESS) return fail(DCG_ERR_CUDA, "cuMemAddressReserve failed");
      std::vector<int> k_of(world, 0);
      for (const UnitRun &r : unit_runs) {
        const int k = k_of[r.owner]++;
        DCG_TRY(map_piece(field_va[f] + (size_t)r.u0 * ub, (size_t)(r.u1 - r.u0) * ub, pieces[r.owner][1 + (size_t)k * kFields + f]));
      }
    }
    vw[0] = reinterpret_cast<float4 *>(field_va[0]); vw[1] = reinterpret_cast<float4 *>(field_va[1]);
    q[0] = reinterpret_cast<float *>(field_va[2]); q[1] = reinterpret_cast<float *>(field_va[3]);
    fl = reinterpret_cast<float *>(field_va[4]); p = reinterpret_cast<float *>(field_va[5]);
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
  int need_ready() { return ready ? DCG_OK : fail(DCG_ERR_INVALID, "sharded instance not finalized: call dcg_shard_import_handles first"); }
  // every public entry point: right device, and (one rank per process) peers mapped — before that the field
  // pointers, the barrier flags and the epoch counter are null
  int enter() {
    DCG_CUDA_TRY(cudaSetDevice(device));
    return need_ready();
  }
  int check_barrier_error() {
    if (!vmm || !d_barrier_err) return DCG_OK;
    uint32_t e = 0;
    DCG_CUDA_TRY(cudaMemcpyAsync(&e, d_barrier_err, 4, cudaMemcpyDeviceToHost, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    if (e) return fail(DCG_ERR_CUDA, "shard barrier timed out: a peer rank never arrived (ranks must issue identical call sequences)");
    return DCG_OK;
  }
  int health_check() override { return check_barrier_error(); }  // every dcg_synchronize / accessor: a desynchronised peer is an error, not silently wrong fields

  int on_params_changed() override {
    if (params.gx != gx || params.gy != gy || params.gz != gz)
      return fail(DCG_ERR_INVALID, "grid size is fixed at construction (the reference sizes its pool in the ctor)");
    steady = false;  // scores depend on SimParams: the fixed-point proof no longer holds
    early_ready = false;
    spec_velocity = false;
    drop_graphs();
    return DCG_OK;
  }

  // ---- extensions ------------------------------------------------------------------------------------
  int on_ext_changed() override {
    if (ext_on() && world > 1) return fail(DCG_ERR_UNSUPPORTED, "the extensions are single-GPU only");
    DCG_CUDA_TRY(cudaSetDevice(device));
    DCG_TRY(ensure_ext_storage());
    steady = false;
    early_ready = false;
    spec_velocity = false;
    drop_graphs();
    return DCG_OK;
  }
  int ensure_ext_storage() {
    if (ext.sources && !th[0]) {
      for (int i = 0; i < 2; i++) {
        DCG_CUDA_TRY(cudaMalloc(&th[i], cells * 4));
        DCG_CUDA_TRY(cudaMalloc(&qvp[i], cells * 4));
        DCG_CUDA_TRY(cudaMemsetAsync(th[i], 0, cells * 4, stream));
        DCG_CUDA_TRY(cudaMemsetAsync(qvp[i], 0, cells * 4, stream));
      }
    }
    if ((ext.sources || ext.score_mode) && !vort) {
      DCG_CUDA_TRY(cudaMalloc(&vort, cells * sizeof(float4)));
      DCG_CUDA_TRY(cudaMemsetAsync(vort, 0, cells * sizeof(float4), stream));
    }
    if (ext.advection && !vw_mc) {
      DCG_CUDA_TRY(cudaMalloc(&vw_mc, cells * sizeof(float4)));
      DCG_CUDA_TRY(cudaMalloc(&s_mc, cells * 4));
      DCG_CUDA_TRY(cudaMemsetAsync(vw_mc, 0, cells * sizeof(float4), stream));
      DCG_CUDA_TRY(cudaMemsetAsync(s_mc, 0, cells * 4, stream));
    }
    return DCG_OK;
  }
  int ensure_sel_storage() {
    if (d_sel_keys[0]) return DCG_OK;
    sel_max_n = max_blocks[0];
    for (int l = 1; l < levels; l++) sel_max_n = std::max<size_t>(sel_max_n, 8 * max_blocks[l]);
    for (int i = 0; i < 2; i++) DCG_CUDA_TRY(cudaMalloc(&d_sel_keys[i], sel_max_n * 8));
    DCG_CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, sel_tmp_bytes, d_sel_keys[0], d_sel_keys[1], (int)sel_max_n, 0, 64, stream));
    DCG_CUDA_TRY(cudaMalloc(&d_sel_tmp, sel_tmp_bytes + 16));
    DCG_CUDA_TRY(cudaMalloc(&d_sel_sc, 4 * kMaxLevels * 4));
    DCG_CUDA_TRY(cudaMallocHost(&h_sel_sc, 4 * kMaxLevels * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_sel_cta, (sel_max_n / sel::kSelCta + 2) * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_sel_mc, (size_t)M * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_sel_dc, (size_t)M * 4));
    return DCG_OK;
  }
  void launch_vorticity() {
    ext::k_dc_ext_vorticity<<<blocks_for(M, kBPC), kCTA, 0, stream>>>(hot(), kp, vw[cur_v], vort);
    launches++;
  }
  // the fused source pass + restriction of what it touched (oracle apply_sources)
  int apply_sources() override {
    DCG_TRY(enter());
    if (!ext.sources) return DCG_OK;
    spec_velocity = false;
    launch_vorticity();
    ext::k_dc_ext_sources<<<blocks_for(M, kBPC), kCTA, 0, stream>>>(hot(), kp, ext, vw[cur_v], q[cur_q], th[cur_s], qvp[cur_s], vort);
    launches++;
    accumulate(vw[cur_v], q[cur_q], true);  // the kernel restricted the childless blocks: only the blocks with children are left
    accumulate(nullptr, th[cur_s], true);
    accumulate(nullptr, qvp[cur_s], true);
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  // semi-Lagrangian gather + restriction of the velocity / of one scalar (kind 0 density, 1 temperature, 2 vapor)
  void sl_velocity(const float4 *in, float4 *out) {
    if (use_advect_pipe) {
      launch_advect_pipe(0, in, out, nullptr, nullptr);
      accumulate(out, nullptr, true);
    } else {
      k_dc_advect_velocity<<<blocks_for(M, kBPC), kCTA, 0, stream>>>(hot(), kp, in, out);
      launches++;
      accumulate(out, nullptr, false);
    }
  }
  void sl_scalar(int kind, const float *in, float *out) {
    const unsigned gb = blocks_for(M, kBPC);
    if (kind == 0 && use_advect_pipe) {
      launch_advect_pipe(1, vw[cur_v], nullptr, in, out);
      accumulate(nullptr, out, true);
      return;
    }
    if (kind == 0) k_dc_advect_density<<<gb, kCTA, 0, stream>>>(hot(), kp, vw[cur_v], fl, in, out);
    else if (kind == 1) ext::k_dc_ext_advect_scalar<1, false><<<gb, kCTA, 0, stream>>>(hot(), kp, ext, vw[cur_v], fl, in, nullptr, out);
    else ext::k_dc_ext_advect_scalar<2, false><<<gb, kCTA, 0, stream>>>(hot(), kp, ext, vw[cur_v], fl, in, nullptr, out);
    launches++;
    accumulate(nullptr, out, false);
  }
  // MacCormack (oracle maccormack()): `cur` holds phi, `other` receives hat, the result lands in s_mc and the
  // buffers rotate so that `cur` holds the result
  void maccormack_scalar(int kind, float *&cur, float *other) {
    sl_scalar(kind, cur, other);
    const unsigned gb = blocks_for(M, kBPC);
    if (kind == 0) ext::k_dc_ext_advect_scalar<0, true><<<gb, kCTA, 0, stream>>>(hot(), kp, ext, vw[cur_v], fl, cur, other, s_mc);
    else if (kind == 1) ext::k_dc_ext_advect_scalar<1, true><<<gb, kCTA, 0, stream>>>(hot(), kp, ext, vw[cur_v], fl, cur, other, s_mc);
    else ext::k_dc_ext_advect_scalar<2, true><<<gb, kCTA, 0, stream>>>(hot(), kp, ext, vw[cur_v], fl, cur, other, s_mc);
    launches++;
    std::swap(cur, s_mc);
    accumulate(nullptr, cur, false);
  }

  // ---- reset / init: fluid_simulation_dcgrid.cu:190-261 -----------------------------------------
  int reset() override {
    DCG_TRY(enter());
    drop_graphs();
    steady = false;
    early_ready = false;
    spec_velocity = false;
    k_fill_posl<<<blocks_for(M, 256), 256, 0, stream>>>(T.posl, M);
    DCG_CUDA_TRY(cudaMemsetAsync(d_flags, 0, (size_t)M * 4, stream));
    DCG_CUDA_TRY(cudaMemsetAsync(T.parent, 0xff, ((size_t)M + kB4) * 4, stream));
    DCG_CUDA_TRY(cudaMemsetAsync(T.child, 0xff, (size_t)M * 8 * 4, stream));
    DCG_CUDA_TRY(cudaMemsetAsync(T.face, 0, (size_t)M * 96 * 4, stream));
    DCG_CUDA_TRY(cudaMemsetAsync(T.fd, 0, (size_t)M * 12 * 4, stream));
    for (int l = 0; l < sparse; l++) DCG_CUDA_TRY(cudaMemsetAsync(T.map[l], 0xff, map_size[l] * 4, stream));
    if (!vmm) {
      for (int i = 0; i < 2; i++) {
        DCG_CUDA_TRY(cudaMemsetAsync(vw[i], 0, cells * sizeof(float4), stream));
        DCG_CUDA_TRY(cudaMemsetAsync(q[i], 0, cells * 4, stream));
      }
      DCG_CUDA_TRY(cudaMemsetAsync(fl, 0, cells * 4, stream));
      DCG_CUDA_TRY(cudaMemsetAsync(p, 0, cells * 4, stream));
      DCG_CUDA_TRY(cudaMemsetAsync(tp, 0, cells * 4, stream));
      DCG_CUDA_TRY(cudaMemsetAsync(div, 0, cells * 4, stream));
    } else {
      // every rank clears its own arena (the fields of the units it owns), then all meet: nobody may start
      // writing initial values into cells a peer is still clearing
      DCG_TRY(need_ready());
      barrier();
      for (int f = 0; f < kFields; f++) {
        const size_t ub = (size_t)unit * kBV * (f < 2 ? 16 : 4);
        uint32_t u = 0;
        while (u < nunits) {
          uint32_t e = u + 1;
          while (e < nunits && unit_owner[e] == unit_owner[u]) e++;
          if (unit_owner[u] == rank0)
            DCG_CUDA_TRY(cudaMemsetAsync(reinterpret_cast<void *>(field_va[f] + (size_t)u * ub), 0, (size_t)(e - u) * ub, stream));
          u = e;
        }
      }
      barrier();
    }
    DCG_TRY(ensure_ext_storage());
    for (int i = 0; i < 2; i++) {
      if (th[i]) DCG_CUDA_TRY(cudaMemsetAsync(th[i], 0, cells * 4, stream));
      if (qvp[i]) DCG_CUDA_TRY(cudaMemsetAsync(qvp[i], 0, cells * 4, stream));
    }
    if (vort) DCG_CUDA_TRY(cudaMemsetAsync(vort, 0, cells * sizeof(float4), stream));
    cur_s = 0;
    DCG_CUDA_TRY(cudaMemsetAsync(d_counters, 0, 2 * 4, stream));
    k_fill_u32<<<blocks_for((size_t)M * 8 + 1, 256), 256, 0, stream>>>(reinterpret_cast<uint32_t *>(d_sub_scores), 0xFF7FFFFFu /* -FLT_MAX */,
                                                                       (size_t)M * 8 + 1);
    k_iota_u32<<<blocks_for(M, 256), 256, 0, stream>>>(d_free, M);  // freeBlockIndices[i] = i, :243-249
    k_iota_u32<<<blocks_for(M, 256), 256, 0, stream>>>(d_perm, M);  // field order = slot order until the first re-sort
    mirrored = false;
    changes_since_resort = 0;
    launches += 4;
    cur_v = cur_q = 0;
    for (int l = 0; l < levels; l++) {  // :232-241
      loads[l] = (max_blocks[l] == full_blocks[l]) ? max_blocks[l] : 0;
      move_limit[l] = 0;
    }
    sync_loads();
    return init();
  }

  int init() override {  // :190-210
    DCG_TRY(enter());
    k_dc_init_apron<<<blocks_for((size_t)M * kAV, 256), 256, 0, stream>>>(T);
    launches++;
    for (int l = sparse; l < levels; l++) {
      k_dc_activate_level<<<(unsigned)full_blocks[l], 64, 0, stream>>>(T, kp, l, vw[0], vw[1], q[0], q[1], fl);
      launches++;  // (sharded: every process writes the same values into every cell)
    }
    barrier();
    if (ext.sources) {  // extension: temperature / vapor start from the ambient profile
      ext::k_dc_ext_init_scalars<<<blocks_for(M, kBPC), kCTA, 0, stream>>>(T, kp, ext, th[0], th[1], qvp[0], qvp[1]);
      launches++;
    }
    DCG_TRY(build_face_descriptors());
    DCG_CUDA_TRY(cudaGetLastError());
    for (int i = 0; i < 5; i++) DCG_TRY(adapt_topology());
    return DCG_OK;
  }

  // ---- adaptation: fluid_simulation_dcgrid.cu:320-483 --------------------------------------------
  void sync_loads() {
    for (int l = 0; l < levels; l++) T.loads[l] = Tf.loads[l] = (uint32_t)loads[l];
  }
  int build_face_descriptors() {
    cudaMemsetAsync(d_counters + 1, 0, 4, stream);
    k_dc_build_fdesc<<<blocks_for((size_t)M * 8, 256), 256, 0, stream>>>(hot(), d_counters + 1);
    k_dc_build_fmap<<<blocks_for(fmap_size, 256), 256, 0, stream>>>(hot(), kp);
    launches += 2;
    return rebuild_order();
  }
  // ---- field order (dcgrid_kernels.cuh, "field order") ----------------------------------------------------
  int ensure_mirror_storage() {
    if (Tf.posl) return DCG_OK;
    DCG_CUDA_TRY(cudaMalloc(&Tf.posl, (size_t)M * sizeof(int4)));
    DCG_CUDA_TRY(cudaMalloc(&Tf.parent, ((size_t)M + kB4) * 4));
    DCG_CUDA_TRY(cudaMalloc(&Tf.child, (size_t)M * 8 * 4));
    DCG_CUDA_TRY(cudaMalloc(&Tf.apron, (size_t)M * kAV * 4));
    for (int l = 0; l < sparse; l++) DCG_CUDA_TRY(cudaMalloc(&Tf.map[l], map_size[l] * 4));
    if (sort_ordered)
      for (int l = sparse; l < levels; l++) DCG_CUDA_TRY(cudaMalloc(&Tf.map[l], full_blocks[l] * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_perm_new, (size_t)M * 4));
    DCG_CUDA_TRY(cudaMalloc(&d_sort_keys64[0], (size_t)M * 8));
    DCG_CUDA_TRY(cudaMalloc(&d_sort_keys64[1], (size_t)M * 8));
    DCG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp64_bytes, d_sort_keys64[0], d_sort_keys64[1], d_order_vals, d_perm_new, (int)M, 0, 64,
                                                 stream));
    DCG_CUDA_TRY(cudaMalloc(&d_sort_tmp64, sort_tmp64_bytes + 16));
    return DCG_OK;
  }
  // rebuilds the field-order mirror of the pool from the reference-order pool through d_perm
  // full = false: after a topology change with an unchanged permutation — k_dc_refresh_apron has already rewritten
  // the mirrored apron entries it changed, only the small per-block arrays and the level maps are redone
  void mirror(bool full = true) {
    if (!mirrored) return;
    Tf.M = T.M; Tf.levels = T.levels; Tf.sparse_levels = T.sparse_levels;
    for (int l = 0; l < kMaxLevels; l++) { Tf.offsets[l] = T.offsets[l]; Tf.max_blocks[l] = T.max_blocks[l]; Tf.loads[l] = T.loads[l]; }
    Tf.flags = T.flags; Tf.fd = T.fd; Tf.face = T.face;  // descriptors exist once: for whichever pool the field kernels see
    cudaMemsetAsync(Tf.parent + M, 0xff, kB4 * 4, stream);
    k_dc_mirror_blocks<<<blocks_for(M, 256), 256, 0, stream>>>(T, Tf, d_perm);
    if (full) k_dc_mirror_apron<<<blocks_for((size_t)M * kAV, 256), 256, 0, stream>>>(T, Tf, d_perm);
    for (int l = 0; l < sparse; l++) k_dc_mirror_map<<<blocks_for(map_size[l], 256), 256, 0, stream>>>(T.map[l], Tf.map[l], map_size[l], d_perm);
    launches += 2 + sparse;
    if (sort_ordered) {  // the mirror looks every level up through a dense map
      Tf.sparse_levels = levels;
      if (full) {
        for (int l = sparse; l < levels; l++) cudaMemsetAsync(Tf.map[l], 0xff, full_blocks[l] * 4, stream);
        ext::k_dc_ext_rebuild_maps<<<blocks_for(M, 256), 256, 0, stream>>>(Tf, kp);
        launches++;
      }
    }
  }
  // new permutation (active blocks of every sparse level sorted by position), fields moved into the new order
  int slab_axis = 0;
  int resort() {
    slab_axis = (gz > gx && gz >= gy) ? 2 : (gy > gx ? 1 : 0);  // longest axis; x on ties (the construction order of the ordered levels)
    sort_ordered = world > 1 && slab_axis != 0;
    DCG_TRY(ensure_mirror_storage());
    k_dc_resort_keys<<<blocks_for(M, 256), 256, 0, stream>>>(T, kp, world, slab_axis, sort_ordered ? 1 : 0, d_sort_keys64[0], d_order_vals);
    size_t bytes = sort_tmp64_bytes;
    DCG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_sort_tmp64, bytes, d_sort_keys64[0], d_sort_keys64[1], d_order_vals, d_order_keys[0], (int)M, 0, 64,
                                                 stream));
    k_dc_perm_from_sorted<<<blocks_for(M, 256), 256, 0, stream>>>(d_order_keys[0], M, d_perm_new);
    // fields: old field order -> new field order.  The ping-pong pairs move into their idle halves; the four
    // single buffers go through the accessor scratch.  (Sharded: every process moves the whole pool — the same
    // values into the same cells — and the ranks meet between reading the old and writing the new order.)
    spec_velocity = false;
    barrier();
    const unsigned gb = blocks_for(cells, 256);
    k_dc_permute_field<float4><<<gb, 256, 0, stream>>>(vw[cur_v], vw[cur_v ^ 1], d_perm, d_perm_new, cells);
    k_dc_permute_field<float><<<gb, 256, 0, stream>>>(q[cur_q], q[cur_q ^ 1], d_perm, d_perm_new, cells);
    barrier();
    cur_v ^= 1;
    cur_q ^= 1;
    if (th[0]) {  // extension scalars travel with the density
      k_dc_permute_field<float><<<gb, 256, 0, stream>>>(th[cur_s], th[cur_s ^ 1], d_perm, d_perm_new, cells);
      k_dc_permute_field<float><<<gb, 256, 0, stream>>>(qvp[cur_s], qvp[cur_s ^ 1], d_perm, d_perm_new, cells);
      cur_s ^= 1;
      launches += 2;
    }
    float *single[4] = {fl, p, tp, div};
    for (float *f : single) {
      k_dc_permute_field<float><<<gb, 256, 0, stream>>>(f, scratch, d_perm, d_perm_new, cells);
      barrier();
      k_copy_f32<<<gb, 256, 0, stream>>>(scratch, f, cells);
      barrier();
    }
    launches += 13;
    std::swap(d_perm, d_perm_new);
    mirrored = true;
    mirror();
    n_resorts++;
    changes_since_resort = 0;
    drop_graphs();
    return DCG_OK;
  }

  // called whenever block positions or the set of active blocks changed: per rank, the Morton-ordered list of
  // its active blocks, the lists of its blocks with children, and the tile runs of every level
  int rebuild_order() {
    runs_overflow = false;
    // restriction tree: one GPU walks to the top; sharded ranks own the subtrees below the small levels (k_dc_accumulate_coarse above)
    tree_top = world > 1 ? small_levels_from(512) - 1 : levels - 1;
    if (use_tree) {
      k_dc_tree_expect<<<blocks_for(M, 256), 256, 0, stream>>>(hot(), d_texpect);
      launches++;
    }
    for (int lr = 0; lr < nlocal; lr++) {
      RankWork &w = work[lr];
      const int rank = rank0 + lr;
      const uint8_t *own = world > 1 ? d_unit_owner : nullptr;
      cudaMemsetAsync(d_pcount, 0, (kMaxLevels + 2) * 4, stream);
      k_dc_order_keys<<<blocks_for(M, 256), 256, 0, stream>>>(hot(), order_mode, own, unit, rank, d_order_keys[0], d_order_vals, d_pcount + kMaxLevels);
      size_t bytes = sort_tmp_bytes;
      cub::DeviceRadixSort::SortPairs(d_sort_tmp, bytes, d_order_keys[0], d_order_keys[1], d_order_vals, w.d_order, (int)M, 0, 32, stream);
      k_dc_list_parents<<<blocks_for(M, 256), 256, 0, stream>>>(hot(), own, unit, rank, w.d_plist, d_pcount);
      if (use_tree) {
        k_dc_tree_starts<<<blocks_for(M, 256), 256, 0, stream>>>(hot(), kp, tree_top, world, slab_axis, rank, d_texpect, w.d_tstarts, d_pcount + kMaxLevels + 1);
        launches++;
      }
      cudaMemcpyAsync(h_pcount, d_pcount, (kMaxLevels + 2) * 4, cudaMemcpyDeviceToHost, stream);
      cudaStreamSynchronize(stream);
      for (int l = 0; l < kMaxLevels; l++) w.pcount[l] = h_pcount[l];
      const uint32_t n = h_pcount[kMaxLevels];
      w.n_tstarts = use_tree ? h_pcount[kMaxLevels + 1] : 0;
      w.n_order = (n + kBPC - 1) / kBPC * kBPC;
      if (w.n_order > n) k_dc_order_pad<<<1, 256, 0, stream>>>(w.d_order, n, w.n_order);
      launches += 4;
      w.all = tile_runs(0, 0, (uint32_t)((M64 + kTile - 1) / kTile), rank);
      for (int l = 0; l < levels; l++) w.level[l] = tile_runs(offsets[l], 0, (uint32_t)((loads[l] + kTile - 1) / kTile), rank);
    }
    level_single.assign(levels, 0);
    for (int l = 0; l < levels; l++) {
      if (loads[l] == 0) { level_single[l] = 1; continue; }
      // a tile belongs to the owner of its first slot; the last tile may reach into the next unit
      const uint32_t u0 = (uint32_t)(offsets[l] / unit), u1 = (uint32_t)std::min<uint64_t>(nunits - 1, (offsets[l] + loads[l] - 1) / unit);
      bool one = true;
      for (uint32_t u = u0; u <= u1; u++) one = one && unit_owner[u] == unit_owner[u0];
      level_single[l] = one ? 1 : 0;
    }
    if (runs_overflow) return fail(DCG_ERR_UNSUPPORTED, "ownership pattern needs more than %d tile runs per launch (dcg_options.shard_unit too small)", kMaxRuns);
    return DCG_OK;
  }
  uint32_t finer_full_mask() const {
    uint32_t m = 0;
    for (int l = 0; l < levels; l++)
      if ((l == 0 && loads[0] == full_blocks[0]) || (l > 0 && loads[l - 1] == full_blocks[l - 1])) m |= 1u << l;
    return m;
  }

  // Scores stay on the device; only the 192-byte per-level summary comes back.
  int launch_scores(bool with_block_scores) {
    const uint32_t mask = finer_full_mask();
    if (early_copy_pending) {  // the side stream may still be reading the score buffers of the previous pass
      DCG_CUDA_TRY(cudaStreamWaitEvent(stream, ev_scores_copied, 0));
      early_copy_pending = false;
    }
    DCG_CUDA_TRY(cudaMemcpyAsync(d_summary, h_summary_init, sizeof(ScoreSummary), cudaMemcpyHostToDevice, stream));
    k_dc_subblock_scores<<<blocks_for((size_t)M * 8, 256), 256, 0, stream>>>(T, kp, mask, d_sub_scores, d_summary, ext.score_mode == 1 ? vort : nullptr,
                                                                             d_perm);
    launches++;
    if (with_block_scores) {
      k_dc_block_scores<<<blocks_for(M, 256), 256, 0, stream>>>(T, mask, d_sub_scores, d_block_scores, d_summary);
      launches++;
    }
    return DCG_OK;
  }
  int compute_scores(bool with_block_scores) {
    DCG_TRY(launch_scores(with_block_scores));
    DCG_CUDA_TRY(cudaMemcpyAsync(h_summary, d_summary, sizeof(ScoreSummary), cudaMemcpyDeviceToHost, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    return DCG_OK;
  }
  // the score slices the host selection of `level` reads, device -> pinned host, on stream `st`
  int copy_score_slices(const std::vector<char> &need, cudaStream_t st) {
    std::vector<char> got_bs(levels, 0), got_ss(levels, 0);
    for (int level = 0; level < levels - 1; level++) {
      if (!need[level]) continue;
      // block scores of level l and l+1 (the protect-next-parent write, :417, lands in level l+1), subblock scores of level l+1
      for (int l2 = level; l2 <= level + 1; l2++)
        if (!got_bs[l2]) {
          got_bs[l2] = 1;
          DCG_CUDA_TRY(cudaMemcpyAsync(h_block_scores + offsets[l2], d_block_scores + offsets[l2], max_blocks[l2] * 4, cudaMemcpyDeviceToHost, st));
        }
      if (!got_ss[level + 1]) {
        got_ss[level + 1] = 1;
        // +1 float: dc[matches] may index one past the level's last subblock (App. B-3)
        DCG_CUDA_TRY(cudaMemcpyAsync(h_sub_scores + 8 * offsets[level + 1], d_sub_scores + 8 * offsets[level + 1],
                                     (8 * max_blocks[level + 1] + 1) * 4, cudaMemcpyDeviceToHost, st));
      }
    }
    return DCG_OK;
  }
  // called at the end of an adaptTopology() that leaves the transient running: scores + summary of the settled topology,
  // summary and the slices of every level that has move candidates copied on the side stream (the host selection of the
  // next call then finds them in pinned memory while the GPU is still busy with this step's field kernels)
  int early_scores() {
    if (!use_early_scores || ext.score_mode != 0 || ext.selection == 1) return DCG_OK;
    DCG_TRY(launch_scores(true));
    DCG_CUDA_TRY(cudaEventRecord(ev_scores, stream));
    DCG_CUDA_TRY(cudaStreamWaitEvent(copy_stream, ev_scores, 0));
    DCG_CUDA_TRY(cudaMemcpyAsync(h_summary, d_summary, sizeof(ScoreSummary), cudaMemcpyDeviceToHost, copy_stream));
    std::vector<char> cand(levels, 0);
    for (int level = 0; level < levels - 1; level++) cand[level] = move_candidates(level) > 0 ? 1 : 0;
    DCG_TRY(copy_score_slices(cand, copy_stream));
    DCG_CUDA_TRY(cudaEventRecord(ev_scores_copied, copy_stream));
    early_ready = early_copy_pending = true;
    n_early_scores++;
    return DCG_OK;
  }

  uint64_t move_candidates(int level) const {
    const uint64_t d0 = max_blocks[level], d1 = 8 * max_blocks[level + 1];
    uint64_t l = std::min({d0, d1, loads[level], full_blocks[level] - loads[level]});
    if (move_limit[level] > 0) l = std::min(l, move_limit[level]);
    return l;
  }

  // moveBlocks, :348-437.
  //
  // The reference copies all 9*M scores to the host every step and runs std::nth_element / std::sort
  // over every level (:365-422).  The outcome of a level's greedy match (:410-418) is "no match" —
  // and then nothing of that level's selection is observable — whenever
  //     no subblock of level+1 has a score >= 0,  or  no block of the level has a score >= 0,  or
  //     min{non-negative block scores} >= max{subblock scores}
  // because the loop tests bs[mc[0]] >= 0 && ss[dc[0]] >= 0 && bs[mc[0]] < ss[dc[0]], where ss[dc[0]] is
  // the maximum and bs[mc[0]] is either negative or >= the minimum, whatever permutation the
  // (non-strict-weak, App. B-2) comparator produces.  Those three quantities are reduced on the
  // device (ScoreSummary).  Only levels that can match take the reference's host selection, verbatim
  // (same comparators, same std algorithms, same value sequences => same permutations), on that
  // level's slice of the scores.
  int move_blocks(uint32_t &num_touched) {
    double tm0 = now_ms();
    const bool early = early_ready;  // scores, summary and slices of this topology are already on their way to the host
    early_ready = false;
    if (early) DCG_CUDA_TRY(cudaEventSynchronize(ev_scores_copied));
    else DCG_TRY(compute_scores(true));
    t_sel_scores_ms += now_ms() - tm0;
    DCG_CUDA_TRY(cudaMemsetAsync(d_flags, 0, (size_t)M * 4, stream));  // :369
    std::vector<char> need(levels, 0);
    bool any = false;
    for (int level = 0; level < levels - 1; level++) {
      if (move_candidates(level) == 0) continue;
      const int mx = h_summary->max_ss[level + 1];
      const uint32_t mn = h_summary->min_bs[level];
      if (mx < 0 || mn == 0xFFFFFFFFu || mn >= (uint32_t)mx) {  // non-negative floats order like their bit patterns
        n_levels_shortcut++;
        continue;
      }
      need[level] = 1;
      any = true;
    }
    if (!any) {
      for (int level = 0; level < levels - 1; level++)
        if (move_candidates(level) > 0) move_limit[level] = 0;  // matches == 0 => moveLimit = 0 (:420)
      return DCG_OK;
    }
    if (ext.selection == 1) return move_blocks_device(need, num_touched);
    tm0 = now_ms();
    if (!early) {
      DCG_TRY(copy_score_slices(need, stream));
      DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    }
    t_sel_d2h_ms += now_ms() - tm0;
    tm0 = now_ms();
    const float *bs = h_block_scores, *ss = h_sub_scores;
    float *bsw = h_block_scores;
    // The reference sorts INDEX arrays through comparators that look the scores up (:350-358, :384-407).  Here the same
    // std::nth_element / std::sort calls run on (score, id) pairs with the same comparison outcomes: libstdc++'s
    // algorithms are driven by the comparison results and the element COUNT only, so the permutation of the ids is the
    // one the reference gets — without a dependent random load per comparison (1.9 M candidates at 512^3: 2x faster).
    auto block_order = [](const Cand &a, const Cand &b) { return a.s < 0.f ? false : a.s < b.s; };  // :350-353
    auto sub_order = [](const Cand &a, const Cand &b) { return a.s > b.s; };                        // :356-358
    uint64_t n_move = 0;
    for (int level = 0; level < levels - 1; level++) {
      const uint64_t d0 = max_blocks[level], d1 = 8 * max_blocks[level + 1];
      const uint64_t l = move_candidates(level);
      if (l == 0) continue;
      if (!need[level]) {
        move_limit[level] = 0;
        continue;
      }
      n_host_selections++;
      sel_mc.resize(d0);
      for (uint64_t i = 0; i < d0; i++) { const uint32_t id = (uint32_t)(offsets[level] + i); sel_mc[i] = Cand{bs[id], id}; }  // std::iota, :384
      Cand *mc = sel_mc.data();
      if (d0 <= l)
        std::sort(mc, mc + d0, block_order);
      else {
        std::nth_element(mc, mc + l, mc + d0, block_order);
        std::sort(mc, mc + l, block_order);
      }
      sel_dc.resize(d1);
      for (uint64_t i = 0; i < d1; i++) { const uint32_t id = (uint32_t)(8 * offsets[level + 1] + i); sel_dc[i] = Cand{ss[id], id}; }  // :396
      Cand *dc = sel_dc.data();
      if (d1 <= l)
        std::sort(dc, dc + d1, sub_order);
      else {
        std::nth_element(dc, dc + l, dc + d1, sub_order);
        std::sort(dc, dc + l, sub_order);
      }
      uint64_t matches = 0;
      while (matches < l && mc[matches].s >= 0.f && dc[matches].s >= 0.f && mc[matches].s < dc[matches].s) {
        h_to_move[n_move + matches] = mc[matches].id;
        h_dest[n_move + matches] = dc[matches].id;
        matches++;
        // :417 — protects the NEXT candidate's parent (App. B-3); that parent is a level+1 block, whose
        // scores were fetched above.  (matches == d1 would read past the list in the reference too.)
        if (matches < d1) bsw[dc[matches].id / 8] = -FLT_MAX;
      }
      move_limit[level] = (uint64_t)(matches * 1.2f);  // :420
      n_move += matches;
    }
    t_sel_host_ms += now_ms() - tm0;
    if (n_move > 0) {
      const uint32_t n = (uint32_t)n_move;
      DCG_CUDA_TRY(cudaMemcpyAsync(d_to_move, h_to_move, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
      DCG_CUDA_TRY(cudaMemcpyAsync(d_dest, h_dest, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
      k_dc_move_prepare<<<blocks_for(n, 256), 256, 0, stream>>>(T, d_to_move, d_dest, n, d_new_posl);
      k_dc_move_commit<<<blocks_for(n, 256), 256, 0, stream>>>(T, kp, d_to_move, d_dest, n, d_new_posl, d_flags);
      k_dc_map_insert<<<blocks_for(n, 256), 256, 0, stream>>>(T, kp, d_to_move, n);
      launches += 3;
      DCG_CUDA_TRY(cudaMemcpyAsync(d_touched, d_to_move, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));  // :433-434
      DCG_CUDA_TRY(cudaStreamSynchronize(stream));  // h_to_move / h_dest are reused by refine
      num_touched += n;
      n_moved += n;
    }
    return DCG_OK;
  }

  // moveBlocks with the total-order selection (dcg_ext_params.selection == 1; oracle move_blocks, "EXTENSION"): keys,
  // two radix sorts, the monotone greedy rule and the protection of the receiving parents, level after level on the
  // stream; ONE 4-byte-per-level read-back (the match counts feed moveLimit and the list sizes).
  int move_blocks_device(const std::vector<char> &need, uint32_t &num_touched) {
    DCG_TRY(ensure_sel_storage());
    DCG_CUDA_TRY(cudaMemsetAsync(d_sel_sc, 0, 4 * kMaxLevels * 4, stream));
    std::vector<uint64_t> base(levels, 0), lim(levels, 0);
    uint64_t upper = 0;
    for (int level = 0; level < levels - 1; level++) {
      const uint64_t l = move_candidates(level);
      lim[level] = l;
      base[level] = upper;
      if (l == 0 || !need[level]) continue;
      n_device_selections++;
      const uint32_t d0 = (uint32_t)max_blocks[level], d1 = (uint32_t)(8 * max_blocks[level + 1]);
      uint32_t *sc = d_sel_sc + 4 * level;
      // movable blocks of the level, ascending
      sel::k_sel_keys<false><<<blocks_for(d0, 256), 256, 0, stream>>>(d_block_scores, (uint32_t)offsets[level], d0, d_sel_keys[0], sc + 0);
      size_t bytes = sel_tmp_bytes;
      DCG_CUDA_TRY(cub::DeviceRadixSort::SortKeys(d_sel_tmp, bytes, d_sel_keys[0], d_sel_keys[1], (int)d0, 0, 64, stream));
      // only the first l entries of the sorted block list are needed: keep them in the tail of buffer 0's space? no —
      // copy them aside (l <= d0 <= M)
      DCG_CUDA_TRY(cudaMemcpyAsync(d_sort_keys64_sel(), d_sel_keys[1], std::min<uint64_t>(l, d0) * 8, cudaMemcpyDeviceToDevice, stream));
      // destinations of level + 1, descending
      sel::k_sel_keys<true><<<blocks_for(d1, 256), 256, 0, stream>>>(d_sub_scores, (uint32_t)(8 * offsets[level + 1]), d1, d_sel_keys[0], sc + 1);
      bytes = sel_tmp_bytes;
      DCG_CUDA_TRY(cub::DeviceRadixSort::SortKeys(d_sel_tmp, bytes, d_sel_keys[0], d_sel_keys[1], (int)d1, 0, 64, stream));
      sel::k_sel_match_init<<<1, 1, 0, stream>>>(sc, (uint32_t)l);
      sel::k_sel_match<<<blocks_for(l, 256), 256, 0, stream>>>(d_block_scores, d_sub_scores, d_sort_keys64_sel(), d_sel_keys[1], sc);
      sel::k_sel_apply<<<blocks_for(l, 256), 256, 0, stream>>>(d_block_scores, d_sort_keys64_sel(), d_sel_keys[1], sc, d_sel_mc + upper, d_sel_dc + upper);
      launches += 7;
      upper += l;
    }
    DCG_CUDA_TRY(cudaMemcpyAsync(h_sel_sc, d_sel_sc, 4 * kMaxLevels * 4, cudaMemcpyDeviceToHost, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    uint64_t n_move = 0;
    for (int level = 0; level < levels - 1; level++) {
      if (lim[level] == 0) continue;
      const uint64_t matches = need[level] ? h_sel_sc[4 * level + 3] : 0;
      move_limit[level] = (uint64_t)(matches * 1.2f);  // :420
      if (matches == 0) continue;
      DCG_CUDA_TRY(cudaMemcpyAsync(d_to_move + n_move, d_sel_mc + base[level], matches * 4, cudaMemcpyDeviceToDevice, stream));
      DCG_CUDA_TRY(cudaMemcpyAsync(d_dest + n_move, d_sel_dc + base[level], matches * 4, cudaMemcpyDeviceToDevice, stream));
      n_move += matches;
    }
    if (n_move > 0) {
      const uint32_t n = (uint32_t)n_move;
      k_dc_move_prepare<<<blocks_for(n, 256), 256, 0, stream>>>(T, d_to_move, d_dest, n, d_new_posl);
      k_dc_move_commit<<<blocks_for(n, 256), 256, 0, stream>>>(T, kp, d_to_move, d_dest, n, d_new_posl, d_flags);
      k_dc_map_insert<<<blocks_for(n, 256), 256, 0, stream>>>(T, kp, d_to_move, n);
      launches += 3;
      DCG_CUDA_TRY(cudaMemcpyAsync(d_touched, d_to_move, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));  // :433-434
      num_touched += n;
      n_moved += n;
    }
    return DCG_OK;
  }
  // the first l sorted block keys of a level, kept while buffer 1 is reused for the destinations (8 B x M)
  unsigned long long *sel_mc_keys = nullptr;
  unsigned long long *d_sort_keys64_sel() {
    if (!sel_mc_keys) cudaMalloc(&sel_mc_keys, (size_t)M * 8);
    return sel_mc_keys;
  }

  // Candidate lists of refineSubblocks, compacted on the device (sel::k_sel_*): for every level with room and
  // candidates, the ids with score > 1e-4 in ascending order — the exact sequence of the reference's host scan (:455-463).
  // Where the list fits the room (n <= limit) it IS the reference's result; where it does not, the reference keeps the
  // `limit` largest ids in the order std::nth_element leaves them (libstdc++-specific): that level's compact list goes
  // to the host for the same call on the same sequence (selection == 0), or its tail is taken in ascending order
  // (selection == 1).  No score array crosses PCIe either way.
  int refine_lists_device(const std::vector<uint64_t> &limits, uint64_t &n_ref, RefineGroups &G, std::vector<uint64_t> &added) {
    DCG_TRY(ensure_sel_storage());
    uint32_t *cand = reinterpret_cast<uint32_t *>(d_sel_keys[0]);
    for (int level = 1; level < levels; level++) {
      const uint64_t limit = limits[level];
      G.start[level - 1] = (uint32_t)n_ref;
      G.base[level - 1] = (uint32_t)(offsets[level - 1] + loads[level - 1]);
      const uint64_t n = h_summary->n_refine[level];
      if (limit == 0 || n == 0) continue;
      const uint32_t first = (uint32_t)(8 * offsets[level]), cnt = (uint32_t)(8 * max_blocks[level]);
      const unsigned nct = blocks_for(cnt, sel::kSelCta);
      sel::k_sel_count<<<nct, sel::kSelCta, 0, stream>>>(d_sub_scores, first, cnt, 1e-4f, d_sel_cta);
      sel::k_sel_scan<<<1, sel::kSelCta, 0, stream>>>(d_sel_cta, nct, d_sel_cta + nct);
      sel::k_sel_scatter<<<nct, sel::kSelCta, 0, stream>>>(d_sub_scores, first, cnt, 1e-4f, d_sel_cta, cand);
      launches += 3;
      const uint64_t take = std::min(n, limit);
      if (n <= limit || ext.selection == 1) {
        n_device_selections++;
        DCG_CUDA_TRY(cudaMemcpyAsync(d_dest + n_ref, cand + (n - take), take * 4, cudaMemcpyDeviceToDevice, stream));
      } else {
        n_selection_fallbacks++;
        n_host_selections++;
        uint32_t *di = h_dest + n_ref;
        DCG_CUDA_TRY(cudaMemcpyAsync(di, cand, n * 4, cudaMemcpyDeviceToHost, stream));
        DCG_CUDA_TRY(cudaStreamSynchronize(stream));
        std::nth_element(di, di + limit, di + n, std::greater<uint32_t>{});  // keeps the LARGEST ids (App. B-4)
        DCG_CUDA_TRY(cudaMemcpyAsync(d_dest + n_ref, di, take * 4, cudaMemcpyHostToDevice, stream));
        DCG_CUDA_TRY(cudaStreamSynchronize(stream));  // h_dest is reused by the next level
      }
      added[level - 1] = take;
      n_ref += take;
    }
    return DCG_OK;
  }

  // refineSubblocks, :439-483.  Levels with no room (limit == 0, the normal case once the pool is
  // full) or no candidate (device-counted) are skipped without copying scores.
  int refine_subblocks(uint32_t &num_touched) {
    std::vector<uint64_t> limits(levels, 0);
    bool any = false;
    for (int level = 1; level < levels; level++) {
      limits[level] = std::min(max_blocks[level - 1] - loads[level - 1], 8 * loads[level] - loads[level - 1]);
      any = any || limits[level] > 0;
    }
    if (!any) return DCG_OK;
    DCG_TRY(compute_scores(false));
    any = false;
    for (int level = 1; level < levels; level++) any = any || (limits[level] > 0 && h_summary->n_refine[level] > 0);
    if (!any) return DCG_OK;
    if (!opt.host_selection || ext.selection == 1) {  // candidate lists compacted on the device
      uint64_t n_ref = 0;
      RefineGroups G{};
      std::vector<uint64_t> added(levels, 0);
      DCG_TRY(refine_lists_device(limits, n_ref, G, added));
      if (n_ref > 0) {
        const uint32_t n = (uint32_t)n_ref;
        if ((size_t)num_touched + n > (size_t)2 * M) return fail(DCG_ERR_POOL, "touched list overflow");
        k_dc_refine<<<blocks_for(n, 256), 256, 0, stream>>>(T, kp, d_dest, n, G, d_free, d_flags, d_touched, num_touched, d_counters);
        launches++;
        for (int l = 0; l < levels; l++) loads[l] += added[l];
        sync_loads();
        num_touched += n;
        n_refined += n;
      }
      return DCG_OK;
    }
    for (int level = 1; level < levels; level++) {
      if (limits[level] == 0 || h_summary->n_refine[level] == 0) continue;
      DCG_CUDA_TRY(cudaMemcpyAsync(h_sub_scores + 8 * offsets[level], d_sub_scores + 8 * offsets[level], 8 * max_blocks[level] * 4,
                                   cudaMemcpyDeviceToHost, stream));
    }
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    const float *ss = h_sub_scores;
    uint64_t n_ref = 0;
    RefineGroups G{};
    std::vector<uint64_t> added(levels, 0);
    for (int level = 1; level < levels; level++) {
      const uint64_t limit = limits[level];
      G.start[level - 1] = (uint32_t)n_ref;
      G.base[level - 1] = (uint32_t)(offsets[level - 1] + loads[level - 1]);
      if (limit == 0 || h_summary->n_refine[level] == 0) continue;
      const uint64_t start = 8 * offsets[level], end = start + 8 * max_blocks[level];
      uint32_t *di = h_dest + n_ref;
      uint64_t n = 0;
      for (uint64_t i = start; i < end; i++)
        if (ss[i] > 1e-4f) di[n++] = (uint32_t)i;
      if (n > limit) std::nth_element(di, di + limit, di + n, std::greater<uint32_t>{});  // keeps the LARGEST ids (App. B-4)
      added[level - 1] = std::min(n, limit);
      n_ref += added[level - 1];
    }
    if (n_ref > 0) {
      const uint32_t n = (uint32_t)n_ref;
      if ((size_t)num_touched + n > (size_t)2 * M) return fail(DCG_ERR_POOL, "touched list overflow");
      DCG_CUDA_TRY(cudaMemcpyAsync(d_dest, h_dest, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
      k_dc_refine<<<blocks_for(n, 256), 256, 0, stream>>>(T, kp, d_dest, n, G, d_free, d_flags, d_touched, num_touched, d_counters);
      launches++;
      for (int l = 0; l < levels; l++) loads[l] += added[l];
      sync_loads();
      num_touched += n;
      n_refined += n;
    }
    return DCG_OK;
  }

  int adapt_topology() override {  // :320-346
    DCG_TRY(enter());
    n_adapt++;
    spec_velocity = false;
    if (steady) {
      n_skipped++;
      return DCG_OK;
    }
    const std::vector<uint64_t> limit_before = move_limit;
    uint32_t num_touched = 0;
    if (ext.score_mode == 1) launch_vorticity();  // k_dcgrid_calc_vorticity at the head of adaptTopology (:321): live with the flow-driven score
    double t0 = now_ms();
    DCG_TRY(move_blocks(num_touched));
    double t1 = now_ms();
    t_move_ms += t1 - t0;
    DCG_TRY(refine_subblocks(num_touched));
    t0 = now_ms();
    t_refine_ms += t0 - t1;
    if (num_touched > 0) {
      n_changed++;
      drop_graphs();
      k_dc_flag_bits<<<blocks_for(M, 256), 256, 0, stream>>>(d_flags, M, d_flag_bits);
      k_dc_refresh_apron<<<blocks_for(M, 8), 256, 0, stream>>>(T, kp, d_flags, d_flag_bits, mirrored ? Tf.apron : nullptr, d_perm);
      launches += 2;
      // the field kernels' view of the pool: new blocks take the field slot of their (fresh) reference slot, moved
      // blocks keep theirs; every `resort_every` changes the sparse levels are re-sorted by position
      changes_since_resort++;
      if (timing_sync) cudaStreamSynchronize(stream);
      t1 = now_ms();
      t_apron_ms += t1 - t0;
      if (use_resort && resort_every > 0 && changes_since_resort >= resort_every) DCG_TRY(resort());
      else mirror(false);
      DCG_TRY(build_face_descriptors());
      t0 = now_ms();
      t_layout_ms += t0 - t1;
      for (int l = levels - 2; l >= 0; l--) {
        // sharded: every process interpolates every new block (identical values); lock step between the levels,
        // a level reads what the coarser one wrote
        k_dc_propagate<<<num_touched, 64, 0, stream>>>(hot(), kp, d_touched, d_perm, l, vw[cur_v], q[cur_q], fl);
        launches++;
        if (ext.sources) {
          ext::k_dc_ext_propagate<<<num_touched, 64, 0, stream>>>(hot(), d_touched, d_perm, l, th[cur_s], qvp[cur_s]);
          launches++;
        }
        barrier();
      }
      uint32_t h_cnt[2] = {0, 0};
      DCG_CUDA_TRY(cudaMemcpyAsync(h_cnt, d_counters, 8, cudaMemcpyDeviceToHost, stream));
      DCG_CUDA_TRY(cudaStreamSynchronize(stream));
      n_failed = h_cnt[0];
      n_irregular = h_cnt[1];
      t_propagate_ms += now_ms() - t0;
      DCG_TRY(early_scores());
    } else if (move_limit == limit_before && ext.score_mode == 0) {  // (flow-driven scores change with the fields: never a fixed point)
      steady = true;  // nothing changed and the selection state is unchanged: fixed point
      if (use_resort && changes_since_resort > 0) {  // the layout the steady state will run on, for good
        DCG_TRY(resort());
        DCG_TRY(build_face_descriptors());
      }
    } else {
      DCG_TRY(early_scores());  // nothing moved, the move limits did: the transient goes on
    }
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }

  // ---- fluid stages ----------------------------------------------------------------------------------
  // first level of the "small" tail: every level >= the returned one has at most `cap` active blocks
  int small_levels_from(uint64_t cap) const {
    int l = levels - 1;
    while (l > 0 && loads[l - 1] <= cap) l--;
    return l;
  }
  // `fused`: the kernel that produced the field restricted every block without children itself, so level 0
  // (never refined) needs no pass at all and the other levels only push up blocks that have children.
  template <bool kV, bool kS>
  void launch_restrict_tree(float4 *v, float *ch) {
    each_rank([&](int, RankWork &w) {
      if (w.n_tstarts == 0) return;
      launch_pdl(k_dc_restrict_tree<kV, kS>, dim3(blocks_for(8 * (size_t)w.n_tstarts, 256)), dim3(256), 0, hot(), (const uint32_t *)w.d_tstarts, w.n_tstarts,
                 (const uint8_t *)d_texpect, d_tcount, tree_top, v, ch);
      launches++;
    });
  }
  void accumulate(float4 *v, float *ch, bool fused) {  // :496-515, fine -> coarse
    if (fused && use_tree) {  // blocks with children: one launch walks up the block tree (k_dc_restrict_tree)
      if (v && ch) launch_restrict_tree<true, true>(v, ch);
      else if (v) launch_restrict_tree<true, false>(v, nullptr);
      else launch_restrict_tree<false, true>(nullptr, ch);
      if (world > 1) {  // the levels above the ranks' subtrees
        barrier();
        if (tree_top + 1 < levels - 1) {
          if (has_rank0()) {
            launch_pdl(k_dc_accumulate_coarse, dim3(kAccClusterCTAs), dim3(kAccClusterThreads), 0, hot(), tree_top + 1, v, ch);
            launches++;
          }
          barrier();
        }
      }
      return;
    }
    const int tail = small_levels_from(512);
    for (int l = fused ? 1 : 0; l < levels - 1 && l < tail; l++) {
      if (loads[l] == 0) continue;
      if (fused) {  // only blocks with children are left (k_dc_list_parents)
        each_rank([&](int, RankWork &w) {
          const uint32_t n = w.pcount[l];
          if (n == 0) return;
          const dim3 grid(blocks_for(8 * (size_t)n, 256));
          const uint32_t *list = w.d_plist + offsets[l];
          if (v && ch) launch_pdl(k_dc_accumulate_list<true, true>, grid, dim3(256), 0, hot(), list, n, v, ch);
          else if (v) launch_pdl(k_dc_accumulate_list<true, false>, grid, dim3(256), 0, hot(), list, n, v, (float *)nullptr);
          else launch_pdl(k_dc_accumulate_list<false, true>, grid, dim3(256), 0, hot(), list, n, (float4 *)nullptr, ch);
          launches++;
        });
        barrier();
      } else if (v) {
        launch_pdl(k_dc_accumulate_velocity, dim3(blocks_for(8 * loads[l], 256)), dim3(256), 0, hot(), l, v, 0);
        launches++;
      } else {
        launch_pdl(k_dc_accumulate_scalar, dim3(blocks_for(8 * loads[l], 256)), dim3(256), 0, hot(), l, ch, 0);
        launches++;
      }
    }
    if (tail < levels - 1) {
      if (has_rank0()) {
        launch_pdl(k_dc_accumulate_coarse, dim3(kAccClusterCTAs), dim3(kAccClusterThreads), 0, hot(), tail, v, ch);
        launches++;
      }
      barrier();
    }
  }
  void accumulate_velocity(bool fused) { accumulate(vw[cur_v], nullptr, fused); }
  void accumulate_scalar(float *ch, bool fused) { accumulate(nullptr, ch, fused); }
  const void *advect_fn(int mode, bool with_scalars = false) const {
    if (with_scalars) {  // extension scalars ride along (modes 1 and 2, the default register budget)
      if (mode == 1) return (const void *)k_dc_advect_pipe<1, 3, true>;
      return (const void *)k_dc_advect_pipe<2, 3, true>;
    }
    if (advect_min_blocks == 2) {
      if (mode == 0) return (const void *)k_dc_advect_pipe<0, 2, false>;
      if (mode == 1) return (const void *)k_dc_advect_pipe<1, 2, false>;
      return (const void *)k_dc_advect_pipe<2, 2, false>;
    }
    if (advect_min_blocks == 3) {
      if (mode == 0) return (const void *)k_dc_advect_pipe<0, 3, false>;
      if (mode == 1) return (const void *)k_dc_advect_pipe<1, 3, false>;
      return (const void *)k_dc_advect_pipe<2, 3, false>;
    }
    if (mode == 0) return (const void *)k_dc_advect_pipe<0, 4, false>;
    if (mode == 1) return (const void *)k_dc_advect_pipe<1, 4, false>;
    return (const void *)k_dc_advect_pipe<2, 4, false>;
  }
  // mode 0: vout <- advected velocity; 1: qout <- advected density; 2: both (k_dc_advect_pipe)
  // scalars != nullptr: temperature and vapor advected in the same pass (modes 1 and 2)
  void launch_advect_pipe(int mode, const float4 *vin, float4 *vout, const float *qi, float *qo, const AdvectScalars *scalars = nullptr) {
    each_rank([&](int, RankWork &w) {
      if (w.n_order == 0) return;
      const unsigned grid = std::min<unsigned>(w.n_order / kBPC, (unsigned)((scalars ? advect_ext_per_sm[mode] : advect_per_sm[mode]) * sm_count));
      const float *flp = fl;
      const uint32_t *ord = w.d_order;
      Pool hp = hot();
      AdvectScalars xs{};
      if (scalars) xs = *scalars;
      void *args[] = {&hp, &kp, &ord, &w.n_order, &vin, &vout, &flp, &qi, &qo, &xs};
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kAdvectThreads); cfg.dynamicSmemBytes = kAdvectPipeSmem; cfg.stream = stream;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at.val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = &at;
      cfg.numAttrs = use_pdl ? 1 : 0;
      cudaLaunchKernelExC(&cfg, advect_fn(mode, scalars != nullptr), args);
      launches++;
    });
    barrier();
  }
  int advect_velocity() override {  // :263-268
    DCG_TRY(enter());
    if (ext.advection == 1) {  // extension: MacCormack
      spec_velocity = false;
      sl_velocity(vw[cur_v], vw[cur_v ^ 1]);
      ext::k_dc_ext_maccormack_velocity<<<blocks_for(M, kBPC), kCTA, 0, stream>>>(hot(), kp, vw[cur_v], vw[cur_v ^ 1], vw_mc);
      launches++;
      std::swap(vw[cur_v], vw_mc);
      accumulate(vw[cur_v], nullptr, false);
      DCG_CUDA_TRY(cudaGetLastError());
      return DCG_OK;
    }
    if (spec_velocity) {
      // vw[cur_v ^ 1] already holds this step's advected velocity (written by the previous advect_density())
      spec_velocity = false;
      cur_v ^= 1;
      if (!spec_restricted) accumulate_velocity(true);
      spec_restricted = false;
    } else if (use_advect_pipe) {
      launch_advect_pipe(0, vw[cur_v], vw[cur_v ^ 1], nullptr, nullptr);
      cur_v ^= 1;
      accumulate_velocity(true);
    } else {
      k_dc_advect_velocity<<<blocks_for(M, kBPC), kCTA, 0, stream>>>(hot(), kp, vw[cur_v], vw[cur_v ^ 1]);
      launches++;
      cur_v ^= 1;
      accumulate_velocity(false);
    }
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  int advect_density() override {  // :313-318
    DCG_TRY(enter());
    if (ext.advection == 1) {  // extension: MacCormack, every advected scalar
      spec_velocity = false;
      maccormack_scalar(0, q[cur_q], q[cur_q ^ 1]);
      if (ext.sources) {
        maccormack_scalar(1, th[cur_s], th[cur_s ^ 1]);
        maccormack_scalar(2, qvp[cur_s], qvp[cur_s ^ 1]);
      }
      DCG_CUDA_TRY(cudaGetLastError());
      return DCG_OK;
    }
    const bool ride = ext.sources && use_advect_pipe;  // temperature and vapor in the density's own gather pass
    if (ext.sources && !ride) {  // extension: temperature and vapor ride along (same trajectories; before the velocity buffer flips)
      sl_scalar(1, th[cur_s], th[cur_s ^ 1]);
      sl_scalar(2, qvp[cur_s], qvp[cur_s ^ 1]);
      cur_s ^= 1;
    }
    if (use_advect_pipe) {
      AdvectScalars xs{};
      if (ride) { xs.th_in = th[cur_s]; xs.qv_in = qvp[cur_s]; xs.th_out = th[cur_s ^ 1]; xs.qv_out = qvp[cur_s ^ 1]; xs.E = ext; }
      launch_advect_pipe(fuse_advect ? 2 : 1, vw[cur_v], vw[cur_v ^ 1], q[cur_q], q[cur_q ^ 1], ride ? &xs : nullptr);
      if (ride) {
        cur_s ^= 1;
        accumulate(nullptr, th[cur_s], true);
        accumulate(nullptr, qvp[cur_s], true);
      }
      spec_velocity = fuse_advect;
      cur_q ^= 1;
      // the speculative velocity is restricted in the same launch as the density (advect_velocity() then only flips)
      spec_restricted = fuse_advect;
      if (spec_restricted) accumulate(vw[cur_v ^ 1], q[cur_q], true);
      else accumulate_scalar(q[cur_q], true);
    } else {
      k_dc_advect_density<<<blocks_for(M, kBPC), kCTA, 0, stream>>>(hot(), kp, vw[cur_v], fl, q[cur_q], q[cur_q ^ 1]);
      launches++;
      cur_q ^= 1;
      accumulate_scalar(q[cur_q], false);
    }
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  // one sweep of a level; persistent TMA-ring kernel for levels with enough tiles to fill the machine
  // (Measured and rejected, profiles/README.md r2q / r2z: the barrier between two sweeps of a shared level moved into the ring
  // kernels.  Whole barrier folded (last CTA of a sweep raises the epoch at the peers, every CTA of the next sweep waits before
  // its first ghost load): 3.07 against 2.98 ms per step on 2 B200.  Halo overlap (tile lists with the tiles that touch another
  // rank first, epoch raised once those are done, interior tiles afterwards): 2.885 against 2.874 ms on 2 B200, 3.24-3.27 against
  // 3.26 on 4 — the 18 barrier launches it removes are paid back by the sweep kernel itself (list indirection, boundary-first
  // order against the L2-friendly snake order).  The stand-alone barrier kernel stays.)
  void jacobi_sweep(int l, const float *in, float *out) {
    if (snake) sweep_parity ^= 1;
    each_rank([&](int, RankWork &w) {
      const TileRuns &R = w.level[l];
      const unsigned tiles = run_total(R);
      if (tiles == 0) return;
      // (the tiles THIS rank sweeps decide: half of a level that fills one GPU does not fill two)
      if (use_pipe && jacobi8 && tiles >= pipe_min_tiles) {
        const unsigned grid = std::min<unsigned>(tiles, (unsigned)jacobi8_ctas);
        launch_pdl(k_dc_jacobi_pipe8, dim3(grid), dim3(kJ8Threads), kJacobiPipeSmem, hot(), kp, R, l, in, out, div, snake ? sweep_parity : 0);
      } else if (use_pipe && tiles >= pipe_min_tiles) {
        const unsigned grid = std::min<unsigned>(tiles, (unsigned)jacobi_pipe_ctas);
        launch_pdl(k_dc_jacobi_pipe, dim3(grid), dim3(kCTA4), kJacobiPipeSmem, hot(), kp, R, l, in, out, div, snake ? sweep_parity : 0);
      } else {
        launch_pdl(k_dc_jacobi4, dim3(tiles), dim3(kCTA4), 0, hot(), kp, R, l, in, out, div);
      }
      launches++;
    });
    if (!level_single[l]) barrier();  // a level held by one rank: one barrier after its last sweep (end_level)
  }
  void end_level(int l) {
    if (level_single[l]) barrier();
  }
  void jacobi_pair(int l) {
    if (loads[l] == 0) return;
    jacobi_sweep(l, p, tp);
    jacobi_sweep(l, tp, p);
  }
  void launch_prolongate(int l) {
    // a level that fills the GPU(s): by parent block (every block of the level is a child of exactly one listed block; sharded:
    // a rank writes the children of ITS parents wherever they live, so every rank must take the same path — the test uses the
    // level's global size — and the ranks meet afterwards even if the level itself has one owner)
    // (one GPU: the per-child kernel; the by-parent kernel is faster alone, 36 against 48 us, and slower inside the step, +15 us)
    const bool by_parent = prolong_staged && l + 1 < levels - 1 && (loads[l] + kB4 - 1) / kB4 >= (uint64_t)pipe_min_tiles * (uint64_t)world &&
                           (world > 1) != ((experiment & 1) != 0);
    each_rank([&](int, RankWork &w) {
      const TileRuns &R = w.level[l];
      if (by_parent) {
        if (w.pcount[l + 1] == 0) return;
        launch_pdl(k_dc_prolongate_parents, dim3(std::min<unsigned>(w.pcount[l + 1], 8u * (unsigned)sm_count)), dim3(kPPThreads), 0, hot(),
                   (const uint32_t *)(w.d_plist + offsets[l + 1]), w.pcount[l + 1], p);
      } else {
        if (run_total(R) == 0) return;
        if (prolong_staged) launch_pdl(k_dc_prolongate_staged, dim3(run_total(R)), dim3(kCTA4), 0, hot(), R, l, p);
        else launch_pdl(k_dc_prolongate4, dim3(blocks_for(loads[l], kB4)), dim3(kCTA4), 0, hot(), l, p);
      }
      launches++;
    });
    if (!level_single[l] || (by_parent && world > 1)) barrier();  // single-owner level written by its owner: the same rank sweeps it next
  }
  void launch_divergence(int zero_from) {
    each_rank([&](int, RankWork &w) {
      const unsigned tiles = run_total(w.all);
      if (tiles == 0) return;
      if (use_stencil_pipe)
        launch_pdl(k_dc_divergence_pipe, dim3(std::min<unsigned>(tiles, (unsigned)div_pipe_ctas)), dim3(kStencilThreads), kDivPipeSmem, hot(), kp, w.all, vw[cur_v], div, p,
                                                                                                                            tp, zero_from);
      else
        launch_pdl(k_dc_divergence4, dim3(blocks_for(M, kB4)), dim3(kCTA4), 0, hot(), kp, vw[cur_v], div, p, tp, zero_from);
      launches++;
    });
    barrier();
  }
  void launch_apply() {
    each_rank([&](int, RankWork &w) {
      const unsigned tiles = run_total(w.all);
      if (tiles == 0) return;
      const unsigned grid = std::min<unsigned>(tiles, (unsigned)apply_pipe_ctas);
      if (use_stencil_pipe && apply_min_blocks == 2)
        launch_pdl(k_dc_apply_pipe<2>, dim3(grid), dim3(kStencilThreads), kApplyPipeSmem, hot(), kp, w.all, p, fl, vw[cur_v]);
      else if (use_stencil_pipe)
        launch_pdl(k_dc_apply_pipe<3>, dim3(grid), dim3(kStencilThreads), kApplyPipeSmem, hot(), kp, w.all, p, fl, vw[cur_v]);
      else
        launch_pdl(k_dc_apply_pressure4, dim3(blocks_for(M, kB4)), dim3(kCTA4), 0, hot(), kp, p, fl, vw[cur_v]);
      launches++;
    });
    barrier();
  }
  void divergence_stage(int zero_from) {
    launch_divergence(zero_from);
    accumulate_scalar(div, true);
  }
  void apply_stage() {
    launch_apply();
    accumulate_velocity(true);
  }
  // the coarse tail of the cascade in one single-CTA launch, its fields in shared memory when they fit
  void launch_coarse_cascade(int cf, int prolong_coarsest, int pairs_coarsest, int pairs_level, int prolong_levels) {
    const uint64_t end = offsets[levels - 1] + max_blocks[levels - 1];
    const uint32_t ncell = (uint32_t)((end - offsets[cf]) * kBV);
    const size_t smem = (size_t)ncell * (3 * 4 + 6 * 2);
    if (has_rank0()) {  // sharded: one rank walks the coarse tail, the others wait at the barrier
      if (coarse_in_smem && smem <= kCoarseSmemMax && ncell <= 65535)
        launch_pdl(k_dc_coarse_cascade<true>, dim3(1), dim3(1024), smem, hot(), kp, cf, prolong_coarsest, pairs_coarsest, pairs_level, prolong_levels, p, tp, div, ncell);
      else
        launch_pdl(k_dc_coarse_cascade<false>, dim3(1), dim3(1024), 0, hot(), kp, cf, prolong_coarsest, pairs_coarsest, pairs_level, prolong_levels, p, tp, div, 0);
      launches++;
    }
    barrier();
  }
  int project() override {  // :270-294
    DCG_TRY(enter());
    spec_velocity = false;
    divergence_stage(skip_dead_zeroing && project_level_pairs >= 1 ? levels - 1 : 0);
    // levels with <= kCoarseBlocks blocks: the whole coarse part of the cascade in one single-CTA launch
    const int cf = small_levels_from(kCoarseBlocks);
    launch_coarse_cascade(cf, 0, project_coarsest_pairs, project_level_pairs, 1);
    for (int l = cf - 1; l >= 0; l--) {
      if (loads[l] == 0) continue;
      launch_prolongate(l);
      for (int i = 0; i < project_level_pairs; i++) jacobi_pair(l);
      end_level(l);
    }
    apply_stage();
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }
  int project_local() override {  // :296-311
    DCG_TRY(enter());
    spec_velocity = false;
    divergence_stage(0);
    const int cf = small_levels_from(kCoarseBlocks);
    launch_coarse_cascade(cf, 0, local_pairs, local_pairs, 0);
    for (int l = cf - 1; l >= 0; l--) {
      for (int i = 0; i < local_pairs; i++) jacobi_pair(l);
      end_level(l);
    }
    apply_stage();
    DCG_CUDA_TRY(cudaGetLastError());
    return DCG_OK;
  }

  int step(int n) override {
    DCG_TRY(enter());
    DCG_CUDA_TRY(cudaEventRecord(ev_begin, stream));
    for (int done = 0; done < n; done++) {
      // transient: adaptation needs host round trips, run call by call (MacCormack rotates buffers: never replayed)
      if (!steady || ext.advection || spec_velocity != (fuse_advect && use_advect_pipe)) {
        DCG_TRY(dcg_sim::step(1));
        continue;
      }
      cudaGraphExec_t &ge = step_graph[cur_v * 4 + cur_q * 2 + cur_s];
      for (int attempt = 0; !ge; attempt++) {
        const uint64_t before = launches, adapt_before = n_adapt, skipped_before = n_skipped, barriers_before = n_barriers;
        const int sv = cur_v, sq = cur_q, ss = cur_s, sp = sweep_parity;
        const bool sspec = spec_velocity;
        cudaGraph_t g = nullptr;
        DCG_CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        const int rc = dcg_sim::step(1);  // adapt_topology() is a no-op in the steady state
        cudaError_t ce = cudaStreamEndCapture(stream, &g);
        cur_v = sv; cur_q = sq; cur_s = ss; spec_velocity = sspec;  // capture records, it does not execute
        step_graph_launches = launches - before;
        launches = before; n_adapt = adapt_before; n_skipped = skipped_before; n_barriers = barriers_before;
        if (rc != DCG_OK && ce == cudaSuccess) return rc;
        if (ce == cudaSuccess) ce = cudaGraphInstantiate(&ge, g, 0);
        if (g) cudaGraphDestroy(g);
        if (ce != cudaSuccess) {
          ge = nullptr;
          cudaGetLastError();
          if (use_pdl && attempt == 0) {  // programmatic edges not capturable on this driver: plain stream order
            use_pdl = false;
            pdl_fallbacks++;
            sweep_parity = sp;
            continue;
          }
          DCG_CUDA_TRY(ce);
        }
      }
      DCG_CUDA_TRY(cudaGraphLaunch(ge, stream));
      launches += step_graph_launches;
      n_adapt++; n_skipped++;
      cur_v ^= 1;
      cur_q ^= 1;
      if (ext.sources) cur_s ^= 1;
    }
    DCG_CUDA_TRY(cudaEventRecord(ev_end, stream));
    step_timing_pending = true;
    return DCG_OK;
  }

  // one launch of a single stage, `reps` times, CUDA-event timed on the instance's stream
  int bench_stage(const char *stage, int level, int reps, float *ms_per_launch, double *alg_bytes) override {
    DCG_TRY(enter());
    const std::string st(stage);
    spec_velocity = false;
    if (level < 0 || level >= levels) return fail(DCG_ERR_INVALID, "bench_stage: bad level");
    double call = 0;
    for (int l = 0; l < levels; l++) call += 64.0 * (double)loads[l];
    const double cl = 64.0 * (double)loads[level];
    double bytes = 0;
    DCG_CUDA_TRY(cudaEventRecord(ev_begin, stream));
    for (int r = 0; r < reps; r++) {
      if (st == "jacobi" || st == "jacobi_legacy" || st == "jacobi_pipe") {
        if (loads[level] == 0) return fail(DCG_ERR_INVALID, "bench_stage: level %d has no active blocks", level);
        const bool saved = use_pipe;
        if (st != "jacobi") use_pipe = st == "jacobi_pipe";
        jacobi_sweep(level, (r & 1) ? tp : p, (r & 1) ? p : tp);
        launches--;
        use_pipe = saved;
        bytes = 12.0 * cl;
      } else if (st == "advect_velocity" || st == "advect_velocity_legacy") {
        if (st == "advect_velocity") launch_advect_pipe(0, vw[cur_v], vw[cur_v ^ 1], nullptr, nullptr);
        else k_dc_advect_velocity<<<blocks_for(M, kBPC), kCTA, 0, stream>>>(hot(), kp, vw[cur_v], vw[cur_v ^ 1]), launches++;
        launches--;
        cur_v ^= 1;
        bytes = 28.0 * call;
      } else if (st == "advect_density" || st == "advect_density_legacy") {
        if (st == "advect_density") launch_advect_pipe(1, vw[cur_v], nullptr, q[cur_q], q[cur_q ^ 1]);
        else k_dc_advect_density<<<blocks_for(M, kBPC), kCTA, 0, stream>>>(hot(), kp, vw[cur_v], fl, q[cur_q], q[cur_q ^ 1]), launches++;
        launches--;
        cur_q ^= 1;
        bytes = 24.0 * call;
      } else if (st == "advect_both") {  // density of this step + velocity of the next one (k_dc_advect_pipe<2>)
        launch_advect_pipe(2, vw[cur_v], vw[cur_v ^ 1], q[cur_q], q[cur_q ^ 1]);
        launches--;
        cur_q ^= 1;
        bytes = 52.0 * call;
      } else if (st == "divergence") {
        launch_divergence(skip_dead_zeroing && project_level_pairs >= 1 ? levels - 1 : 0);
        launches--;
        bytes = 28.0 * call;
      } else if (st == "apply_pressure") {
        launch_apply();
        launches--;
        bytes = 32.0 * call;
      } else if (st == "accumulate_velocity") {
        launch_pdl(k_dc_accumulate_velocity, dim3(blocks_for(8 * loads[level], 256)), dim3(256), 0, hot(), level, vw[cur_v], 0);
        bytes = 13.5 * cl;
      } else if (st == "prolongate") {
        launch_prolongate(level);
        launches--;
        bytes = 4.5 * cl;
      } else return fail(DCG_ERR_INVALID, "bench_stage: unknown stage %s", stage);
      launches++;
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

  // ---- stats / accessors --------------------------------------------------------------------------------
  int debug_stats(float *out) override {  // :517-528 (host sums the per-block partials in slot order)
    DCG_TRY(enter());
    k_dc_debug_stats<<<blocks_for(M, 256), 256, 0, stream>>>(hot(), kp, p, div, scratch);
    // per-block partials back in the reference's slot order: the host sum below is order dependent
    k_dc_gather_u32<<<blocks_for(M, 256), 256, 0, stream>>>(scratch, d_perm, scratch + M, M);
    launches += 2;
    std::vector<float> h(M);
    DCG_CUDA_TRY(cudaMemcpyAsync(h.data(), scratch + M, (size_t)M * 4, cudaMemcpyDeviceToHost, stream));
    DCG_TRY(synchronize());
    float sum = 0.f;
    for (uint32_t i = 0; i < M; i++) sum += h[i];
    *out = sum;
    return DCG_OK;
  }
  int total_density(double *out) override {
    DCG_TRY(enter());
    // sum over the cells this instance owns (sharded, one rank per process: the host adds the ranks' partials)
    const int blocks = (int)std::min<size_t>(1024, (cells + 255) / 256);
    double s = 0.0;
    for (int lr = 0; lr < nlocal; lr++) {
      k_dc_total_density<<<blocks, 256, 0, stream>>>(hot(), work[lr].all, q[cur_q], fl, d_partial);
      launches++;
      DCG_CUDA_TRY(cudaMemcpyAsync(h_partial, d_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, stream));
      DCG_TRY(synchronize());
      for (int i = 0; i < blocks; i++) s += h_partial[i];
    }
    *out = s;
    return DCG_OK;
  }

  uint64_t num_cells() const override { return cells; }
  uint64_t max_num_blocks() const override { return M; }
  int num_levels() const override { return levels; }
  int sparse_levels() const override { return sparse; }

  int get_field(int field, int layout, float *dst, uint64_t count) override {
    DCG_TRY(enter());
    const int comps = (field == DCG_FIELD_VELOCITY || field == DCG_FIELD_VORTICITY) ? 3 : 1;
    const float *src = nullptr;
    int stride = 1;
    switch (field) {
      case DCG_FIELD_DENSITY: src = q[cur_q]; break;
      case DCG_FIELD_VELOCITY: src = reinterpret_cast<const float *>(vw[cur_v]); stride = 4; break;
      case DCG_FIELD_FLUIDITY: src = fl; break;
      case DCG_FIELD_PRESSURE: src = p; break;
      case DCG_FIELD_DIVERGENCE: src = div; break;
      case DCG_FIELD_T_PRESSURE: src = tp; break;
      case DCG_FIELD_TEMPERATURE: src = th[cur_s]; break;
      case DCG_FIELD_VAPOR: src = qvp[cur_s]; break;
      case DCG_FIELD_VORTICITY: src = reinterpret_cast<const float *>(vort); stride = 4; break;
      default: return fail(DCG_ERR_INVALID, "get_field: unknown field %d", field);
    }
    if (!src) return fail(DCG_ERR_INVALID, "get_field: field %d belongs to an extension that is not switched on (dcg_set_ext_params)", field);
    if (layout == DCG_LAYOUT_DENSE_L0) {
      const size_t n = (size_t)gx * gy * gz;
      if (!dst || count < n * comps) return fail(DCG_ERR_INVALID, "get_field: destination too small");
      if (n * comps > scratch_floats) {  // 8.6 G cells at 2048^3: only allocated when somebody asks for it
        DCG_CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(scratch);
        scratch = nullptr;
        scratch_floats = 0;
        DCG_CUDA_TRY(cudaMalloc(&scratch, n * comps * 4));
        scratch_floats = n * comps;
      }
      k_dc_dense_l0<<<blocks_for(n, 256), 256, 0, stream>>>(hot(), kp, src, comps, stride, scratch);
      launches++;
      DCG_CUDA_TRY(cudaMemcpyAsync(dst, scratch, n * comps * 4, cudaMemcpyDeviceToHost, stream));
      return synchronize();
    }
    if (layout != DCG_LAYOUT_NATIVE) return fail(DCG_ERR_INVALID, "get_field: unknown layout %d", layout);
    if (!dst || count < cells * comps) return fail(DCG_ERR_INVALID, "get_field: destination too small");
    // fields are stored in field order (and, sharded, span every rank's memory): gather into the reference's slot
    // order in local memory first
    k_dc_unpermute_f32<<<blocks_for(cells, 256), 256, 0, stream>>>(src, stride, comps, d_perm, scratch, cells);
    launches++;
    src = scratch;
    DCG_CUDA_TRY(cudaMemcpyAsync(dst, src, cells * comps * 4, cudaMemcpyDeviceToHost, stream));
    return synchronize();
  }

  // field pointer (field order), components and float stride of an accessor field; nullptr = not available
  const float *field_src(int field, int &comps, int &stride) const {
    comps = 1; stride = 1;
    switch (field) {
      case DCG_FIELD_DENSITY: return q[cur_q];
      case DCG_FIELD_VELOCITY: comps = 3; stride = 4; return reinterpret_cast<const float *>(vw[cur_v]);
      case DCG_FIELD_FLUIDITY: return fl;
      case DCG_FIELD_PRESSURE: return p;
      case DCG_FIELD_DIVERGENCE: return div;
      case DCG_FIELD_T_PRESSURE: return tp;
      case DCG_FIELD_TEMPERATURE: return th[cur_s];
      case DCG_FIELD_VAPOR: return qvp[cur_s];
      case DCG_FIELD_VORTICITY: comps = 3; stride = 4; return reinterpret_cast<const float *>(vort);
    }
    return nullptr;
  }
  int sample_field(int field, int mode, const float *positions, uint64_t n, float *out) override {
    DCG_TRY(enter());
    int comps, stride;
    const float *src = field_src(field, comps, stride);
    if (!src) return fail(DCG_ERR_INVALID, "sample_field: field %d is unknown or belongs to an extension that is not switched on", field);
    if (n == 0) return DCG_OK;
    struct Tmp {
      void *p = nullptr;
      ~Tmp() { cudaFree(p); }
    } t_pos, t_out;
    DCG_CUDA_TRY(cudaMalloc(&t_pos.p, n * 12));
    DCG_CUDA_TRY(cudaMalloc(&t_out.p, n * comps * 4));
    DCG_CUDA_TRY(cudaMemcpyAsync(t_pos.p, positions, n * 12, cudaMemcpyHostToDevice, stream));
    ext::k_dc_ext_sample<<<blocks_for(n, 256), 256, 0, stream>>>(hot(), kp, src, comps, stride, mode, static_cast<const float *>(t_pos.p), n,
                                                                 static_cast<float *>(t_out.p));
    launches++;
    DCG_CUDA_TRY(cudaMemcpyAsync(out, t_out.p, n * comps * 4, cudaMemcpyDeviceToHost, stream));
    return synchronize();
  }

  // ---- state dump / load (format: DESIGN.md "State file") ------------------------------------------------------
  struct StateHeader {
    char magic[8];  // "DCGB200S"
    uint32_t version, header_bytes;
    int32_t gx, gy, gz, levels, sparse, has_scalars;
    uint64_t M;
    uint64_t counters[8];
    uint64_t loads[kMaxLevels], move_limit[kMaxLevels];
    dcg_sim_params params;
    dcg_ext_params ext;
    int32_t pairs[3];
  };
  int save_state(const char *path) override {
    DCG_TRY(enter());
    if (world > 1) return fail(DCG_ERR_UNSUPPORTED, "save_state: single-GPU instances only");
    DCG_TRY(synchronize());
    FILE *f = std::fopen(path, "wb");
    if (!f) return fail(DCG_ERR_INVALID, "save_state: cannot open %s", path);
    struct Closer {
      FILE *f;
      ~Closer() { std::fclose(f); }
    } closer{f};
    StateHeader H{};
    std::memcpy(H.magic, "DCGB200S", 8);
    H.version = 1; H.header_bytes = (uint32_t)sizeof H;
    H.gx = gx; H.gy = gy; H.gz = gz; H.levels = levels; H.sparse = sparse; H.has_scalars = th[0] ? 1 : 0;
    H.M = M;
    get_counters(H.counters);
    for (int l = 0; l < levels; l++) { H.loads[l] = loads[l]; H.move_limit[l] = move_limit[l]; }
    H.params = params; H.ext = ext;
    H.pairs[0] = project_coarsest_pairs; H.pairs[1] = project_level_pairs; H.pairs[2] = local_pairs;
    bool ok = std::fwrite(&H, sizeof H, 1, f) == 1;
    std::vector<unsigned char> h;
    auto put_dev = [&](const void *dev, size_t bytes) {
      h.resize(bytes);
      if (cudaMemcpyAsync(h.data(), dev, bytes, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) { ok = false; return; }
      ok = ok && std::fwrite(h.data(), 1, bytes, f) == bytes;
    };
    // block pool, reference numbering
    put_dev(T.posl, (size_t)M * sizeof(int4));
    put_dev(T.parent, (size_t)M * 4);
    put_dev(T.child, (size_t)M * 8 * 4);
    put_dev(T.apron, (size_t)M * kAV * 4);
    // fields, reference slot order (through the accessor path)
    const int fields[] = {DCG_FIELD_VELOCITY, DCG_FIELD_DENSITY, DCG_FIELD_FLUIDITY, DCG_FIELD_PRESSURE, DCG_FIELD_T_PRESSURE, DCG_FIELD_DIVERGENCE,
                          DCG_FIELD_TEMPERATURE, DCG_FIELD_VAPOR};
    for (int fi = 0; fi < (H.has_scalars ? 8 : 6) && ok; fi++) {
      int comps, stride;
      const float *src = field_src(fields[fi], comps, stride);
      k_dc_unpermute_f32<<<blocks_for(cells, 256), 256, 0, stream>>>(src, stride, comps, d_perm, scratch, cells);
      launches++;
      if (cudaStreamSynchronize(stream) != cudaSuccess) { ok = false; break; }
      put_dev(scratch, cells * comps * 4);
    }
    if (!ok) return fail(DCG_ERR_CUDA, "save_state: write to %s failed", path);
    return DCG_OK;
  }
  int load_state(const char *path) override {
    DCG_TRY(enter());
    if (world > 1) return fail(DCG_ERR_UNSUPPORTED, "load_state: single-GPU instances only");
    FILE *f = std::fopen(path, "rb");
    if (!f) return fail(DCG_ERR_INVALID, "load_state: cannot open %s", path);
    struct Closer {
      FILE *f;
      ~Closer() { std::fclose(f); }
    } closer{f};
    StateHeader H{};
    if (std::fread(&H, sizeof H, 1, f) != 1 || std::memcmp(H.magic, "DCGB200S", 8) != 0 || H.version != 1 || H.header_bytes != sizeof H)
      return fail(DCG_ERR_INVALID, "load_state: %s is not a dcgrid_b200 state file of this version", path);
    if (H.gx != gx || H.gy != gy || H.gz != gz || H.M != M || H.levels != levels || H.sparse != sparse)
      return fail(DCG_ERR_INVALID, "load_state: the file holds a %dx%dx%d grid with %llu blocks, this instance %dx%dx%d with %u", H.gx, H.gy, H.gz,
                  (unsigned long long)H.M, gx, gy, gz, M);
    DCG_TRY(synchronize());
    drop_graphs();
    params = H.params;
    ext = H.ext;
    kp = make_kparams(params, ext);
    DCG_TRY(ensure_ext_storage());
    if (H.has_scalars && !th[0]) return fail(DCG_ERR_INVALID, "load_state: the file carries temperature / vapor but its parameters do not enable them");
    project_coarsest_pairs = H.pairs[0]; project_level_pairs = H.pairs[1]; local_pairs = H.pairs[2];
    for (int l = 0; l < levels; l++) { loads[l] = H.loads[l]; move_limit[l] = H.move_limit[l]; }
    n_adapt = H.counters[0]; n_changed = H.counters[1]; n_moved = H.counters[2]; n_refined = H.counters[3]; n_skipped = H.counters[4];
    n_failed = H.counters[5];
    sync_loads();
    std::vector<unsigned char> h;
    bool ok = true;
    auto get_dev = [&](void *dev, size_t bytes) {
      h.resize(bytes);
      if (std::fread(h.data(), 1, bytes, f) != bytes) { ok = false; return; }
      // on the instance's stream: a blocking cudaMemcpy from pageable memory may return before its DMA has landed, and
      // the non-blocking stream the kernels run on is not ordered behind it
      ok = ok && cudaMemcpyAsync(dev, h.data(), bytes, cudaMemcpyHostToDevice, stream) == cudaSuccess && cudaStreamSynchronize(stream) == cudaSuccess;
    };
    get_dev(T.posl, (size_t)M * sizeof(int4));
    get_dev(T.parent, (size_t)M * 4);
    get_dev(T.child, (size_t)M * 8 * 4);
    get_dev(T.apron, (size_t)M * kAV * 4);
    if (!ok) return fail(DCG_ERR_INVALID, "load_state: %s is truncated", path);
    // field order = slot order until the next re-sort; level maps rebuilt from the positions
    mirrored = false;
    changes_since_resort = 1;
    steady = false;
    early_ready = false;
    spec_velocity = false;
    cur_v = cur_q = cur_s = 0;
    k_iota_u32<<<blocks_for(M, 256), 256, 0, stream>>>(d_perm, M);
    DCG_CUDA_TRY(cudaMemsetAsync(d_flags, 0, (size_t)M * 4, stream));
    for (int l = 0; l < sparse; l++) DCG_CUDA_TRY(cudaMemsetAsync(T.map[l], 0xff, map_size[l] * 4, stream));
    ext::k_dc_ext_rebuild_maps<<<blocks_for(M, 256), 256, 0, stream>>>(T, kp);
    launches += 2;
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    // velocity (3 floats / cell) and fluidity arrive separately and are packed; scratch holds 3 * cells floats
    get_dev(scratch, cells * 3 * 4);
    get_dev(q[0], cells * 4);
    get_dev(fl, cells * 4);
    if (!ok) return fail(DCG_ERR_INVALID, "load_state: %s is truncated", path);
    ext::k_dc_ext_pack_vw<<<blocks_for(cells, 256), 256, 0, stream>>>(scratch, fl, vw[0], vw[1], cells);
    launches++;
    DCG_CUDA_TRY(cudaMemcpyAsync(q[1], q[0], cells * 4, cudaMemcpyDeviceToDevice, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    get_dev(p, cells * 4);
    get_dev(tp, cells * 4);
    get_dev(div, cells * 4);
    if (H.has_scalars) {
      get_dev(th[0], cells * 4);
      get_dev(qvp[0], cells * 4);
      if (ok) {
        DCG_CUDA_TRY(cudaMemcpyAsync(th[1], th[0], cells * 4, cudaMemcpyDeviceToDevice, stream));
        DCG_CUDA_TRY(cudaMemcpyAsync(qvp[1], qvp[0], cells * 4, cudaMemcpyDeviceToDevice, stream));
      }
    }
    if (!ok) return fail(DCG_ERR_INVALID, "load_state: %s is truncated", path);
    DCG_TRY(build_face_descriptors());
    DCG_CUDA_TRY(cudaGetLastError());
    return synchronize();
  }

  int get_level_table(uint64_t *mx, uint64_t *full, uint64_t *ld, uint64_t *offs) override {
    for (int l = 0; l < levels; l++) {
      if (mx) mx[l] = max_blocks[l];
      if (full) full[l] = full_blocks[l];
      if (ld) ld[l] = loads[l];
      if (offs) offs[l] = offsets[l];
    }
    return DCG_OK;
  }

  int get_topology(int32_t *positions, uint8_t *lv, uint64_t *parent, uint64_t *children, uint64_t *apron) override {
    DCG_TRY(enter());
    DCG_TRY(synchronize());
    auto widen = [](const std::vector<uint32_t> &in, uint64_t *out) {
      for (size_t i = 0; i < in.size(); i++) out[i] = in[i] == kNone ? UINT64_MAX : (uint64_t)in[i];
    };
    if (positions || lv) {
      std::vector<int4> h(M);
      DCG_CUDA_TRY(cudaMemcpy(h.data(), T.posl, (size_t)M * sizeof(int4), cudaMemcpyDeviceToHost));
      for (uint32_t i = 0; i < M; i++) {
        if (positions) { positions[3 * i] = h[i].x; positions[3 * i + 1] = h[i].y; positions[3 * i + 2] = h[i].z; }
        if (lv) lv[i] = (uint8_t)h[i].w;
      }
    }
    if (parent) {
      std::vector<uint32_t> h(M);
      DCG_CUDA_TRY(cudaMemcpy(h.data(), T.parent, (size_t)M * 4, cudaMemcpyDeviceToHost));
      widen(h, parent);
    }
    if (children) {
      std::vector<uint32_t> h((size_t)M * 8);
      DCG_CUDA_TRY(cudaMemcpy(h.data(), T.child, (size_t)M * 8 * 4, cudaMemcpyDeviceToHost));
      widen(h, children);
    }
    if (apron) {
      std::vector<uint32_t> h((size_t)M * kAV);
      DCG_CUDA_TRY(cudaMemcpy(h.data(), T.apron, (size_t)M * kAV * 4, cudaMemcpyDeviceToHost));
      widen(h, apron);
    }
    return DCG_OK;
  }

  int lookup_blocks(const int32_t *positions, uint64_t n, uint64_t *out_slot, uint8_t *out_level) override {
    DCG_TRY(enter());
    if (n == 0) return DCG_OK;
    struct Tmp {  // freed on every exit path
      void *p = nullptr;
      ~Tmp() { cudaFree(p); }
    } t_pos, t_slot, t_lvl;
    DCG_CUDA_TRY(cudaMalloc(&t_pos.p, n * 12));
    DCG_CUDA_TRY(cudaMalloc(&t_slot.p, n * 4));
    DCG_CUDA_TRY(cudaMalloc(&t_lvl.p, n));
    int *d_pos = static_cast<int *>(t_pos.p);
    uint32_t *d_slot = static_cast<uint32_t *>(t_slot.p);
    uint8_t *d_lvl = static_cast<uint8_t *>(t_lvl.p);
    DCG_CUDA_TRY(cudaMemcpyAsync(d_pos, positions, n * 12, cudaMemcpyHostToDevice, stream));
    k_dc_lookup<<<blocks_for(n, 256), 256, 0, stream>>>(T, kp, d_pos, n, d_slot, d_lvl);
    launches++;
    std::vector<uint32_t> h(n);
    DCG_CUDA_TRY(cudaMemcpyAsync(h.data(), d_slot, n * 4, cudaMemcpyDeviceToHost, stream));
    DCG_CUDA_TRY(cudaMemcpyAsync(out_level, d_lvl, n, cudaMemcpyDeviceToHost, stream));
    DCG_CUDA_TRY(cudaStreamSynchronize(stream));
    for (uint64_t i = 0; i < n; i++) out_slot[i] = h[i] == kNone ? UINT64_MAX : h[i];
    return DCG_OK;
  }

  int get_counters(uint64_t out[8]) override {
    out[0] = n_adapt; out[1] = n_changed; out[2] = n_moved; out[3] = n_refined;
    out[4] = n_skipped; out[5] = n_failed; out[6] = launches; out[7] = steady ? 1 : 0;
    return DCG_OK;
  }

  int get_info(const char *key, double *out) override {
    const std::string k(key);
    if (k == "pdl") *out = use_pdl ? 1 : 0;
    else if (k == "pdl_fallbacks") *out = (double)pdl_fallbacks;
    else if (k == "host_selections") *out = (double)n_host_selections;
    else if (k == "device_selections") *out = (double)n_device_selections;
    else if (k == "selection_fallbacks") *out = (double)n_selection_fallbacks;
    else if (k == "levels_shortcut") *out = (double)n_levels_shortcut;
    else if (k == "resorts") *out = (double)n_resorts;
    else if (k == "irregular_blocks") *out = (double)n_irregular;
    else if (k == "barriers") *out = (double)n_barriers;
    else if (k == "graph_launches_per_step") *out = (double)step_graph_launches;
    else if (k == "adapt_move_ms") *out = t_move_ms;
    else if (k == "adapt_refine_ms") *out = t_refine_ms;
    else if (k == "adapt_apron_ms") *out = t_apron_ms;
    else if (k == "adapt_layout_ms") *out = t_layout_ms;
    else if (k == "adapt_propagate_ms") *out = t_propagate_ms;
    else if (k == "early_scores") *out = (double)n_early_scores;
    else if (k == "select_scores_ms") *out = t_sel_scores_ms;
    else if (k == "select_d2h_ms") *out = t_sel_d2h_ms;
    else if (k == "select_host_ms") *out = t_sel_host_ms;
    else if (k == "timing_sync") { timing_sync = !timing_sync; *out = timing_sync ? 1 : 0; }
    else return dcg_sim::get_info(key, out);
    return DCG_OK;
  }

  // C*(28+28+12*sweeps+32+24) + (C - C_top)*(13.5+4.5+4.5+13.5+4.5), C = 64*active blocks,
  // sweeps = 2*pairs per level (10 by default => 120 B).
  int algorithmic_bytes(double *bytes, uint64_t *active_blocks) override {
    uint64_t active = 0;
    double b = 0.0;
    for (int l = 0; l < levels; l++) {
      const double c = 64.0 * (double)loads[l];
      const int pairs = (l == levels - 1) ? project_coarsest_pairs : project_level_pairs;
      b += c * (28.0 + 28.0 + 12.0 * 2 * pairs + 32.0 + 24.0);
      if (l < levels - 1) b += c * (13.5 + 4.5 + 4.5 + 13.5 + 4.5);
      active += loads[l];
    }
    if (bytes) *bytes = b;
    if (active_blocks) *active_blocks = active;
    return DCG_OK;
  }
};

}  // namespace
}  // namespace dcg

dcg_sim *dcg_make_dcgrid(uint64_t max_num_blocks) { return new dcg::DCGridSim(max_num_blocks); }
dcg_sim *dcg_make_dcgrid_sharded(uint64_t max_num_blocks, int rank, int world, int nlocal) {
  auto *s = new dcg::DCGridSim(max_num_blocks);
  s->world = world; s->rank0 = rank; s->nlocal = nlocal;
  return s;
}
