/*
 * dcgrid_b200 — C ABI of the B200-native per-timestep fluid solve.
 *
 * This is the drop-in boundary for the reference's solver interface
 *   class FluidSimulation            (reference src/fluid_simulation.h:4-27)
 * and its two implementations
 *   FluidSimulationUniform           (src/uniformgrid/fluid_simulation_uniform.h:5-39)
 *   FluidSimulationDCGrid            (src/dcgrid/fluid_simulation_dcgrid.h:5-61)
 * One opaque handle == one FluidSimulation instance.  Every reference virtual has
 * one entry point here; `dcg_step` is the 4-call sequence the reference's UI layer
 * issues once per frame (src/simulation.cpp:104-111).  All calls return a status
 * (the reference prints and exit()s instead, include/cuda/helper_cuda.h:771-781).
 *
 * No torch types, plain pointers and sizes only.  A maintainer binds this from the
 * reference tree with the adapter class in include/fluid_simulation_b200.h
 * (see INTEGRATION.md).
 */
#ifndef DCGRID_B200_H
#define DCGRID_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCG_API __attribute__((visibility("default")))

/* ---- status codes ------------------------------------------------------ */
enum {
  DCG_OK = 0,
  DCG_ERR_INVALID = 1,   /* bad argument / null handle                               */
  DCG_ERR_CUDA = 2,      /* a CUDA runtime call failed; see dcg_last_error           */
  DCG_ERR_POOL = 3,      /* block pool cannot fill the coarsest level / reach level 0
                            (reference: printf + exit(1),
                            src/dcgrid/fluid_simulation_dcgrid.cu:35-39,50-53)       */
  DCG_ERR_UNSUPPORTED = 4,
  DCG_ERR_NO_DEVICE = 5  /* no CUDA device: there is NO CPU fallback                 */
};

/* ---- SimParams ---------------------------------------------------------
 * Layout-identical (field order, types, 100 bytes) to the reference's
 * `struct SimParams` (src/data/sim_params.h:14-47), so a caller can pass its own
 * struct by pointer.  Only the first ten fields influence the solve; the render
 * fields are carried for layout compatibility (rendering is out of scope).
 * NOTE: as in the reference, `rdx` is NOT derived from `dx` (src/main.cpp:17 sets
 * it); dcg_default_params() sets both.                                            */
typedef struct dcg_float3 { float x, y, z; } dcg_float3;

typedef struct dcg_sim_params {
  float dt;                       /* timestep                                        */
  int gx, gy, gz;                 /* effective (level-0) grid size                   */
  float dx;                       /* cell size                                       */
  float rdx;                      /* 1/dx, set by the caller                         */
  float velocity_emission_rate;   /* inlet velocity                                  */
  float density_emission_rate;    /* inlet density                                   */
  float emission_radius;          /* inlet disc radius (world units)                 */
  bool enable_additional_solids;  /* sphere SDF obstacle (src/sdf.cuh:8-20)          */
  bool render_solids, render_shadows, render_precise;
  int32_t render_channel;         /* enum class RenderChannel                        */
  float aa_samples;
  float ambient;
  dcg_float3 background_color, floor_color, smoke_color, scene_color;
} dcg_sim_params;

typedef struct dcg_sim dcg_sim;   /* opaque: one FluidSimulation instance            */

/* ---- creation-time options ----------------------------------------------
 * Kernel variants, re-sort cadence and sharding granularity of ONE instance, fixed when it is created
 * (the reference has compile-time constants only).  Every field: 0 = default, so a zeroed struct (or a
 * NULL pointer) is the default configuration.  Every variant computes bit-identical results
 * (tests/test_dcgrid_gpu.py runs the combinations against the oracle); they exist for testing small
 * cases on the code paths big cases take, and for A/B measurements.                              */
typedef struct dcg_options {
  uint32_t struct_size;        /* sizeof(dcg_options) of the caller; 0 = this header's                */
  int32_t jacobi;              /* 0 ring kernel (8 cells/thread) on levels that fill the GPU, one CTA
                                  per tile below; 1 one CTA per tile everywhere; 2 ring on every level;
                                  3 / 4 = 0 / 2 with the 4-cells-per-thread ring kernel                 */
  int32_t jacobi_ctas_per_sm;  /* resident CTAs per SM of the ring kernel (0 = occupancy limit)         */
  int32_t no_snake;            /* 1: sweeps do not alternate direction                                  */
  int32_t advect;              /* 0 persistent ring kernel, 1 one CTA per 4 blocks                       */
  int32_t advect_slot_order;   /* 1: process blocks in pool-slot order instead of Morton order          */
  int32_t advect_no_fuse;      /* 1: density(n) and velocity(n+1) advection as separate passes          */
  int32_t advect_min_blocks;   /* __launch_bounds__ variant: 2, 3 (default) or 4 CTAs per SM            */
  int32_t advect_ctas_per_sm;  /* resident CTAs per SM (0 = occupancy limit)                            */
  int32_t stencil;             /* divergence / gradient / prolongation: 0 ring + staged, 1 one CTA/tile */
  int32_t stencil_ctas_per_sm;
  int32_t apply_min_blocks;    /* 2 (default) or 3                                                      */
  int32_t coarse_in_gmem;      /* 1: coarse cascade works in global memory instead of shared            */
  int32_t zero_all;            /* 1: clear pressure / t_pressure of every level like the reference      */
  int32_t no_resort;           /* 1: field order = the reference's slot order throughout                */
  int32_t resort_every;        /* topology changes between re-sorts: 0 = 32, -1 = only at the fixed point */
  uint32_t shard_unit;         /* sharded instances: slots per ownership unit (0 = 8192 / one tile)     */
  int32_t no_pdl;              /* 1: plain stream order between the kernels of a step (no programmatic
                                  dependent launch)                                                     */
  int32_t host_selection;      /* 1: adaptTopology always takes the reference's host selection
                                  (std::nth_element / std::sort on scores copied D2H)                    */
  int32_t jacobi_max_ctas;     /* > 0: cap on the resident CTAs of the ring kernel (tests: few CTAs walk many
                                  tiles each, the regime of the big scenes, on a small one)               */
  int32_t experiment;          /* variant bits for A/B measurements and tests; 0 = the shipped choice.  1: the other
                                  prolongation kernel (by parent block on one GPU, per child block sharded); 8: the
                                  other restriction schedule (one walk up the block tree on one GPU, per-level list
                                  passes sharded); 32: no early scores (adaptTopology computes its scores on entry) */
  int32_t reserved[12];
} dcg_options;
DCG_API int dcg_default_options(dcg_options *out);


/* ---- extensions beyond the reference snapshot ------------------------------------
 * BASELINE.json's north_star names features the snapshot under /root/reference does not contain:
 * a flow-driven refinement score, MacCormack advection, temperature / vapor fields with buoyancy, vorticity
 * confinement and condensation source terms, a terrain SDF.  They are specified by the CPU oracle
 * (oracle/dcgrid_oracle.cpp, "extensions") — the ONLY pin there is: "parity unpinned" against the reference — and
 * switched on per instance through this struct.  All zero (the default) = the reference snapshot's behaviour, bit for
 * bit.  Arithmetic uses + - * / sqrtf floorf fminf fmaxf only, in a fixed order, so CPU spec and CUDA agree bitwise.
 * DCGrid instances only (the uniform solver ignores everything but `terrain`).                                   */
typedef struct dcg_ext_params {
  uint32_t struct_size;          /* sizeof(dcg_ext_params) of the caller; 0 = this header's                  */
  int32_t score_mode;            /* 0: geometric score of the snapshot (dcgrid_adaptation.cu:26-34);
                                    1: flow-driven = sum over the subblock's 8 cells of |vorticity|, the code the
                                       reference carries commented out (dcgrid_adaptation.cu:6-8,36-39) fed by
                                       k_dcgrid_calc_vorticity (dcgrid_fluid.cu:146-172)                      */
  int32_t advection;             /* 0: semi-Lagrangian (reference); 1: MacCormack (forward SL, backward SL of the
                                       result, half the difference added back, clamped to the 8 corners of the
                                       forward sample; velocity, density, temperature and vapor)              */
  int32_t sources;               /* 1: temperature / vapor are advected with the density and dcg_apply_sources()
                                       (called by dcg_step between adaptTopology and project) adds buoyancy,
                                       vorticity confinement and condensation in ONE fused pass               */
  int32_t terrain;               /* 1: solids = height-field terrain instead of the sphere (src/sdf.cuh:20);
                                       needs SimParams.enable_additional_solids                               */
  int32_t selection;             /* move / refine selection of adaptTopology.
                                    0: the reference's, verbatim (std::nth_element / std::sort with its comparators on
                                       the host, fluid_simulation_dcgrid.cu:348-483: the outcome depends on libstdc++'s
                                       treatment of ties and of a comparator that is not a strict weak order, so only the
                                       same algorithm on the same sequence reproduces it);
                                    1: a TOTAL order, evaluated entirely on the device: blocks by (score ascending, slot
                                       ascending), destinations by (score descending, subblock id ascending), negative
                                       scores excluded, the parents of ALL matched destinations protected, refinement
                                       candidates in ascending id order keeping the largest ids.  Same greedy rule and
                                       limits; a different, well-defined topology evolution                       */
  float buoyancy;                /* m/s^2 per unit of (theta - theta_ambient(h)) / ambient_temperature       */
  float vapor_buoyancy;          /* m/s^2 per unit of vapor mixing ratio                                     */
  float smoke_weight;            /* m/s^2 per unit of density (condensed water / smoke loading)              */
  float ambient_temperature;     /* potential temperature at the floor, K                                    */
  float ambient_lapse;           /* d(theta_ambient)/dh, K per world unit                                    */
  float adiabatic_lapse;         /* T = theta - adiabatic_lapse * h, K per world unit                        */
  float vorticity_confinement;   /* epsilon of the confinement force eps * dx * (N x omega)                  */
  float saturation_base;         /* qs(T) = max(0, saturation_base + saturation_slope * (T - ambient_temperature)) */
  float saturation_slope;
  float condensation_rate;       /* fraction of (qv - qs) exchanged per step, 0..1                           */
  float latent_heat;             /* K of theta per unit of condensed vapor                                   */
  float temperature_emission;    /* inlet: theta = theta_ambient(0) + temperature_emission                   */
  float vapor_emission;          /* inlet: vapor mixing ratio                                                */
  float ambient_vapor;           /* initial / far-field vapor mixing ratio                                   */
  float terrain_height;          /* peak height of the terrain, level-0 cells                                */
  float terrain_wavelength;      /* hill spacing, level-0 cells                                              */
  int32_t reserved[10];
} dcg_ext_params;
DCG_API int dcg_default_ext_params(dcg_ext_params *out);   /* all switches off, plausible coefficients       */
/* Takes effect at the next call; changing `terrain`, `sources` or the ambient profile re-initialises nothing by
 * itself: call dcg_reset() afterwards for a consistent start.                                                 */
DCG_API int dcg_set_ext_params(dcg_sim *sim, const dcg_ext_params *ext);
DCG_API int dcg_get_ext_params(const dcg_sim *sim, dcg_ext_params *out);
/* The fused source pass (sources == 1), on leaf cells, followed by the restriction of the touched fields. */
DCG_API int dcg_apply_sources(dcg_sim *sim);

/* Point samples of a field at `n` positions (3*n floats, level-0 cell units): mode 0 = value of the finest
 * covering cell (sampleCoarse, src/dcgrid/dcgrid_rendering.cu:6-24), mode 1 = trilinear through the covering
 * block's apron map (samplePrecise, :26-58, interpolate() of src/raymarching.cuh:26-40).  `out`: n floats
 * (3*n for the velocity).  Uniform instances: mode 0 only.                                                  */
DCG_API int dcg_sample_field(dcg_sim *sim, int field, int mode, const float *positions, uint64_t n, float *out);

/* State dump / load: the whole simulation state (block pool, level tables, move limits, every field, the
 * extension parameters) in the reference's slot order, little-endian, format documented in DESIGN.md.  A run
 * continued from a loaded state is bit-identical to the uninterrupted run.  The target of dcg_load_state must
 * have been created with the same grid size and pool size.                                                  */
DCG_API int dcg_save_state(dcg_sim *sim, const char *path);
DCG_API int dcg_load_state(dcg_sim *sim, const char *path);

/* ---- field / layout selectors for the accessors ------------------------- */
enum {
  DCG_FIELD_DENSITY = 0,   /* 1 float / cell                                         */
  DCG_FIELD_VELOCITY = 1,  /* 3 floats / cell (x,y,z interleaved, like float3[])     */
  DCG_FIELD_FLUIDITY = 2,  /* 1 float / cell (uniform: level-0 part of the pyramid)  */
  DCG_FIELD_PRESSURE = 3,  /* 1 float / cell; valid between project() and the next
                              advectVelocity() (buffers alias in the reference,
                              src/dcgrid/dcgrid_structure.cu:94-102)                 */
  DCG_FIELD_DIVERGENCE = 4,/* same validity window as pressure                       */
  DCG_FIELD_T_PRESSURE = 5,/* the penultimate Jacobi iterate                         */
  DCG_FIELD_TEMPERATURE = 6,/* extension (dcg_ext_params.sources): potential temperature, 1 float / cell */
  DCG_FIELD_VAPOR = 7,     /* extension: vapor mixing ratio, 1 float / cell          */
  DCG_FIELD_VORTICITY = 8  /* extension: 3 floats / cell, as of the last adaptTopology() / dcg_apply_sources() */
};
enum {
  DCG_LAYOUT_NATIVE = 0,   /* reference memory order: uniform idx=(z*gy+y)*gx+x
                              (src/utils/grid_math.cuh:10); DCGrid cell id =
                              64*slot + in-block bits (src/dcgrid/dcgrid_utils.cuh:31-81),
                              numCells = 64*maxNumBlocks entries                     */
  DCG_LAYOUT_DENSE_L0 = 1  /* DCGrid only: finest covering cell resampled onto the
                              gx*gy*gz level-0 grid, uniform memory order            */
};

/* ---- lifetime ----------------------------------------------------------- */
/* SimParams::defaultParams() (src/data/sim_params.cpp:4-34) plus rdx = 1/dx. */
DCG_API int dcg_default_params(dcg_sim_params *out);

/* new FluidSimulationUniform(size): allocates and reset()s
 * (src/uniformgrid/fluid_simulation_uniform.cu:6-57).  `device` = CUDA ordinal. */
DCG_API int dcg_create_uniform(const dcg_sim_params *params, int device, dcg_sim **out);

/* new FluidSimulationDCGrid(size, maxNumBlocks): level/pool sizing, allocation,
 * reset() incl. its 5 adaptTopology passes
 * (src/dcgrid/fluid_simulation_dcgrid.cu:9-140,190-261).                          */
DCG_API int dcg_create_dcgrid(const dcg_sim_params *params, uint64_t max_num_blocks,
                              int device, dcg_sim **out);

/* the same with creation-time options (NULL = defaults) */
DCG_API int dcg_create_uniform_opt(const dcg_sim_params *params, int device, const dcg_options *opt, dcg_sim **out);
DCG_API int dcg_create_dcgrid_opt(const dcg_sim_params *params, uint64_t max_num_blocks, int device,
                                  const dcg_options *opt, dcg_sim **out);

DCG_API int dcg_destroy(dcg_sim *sim);

/* copySimParamsToDevice (src/utils/sim_utils.cu:7-9); per-instance here. */
DCG_API int dcg_set_params(dcg_sim *sim, const dcg_sim_params *params);
DCG_API int dcg_get_params(const dcg_sim *sim, dcg_sim_params *out);

/* ---- the 9 virtuals of FluidSimulation (render is a stub) --------------- */
DCG_API int dcg_init(dcg_sim *sim);             /* fluid_simulation.h:9   */
DCG_API int dcg_reset(dcg_sim *sim);            /* fluid_simulation.h:10  */
DCG_API int dcg_adapt_topology(dcg_sim *sim);   /* fluid_simulation.h:11  */
DCG_API int dcg_advect_velocity(dcg_sim *sim);  /* fluid_simulation.h:13  */
DCG_API int dcg_project(dcg_sim *sim);          /* fluid_simulation.h:14  */
DCG_API int dcg_project_local(dcg_sim *sim);    /* fluid_simulation.h:15  */
DCG_API int dcg_advect_density(dcg_sim *sim);   /* fluid_simulation.h:16  */
DCG_API int dcg_render(dcg_sim *sim);           /* fluid_simulation.h:18-20: out of
                                                   scope, returns DCG_ERR_UNSUPPORTED */
/* debugStats (fluid_simulation.h:22): uniform = sum(density*fluidity)
 * (uniformgrid_structure.cu:33-43, fluid_simulation_uniform.cu:160-176); DCGrid = L1
 * pressure residual over leaf cells (dcgrid_structure.cu:224-251,
 * fluid_simulation_dcgrid.cu:517-528).  Same summation order as the reference.    */
DCG_API int dcg_debug_stats(dcg_sim *sim, float *out);

/* ---- additions the reference lacks -------------------------------------- */
/* n x { advectVelocity; adaptTopology; project; advectDensity }
 * (src/simulation.cpp:104-111).  Asynchronous; dcg_synchronize() waits.          */
DCG_API int dcg_step(dcg_sim *sim, int n);
DCG_API int dcg_synchronize(dcg_sim *sim);

/* Jacobi schedule (pairs = jacobi + jacobi_inv).  Defaults are the reference's
 * compile-time constants: uniform project 2 coarsest + 1/level, projectLocal 5
 * (fluid_simulation_uniform.cu:103-121,129-132); DCGrid project 5/level,
 * projectLocal 10/level (fluid_simulation_dcgrid.cu:274-289,300-307).            */
DCG_API int dcg_set_jacobi_schedule(dcg_sim *sim, int project_coarsest_pairs,
                                    int project_level_pairs, int local_pairs);

/* sum over all cells of density*fluidity reduced on the device (8 bytes D2H). */
DCG_API int dcg_total_density(dcg_sim *sim, double *out);

/* ---- observable state ---------------------------------------------------- */
DCG_API int dcg_is_dcgrid(const dcg_sim *sim);
DCG_API uint64_t dcg_num_cells(const dcg_sim *sim);        /* N or 64*maxNumBlocks   */
DCG_API uint64_t dcg_max_num_blocks(const dcg_sim *sim);   /* 0 for uniform          */
DCG_API int dcg_num_levels(const dcg_sim *sim);            /* mipmapLevels / levels  */
DCG_API int dcg_sparse_levels(const dcg_sim *sim);

/* Copies a field to HOST memory; `count` = number of floats `dst` can hold. */
DCG_API int dcg_get_field(dcg_sim *sim, int field, int layout, float *dst, uint64_t count);

/* Per-level tables, each `levels` entries (any pointer may be NULL):
 * maxNumBlocksLevel, fullBlocksLevel, blockLoads, levelOffsets
 * (fluid_simulation_dcgrid.cu:24-58,232-241).                                     */
DCG_API int dcg_get_level_table(dcg_sim *sim, uint64_t *max_blocks, uint64_t *full_blocks,
                                uint64_t *block_loads, uint64_t *level_offsets);

/* Block pool in the reference's own types/meaning (struct DCGrid, dcgrid.h:5-43);
 * any pointer may be NULL.  positions: 3*M ints; levels: M (0xFF = free slot);
 * parent: M (8*parentSlot+subblock, UINT64_MAX = none); children: 8*M;
 * apron: 216*M cell ids (z fastest).                                               */
DCG_API int dcg_get_topology(dcg_sim *sim, int32_t *positions, uint8_t *levels,
                             uint64_t *parent, uint64_t *children, uint64_t *apron);

/* Finest block covering a level-0 cell: getBlockIndexDeep(pos, 0)
 * (dcgrid_utils.cuh:201-233), evaluated for `n` positions (3*n ints) on the
 * device.  out_slot[i] = pool slot, out_level[i] = its level.                     */
DCG_API int dcg_lookup_blocks(dcg_sim *sim, const int32_t *positions, uint64_t n,
                              uint64_t *out_slot, uint8_t *out_level);

/* Counters: [0] adaptTopology calls, [1] calls that changed the topology,
 * [2] blocks moved, [3] subblocks refined, [4] calls skipped at a proven fixed
 * point, [5] failed allocations (pool exhausted; the reference device-printf's,
 * dcgrid_utils.cuh:120-124), [6] kernel launches issued so far, [7] reserved.     */
DCG_API int dcg_get_counters(dcg_sim *sim, uint64_t out[8]);

/* Named diagnostics (what the instance did, not part of the result): "pdl" (1 = kernels of a step are chained by
 * programmatic dependent launch), "pdl_fallbacks", "host_selections" (levels that took the reference's host
 * selection in adaptTopology), "device_selections", "selection_fallbacks" (levels where the device selection found a
 * tie group straddling the cut and handed over to the host selection), "levels_shortcut", "resorts",
 * "irregular_blocks", "barriers", "graph_launches_per_step".  Unknown key: DCG_ERR_INVALID.                  */
DCG_API int dcg_get_info(dcg_sim *sim, const char *key, double *out);

/* Device time (ms, CUDA events on the instance's stream) of the last dcg_step
 * batch; valid after dcg_synchronize().                                            */
DCG_API int dcg_last_step_ms(dcg_sim *sim, float *out);

/* Algorithmic field bytes of one full step at the current topology
 * (SURVEY.md §8d / DESIGN.md formulas) and active-block count per level.           */
DCG_API int dcg_algorithmic_bytes(dcg_sim *sim, double *bytes_per_step,
                                  uint64_t *active_blocks_total);

/* Measurement hook (bench.py's roofline leg): launches ONE stage kernel `reps` times back to back,
 * CUDA-event timed on the instance's stream.  stage = "jacobi" | "advect_velocity" |
 * "advect_density" | "divergence" | "apply_pressure" | "accumulate_velocity" | "prolongate";
 * `level` selects the level for per-level stages.  Returns the mean device time per launch and the
 * stage's algorithmic bytes per launch (SURVEY.md §8d per-cell figure x cells it processes).
 * Scribbles over the pressure / ping-pong buffers: call it after the timed steps.          */
DCG_API int dcg_bench_stage(dcg_sim *sim, const char *stage, int level, int reps,
                            float *ms_per_launch, double *alg_bytes_per_launch);

/* ---- multi-GPU: z-slab sharding of the uniform grid -------------------------------------------
 * The reference is single-GPU (src/main.cpp:46-48); these entry points have no counterpart there.
 * One instance holds `nlocal` consecutive ranks [rank, rank+nlocal) of a `world`-rank decomposition of
 * the gx*gy*gz grid into equal z-slabs (gz % world == 0).  nlocal == 1: one rank per process/GPU
 * (torchrun); the ranks exchange dcg_shard_export_handle() blobs (dcg_shard_handle_bytes() each,
 * rank order) through any host channel and pass the concatenation to dcg_shard_import_handles(),
 * which maps every peer's slab over NVLink (CUDA IPC) and resets.  nlocal == world: all ranks live in
 * this instance on one device (used to test the decomposition on a single GPU); ready after create.
 * All ranks must then issue the same sequence of solver calls (lock-step flag barriers over peer
 * memory).  Accessors return the LOCAL slab(s); dcg_total_density / dcg_debug_stats return local
 * partial sums to be combined by the host (sum over ranks, e.g. an allreduce).
 * dcg_get_counters()[7] = number of barriers executed.                                          */
DCG_API int dcg_create_uniform_sharded(const dcg_sim_params *params, int device, int rank, int world,
                                       int nlocal, dcg_sim **out);
/* Slab decomposition of the adaptive grid.  The block pool is cut into pieces of 8,192 slots (one tile of 16
 * slots when nlocal == world), each level's slot range shared equally among the ranks; a rank runs every field
 * kernel on the tiles it owns.  Block topology is replicated: every process updates its own copy identically
 * (same kernels, same host selection).  nlocal == 1: one rank per process/GPU; the eight field arrays live in
 * one virtual range per field stitched from every rank's arena (CUDA VMM, POSIX-fd handles passed over abstract
 * AF_UNIX sockets), so a cell id addresses the same cell on every GPU and neighbour / gather accesses to cells
 * of other ranks travel over NVLink inside the kernels; exchange the dcg_shard_export_handle() blobs and call
 * dcg_shard_import_handles() (maps the peers, then resets in lock step).  nlocal == world: all ranks in this
 * instance on one device.  Accessors return the WHOLE pool (any rank can read every cell);
 * dcg_total_density returns the sum over the cells this instance owns.                                      */
DCG_API int dcg_create_dcgrid_sharded(const dcg_sim_params *params, uint64_t max_num_blocks, int device,
                                      int rank, int world, int nlocal, dcg_sim **out);
DCG_API int dcg_create_dcgrid_sharded_opt(const dcg_sim_params *params, uint64_t max_num_blocks, int device,
                                          int rank, int world, int nlocal, const dcg_options *opt, dcg_sim **out);
DCG_API uint64_t dcg_shard_handle_bytes(void);
DCG_API int dcg_shard_export_handle(dcg_sim *sim, void *out, uint64_t capacity);
DCG_API int dcg_shard_import_handles(dcg_sim *sim, const void *handles, int count);

/* FNV-1a (64-bit, offset basis 14695981039346656037 when seed == 0) of a HOST buffer; chain calls by passing the
 * previous result as `seed`.  oracle/ref_harness prints this digest of the reference's raw density + velocity
 * arrays, started from ITS offset basis 1469598103934665603 (pass that as `seed`; dcgrid_b200.fnv1a64 does):
 * bench.py and the tests compare fields of the big scenes with committed reference digests through it.       */
DCG_API uint64_t dcg_fnv1a64(const void *data, uint64_t bytes, uint64_t seed);

/* Last error text of this instance (or of creation when sim == NULL). */
DCG_API const char *dcg_last_error(const dcg_sim *sim);
DCG_API const char *dcg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DCGRID_B200_H */
