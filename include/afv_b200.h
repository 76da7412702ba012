/*
 * afv_b200.h -- C-ABI of the B200-native active-finite-Voronoi (AFV) build / force / step path.
 *
 * This is the drop-in boundary for PyAFV's per-step hot path.  Every entry point takes plain pointers and
 * sizes; no C++ or torch types cross it.  A handle owns all device memory and one CUDA stream; host pointers
 * are borrowed for the duration of the call only.  A handle is not re-entrant.  Errors are negative return
 * codes plus afv_last_error(); nothing throws across the ABI.  There is NO CPU fallback: every entry point
 * fails with AFV_ERR_CUDA when no sm_100-class device is usable.
 *
 * Reference interfaces replaced (paths into wwang721/pyafv v0.4.16):
 *   afv_create / afv_set_params / afv_set_positions / afv_set_preferred_areas
 *       <- FiniteVoronoiSimulator.__init__            pyafv/finite_voronoi.py:57-93
 *          update_positions / update_params / update_preferred_areas   :1153-1230
 *   afv_build + afv_get_*
 *       <- FiniteVoronoiSimulator.build()             pyafv/finite_voronoi.py:866-945, i.e. the whole of
 *          _build_voronoi_with_extensions (:96-289, scipy/Qhull), _per_cell_geometry (:292-509),
 *          backend_impl.build_point_edges / compute_vertex_derivatives (pyafv/cell_geom.pyx:128-289,307-547),
 *          _assemble_forces (:512-754) and _get_connections (:1110-1150)
 *   afv_step
 *       <- the user-level Euler / active-Langevin loop  examples/active_FV.py:43-54, examples/relaxation.py:41-45
 *   box_L > 0 (native periodic boundaries)
 *       <- tile_pbc + build + [:N]                     pyafv/_connectutil.py:12-96
 *   afv_set_slab / afv_pack_exchange / afv_finish_exchange / afv_late_step_* / afv_pack_owned (slab sharding)
 *       <- decompose_points / _build_domain / _merge_results   pyafv/parallel_voronoi.py:125-338,351-418,590-633
 *   afv_cluster_analysis / afv_get_cluster_labels / afv_get_cluster_sizes / afv_components_of_edges
 *       <- rebuild_connection_matrix / get_cluster_sizes / select_daughter_cluster   pyafv/_connectutil.py:101-172
 *   afv_get_energy / afv_get_cell_energies / afv_relax
 *       <- the energy behind the gradient code (finite_voronoi.py:529,746; physical_params.py:160-171) and the
 *          fixed-dt relaxation loop of examples/relaxation.py:41-45
 */
#ifndef AFV_B200_H
#define AFV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct AfvParams {
    double r;      /* maximal cell radius (l)                         physical_params.py:72 */
    double A0;     /* default preferred area                          :73 */
    double P0;     /* preferred perimeter                             :74 */
    double KA;     /* area elasticity                                 :75 */
    double KP;     /* perimeter elasticity                            :76 */
    double Lambda; /* tension of non-contacting (arc) edges           :77 */
    double delta;  /* contact truncation threshold (already resolved; the host maps None -> 0.45 r, :92-93) */
} AfvParams;

typedef struct AfvContext *AfvHandle;

/* return codes */
#define AFV_OK 0
#define AFV_ERR_CUDA (-1)     /* CUDA runtime error, or no usable device */
#define AFV_ERR_ARG (-2)      /* invalid argument */
#define AFV_ERR_CAPACITY (-3) /* a cell exceeded the ring / polygon capacity of every code path */
#define AFV_ERR_STATE (-4)    /* call order (e.g. getter before build) */

/* afv_build flags */
#define AFV_BUILD_FORCES 1u      /* always implied */
#define AFV_BUILD_CONNECT 2u     /* also produce the contact list (finite_voronoi.py:1110-1150) */
#define AFV_BUILD_RINGS 4u       /* also produce the compact ragged rings (vertices / regions / edges_type) */

/* fields for afv_device_ptr (device arrays in the CURRENT SORTED ORDER; afv_device_ptr(AFV_F_GID) maps a
 * sorted slot back to the caller's cell index) */
enum AfvField {
    AFV_F_XY = 0 /* (n,2) interleaved */, AFV_F_THETA = 2, AFV_F_A0 = 3, AFV_F_GID = 4,
    AFV_F_FX = 5, AFV_F_FY = 6, AFV_F_AREA = 7, AFV_F_PERIM = 8, AFV_F_ARCLEN = 9, AFV_F_COORD = 10,
    AFV_F_OWNED = 11
};

/* ---- life cycle ---------------------------------------------------------------------------------- */
/* device: CUDA ordinal.  box_L: 0 for open boundaries, else the side of the periodic square [0,L)^2
 * (must be >= 4 r).  stream: a cudaStream_t to run on (e.g. torch's current stream; pass cudaStreamLegacy, not 0, for
 * the default stream), or NULL to let the handle create its own non-blocking stream. */
int afv_create(int device, const AfvParams *params, double box_L, void *stream, AfvHandle *out);
void afv_destroy(AfvHandle h);
/* last error text of the handle (or of the failed afv_create when h is NULL); never NULL */
const char *afv_last_error(AfvHandle h);
/* static description: "sm_100a; ring_stride=..; ..." */
const char *afv_build_info(void);

/* ---- state --------------------------------------------------------------------------------------- */
int afv_set_params(AfvHandle h, const AfvParams *params);
/* xy_host: (N,2) row-major float64.  If N differs from the current count, preferred areas are reset to
 * params.A0 and theta to 0 (finite_voronoi.py:1177-1182).  Cell order is reset to the caller's order. */
int afv_set_positions(AfvHandle h, const double *xy_host, int64_t N);
int afv_set_preferred_areas(AfvHandle h, const double *a0_host, int64_t N);
int afv_set_theta(AfvHandle h, const double *theta_host, int64_t N);
int64_t afv_num_cells(AfvHandle h);

/* ---- the hot path -------------------------------------------------------------------------------- */
/* bin + sort + geometry + force for the current positions.  Results stay on the device. */
int afv_build(AfvHandle h, unsigned flags);
/* n_steps of: build, then r += (mu F + v0 (cos th, sin th)) dt, th += sqrt(2 Dr dt) xi   (examples/active_FV.py:43-54).
 * noise_host: NULL -> xi from Philox4x32-10 keyed by (seed, cell id, step0 + s);
 *             else (n_steps, N) float64 in the caller's cell order (parity runs with a recorded stream).
 * Positions are wrapped into [0,L) when the handle is periodic.  The geometry left on the device afterwards is
 * that of the LAST build, i.e. of the positions before the last update.
 * Periodic handles with the built-in generator replay calls of n_steps >= 6 as a CUDA graph of two steps (kept in the
 * handle; environment AFV_NO_GRAPH=1 turns that off): same results, less idle time between the launches. */
int afv_step(AfvHandle h, double dt, double mu, double v0, double Dr, uint64_t seed, int64_t step0, int n_steps,
             const double *noise_host);
/* the xi the device generator produces for (seed, cell id, step): for host-side verification of the stream */
double afv_philox_normal(uint64_t seed, int64_t cell_id, int64_t step);

/* ---- results, in the caller's cell order ----------------------------------------------------------- */
int afv_get_positions(AfvHandle h, double *xy_out /* (N,2) */);
int afv_get_theta(AfvHandle h, double *out /* (N,) */);
int afv_get_forces(AfvHandle h, double *out /* (N,2) */);
int afv_get_areas(AfvHandle h, double *out);
int afv_get_perimeters(AfvHandle h, double *out);
int afv_get_arclens(AfvHandle h, double *out);
int afv_get_coord_nums(AfvHandle h, int64_t *out);
/* contact pairs (i<j): first query the count, then copy (K,2) int64 (unsorted rows) */
int afv_get_num_connections(AfvHandle h, int64_t *K);
int afv_get_connections(AfvHandle h, int64_t *pairs_out);
/* compact rings: total entry count T; then region_off (N+1), edge_types (T) int8 (1 straight, 0 arc),
 * vertices (T,2) absolute coordinates, clockwise per cell (finite_voronoi.py:441-459) */
int afv_get_num_ring_entries(AfvHandle h, int64_t *T);
int afv_get_rings(AfvHandle h, int64_t *region_off, int8_t *edge_types, double *vertices_xy);

/* diagnostics of the last build: [0] cells that needed the large-capacity path, [1] ring look-ups that found
 * no partner (inconsistent near-cocircular topology), [2] longest ring when it exceeds 12 entries (else 0), [3] number
 * of bins (periodic handles) */
int afv_get_diagnostics(AfvHandle h, int64_t out[4]);

/* ---- measurement --------------------------------------------------------------------------------- */
/* when enabled, every stage of build/step is bracketed by CUDA events on the handle's stream */
int afv_set_profiling(AfvHandle h, int enabled);
/* accumulated milliseconds and span counts since the last reset, per stage:
 * 0 bin + sort + reorder + bin starts, 1 k_nbrs (neighbour lists), 2 k_clip (polygon, ring, measures), 3 k_force
 * (+ update), 4 connections / ring export, 5 k_geometry_slow (queued cells) */
#define AFV_N_STAGES 6
int afv_get_stage_times(AfvHandle h, double ms_out[AFV_N_STAGES], int64_t launches_out[AFV_N_STAGES], int reset);
/* number of kernel launches issued by this handle since creation */
int64_t afv_launch_count(AfvHandle h);
/* pinned host memory for end-to-end runs */
void *afv_host_alloc(int64_t bytes);
void afv_host_free(void *p);
/* sustained FP64 FMA throughput of the device in TFLOP/s (a DFMA micro-benchmark; roofline denominator) */
int afv_measure_fp64_peak(AfvHandle h, double *tflops_out);

/* ---- zero-copy + slab sharding primitives (one handle = one GPU = one x-slab) ---------------------- */
void *afv_device_ptr(AfvHandle h, int field);
void *afv_stream(AfvHandle h);
int afv_synchronize(AfvHandle h);
/* Deferred checks (sharded time loops): afv_step returns without synchronising with the host; capacity errors of the
 * steps issued since are reported by the next afv_finish_exchange or afv_synchronize, which synchronise anyway.  Not
 * used when afv_step borrows a host noise array. */
int afv_set_deferred_checks(AfvHandle h, int enabled);
/* Turn a periodic handle (box_L > 0) into one x-slab of that box: the bin grid covers x in [grid_x0, grid_x1) without
 * wrap, y stays periodic.  Cells -- owned and halo -- keep the box's own coordinates in [0, L): bins are taken from the
 * offset to grid_x0 wrapped by L, every displacement is a minimum image, and afv_step wraps x like the single-GPU
 * run, so that every rank forms a displacement from the same two operands as one GPU would: results do not depend on
 * the number of slabs, bit for bit.  Cells leaving the slab are migrated by the exchange. */
int afv_set_slab(AfvHandle h, double grid_x0, double grid_x1);
/* Replace the device state by (x, y, theta, A0, gid) rows given in DEVICE memory: n_owned owned cells followed
 * by n_halo halo cells (halo cells take part in the geometry but receive no force / update).  rows: (n,5)
 * float64 row-major with gid stored as a float64-exact integer. */
int afv_set_cells_device(AfvHandle h, const double *rows_dev, int64_t n_owned, int64_t n_halo);
/* Pack the OWNED cells with x in [x_lo, x_hi) into rows_dev_out (capacity rows), adding x_shift to x; returns the
 * count (halo send when keep != 0; migration when keep == 0, in which case the cells are removed). */
int afv_extract_slab(AfvHandle h, double x_lo, double x_hi, double x_shift, int keep, double *rows_dev_out,
                     int64_t capacity, int64_t *count);
/* drop all halo cells now (afv_pack_exchange retires them on its own) */
int afv_drop_halo(AfvHandle h);
/* append rows (device memory) as owned (is_halo = 0) or halo (is_halo = 1) cells */
int afv_append_cells(AfvHandle h, const double *rows_dev, int64_t n, int is_halo);
int64_t afv_num_owned(AfvHandle h);
/* Pair exchange with the two ring neighbours, no host round trip on the sending side.  A message is a
 * (cap_rows+1, 5) float64 device buffer: row 0 = [count, 0, 0, 0, 0], rows 1..count = cells.
 * afv_pack_pair packs the owned cells with x in [lo0,hi0) / [lo1,hi1) (x shifted by shift0 / shift1).
 * afv_finish_pair, called after the messages were exchanged, reads all four counts with one synchronisation,
 * removes the cells packed before when remove_packed != 0 (migration), and appends the received cells as halo
 * (is_halo != 0) or owned cells.  AFV_ERR_CAPACITY when a count exceeds cap_rows. */
int afv_pack_pair(AfvHandle h, double lo0, double hi0, double shift0, double lo1, double hi1, double shift1,
                  double *msg0_dev, double *msg1_dev, int64_t cap_rows);
int afv_finish_pair(AfvHandle h, const double *in0_dev, const double *in1_dev, int64_t cap_rows, int is_halo, int remove_packed);
/* One exchange per step: ownership transfers and halo cells in the same message.  Message = (cap_rows+1, 5): header
 * [n_transfer, n_halo, n_deep, 0, 0]; the transfer rows fill rows 1 .. n_transfer, the halo rows fill the buffer from
 * the back (halo row i at row cap_rows - i); the order inside either group is arbitrary (one launch packs both messages
 * with atomics).  afv_pack_exchange selects, for the slab
 * [lo, hi) and with d = the offset of x to lo (on a slab handle: the minimum image about the slab centre): left
 * message = owned cells with d < 0 (transfers) and 0 <= d < halo (halo), x + shift_left; right message = d >= hi-lo
 * (transfers) and hi-lo-halo <= d < hi-lo (halo), x + shift_right (both shifts are 0 for slab handles).  The same pass retires the halo
 * copies of the previous exchange (the next sort drops them; afv_drop_halo is the explicit form) and keeps this rank's
 * transferred cells as halo copies.  afv_finish_exchange (one synchronisation) appends the received rows.
 * Replaces parallel_voronoi.py:125-338 (decompose_points) + :590-633 (_merge_results) for a time loop. */
int afv_pack_exchange(AfvHandle h, double lo, double hi, double halo, double shift_left, double shift_right,
                      double *msg_left_dev, double *msg_right_dev, int64_t cap_rows);
int afv_finish_exchange(AfvHandle h, const double *in_left_dev, const double *in_right_dev, int64_t cap_rows);
/* afv_finish_exchange followed by one step (afv_step with n_steps = 1 and the built-in generator), ordered so that the
 * GPU never waits for the host: the received rows are appended and the cell list of the step is built with the counts
 * still on the device, and the host's one synchronisation of the step (to read them) overlaps those launches. */
int afv_finish_exchange_step(AfvHandle h, const double *in_left_dev, const double *in_right_dev, int64_t cap_rows, double dt, double mu,
                             double v0, double Dr, uint64_t seed, int64_t step);
/* Exchange overlapped with the geometry of the interior cells, exact for any displacement ("late cells").  Per step:
 *   afv_late_step_begin   handle's stream: sort of the cells on the device, geometry of the interior cells;
 *   afv_late_finish       exchange stream: one synchronisation (of that stream only) for the counts of the messages
 *                         packed after the previous update, then the received rows are appended behind the sorted
 *                         cells as "late" cells with a bin table of their own;
 *   afv_late_step_end     handle's stream: geometry of the cells in the face columns and of the late cells (both
 *                         tables in view; of every cell when a received transfer landed beyond the face columns, which
 *                         the sender flags in header slot 2), force gather + update of every owned cell (as afv_step
 *                         with n_steps = 1 and the built-in generator); then, on the exchange stream, the messages of
 *                         the next exchange.  The caller moves them on that stream and loops.
 * Before the first afv_late_step_begin: afv_pack_exchange + transport on the handle's stream.  To leave the loop:
 * afv_stream_wait(h, exchange_stream), then afv_finish_exchange with the last messages received.  Needs afv_set_slab
 * and afv_set_deferred_checks(1); cap_rows as for afv_pack_exchange (room for 2 cap_rows cells is reserved). */
int afv_late_step_begin(AfvHandle h, int64_t cap_rows);
int afv_late_finish(AfvHandle h, const double *in_left_dev, const double *in_right_dev, int64_t cap_rows, void *exchange_stream);
int afv_late_step_end(AfvHandle h, double dt, double mu, double v0, double Dr, uint64_t seed, int64_t step, double lo, double hi,
                      double halo, double shift_left, double shift_right, double *msg_left_dev, double *msg_right_dev,
                      int64_t cap_rows, void *exchange_stream);
/* [0] steps taken in the late form, [1] of those, steps whose face pass covered every cell (a received transfer had
 * landed beyond the face columns), [2] late cells of the last step, [3] face columns per side */
int afv_get_exchange_stats(AfvHandle h, int64_t out[4]);
/* make the handle's stream wait for everything issued so far on other_stream */
int afv_stream_wait(AfvHandle h, void *other_stream);

/* ---- the exchange over peer memory (NVLink) instead of a send / receive pair ---------------------------------
 * Replaces the transport of pyafv/parallel_voronoi.py:536-633 (there: pickled sub-arrays through a process pool) between
 * the GPUs of one node.  The caller maps the inboxes of the two ring neighbours into its address space (CUDA IPC) and
 * passes them to afv_pack_exchange / afv_late_step_end as msg_left_dev / msg_right_dev: the pack kernel then writes
 * its messages straight into the neighbours' memory.  afv_flag_signal, issued behind the pack on the same stream,
 * publishes a sequence number in the neighbours' memory (two flag words per launch: a rank has two ring neighbours) once
 * those writes are complete; afv_flag_wait holds a stream until the numbers in this rank's own flag words have reached
 * `value` (sequence numbers compare wrap-safe).  Two
 * inboxes per side, used alternately, are enough: a rank cannot be more than one exchange ahead of a neighbour it also
 * receives from.  A wait that lasts longer than 20 s gives up: the next status read reports AFV_ERR_CAPACITY with
 * "... did not arrive", and the cells are not moved.
 * afv_peer_buffer_create: zero-filled device memory owned by the handle (released by afv_destroy) and its CUDA IPC
 * handle (64 bytes, to be passed to the neighbours by any means); afv_peer_buffer_open: map a neighbour's buffer into
 * this process for this handle's device (unmapped by afv_destroy; not for a buffer of the same process).
 * stream: NULL = the handle's stream. */
int afv_peer_buffer_create(AfvHandle h, int64_t bytes, void **dev_out, unsigned char ipc_handle_out[64]);
int afv_peer_buffer_open(AfvHandle h, const unsigned char ipc_handle[64], void **dev_out);
int afv_flag_signal(AfvHandle h, void *stream, unsigned *flag_a_dev, unsigned *flag_b_dev, unsigned value);       /* flag_b_dev may be NULL */
int afv_flag_wait(AfvHandle h, void *stream, const unsigned *flag_a_dev, const unsigned *flag_b_dev, unsigned value);
/* after afv_build: rows (count,11) of the OWNED cells in device memory:
 * x, y, theta, A0, gid, fx, fy, area, perimeter, arclen, coord_num */
int afv_pack_owned(AfvHandle h, double *rows_dev_out, int64_t capacity, int64_t *count);


/* ---- cluster analysis on the device (the callers of `connections`, pyafv/_connectutil.py:101-172) ------------- */
/* Connected components of the contact graph of the last build, straight from the rings on the device (no (K,2) list
 * is built or copied).  Component labels are numbered by ascending smallest cell index, i.e. exactly like
 * scipy.sparse.csgraph.connected_components numbers them.  Single-GPU handles (afv_set_positions) only. */
int afv_cluster_analysis(AfvHandle h, int64_t *n_components);
int afv_get_cluster_labels(AfvHandle h, int64_t *labels_out /* (N,) caller order */);
int afv_get_cluster_sizes(AfvHandle h, int64_t *sizes_out /* (n_components,) */);
/* the same for an explicit edge list in host memory (get_cluster_sizes(N, connect)): labels_out (N) and sizes_out (N;
 * the first *n_components entries are set) may be NULL */
int afv_components_of_edges(AfvHandle h, int64_t N, const int64_t *pairs_host, int64_t K, int64_t *labels_out, int64_t *sizes_out,
                            int64_t *n_components);

/* ---- energy and relaxation ------------------------------------------------------------------------------------ */
/* E = sum_i KA (A_i - A0_i)^2 + KP (P_i - P0)^2 + Lambda L_arc,i of the last build, reduced on the device in a fixed
 * order (bit-reproducible run to run) */
int afv_get_energy(AfvHandle h, double *energy_out);
int afv_get_cell_energies(AfvHandle h, double *out /* (N,) caller order */);
/* FIRE relaxation (Bitzek et al., PRL 97, 170201) entirely on the device: per iteration build + forces, the three
 * global sums F.v, |v|^2, |F|^2, the time-step / mixing update and the move, without a host round trip; the host
 * looks at max|F| every `check_every` iterations.  Stops when max|F| <= f_tol or after max_steps.  dt0: initial time
 * step, dt_max: largest one.  out[0] iterations done, [1] energy, [2] max|F|, [3] last dt, [4] 1 if converged.
 * energies_out (may be NULL): energy at every check, at most n_energies entries; *n_checks = number written.
 * Single-GPU handles only; replaces the fixed-dt loop of examples/relaxation.py:41-45. */
int afv_relax(AfvHandle h, double dt0, double dt_max, double f_tol, int max_steps, int check_every, double out[5],
              double *energies_out, int n_energies, int *n_checks);

#ifdef __cplusplus
}
#endif
#endif /* AFV_B200_H */
