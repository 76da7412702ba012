// C-ABI implementation (include/afv_b200.h): context, memory, stage scheduling on one CUDA stream.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a profiler is attached

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/afv_b200.h"
#include "afv_kernels.cuh"

using namespace afv;

namespace {

thread_local std::string g_create_error = "";

constexpr int TPB = 128;
inline int nblk(int64_t n, int tpb = TPB) { return (int)((n + tpb - 1) / tpb); }

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct AfvContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    AfvParams prm{};
    double L = 0.0;
    bool slab = false;          // x-slab of a periodic box: grid open in x over [slab_x0, slab_x1), periodic in y
    double slab_x0 = 0.0, slab_x1 = 0.0;
    int64_t n = 0, n_owned = 0, cap = 0;
    GridDesc grid_host{};        // last grid description uploaded to d_grid (periodic / slab mode)
    bool grid_uploaded = false;
    int64_t n_dead = 0;          // cells marked CELL_DEAD (counted in n) that the next sort drops
    // late cells: appended behind the sorted cells after the sort of the step (overlapped exchange), slots n_main .. n-1
    int64_t n_main = 0, n_late = 0;
    int *bin_start_late = nullptr;     // their bin table (entries relative to n_main)
    bool late_deep = false;            // a received cell lies beyond the face columns: the face pass covers every cell
    int face_cols = 0;
    int64_t face_len0 = 0, face_s1 = 0;   // sorted cells in the first face_cols columns / first slot of the last face_cols columns
    DevBuf late_keys, late_keys_sorted, late_perm, late_tmp;
    cudaEvent_t ev_sorted = nullptr, ev_late = nullptr, ev_force = nullptr, ev_pack = nullptr, ev_interior = nullptr;
    int64_t late_steps = 0, late_deep_steps = 0;   // steps taken in the late form / of those, with the face pass over every cell
    bool late_mode = false;            // stepping with afv_late_*: column-major bins and the slab variants of the geometry kernels
    bool pack_pending = false;         // ev_pack was recorded: the next sort waits for the ownership marks of that pack
    bool gid_is_perm = true;  // gid is a permutation of 0..n-1 (single-GPU use); false after set_cells_device
    // SoA state, double-buffered for the reorder
    double2 *xy[2] = {nullptr, nullptr};
    double *th[2] = {nullptr, nullptr}, *a0[2] = {nullptr, nullptr};
    int64_t *gid[2] = {nullptr, nullptr};
    unsigned char *owned[2] = {nullptr, nullptr};
    int cur = 0;
    // cell list
    unsigned *keys = nullptr, *keys_sorted = nullptr;
    int *iota = nullptr, *perm = nullptr, *bin_start = nullptr, *bin_cnt = nullptr;
    bool bin_cnt_zero = false;     // every counter is zero (k_bin_scatter counts them back down)
    int *slot_of_gid = nullptr;     // where the caller's cell gid sits in the sorted order (handles filled by afv_set_positions)
    int nbins_cap = 0, key_bits = 0;
    GridDesc *d_grid = nullptr;
    double *bbox_part = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;
    // results (sorted order)
    double *fx = nullptr, *fy = nullptr, *area = nullptr, *perim = nullptr, *arcl = nullptr;
    int *coord = nullptr;
    unsigned char *ring_len = nullptr, *n_arc = nullptr;
    double2 *ring_h = nullptr;    // slot-major ring: vertices
    int *ring_nbr = nullptr;      //                  neighbour slot across the leaving edge (ARC = arc)
    int2 *arc_key = nullptr;      // slot-major arcs (afv_device.cuh)
    double4 *arc_pts = nullptr;
    CellRec *rec = nullptr;       // per cell: position + energy prefactors, gathered by the neighbours in k_force
    DevBuf nb_list;               // slot-major neighbour lists, nb_rows rows of `cap` slots
    int nb_rows = 0;
    unsigned char *nb_cnt = nullptr;
    int *nb_min = nullptr;
    int *flags = nullptr;         // device: [0] slow-path cells, [1] lookup misses, [2] max ring length, [3] hard overflows
    int *ovf_cells = nullptr, *ovf_index = nullptr, *ovf_len = nullptr, *ovf_hull = nullptr, *ovf_m = nullptr, *ovf_ring_nbr = nullptr;
    double2 *ovf_ring_h = nullptr;
    int2 *ovf_arc_key = nullptr;
    double4 *ovf_arc_pts = nullptr;
    int h_flags[4] = {0, 0, 0, 0};
    // scratch
    DevBuf scratch_a, scratch_b, noise_buf, conn_buf, ring_types, ring_xy, ring_off;
    DevBuf cc_parent, cc_flag, cc_rank, cc_labels, cc_sizes, cc_tmp, cc_edges;   // cluster analysis (afv_cluster.inc)
    int64_t n_clusters = 0;
    DevBuf sums_partial, fire_state, fire_vel;   // energy / FIRE relaxation (afv_relax.inc)
    int *cnt = nullptr, *off = nullptr;  // per-cell counts / offsets (cap + 1)
    int64_t n_conn = -1, n_ring = -1;
    bool built = false;
    int cmax_hint = 32;
    bool defer_checks = false;  // afv_step returns without synchronising; the flags are read at the next exchange / synchronize
    // pair exchange state (afv_pack_pair -> afv_finish_pair)
    int *pair_cnt = nullptr;       // device: list lengths of the last pack (2 for the pair form, 4 for the exchange)
    int *xch_work = nullptr;       // device int[8]: working counters of k_exchange_pack (zero between launches)
    int *n_all_dev = nullptr;      // device: cell count after an append whose counts the host has not read yet
    cudaEvent_t ev_counts = nullptr, ev_copy = nullptr;
    double *pair_host = nullptr;   // pinned [20]: counts sent / received, the deferred flags, the displacement slots
    double *pair_dev = nullptr;    // device [20]
    double pair_rng[4] = {0, 0, 0, 0};
    // profiling
    bool profiling = false;
    // time loop of afv_step replayed as a CUDA graph (two steps per replay: the double buffers are back where they started)
    long long *step_dev = nullptr;     // the step index on the device: k_bin_count advances it, k_force reads it
    bool step_on_device = false;       // the launches being issued take the step index from step_dev
    uint64_t alloc_epoch = 0;          // bumped by every (re)allocation: a captured graph holds the old pointers
    cudaGraphExec_t step_graph = nullptr;
    bool step_graph_off = false;       // capturing failed once on this handle's stream: plain launches from then on
    std::string step_graph_key;
    struct Span { int stage; cudaEvent_t a, b; };
    std::vector<void *> peer_own, peer_open;   // exchange over peer memory: this rank's inboxes; the neighbours' (CUDA IPC mappings)
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    double stage_ms[AFV_N_STAGES] = {0, 0, 0, 0, 0, 0};
    int64_t stage_launches[AFV_N_STAGES] = {0, 0, 0, 0, 0, 0};
    int64_t launches = 0;
    std::string err;
};

namespace {

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
            return AFV_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

#define CHECK_H(h)                                                                                       \
    if (!(h)) return AFV_ERR_ARG;                                                                        \
    {                                                                                                    \
        cudaError_t e_ = cudaSetDevice((h)->device);                                                     \
        if (e_ != cudaSuccess) { (h)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e_); return AFV_ERR_CUDA; } \
    }

template <typename T>
int dev_alloc(AfvHandle h, T **p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) count = 1;
    ++h->alloc_epoch;
    CK(cudaMalloc((void **)p, count * sizeof(T)));
    return AFV_OK;
}

int ensure_buf(AfvHandle h, DevBuf &b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return AFV_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.bytes = 0;
    ++h->alloc_epoch;
    size_t want = std::max(bytes, (size_t)256);
    CK(cudaMalloc(&b.p, want));
    b.bytes = want;
    return AFV_OK;
}

void free_all(AfvHandle h) {
    for (int k = 0; k < 2; ++k) {
        cudaFree(h->xy[k]); cudaFree(h->th[k]); cudaFree(h->a0[k]); cudaFree(h->gid[k]); cudaFree(h->owned[k]);
        h->xy[k] = nullptr; h->th[k] = h->a0[k] = nullptr; h->gid[k] = nullptr; h->owned[k] = nullptr;
    }
    cudaFree(h->keys); cudaFree(h->keys_sorted); cudaFree(h->iota); cudaFree(h->perm); cudaFree(h->bin_start);
    cudaFree(h->bin_cnt); h->bin_cnt = nullptr; cudaFree(h->slot_of_gid); h->slot_of_gid = nullptr;
    cudaFree(h->cub_tmp); cudaFree(h->fx); cudaFree(h->fy); cudaFree(h->area); cudaFree(h->perim); cudaFree(h->arcl);
    cudaFree(h->coord); cudaFree(h->ring_len); cudaFree(h->cnt); cudaFree(h->off);
    cudaFree(h->ring_h); h->ring_h = nullptr; cudaFree(h->ring_nbr); h->ring_nbr = nullptr;
    cudaFree(h->arc_key); h->arc_key = nullptr; cudaFree(h->arc_pts); h->arc_pts = nullptr;
    cudaFree(h->n_arc); h->n_arc = nullptr; cudaFree(h->rec); h->rec = nullptr;
    cudaFree(h->nb_cnt); h->nb_cnt = nullptr; cudaFree(h->nb_min); h->nb_min = nullptr;
    cudaFree(h->nb_list.p); h->nb_list.p = nullptr; h->nb_list.bytes = 0; h->nb_rows = 0;
    cudaFree(h->ovf_index); h->ovf_index = nullptr;
    cudaFree(h->bin_start_late); h->bin_start_late = nullptr;
    h->keys = h->keys_sorted = nullptr; h->iota = h->perm = h->bin_start = nullptr; h->cub_tmp = nullptr;
    h->fx = h->fy = h->area = h->perim = h->arcl = nullptr; h->coord = nullptr; h->ring_len = nullptr;
    h->cnt = h->off = nullptr;
    h->cap = 0;
}

// (re)allocate everything for `cap` cells, preserving the first `keep` cells of the current SoA state
int reserve(AfvHandle h, int64_t want, int64_t keep) {
    if (want <= h->cap) return AFV_OK;
    if (want >= (int64_t)1 << 31) { h->err = "more than 2^31-1 cells per GPU are not supported"; return AFV_ERR_ARG; }
    int64_t cap = std::max<int64_t>(want + want / 8, 1024);
    cap = std::min<int64_t>(cap, ((int64_t)1 << 31) - 1);
    // keep old state
    double2 *oxy = h->xy[h->cur];
    double *ot = h->th[h->cur], *oa = h->a0[h->cur];
    int64_t *og = h->gid[h->cur];
    unsigned char *oo = h->owned[h->cur];
    h->xy[h->cur] = nullptr; h->th[h->cur] = h->a0[h->cur] = nullptr; h->gid[h->cur] = nullptr; h->owned[h->cur] = nullptr;
    free_all(h);
    int rc = AFV_OK;
    for (int k = 0; k < 2 && rc == AFV_OK; ++k) {
        if ((rc = dev_alloc(h, &h->xy[k], cap))) break;
        if ((rc = dev_alloc(h, &h->th[k], cap))) break;
        if ((rc = dev_alloc(h, &h->a0[k], cap))) break;
        if ((rc = dev_alloc(h, &h->gid[k], cap))) break;
        if ((rc = dev_alloc(h, &h->owned[k], cap))) break;
    }
    cudaError_t ce = cudaSuccess;
    if (rc == AFV_OK && keep > 0 && oxy) {
        h->cur = 0;
        auto cp = [&](void *dst, const void *src, size_t bytes) { if (ce == cudaSuccess) ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, h->stream); };
        cp(h->xy[0], oxy, keep * sizeof(double2)); cp(h->th[0], ot, keep * sizeof(double)); cp(h->a0[0], oa, keep * sizeof(double));
        cp(h->gid[0], og, keep * sizeof(int64_t)); cp(h->owned[0], oo, keep);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    }
    cudaFree(oxy); cudaFree(ot); cudaFree(oa); cudaFree(og); cudaFree(oo);
    if (rc == AFV_OK && ce != cudaSuccess) { h->err = std::string("reserve: ") + cudaGetErrorString(ce); rc = AFV_ERR_CUDA; }
    if (rc) { h->n = h->n_owned = h->n_dead = 0; h->built = false; return rc; }   // the old state is gone: the handle is empty, not stale
    // bins: up to 2 per cell, power of two so that the sort width is fixed
    int bits = 6;
    while (((int64_t)1 << bits) < 2 * cap && bits < 31) ++bits;
    h->key_bits = bits;
    h->nbins_cap = (int)std::min<int64_t>(((int64_t)1 << bits) - 1, ((int64_t)1 << 31) - 2);   // the all-ones key marks dead cells
    if ((rc = dev_alloc(h, &h->keys, cap))) return rc;
    if ((rc = dev_alloc(h, &h->keys_sorted, cap))) return rc;
    if ((rc = dev_alloc(h, &h->iota, cap))) return rc;
    if ((rc = dev_alloc(h, &h->perm, cap))) return rc;
    if ((rc = dev_alloc(h, &h->bin_start, (size_t)h->nbins_cap + 2))) return rc;
    if ((rc = dev_alloc(h, &h->bin_cnt, (size_t)h->nbins_cap + 2))) return rc;
    h->bin_cnt_zero = false;
    if ((rc = dev_alloc(h, &h->slot_of_gid, cap))) return rc;
    if ((rc = dev_alloc(h, &h->bin_start_late, (size_t)h->nbins_cap + 2))) return rc;
    if ((rc = dev_alloc(h, &h->fx, cap))) return rc;
    if ((rc = dev_alloc(h, &h->fy, cap))) return rc;
    if ((rc = dev_alloc(h, &h->area, cap))) return rc;
    if ((rc = dev_alloc(h, &h->perim, cap))) return rc;
    if ((rc = dev_alloc(h, &h->arcl, cap))) return rc;
    if ((rc = dev_alloc(h, &h->coord, cap))) return rc;
    if ((rc = dev_alloc(h, &h->ring_len, cap))) return rc;
    if ((rc = dev_alloc(h, &h->n_arc, cap))) return rc;
    if ((rc = dev_alloc(h, &h->rec, cap))) return rc;
    if ((rc = dev_alloc(h, &h->nb_cnt, cap))) return rc;
    if ((rc = dev_alloc(h, &h->nb_min, cap))) return rc;
    if ((rc = dev_alloc(h, &h->ring_h, (size_t)cap * RS))) return rc;
    if ((rc = dev_alloc(h, &h->ring_nbr, (size_t)cap * RS))) return rc;
    if ((rc = dev_alloc(h, &h->arc_key, (size_t)cap * ARC_MAX))) return rc;
    if ((rc = dev_alloc(h, &h->arc_pts, (size_t)cap * ARC_MAX))) return rc;
    if ((rc = dev_alloc(h, &h->ovf_index, cap))) return rc;
    if ((rc = dev_alloc(h, &h->cnt, cap + 1))) return rc;
    if ((rc = dev_alloc(h, &h->off, cap + 1))) return rc;
    k_iota<<<nblk(cap), TPB, 0, h->stream>>>(h->iota, (int)cap);
    // cub temp storage for the largest sort / scan we issue
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, h->keys, h->keys_sorted, h->iota, h->perm, (int)cap, 0, h->key_bits, h->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, h->cnt, h->off, std::max((int)cap + 1, h->nbins_cap + 2), h->stream);
    size_t b3 = 0;
    cub::DeviceSelect::Flagged(nullptr, b3, h->iota, h->cnt, h->perm, h->off, (int)cap, h->stream);
    h->cub_tmp_bytes = std::max(std::max(b1, b2), b3) + 256;
    CK(cudaMalloc(&h->cub_tmp, h->cub_tmp_bytes));
    CK(cudaGetLastError());
    h->cap = cap;
    return AFV_OK;
}

// One stage of a build / step: an NVTX range around its launches (tracing) and, when profiling is
// switched on, a pair of CUDA events on the launching stream.
static const char *const STAGE_NAMES[AFV_N_STAGES] = {"afv:cell_list", "afv:k_nbrs", "afv:k_clip", "afv:k_force", "afv:export", "afv:k_geometry_slow"};
struct StageTimer {
    AfvHandle h;
    int idx = -1;
    cudaStream_t st;
    StageTimer(AfvHandle h_, int stage, cudaStream_t stream = nullptr) : h(h_), st(stream ? stream : h_->stream) {
        nvtxRangePushA(STAGE_NAMES[stage]);
        if (!h->profiling) return;
        AfvContext::Span sp;
        sp.stage = stage;
        auto get = [&]() {
            cudaEvent_t e;
            if (!h->event_pool.empty()) { e = h->event_pool.back(); h->event_pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        sp.a = get(); sp.b = get();
        cudaEventRecord(sp.a, st);
        h->spans.push_back(sp);
        idx = (int)h->spans.size() - 1;
    }
    ~StageTimer() {
        if (idx >= 0) cudaEventRecord(h->spans[idx].b, st);
        nvtxRangePop();
    }
};

void collect_spans(AfvHandle h) {
    if (h->spans.empty()) return;
    cudaStreamSynchronize(h->stream);
    for (auto &sp : h->spans) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, sp.a, sp.b);
        h->stage_ms[sp.stage] += ms;
        h->stage_launches[sp.stage] += 1;
        h->event_pool.push_back(sp.a); h->event_pool.push_back(sp.b);
    }
    h->spans.clear();
}

RingView ring_view(AfvHandle h) { return RingView{h->ring_h, h->ring_nbr, h->ring_len, h->ovf_index, h->ovf_len, h->ovf_ring_h, h->ovf_ring_nbr, (size_t)h->cap}; }

PhysDev phys_dev(const AfvParams &p) { return PhysDev{p.r, p.P0, p.KA, p.KP, p.Lambda, p.delta}; }

// grid of a periodic box (or of an x-slab of it): returns the bin counts
void periodic_grid(AfvHandle h, GridDesc &g) {
    const double cut = 2.0 * h->prm.r;
    const double wx = h->slab ? (h->slab_x1 - h->slab_x0) : h->L;
    double fy = std::max(1.0, std::floor(h->L / cut)), fx = std::max(1.0, std::floor(wx / cut));
    if (fx * fy > (double)h->nbins_cap) {   // very dilute system: coarser bins (still >= 2r) so that the grid fits
        const double sc = std::sqrt((double)h->nbins_cap / (fx * fy));
        fx = std::max(1.0, std::floor(fx * sc)); fy = std::max(1.0, std::floor(fy * sc));
    }
    int nbx = (int)fx, nby = (int)fy;
    while ((int64_t)nbx * nby > h->nbins_cap) { if (nbx >= nby) --nbx; else --nby; }
    g.x0 = h->slab ? h->slab_x0 : 0.0; g.y0 = 0.0;
    g.inv_bx = (double)nbx / wx; g.inv_by = (double)nby / h->L;
    g.Lx = h->L; g.halfLx = 0.5 * h->L; g.Ly = h->L; g.halfLy = 0.5 * h->L;
    g.nbx = nbx; g.nby = nby; g.perx = h->slab ? 0 : 1; g.pery = 1;
    // late-cell stepping of an x-slab: columns contiguous, so that the face columns are the two ends of the sorted order
    g.swap = (h->slab && h->late_mode) ? 1 : 0;
    g.mix = h->slab ? 1 : 0;
    g.seam_lo = 0; g.seam_hi = -1;
    if (h->slab) {   // where the box seam x = 0 = L lies in this grid: wrapped offset of x = 0 to the grid origin
        double so = std::fmod(-g.x0, h->L);
        if (so < 0.0) so += h->L;
        if (so < wx + 1.0 / g.inv_bx) {
            const int sc = (int)std::floor(so * g.inv_bx);
            g.seam_lo = std::max(sc - 1, 0); g.seam_hi = std::min(sc + 1, nbx - 1);
        }
    }
}

int upload_grid_periodic(AfvHandle h) {
    GridDesc g{};
    periodic_grid(h, g);
    if (h->grid_uploaded && std::memcmp(&g, &h->grid_host, sizeof g) == 0) return AFV_OK;   // unchanged since the last step
    h->grid_host = g;
    CK(cudaMemcpyAsync(h->d_grid, &h->grid_host, sizeof g, cudaMemcpyHostToDevice, h->stream));
    h->grid_uploaded = true;
    return AFV_OK;
}

// The cell list: counting sort by bin (k_bin_count, exclusive scan, k_bin_scatter).  Cells marked dead by the exchange land
// behind the live ones and are dropped here.
// n_dev != nullptr: the number of cells (h->n is then only what the host knows so far) is read on the device, the launches
// cover n_upper cells; the caller finishes the bookkeeping with sort_booked() once it has the count.
void sort_booked(AfvHandle h, int64_t n_all) {
    h->n = n_all - h->n_dead; h->n_dead = 0;
    h->n_main = h->n; h->n_late = 0; h->late_deep = false;
}
int stage_sort(AfvHandle h, const int *n_dev = nullptr, int64_t n_upper = 0) {
    StageTimer t(h, 0);
    const int n_all = (int)(n_dev ? n_upper : h->n);
    const int c = h->cur, o = 1 - c;
    if (h->L > 0.0) {
        int rc = upload_grid_periodic(h);
        if (rc) return rc;
    } else {
        int nb = std::min(nblk(n_all, 256), 1024);
        k_bbox_partial<<<nb, 256, 0, h->stream>>>(h->xy[c], n_all, h->bbox_part);
        k_bbox_final<<<1, 32, 0, h->stream>>>(h->bbox_part, nb, h->prm.r, h->nbins_cap, h->d_grid);
        h->launches += 2;
    }
    // bins: those actually in use when the host knows them (periodic box / slab), else the capacity (the grid of an open
    // system is made on the device); dead cells go to the bin behind the last one
    int nbins = h->nbins_cap;
    if (h->L > 0.0) nbins = h->grid_host.nbx * h->grid_host.nby;
    const unsigned dead_key = (unsigned)nbins;
    if (!h->bin_cnt_zero) {   // fresh allocation; afterwards k_bin_scatter leaves every counter at zero
        CK(cudaMemsetAsync(h->bin_cnt, 0, ((size_t)h->nbins_cap + 2) * sizeof(int), h->stream));
        h->bin_cnt_zero = true;
    }
    // (thread 0 also restarts the queue of the slow path: flags[0])
    k_bin_count<<<nblk(n_all, 256), 256, 0, h->stream>>>(h->xy[c], h->owned[c], n_all, n_dev, h->d_grid, dead_key, h->keys, h->bin_cnt, h->flags,
                                                          h->step_on_device ? h->step_dev : nullptr);
    size_t tmp = h->cub_tmp_bytes;
    CK(cub::DeviceScan::ExclusiveSum(h->cub_tmp, tmp, h->bin_cnt, h->bin_start, nbins + 2, h->stream));
    k_bin_scatter<<<nblk(n_all, 256), 256, 0, h->stream>>>(h->keys, n_all, n_dev, dead_key, h->bin_start, h->bin_cnt, h->xy[c], h->th[c], h->a0[c],
                                                            h->gid[c], h->owned[c], h->xy[o], h->th[o], h->a0[o], h->gid[o], h->owned[o],
                                                            h->keys_sorted, h->gid_is_perm ? h->slot_of_gid : nullptr);
    h->cur = o;
    if (!n_dev) sort_booked(h, h->n);
    if (h->L > 0.0) h->h_flags[3] = nbins;
    h->launches += 3;  // count, scan, scatter
    CK(cudaGetLastError());
    return AFV_OK;
}

// which cells a geometry launch covers (slab sharding with the exchange overlapped; the default is every cell)
struct GeomPass {
    int filter = 0;        // 0 every cell, 1 interior cells only (placed by the sort, away from the face columns), 2 the others
    bool late = false;     // look at the late cells' bin table as well
    bool slow = true;      // run the queue of the slow path afterwards (once per step, after the last pass)
    bool fast = true;      // run k_hull / k_ring (false: only the slow-path queue)
    cudaStream_t stream = nullptr;   // nullptr: the handle's stream
};

int stage_geometry(AfvHandle h, GeomPass gp = GeomPass()) {
    // threads with work: every cell; the sorted cells (interior pass: the face columns drop out per thread); or, dense, the
    // cells of the face columns + the late cells
    const int64_t face_len1 = h->n_main - h->face_s1;
    const int n = (int)(gp.filter == 1 ? h->n_main : (gp.filter == 2 ? h->face_len0 + face_len1 + h->n_late : h->n));
    const int c = h->cur;
    GeomArgs a;
    a.face_len0 = (int)h->face_len0; a.face_len1 = (int)face_len1; a.face_s1 = (int)h->face_s1;
    a.xy = h->xy[c]; a.a0 = h->a0[c]; a.keys = h->keys_sorted; a.bin_start = h->bin_start; a.grid = h->d_grid;
    a.n = n; a.ph = phys_dev(h->prm); a.stride = (size_t)h->cap;
    a.ring_h = h->ring_h; a.ring_nbr = h->ring_nbr; a.ring_len = h->ring_len; a.arc_key = h->arc_key; a.arc_pts = h->arc_pts; a.n_arc = h->n_arc;
    a.rec = h->rec; a.area = h->area; a.perim = h->perim; a.arcl = h->arcl; a.coord = h->coord; a.flags = h->flags;
    a.ovf_cells = h->ovf_cells; a.ovf_index = h->ovf_index; a.ovf_len = h->ovf_len; a.ovf_hull = h->ovf_hull; a.ovf_m = h->ovf_m;
    a.ovf_ring_h = h->ovf_ring_h; a.ovf_ring_nbr = h->ovf_ring_nbr; a.ovf_arc_key = h->ovf_arc_key; a.ovf_arc_pts = h->ovf_arc_pts;
    a.n_main = (int)h->n_main; a.bin_start_late = (gp.late && h->n_late > 0) ? h->bin_start_late : nullptr;
    a.filter = gp.filter; a.face_cols = h->face_cols;
    a.mi_Lx = a.mi_Ly = h->L > 0.0 ? h->L : 0.0; a.mi_hx = a.mi_hy = 0.5 * a.mi_Lx;
    // candidate-list capacity per cell (shared memory scales with it): for a periodic box the mean number of cells
    // within 2r is known, 4 pi r^2 N / L^2; otherwise start at 32 and grow when the slow path gets busy
    if (h->L > 0.0) {
        const double area = h->L * (h->slab ? (h->slab_x1 - h->slab_x0) : h->L);
        const double mean = 4.0 * 3.141592653589793 * h->prm.r * h->prm.r * (double)h->n / area;
        h->cmax_hint = std::max(h->cmax_hint, std::min(CMAX_HARD, std::max(24, (int)(1.8 * mean) + 12)));
    }
    int cmax = h->cmax_hint;
    cudaStream_t st = gp.stream ? gp.stream : h->stream;
    a.nb_list = (int *)h->nb_list.p; a.nb_cnt = h->nb_cnt; a.nb_min = h->nb_min;
    if (n > 0 && gp.fast) {
        if (h->nb_rows < cmax) {   // the neighbour lists grow with the candidate capacity (the other stream of a slab may still read them)
            CK(cudaDeviceSynchronize());
            int rc = ensure_buf(h, h->nb_list, (size_t)cmax * (size_t)h->cap * sizeof(int));
            if (rc) return rc;
            h->nb_rows = cmax;
            a.nb_list = (int *)h->nb_list.p;
        }
        {
            StageTimer t(h, 1, st);
            if (h->late_mode) k_nbrs<true><<<nblk(n, 128), 128, 0, st>>>(a, cmax);
            else k_nbrs<false><<<nblk(n, 128), 128, 0, st>>>(a, cmax);
        }
        {
            StageTimer t(h, 2, st);
            if (h->late_mode) k_clip<true><<<nblk(n, GEOM_TPB), GEOM_TPB, clip_smem_bytes(), st>>>(a);
            else k_clip<false><<<nblk(n, GEOM_TPB), GEOM_TPB, clip_smem_bytes(), st>>>(a);
        }
        h->launches += 2;
    }
    if (gp.slow) {
        StageTimer t(h, 5, st);
        // cells queued by the fast kernels (usually none): large-capacity path, reads the queue length on the device
        a.filter = 0;
        k_geometry_slow<<<148, 64, 0, st>>>(a);
        h->launches += 1;
    }
    CK(cudaGetLastError());
    return AFV_OK;
}

int stage_force(AfvHandle h, bool do_step, double dt, double mu, double v0, double Dr, uint64_t seed, int64_t step,
                const double *noise_row) {
    StageTimer t(h, 3);
    const int n = (int)h->n;
    const int c = h->cur;
    ForceArgs a;
    a.ring_h = h->ring_h; a.ring_nbr = h->ring_nbr; a.stride = (size_t)h->cap; a.ring_len = h->ring_len; a.rec = h->rec;
    a.n_arc = h->n_arc; a.arc_key = h->arc_key; a.arc_pts = h->arc_pts; a.grid = h->d_grid; a.owned = h->owned[c]; a.gid = h->gid[c]; a.n = n;
    a.ovf_index = h->ovf_index; a.ovf_len = h->ovf_len; a.ovf_ring_h = h->ovf_ring_h; a.ovf_ring_nbr = h->ovf_ring_nbr;
    a.ovf_arc_key = h->ovf_arc_key; a.ovf_arc_pts = h->ovf_arc_pts;
    a.ph = phys_dev(h->prm); a.fx = h->fx; a.fy = h->fy; a.flags = h->flags;
    a.do_step = do_step ? 1 : 0; a.xy = h->xy[c]; a.theta = h->th[c];
    a.dt = dt; a.mu = mu; a.v0 = v0; a.noise_amp = std::sqrt(2.0 * Dr * dt); a.seed = seed; a.step = step;
    a.step_dev = h->step_on_device ? h->step_dev : nullptr;
    a.noise = noise_row; a.Lx = h->L; a.Ly = h->L;   // a slab keeps the box's own coordinates, wrapped like the single-GPU run
    // (a periodic box -- whole or an x-slab of it -- takes the minimum image in both directions; an open system in none)
    a.mi_Lx = a.mi_Ly = h->L > 0.0 ? h->L : 0.0; a.mi_hx = a.mi_hy = 0.5 * a.mi_Lx;
    k_force<<<nblk(n, FB_CELLS), FB_THREADS, 0, h->stream>>>(a);
    h->launches += 1;
    CK(cudaGetLastError());
    return AFV_OK;
}

int check_flags(AfvHandle h) {
    int f[4];
    CK(cudaMemcpyAsync(f, h->flags, sizeof f, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->h_flags[0] = f[0]; h->h_flags[1] = f[1]; h->h_flags[2] = f[2];
    if ((int64_t)f[0] * 100 > h->n + 400) h->cmax_hint = std::min(CMAX_HARD, h->cmax_hint + 12);  // busy slow path: longer lists next time
    if (f[3] > 0) {
        char buf[200];
        if (f[3] >= FLAG_TIMEOUT_MARK) snprintf(buf, sizeof buf, "the exchange message of a ring neighbour did not arrive (peer-memory flag wait timed out)");
        else snprintf(buf, sizeof buf, "%d cell(s) exceeded every capacity (slow path: ring %d, polygon %d, at most %d cells per build)",
                 f[3], RS_SLOW, KPOLY_SLOW, OVF_MAX);
        h->err = buf;
        return AFV_ERR_CAPACITY;
    }
    return AFV_OK;
}

// Geometry stage with a synchronous capacity check: when more cells were queued for the slow path than it can take
// (the candidate lists were too short for the local density), double the list capacity and run the stage again.
// Used where a host round trip per build is acceptable: afv_build, and every step of an open-boundary run (whose
// density is not known in advance).
int geometry_checked(AfvHandle h) {
    for (;;) {
        int rc = stage_geometry(h);
        if (rc) return rc;
        int f[4];
        CK(cudaMemcpyAsync(f, h->flags, sizeof f, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (f[3] == 0 || h->cmax_hint >= CMAX_HARD) return AFV_OK;      // fine, or nothing left to grow: check_flags reports it
        h->cmax_hint = std::min(CMAX_HARD, 2 * h->cmax_hint);
        CK(cudaMemsetAsync(h->flags, 0, 4 * sizeof(int), h->stream));
    }
}

int require_built(AfvHandle h) {
    if (!h->built) { h->err = "no geometry on the device: call afv_build first"; return AFV_ERR_STATE; }
    return AFV_OK;
}

int get_unsorted(AfvHandle h, const double *src, double *out_host, int stride, int comp_count, const double *src2) {
    // scatter to caller order into scratch, then D2H
    const int n = (int)h->n;
    if (!h->gid_is_perm) { h->err = "results in caller order are only defined for handles filled by afv_set_positions"; return AFV_ERR_STATE; }
    int rc = ensure_buf(h, h->scratch_a, (size_t)n * stride * sizeof(double));
    if (rc) return rc;
    double *tmp = (double *)h->scratch_a.p;
    k_unsort_f64<<<nblk(n, 256), 256, 0, h->stream>>>(src, h->gid[h->cur], n, tmp, stride, 0);
    if (comp_count == 2) k_unsort_f64<<<nblk(n, 256), 256, 0, h->stream>>>(src2, h->gid[h->cur], n, tmp, stride, 1);
    h->launches += comp_count;
    CK(cudaMemcpyAsync(out_host, tmp, (size_t)n * stride * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return AFV_OK;
}

}  // namespace

// ======================================================================================================
extern "C" {

const char *afv_build_info(void) {
    static char buf[160];
    snprintf(buf, sizeof buf, "afv_b200 sm_100a; ring_stride=%d; polygon_capacity=%d; slow_path=%d/%d; cuda_runtime=%d", RS, KP, KPOLY_SLOW, RS_SLOW, CUDART_VERSION);
    return buf;
}

const char *afv_last_error(AfvHandle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int afv_create(int device, const AfvParams *params, double box_L, void *stream, AfvHandle *out) {
    if (!params || !out) { g_create_error = "null argument"; return AFV_ERR_ARG; }
    if (!(params->r > 0.0)) { g_create_error = "r must be positive"; return AFV_ERR_ARG; }
    if (box_L < 0.0 || (box_L > 0.0 && box_L < 4.0 * params->r)) { g_create_error = "periodic box side must be >= 4 r"; return AFV_ERR_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        return AFV_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device ordinal out of range"; return AFV_ERR_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return AFV_ERR_CUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) {
        g_create_error = std::string("device '") + prop.name + "' is not sm_100-class; this library is built for sm_100a only";
        return AFV_ERR_CUDA;
    }
    AfvHandle h = new AfvContext();
    h->device = device; h->prm = *params; h->L = box_L;
    if (stream) { h->stream = (cudaStream_t)stream; h->own_stream = false; }
    else {
        e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); h->stream = nullptr; afv_destroy(h); return AFV_ERR_CUDA; }
        h->own_stream = true;
    }
    if (cudaMalloc((void **)&h->d_grid, sizeof(GridDesc)) != cudaSuccess || cudaMalloc((void **)&h->flags, 8 * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&h->bbox_part, 4 * 1024 * sizeof(double)) != cudaSuccess) {
        g_create_error = "cudaMalloc failed"; afv_destroy(h); return AFV_ERR_CUDA;
    }
    if (cudaMalloc((void **)&h->ovf_cells, OVF_MAX * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&h->ovf_len, OVF_MAX * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&h->ovf_hull, (size_t)OVF_MAX * KPOLY_SLOW * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&h->ovf_m, (size_t)OVF_MAX * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&h->ovf_ring_h, (size_t)OVF_MAX * RS_SLOW * sizeof(double2)) != cudaSuccess ||
        cudaMalloc((void **)&h->ovf_ring_nbr, (size_t)OVF_MAX * RS_SLOW * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&h->ovf_arc_key, (size_t)OVF_MAX * ARC_SLOW * sizeof(int2)) != cudaSuccess ||
        cudaMalloc((void **)&h->ovf_arc_pts, (size_t)OVF_MAX * ARC_SLOW * sizeof(double4)) != cudaSuccess) {
        g_create_error = "cudaMalloc failed"; afv_destroy(h); return AFV_ERR_CUDA;
    }
    if (cudaMalloc((void **)&h->n_all_dev, sizeof(int)) != cudaSuccess || cudaMalloc((void **)&h->xch_work, 8 * sizeof(int)) != cudaSuccess || cudaMemset(h->xch_work, 0, 8 * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&h->pair_cnt, 4 * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&h->pair_dev, 20 * sizeof(double)) != cudaSuccess ||
        cudaHostAlloc((void **)&h->pair_host, 20 * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
        g_create_error = "cudaMalloc failed"; afv_destroy(h); return AFV_ERR_CUDA;
    }
    cudaMemsetAsync(h->flags, 0, 8 * sizeof(int), h->stream);

    *out = h;
    return AFV_OK;
}

void afv_destroy(AfvHandle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_all(h);
    cudaFree(h->d_grid); cudaFree(h->flags); cudaFree(h->bbox_part);
    cudaFree(h->ovf_cells); cudaFree(h->ovf_len); cudaFree(h->ovf_hull); cudaFree(h->ovf_m); cudaFree(h->ovf_ring_h); cudaFree(h->ovf_ring_nbr);
    cudaFree(h->ovf_arc_key); cudaFree(h->ovf_arc_pts);
    cudaFree(h->pair_cnt); cudaFree(h->xch_work); cudaFree(h->n_all_dev); cudaFree(h->pair_dev); cudaFreeHost(h->pair_host);
    for (cudaEvent_t e : {h->ev_sorted, h->ev_late, h->ev_force, h->ev_pack, h->ev_interior, h->ev_counts, h->ev_copy}) if (e) cudaEventDestroy(e);
    for (DevBuf *b : {&h->late_keys, &h->late_keys_sorted, &h->late_perm, &h->late_tmp}) cudaFree(b->p);
    for (DevBuf *b : {&h->scratch_a, &h->scratch_b, &h->noise_buf, &h->conn_buf, &h->ring_types, &h->ring_xy, &h->ring_off,
                      &h->cc_parent, &h->cc_flag, &h->cc_rank, &h->cc_labels, &h->cc_sizes, &h->cc_tmp, &h->cc_edges,
                      &h->sums_partial, &h->fire_state, &h->fire_vel}) cudaFree(b->p);
    for (auto &sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : h->event_pool) cudaEventDestroy(e);
    if (h->step_graph) cudaGraphExecDestroy(h->step_graph);
    cudaFree(h->step_dev);
    for (void *q : h->peer_open) cudaIpcCloseMemHandle(q);
    for (void *q : h->peer_own) cudaFree(q);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int afv_set_params(AfvHandle h, const AfvParams *p) {
    CHECK_H(h);
    if (!p || !(p->r > 0.0)) { h->err = "invalid params"; return AFV_ERR_ARG; }
    if (h->L > 0.0 && h->L < 4.0 * p->r) { h->err = "periodic box side must be >= 4 r"; return AFV_ERR_ARG; }
    h->prm = *p;
    h->built = false;
    return AFV_OK;
}

int64_t afv_num_cells(AfvHandle h) { return h ? h->n - h->n_dead : 0; }
int64_t afv_num_owned(AfvHandle h) { return h ? h->n_owned : 0; }

int afv_set_positions(AfvHandle h, const double *xy_host, int64_t N) {
    CHECK_H(h);
    if (N < 0 || (N > 0 && !xy_host)) { h->err = "invalid positions"; return AFV_ERR_ARG; }
    const bool same = (N == h->n) && h->gid_is_perm && h->n_owned == h->n && h->cap > 0;
    int rc;
    if (!same) {
        if ((rc = reserve(h, N, 0))) return rc;
    }
    const int n = (int)N;
    const int c = h->cur, o = 1 - c;
    const double wrapL = (h->L > 0.0 && !h->slab) ? h->L : 0.0;   // cells outside the box would miss their neighbours across the other face
    if (N > 0) {
        // The caller's (N,2) row-major array is exactly the device layout.  The copy goes first and the host waits for it alone
        // (xy_host is borrowed for the call only); the kernels behind it run while the caller is already issuing the next call.
        double2 *const dst = same ? h->xy[o] : h->xy[c];
        CK(cudaMemcpyAsync(dst, xy_host, (size_t)N * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
        if (!h->ev_copy) CK(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
        CK(cudaEventRecord(h->ev_copy, h->stream));
        if (same) {
            // theta and A0 back to the caller's order, ids and marks reset, positions wrapped: one pass
            k_reset_order<<<nblk(n, 256), 256, 0, h->stream>>>(n, h->th[c], h->a0[c], h->gid[c], h->th[o], h->a0[o], h->gid[o], h->owned[o], h->xy[o], wrapL);
            h->cur = o;
            h->launches += 1;
        } else {
            k_fill<<<nblk(n, 256), 256, 0, h->stream>>>(h->th[c], n, 0.0);
            k_fill<<<nblk(n, 256), 256, 0, h->stream>>>(h->a0[c], n, h->prm.A0);
            k_init_cells<<<nblk(n, 256), 256, 0, h->stream>>>(n, h->gid[c], h->owned[c]);
            h->launches += 3;
            if (wrapL > 0.0) {
                k_wrap_box<<<nblk(n, 256), 256, 0, h->stream>>>(h->xy[c], n, wrapL);
                h->launches += 1;
            }
        }
    }
    h->n = N; h->n_owned = N; h->n_dead = 0; h->gid_is_perm = true; h->built = false;
    CK(cudaGetLastError());
    if (N > 0) CK(cudaEventSynchronize(h->ev_copy));
    return AFV_OK;
}

static int set_per_cell(AfvHandle h, const double *host, int64_t N, double *const *field) {
    if (N != h->n || !host) { h->err = "array length does not match the number of cells"; return AFV_ERR_ARG; }
    if (!h->gid_is_perm) { h->err = "per-cell setters need a handle filled by afv_set_positions"; return AFV_ERR_STATE; }
    if (N == 0) return AFV_OK;
    int rc = ensure_buf(h, h->scratch_a, (size_t)N * sizeof(double));
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->scratch_a.p, host, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_gather_by_gid<<<nblk(N, 256), 256, 0, h->stream>>>((const double *)h->scratch_a.p, h->gid[h->cur], (int)N, field[h->cur]);
    h->launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return AFV_OK;
}

int afv_set_preferred_areas(AfvHandle h, const double *a0_host, int64_t N) {
    CHECK_H(h);
    h->built = false;
    return set_per_cell(h, a0_host, N, h->a0);
}

int afv_set_theta(AfvHandle h, const double *theta_host, int64_t N) {
    CHECK_H(h);
    return set_per_cell(h, theta_host, N, h->th);
}

int afv_build(AfvHandle h, unsigned flags) {
    CHECK_H(h);
    h->n_conn = -1; h->n_ring = -1; h->n_clusters = 0;
    if (h->n == h->n_dead) { h->n = h->n_dead = 0; }   // only stale halo copies left
    h->late_mode = false;
    if (h->n == 0) { h->built = true; h->n_conn = 0; h->n_ring = 0; return AFV_OK; }
    int rc;
    CK(cudaMemsetAsync(h->flags, 0, 4 * sizeof(int), h->stream));
    if ((rc = stage_sort(h))) return rc;
    if ((rc = geometry_checked(h))) return rc;
    if ((rc = stage_force(h, false, 0, 0, 0, 0, 0, 0, nullptr))) return rc;
    if ((rc = check_flags(h))) return rc;
    h->built = true;
    const int n = (int)h->n;
    const int c = h->cur;
    if (flags & AFV_BUILD_CONNECT) {
        StageTimer t(h, 4);
        k_conn_count<<<nblk(n), TPB, 0, h->stream>>>(ring_view(h), h->gid[c], h->owned[c], n, h->cnt);
        size_t tmp = h->cub_tmp_bytes;
        CK(cub::DeviceScan::ExclusiveSum(h->cub_tmp, tmp, h->cnt, h->off, n + 1, h->stream));
        int total = 0;
        CK(cudaMemcpyAsync(&total, h->off + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if ((rc = ensure_buf(h, h->conn_buf, (size_t)std::max(total, 1) * 2 * sizeof(int64_t)))) return rc;
        k_conn_emit<<<nblk(n), TPB, 0, h->stream>>>(ring_view(h), h->gid[c], h->owned[c], n, h->off, (int64_t *)h->conn_buf.p);
        h->n_conn = total;
        h->launches += 3;
    }
    if (flags & AFV_BUILD_RINGS) {
        if (!h->gid_is_perm) { h->err = "ring export needs a handle filled by afv_set_positions"; return AFV_ERR_STATE; }
        StageTimer t(h, 4);
        k_ring_len_by_gid<<<nblk(n, 256), 256, 0, h->stream>>>(ring_view(h), h->gid[c], n, h->cnt);
        CK(cudaMemsetAsync(h->cnt + n, 0, sizeof(int), h->stream));
        size_t tmp = h->cub_tmp_bytes;
        CK(cub::DeviceScan::ExclusiveSum(h->cub_tmp, tmp, h->cnt, h->off, n + 1, h->stream));
        int total = 0;
        CK(cudaMemcpyAsync(&total, h->off + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if ((rc = ensure_buf(h, h->ring_types, (size_t)std::max(total, 1)))) return rc;
        if ((rc = ensure_buf(h, h->ring_xy, (size_t)std::max(total, 1) * 2 * sizeof(double)))) return rc;
        if ((rc = ensure_buf(h, h->ring_off, (size_t)(n + 1) * sizeof(int)))) return rc;
        CK(cudaMemcpyAsync(h->ring_off.p, h->off, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
        k_ring_emit<<<nblk(n), TPB, 0, h->stream>>>(ring_view(h), h->gid[c], h->xy[c], n, (const int *)h->ring_off.p,
                                                    (signed char *)h->ring_types.p, (double *)h->ring_xy.p);
        h->n_ring = total;
        h->launches += 3;
    }
    CK(cudaGetLastError());
    return AFV_OK;
}

int afv_step(AfvHandle h, double dt, double mu, double v0, double Dr, uint64_t seed, int64_t step0, int n_steps,
             const double *noise_host) {
    CHECK_H(h);
    if (n_steps < 0 || Dr < 0.0 || !(dt >= 0.0)) { h->err = "invalid step arguments"; return AFV_ERR_ARG; }
    if (h->n == h->n_dead) { h->n = h->n_dead = 0; }
    if (h->n == 0 || n_steps == 0) return AFV_OK;
    h->late_mode = false;
    int rc;
    const double *noise_dev = nullptr;
    if (noise_host) {
        if (!h->gid_is_perm) { h->err = "a recorded noise stream needs a handle filled by afv_set_positions"; return AFV_ERR_STATE; }
        size_t bytes = (size_t)n_steps * (size_t)h->n * sizeof(double);
        if ((rc = ensure_buf(h, h->noise_buf, bytes))) return rc;
        CK(cudaMemcpyAsync(h->noise_buf.p, noise_host, bytes, cudaMemcpyHostToDevice, h->stream));
        noise_dev = (const double *)h->noise_buf.p;
    }
    const bool deferred = h->defer_checks && !noise_host;
    if (!deferred) CK(cudaMemsetAsync(h->flags, 0, 4 * sizeof(int), h->stream));
    int s_first = 0;
    // A periodic time loop issues the same launches every step and never asks the host anything: after two plain steps
    // (which settle every buffer) the rest is replayed as a CUDA graph of two steps -- the double buffers are back where
    // they started -- with the step index kept on the device.  What it saves is the idle time in front of every launch
    // (8 launches, ~2 us each, per step at N = 10^6; most of the step at N = 10^4).
    static const bool no_graph = getenv("AFV_NO_GRAPH") != nullptr;
    if (!no_graph && !h->step_graph_off && h->L > 0.0 && !noise_host && !deferred && !h->slab && !h->profiling && n_steps >= 6) {
        for (; s_first < 2; ++s_first) {
            if ((rc = stage_sort(h))) return rc;
            if ((rc = stage_geometry(h))) return rc;
            if ((rc = stage_force(h, true, dt, mu, v0, Dr, seed, step0 + s_first, nullptr))) return rc;
        }
        char key[512];
        snprintf(key, sizeof key, "%lld %lld %d %d %d %llu %a %a %a %a %llu %a %a %a %a %a %a %a %a", (long long)h->n, (long long)h->cap, h->cur, h->cmax_hint, (int)h->gid_is_perm,
                 (unsigned long long)h->alloc_epoch, dt, mu, v0, Dr, (unsigned long long)seed, h->L, h->prm.r, h->prm.A0, h->prm.P0, h->prm.KA, h->prm.KP,
                 h->prm.Lambda, h->prm.delta);
        if (!h->step_dev) CK(cudaMalloc((void **)&h->step_dev, sizeof(long long)));
        if (!h->step_graph || h->step_graph_key != key) {
            if (h->step_graph) { cudaGraphExecDestroy(h->step_graph); h->step_graph = nullptr; }
            cudaGraph_t g = nullptr;
            const int64_t launches0 = h->launches;
            cudaError_t ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
            if (ce == cudaSuccess) {
                h->step_on_device = true;
                rc = AFV_OK;
                for (int k = 0; k < 2 && rc == AFV_OK; ++k) {
                    if ((rc = stage_sort(h)) == AFV_OK && (rc = stage_geometry(h)) == AFV_OK) rc = stage_force(h, true, dt, mu, v0, Dr, seed, 0, nullptr);
                }
                ce = cudaStreamEndCapture(h->stream, &g);
                h->step_on_device = false;
                h->launches = launches0;               // (nothing has run yet)
                if (ce == cudaSuccess && rc == AFV_OK) ce = cudaGraphInstantiate(&h->step_graph, g, 0);
                if (g) cudaGraphDestroy(g);
            }
            if (rc != AFV_OK) return rc;               // (a stage refused while being captured: its message stands)
            if (ce != cudaSuccess) {                   // a stream that cannot be captured, ...: this handle steps the plain way from now on
                cudaGetLastError();
                h->step_graph = nullptr; h->step_graph_off = true;
            } else h->step_graph_key = key;
        }
        const int pairs = h->step_graph ? (n_steps - s_first) / 2 : 0;
        if (pairs > 0) {
            const long long before = step0 + s_first - 1;  // k_bin_count advances the index at the start of every step
            CK(cudaMemcpyAsync(h->step_dev, &before, sizeof before, cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));      // (`before` lives on this stack frame)
        }
        for (int p = 0; p < pairs; ++p) CK(cudaGraphLaunch(h->step_graph, h->stream));
        h->launches += (int64_t)pairs * 2 * 7;         // per step, as counted outside the graph: cell list 3, k_nbrs, k_clip, k_geometry_slow, k_force
        s_first += 2 * pairs;
    }
    for (int s = s_first; s < n_steps; ++s) {
        // (the slow-path queue, flags[0], restarts with every cell list: k_bin_count; the error counters run on)
        if ((rc = stage_sort(h))) return rc;
        // open boundaries: the local density is unknown, so the capacity is checked (and grown) before the positions move
        if ((rc = (h->L > 0.0 ? stage_geometry(h) : geometry_checked(h)))) return rc;
        if ((rc = stage_force(h, true, dt, mu, v0, Dr, seed, step0 + s, noise_dev ? noise_dev + (size_t)s * h->n : nullptr))) return rc;
    }
    if (!deferred && (rc = check_flags(h))) return rc;   // also synchronises: noise_host is borrowed for the call only
    h->built = true;
    h->n_conn = -1; h->n_ring = -1; h->n_clusters = 0;
    return AFV_OK;
}

double afv_philox_normal(uint64_t seed, int64_t cell_id, int64_t step) { return philox_normal(seed, cell_id, step); }

int afv_get_positions(AfvHandle h, double *out) {
    CHECK_H(h);
    if (h->n == 0) return AFV_OK;
    if (!h->gid_is_perm) { h->err = "results in caller order are only defined for handles filled by afv_set_positions"; return AFV_ERR_STATE; }
    const int n = (int)h->n;
    int rc = ensure_buf(h, h->scratch_a, (size_t)n * sizeof(double2));
    if (rc) return rc;
    k_unsort_xy<<<nblk(n, 256), 256, 0, h->stream>>>(h->xy[h->cur], h->gid[h->cur], n, (double2 *)h->scratch_a.p);
    h->launches += 1;
    CK(cudaMemcpyAsync(out, h->scratch_a.p, (size_t)n * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return AFV_OK;
}
int afv_get_theta(AfvHandle h, double *out) {
    CHECK_H(h);
    if (h->n == 0) return AFV_OK;
    return get_unsorted(h, h->th[h->cur], out, 1, 1, nullptr);
}
int afv_get_forces(AfvHandle h, double *out) {
    CHECK_H(h);
    int rc = require_built(h); if (rc) return rc;
    if (h->n == 0) return AFV_OK;
    return get_unsorted(h, h->fx, out, 2, 2, h->fy);
}
int afv_get_areas(AfvHandle h, double *out) {
    CHECK_H(h);
    int rc = require_built(h); if (rc) return rc;
    if (h->n == 0) return AFV_OK;
    return get_unsorted(h, h->area, out, 1, 1, nullptr);
}
int afv_get_perimeters(AfvHandle h, double *out) {
    CHECK_H(h);
    int rc = require_built(h); if (rc) return rc;
    if (h->n == 0) return AFV_OK;
    return get_unsorted(h, h->perim, out, 1, 1, nullptr);
}
int afv_get_arclens(AfvHandle h, double *out) {
    CHECK_H(h);
    int rc = require_built(h); if (rc) return rc;
    if (h->n == 0) return AFV_OK;
    return get_unsorted(h, h->arcl, out, 1, 1, nullptr);
}
int afv_get_coord_nums(AfvHandle h, int64_t *out) {
    CHECK_H(h);
    int rc = require_built(h); if (rc) return rc;
    if (h->n == 0) return AFV_OK;
    if (!h->gid_is_perm) { h->err = "results in caller order are only defined for handles filled by afv_set_positions"; return AFV_ERR_STATE; }
    const int n = (int)h->n;
    if ((rc = ensure_buf(h, h->scratch_a, (size_t)n * sizeof(int64_t)))) return rc;
    k_unsort_i32_to_i64<<<nblk(n, 256), 256, 0, h->stream>>>(h->coord, h->gid[h->cur], n, (int64_t *)h->scratch_a.p);
    h->launches += 1;
    CK(cudaMemcpyAsync(out, h->scratch_a.p, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return AFV_OK;
}

int afv_get_num_connections(AfvHandle h, int64_t *K) {
    CHECK_H(h);
    if (h->n_conn < 0) { h->err = "no contact list: call afv_build with AFV_BUILD_CONNECT"; return AFV_ERR_STATE; }
    *K = h->n_conn;
    return AFV_OK;
}
int afv_get_connections(AfvHandle h, int64_t *pairs) {
    CHECK_H(h);
    if (h->n_conn < 0) { h->err = "no contact list: call afv_build with AFV_BUILD_CONNECT"; return AFV_ERR_STATE; }
    if (h->n_conn > 0) {
        CK(cudaMemcpyAsync(pairs, h->conn_buf.p, (size_t)h->n_conn * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return AFV_OK;
}

int afv_get_num_ring_entries(AfvHandle h, int64_t *T) {
    CHECK_H(h);
    if (h->n_ring < 0) { h->err = "no ring export: call afv_build with AFV_BUILD_RINGS"; return AFV_ERR_STATE; }
    *T = h->n_ring;
    return AFV_OK;
}
int afv_get_rings(AfvHandle h, int64_t *region_off, int8_t *edge_types, double *vertices_xy) {
    CHECK_H(h);
    if (h->n_ring < 0) { h->err = "no ring export: call afv_build with AFV_BUILD_RINGS"; return AFV_ERR_STATE; }
    const int64_t n = h->n;
    if (n == 0) { region_off[0] = 0; return AFV_OK; }
    std::vector<int> off((size_t)n + 1);
    CK(cudaMemcpyAsync(off.data(), h->ring_off.p, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (h->n_ring > 0) {
        CK(cudaMemcpyAsync(edge_types, h->ring_types.p, (size_t)h->n_ring, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(vertices_xy, h->ring_xy.p, (size_t)h->n_ring * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i <= n; ++i) region_off[i] = off[(size_t)i];
    return AFV_OK;
}

int afv_get_diagnostics(AfvHandle h, int64_t out[4]) {
    CHECK_H(h);
    for (int i = 0; i < 4; ++i) out[i] = h->h_flags[i];
    return AFV_OK;
}

int afv_set_profiling(AfvHandle h, int enabled) {
    CHECK_H(h);
    collect_spans(h);
    h->profiling = enabled != 0;
    return AFV_OK;
}

int afv_get_stage_times(AfvHandle h, double ms_out[AFV_N_STAGES], int64_t launches_out[AFV_N_STAGES], int reset) {
    CHECK_H(h);
    collect_spans(h);
    for (int i = 0; i < AFV_N_STAGES; ++i) { ms_out[i] = h->stage_ms[i]; launches_out[i] = h->stage_launches[i]; }
    if (reset) for (int i = 0; i < AFV_N_STAGES; ++i) { h->stage_ms[i] = 0; h->stage_launches[i] = 0; }
    return AFV_OK;
}

int64_t afv_launch_count(AfvHandle h) { return h ? h->launches : 0; }

void *afv_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (bytes <= 0) bytes = 1;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void afv_host_free(void *p) { if (p) cudaFreeHost(p); }

int afv_measure_fp64_peak(AfvHandle h, double *tflops_out) {
    CHECK_H(h);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    int rc = ensure_buf(h, h->scratch_b, (size_t)blocks * threads * sizeof(double));
    if (rc) return rc;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, h->stream);
        k_dfma_peak<<<blocks, threads, 0, h->stream>>>((double *)h->scratch_b.p, iters);
        cudaEventRecord(b, h->stream);
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        double fl = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    h->launches += 5;
    *tflops_out = best;
    return AFV_OK;
}

void *afv_device_ptr(AfvHandle h, int field) {
    if (!h) return nullptr;
    const int c = h->cur;
    switch (field) {
        case AFV_F_XY: return h->xy[c];
        case AFV_F_THETA: return h->th[c];
        case AFV_F_A0: return h->a0[c];
        case AFV_F_GID: return h->gid[c];
        case AFV_F_FX: return h->fx;
        case AFV_F_FY: return h->fy;
        case AFV_F_AREA: return h->area;
        case AFV_F_PERIM: return h->perim;
        case AFV_F_ARCLEN: return h->arcl;
        case AFV_F_COORD: return h->coord;
        case AFV_F_OWNED: return h->owned[c];
        default: return nullptr;
    }
}

void *afv_stream(AfvHandle h) { return h ? (void *)h->stream : nullptr; }

int afv_synchronize(AfvHandle h) {
    CHECK_H(h);
    if (h->defer_checks) return check_flags(h);   // synchronises and reports deferred capacity errors
    CK(cudaStreamSynchronize(h->stream));
    return AFV_OK;
}

int afv_set_deferred_checks(AfvHandle h, int enabled) {
    CHECK_H(h);
    h->defer_checks = enabled != 0;
    return AFV_OK;
}

// ---- slab sharding primitives ----
#include "afv_shard.inc"

// ---- cluster analysis on the device ----
#include "afv_cluster.inc"

// ---- energy and FIRE relaxation on the device ----
#include "afv_relax.inc"

}  // extern "C"
