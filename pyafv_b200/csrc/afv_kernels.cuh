// The kernels of the AFV hot path (sm_100a): bin -> (radix sort) -> reorder -> geometry -> force(+step).
#pragma once
#include "afv_device.cuh"

namespace afv {

struct PhysDev {
    double r, P0, KA, KP, Lambda, delta;
};

// ------------------------------------------------------------------------------------------------------
// K0: bounding box (open boundaries) -> grid description, entirely on the device
// ------------------------------------------------------------------------------------------------------
__global__ void k_bbox_partial(const double2 *__restrict__ xy, int n, double *__restrict__ part) {
    // part: [4 * gridDim.x] = xmin, xmax, ymin, ymax per block
    __shared__ double sm[4][32];
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 q = xy[i];
        const double a = q.x, b = q.y;
        xmin = fmin(xmin, a); xmax = fmax(xmax, a); ymin = fmin(ymin, b); ymax = fmax(ymax, b);
    }
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sm[0][w] = xmin; sm[1][w] = xmax; sm[2][w] = ymin; sm[3][w] = ymax; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        xmin = l < nw ? sm[0][l] : 1e300; xmax = l < nw ? sm[1][l] : -1e300;
        ymin = l < nw ? sm[2][l] : 1e300; ymax = l < nw ? sm[3][l] : -1e300;
        for (int o = 16; o > 0; o >>= 1) {
            xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
        if (l == 0) { part[4 * blockIdx.x] = xmin; part[4 * blockIdx.x + 1] = xmax; part[4 * blockIdx.x + 2] = ymin; part[4 * blockIdx.x + 3] = ymax; }
    }
}

__global__ void k_bbox_final(const double *__restrict__ part, int nblocks, double r, int nbins_cap, GridDesc *__restrict__ g) {
    // one warp
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int i = threadIdx.x; i < nblocks; i += 32) {
        xmin = fmin(xmin, part[4 * i]); xmax = fmax(xmax, part[4 * i + 1]);
        ymin = fmin(ymin, part[4 * i + 2]); ymax = fmax(ymax, part[4 * i + 3]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (threadIdx.x == 0) {
        const double cut = 2.0 * r;
        double sx = xmax - xmin, sy = ymax - ymin;
        double fx = floor(sx / cut) + 1.0, fy = floor(sy / cut) + 1.0;
        // keep nbx * nby <= nbins_cap by coarsening the larger direction
        while (fx * fy > (double)nbins_cap) { if (fx >= fy) fx = ceil(fx * 0.5); else fy = ceil(fy * 0.5); }
        int nbx = (int)fx, nby = (int)fy;
        double bsx = fmax(cut, sx / (double)nbx), bsy = fmax(cut, sy / (double)nby);
        g->x0 = xmin; g->y0 = ymin; g->inv_bx = 1.0 / bsx; g->inv_by = 1.0 / bsy;
        g->Lx = g->Ly = 0.0; g->halfLx = g->halfLy = 0.0; g->nbx = nbx; g->nby = nby; g->perx = 0; g->pery = 0; g->swap = 0; g->mix = 0; g->seam_lo = 0; g->seam_hi = -1;
    }
}

// ------------------------------------------------------------------------------------------------------
// K1: the cell list is a COUNTING sort by bin -- three launches: (a) bin key of every cell + histogram, (b) exclusive
// scan of the histogram = first slot of every bin, (c) scatter of the SoA state into bin order.  key = by * nbx + bx
// (row-major, so that the 3 bins of a row are one contiguous run of slots).  The order of the cells INSIDE a bin is
// whatever the atomics make it; nothing downstream depends on it (k_clip's results are a function of the set of
// neighbours, not of the order they are met in -- test_late_cell_exchange_equals_plain).
// Cells marked CELL_DEAD (stale halo copies, see k_exchange_pack) get `dead_key`, the bin behind every real one: they
// land behind the live cells, where the scatter drops them.
// ------------------------------------------------------------------------------------------------------
// n_dev (may be null): the cell count lives on the device (the host launched for an upper bound and has not read it yet)
__global__ void k_bin_count(const double2 *__restrict__ xy, const unsigned char *__restrict__ owned, int n, const int *__restrict__ n_dev,
                            const GridDesc *__restrict__ gp, unsigned dead_key, unsigned *__restrict__ keys, int *__restrict__ bin_cnt,
                            int *__restrict__ queue_count, long long *__restrict__ step_counter) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        *queue_count = 0;                    // the queue of the slow path restarts with every cell list
        if (step_counter) *step_counter += 1;   // replayed time loops (CUDA graph): the step index lives on the device; k_force reads it
    }
    if (n_dev) n = *n_dev;
    if (i >= n) return;
    unsigned key = dead_key;
    if (owned[i] != CELL_DEAD) {
        const GridDesc g = *gp;
        const double2 q = xy[i];
        const int bx = bin_col(g, q.x);
        int by = (int)floor((q.y - g.y0) * g.inv_by);
        by = min(max(by, 0), g.nby - 1);
        key = bin_key(g, bx, by);
    }
    keys[i] = key;
    // (dead cells are not counted: they would all hit ONE counter -- the 23 000 retired halo copies of an 8-GPU slab cost
    // 0.03 ms per step that way, here and in the scatter -- and nothing reads their count)
    if (key != dead_key) atomicAdd(&bin_cnt[key], 1);
}

// (c): cell i takes the slot bin_start[key] + (count left in its bin) - 1; the counts are back at zero afterwards.  Dead
// cells are dropped.  slot_of_gid (may be null): where the caller's cell gid went (fixed-order reductions).
__global__ void k_bin_scatter(const unsigned *__restrict__ keys, int n, const int *__restrict__ n_dev, unsigned dead_key,
                              const int *__restrict__ bin_start, int *__restrict__ bin_cnt,
                              const double2 *__restrict__ xy0, const double *__restrict__ t0, const double *__restrict__ a0,
                              const int64_t *__restrict__ g0, const unsigned char *__restrict__ o0, double2 *__restrict__ xy1,
                              double *__restrict__ t1, double *__restrict__ a1, int64_t *__restrict__ g1, unsigned char *__restrict__ o1,
                              unsigned *__restrict__ keys_sorted, int *__restrict__ slot_of_gid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n = *n_dev;
    if (i >= n) return;
    const unsigned key = keys[i];
    if (key == dead_key) return;
    const int left = atomicSub(&bin_cnt[key], 1);
    const int s = bin_start[key] + left - 1;
    const int64_t g = g0[i];
    xy1[s] = xy0[i]; t1[s] = t0[i]; a1[s] = a0[i]; g1[s] = g; o1[s] = o0[i];
    keys_sorted[s] = key;
    if (slot_of_gid) slot_of_gid[g] = s;
}

// gather the SoA state through a permutation (compaction paths)
__global__ void k_reorder(const int *__restrict__ perm, int n, const double2 *__restrict__ xy0,
                          const double *__restrict__ t0, const double *__restrict__ a0, const int64_t *__restrict__ g0,
                          const unsigned char *__restrict__ o0, double2 *__restrict__ xy1,
                          double *__restrict__ t1, double *__restrict__ a1, int64_t *__restrict__ g1,
                          unsigned char *__restrict__ o1) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int p = perm[s];
    xy1[s] = xy0[p]; t1[s] = t0[p]; a1[s] = a0[p]; g1[s] = g0[p]; o1[s] = o0[p];
}


#include "afv_geometry.cuh"

// ------------------------------------------------------------------------------------------------------
// K3: force gather (+ fused active-Langevin update).  The ring of every owned cell (slot-major, coalesced) is walked
// and ONE 32-byte record (position, cA = 2KA(A-A0), cP = 2KP(P-P0)) is gathered per contact neighbour.
//
// The reference assembles F_i = -sum_h (dE/dh)(dh/dr_i) with dE/dh = sum over the (<= 3) cells c around vertex h of
// cA_c dA_c/dh + cP_c dP_c/dh (cell_geom.pyx:478-518, finite_voronoi.py:537-634).  dA_c/dh, dP_c/dh involve c's two
// ring neighbours of h, one of which cell i does not see (the far end b of the edge between the two OTHER cells).
// For an inner vertex h = circumcentre(i, a, b) that end is not needed: dh/dr_i is parallel to the bisector of (a, b),
// which carries h and b, so   cross(b, dh) = cross(h, dh)   and   d|h - b| = -(d^ . dh)   with d^ the unit vector
// along that bisector pointing away from cell i.  Everything else -- the far ends of i's own two edges, the three
// positions -- cell i holds.  Outer vertices (circle-circle intersections) keep the reference's regularised Jacobian
// literally (finite_voronoi.py:653-746), so there the neighbour's arc end IS needed: it comes from the neighbour's
// arc table (<= 1.1 look-ups per cell at rho l^2 = 1, none in a confluent tissue).
// No atomics; the summation order is the ring order, which starts at the nearest neighbour: results do not depend on
// the sort order.
// ------------------------------------------------------------------------------------------------------
struct ForceArgs {
    const double2 *ring_h;         // slot-major
    const int *ring_nbr;
    size_t stride;
    const unsigned char *ring_len;
    const CellRec *rec;
    const unsigned char *n_arc;
    const int2 *arc_key;
    const double4 *arc_pts;
    const int *ovf_index, *ovf_len;
    const double2 *ovf_ring_h;
    const int *ovf_ring_nbr;
    const int2 *ovf_arc_key;
    const double4 *ovf_arc_pts;
    const GridDesc *grid;
    const unsigned char *owned;
    const int64_t *gid;
    int n;
    PhysDev ph;
    double *fx, *fy;
    int *flags;
    // step (do_step != 0)
    int do_step;
    double2 *xy;
    double *theta;
    double dt, mu, v0, noise_amp;  // noise_amp = sqrt(2 Dr dt)
    uint64_t seed;
    int64_t step;
    const long long *step_dev;     // not null: the step index is read from the device (time loop replayed as a CUDA graph)
    const double *noise;           // nullptr -> Philox; else row of this step, indexed by gid
    double Lx, Ly;                 // periodic wrap of the updated positions per direction (0 = none)
    // minimum image of the displacements per direction (period, half period; 0 = none).  By value: the kernel reads them
    // from the constant bank instead of keeping a copy of the grid description in registers
    double mi_Lx, mi_hx, mi_Ly, mi_hy;
};

// offset of x to the lower face `lo` of a slab of width W; with a box period L > 0 the positions are the box's own
// (wrapped into [0, L)) and the offset is the one whose distance to the slab CENTRE is the minimum image (so that "left
// of the slab" and "right of the slab" stay apart even when the slab is half the box)
__device__ __forceinline__ double slab_offset(double x, double lo, double W, double L) {
    double d = x - (lo + 0.5 * W);
    if (L > 0.0) { if (d >= 0.5 * L) d -= L; else if (d < -0.5 * L) d += L; }
    return d + 0.5 * W;
}
// exchange lists of an owned cell at offset xr = slab_offset(x) in a slab of width W (bit q set = member of list q, see
// k_exchange_count)
__device__ __forceinline__ unsigned xch_lists(double xr, unsigned char own, double W, double halo) {
    if (own != CELL_OWNED) return 0u;
    return (xr < 0.0 ? 1u : 0u) | ((xr >= 0.0 && xr < halo) ? 2u : 0u) | (xr >= W ? 4u : 0u) | ((xr >= W - halo && xr < W) ? 8u : 0u);
}

// far end (relative to cell j) of j's arc that STARTS where j's contact edge with `self` ends (starts_there) or that
// ENDS where j's contact edge with `self` starts (!starts_there)
__device__ __forceinline__ bool arc_far_end(const ForceArgs &a, int j, int self, bool starts_there, double &qx, double &qy) {
    // everything that depends on j alone is requested at once: one round trip, then the matching arc's end points
    const int na = a.n_arc[j];
    const int rl = a.ring_len[j];
    const int4 *const kp = reinterpret_cast<const int4 *>(a.arc_key + (size_t)j * ARC_MAX);
    const int4 k01 = kp[0], k23 = kp[1];
    qx = qy = 0.0;
    if (rl != RING_IN_POOL) {
        const int m0 = starts_there ? k01.x : k01.y, m1 = starts_there ? k01.z : k01.w, m2 = starts_there ? k23.x : k23.y, m3 = starts_there ? k23.z : k23.w;
        int q = -1;
        if (na > 4) {   // (rare: the second half of the table)
            const int4 k45 = kp[2], k67 = kp[3];
            const int m4 = starts_there ? k45.x : k45.y, m5 = starts_there ? k45.z : k45.w, m6 = starts_there ? k67.x : k67.y, m7 = starts_there ? k67.z : k67.w;
            if (na > 7 && m7 == self) q = 7;
            if (na > 6 && m6 == self) q = 6;
            if (na > 5 && m5 == self) q = 5;
            if (m4 == self) q = 4;
        }
        if (na > 3 && m3 == self) q = 3;
        if (na > 2 && m2 == self) q = 2;
        if (na > 1 && m1 == self) q = 1;
        if (na > 0 && m0 == self) q = 0;
        if (q < 0) return false;
        const double4 c = a.arc_pts[(size_t)j * ARC_MAX + q];
        qx = starts_there ? c.z : c.x; qy = starts_there ? c.w : c.y;
        return true;
    }
    const int pi = a.ovf_index[j];
    const int2 *const ak = a.ovf_arc_key + (size_t)pi * ARC_SLOW;
    const double4 *const ap = a.ovf_arc_pts + (size_t)pi * ARC_SLOW;
    for (int q = 0; q < na; ++q) {
        const int2 k = ak[q];
        if ((starts_there ? k.x : k.y) == self) {
            const double4 c = ap[q];
            qx = starts_there ? c.z : c.x; qy = starts_there ? c.w : c.y;
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------------
// K3, entry-parallel.  The terms of dE/dr_i are independent per ring vertex, so a cell's ring is split over FB_ROWS
// threads: a block takes FB_CELLS consecutive cells, thread (cell c, row w) does the entries w, w + FB_ROWS, ... of
// cell c (entry e of 32 consecutive cells is one coalesced load per operand in the slot-major ring).  Every entry is
// loaded from scratch -- no rolling window over the ring -- so the kernel needs 64 registers instead of 80 + spills
// and a thread keeps all the gathers of an entry in flight at once; the serial ring walk it replaces was bound by its
// chain of dependent loads (52 % long-scoreboard stalls at 30 % issue utilisation).  Measured at N = 1e6 (ms):
// rows x cells x blocks/SM  1x128x6 0.246, 2x64x8 0.238, 2x32x16 0.237, 4x64x4 0.329, 8x32x4 0.491; serial walk 0.272.
// Inner vertices (circumcentres) are finished on the spot; outer vertices (circle-circle intersections: a long code
// path that one entry in six takes) are queued in shared memory and then done by the whole block with all lanes
// busy.  The terms meet in a shared-memory table [entry][cell]; one thread per cell adds them up IN RING ORDER -- the
// sum does not depend on who computed what -- and does the Langevin update.  No atomics on the result.
// ------------------------------------------------------------------------------------------------------
#ifndef FB_CELLS_
#define FB_CELLS_ 64
#endif
#ifndef FB_ROWS_
#define FB_ROWS_ 2
#endif
#ifndef FORCE_MIN_BLOCKS
#define FORCE_MIN_BLOCKS 8
#endif
constexpr int FB_CELLS = FB_CELLS_, FB_ROWS = FB_ROWS_, FB_THREADS = FB_CELLS * FB_ROWS;

struct RingRef { const double2 *h; const int *nb; size_t st; int E; };

__device__ __forceinline__ double force_inv_len(double ax, double ay, double bx, double by) {
    const double sx = ax - bx, sy = ay - by, l2 = sx * sx + sy * sy;
    return l2 > 0.0 ? rsqrt(l2) : 0.0;
}
// neighbour j seen from the cell with record mec: minimum-image displacement and prefactors
__device__ __forceinline__ double4 force_fetch(const ForceArgs &a, const double4 mec, int j) {
    double4 v = *reinterpret_cast<const double4 *>(a.rec + j);
    v.x -= mec.x; v.y -= mec.y;
    if (a.mi_Lx > 0.0) v.x = min_image(v.x, a.mi_Lx, a.mi_hx);
    if (a.mi_Ly > 0.0) v.y = min_image(v.y, a.mi_Ly, a.mi_hy);
    return v;
}
// inner vertex h = hc = circumcentre(0, pa, pb) between the contact edges with bp (arriving) and bc (leaving):
// dh = eps (h . dr_i) / (pa . eps), eps = perp(pb - pa)   (see the derivation at the top of K3)
__device__ __forceinline__ double2 force_inner_term(const double4 mec, const double4 bp, const double4 bc, const double2 hp, const double2 hc,
                                                    const double2 hn) {
    const double cAi = mec.z, cPi = mec.w;
    const double hx = hc.x, hy = hc.y;
    const double ilp = force_inv_len(hp.x, hp.y, hx, hy), ilc = force_inv_len(hx, hy, hn.x, hn.y);
    const double epx = -(bc.y - bp.y), epy = bc.x - bp.x;
    const double den = bp.x * epx + bp.y * epy;
    // sum_c cA_c dA_c/dh . eps  with the unseen far end replaced by h
    const double Sx = cAi * (hp.x - hn.x) + bp.z * (hx - hp.x) + bc.z * (hn.x - hx);
    const double Sy = cAi * (hp.y - hn.y) + bp.z * (hy - hp.y) + bc.z * (hn.y - hy);
    // unit vectors from the far ends of i's two edges towards h
    const double eax = (hx - hp.x) * ilp, eay = (hy - hp.y) * ilp;
    const double ebx = (hx - hn.x) * ilc, eby = (hy - hn.y) * ilc;
    const double e2n = epx * epx + epy * epy;
    const double elen = e2n * rsqrt(e2n);
    const double Ge = -0.5 * (Sx * epy - Sy * epx) + (cPi + bp.w) * (eax * epx + eay * epy) + (cPi + bc.w) * (ebx * epx + eby * epy);
    // edge (a, b) leaves h away from cell i: its unit vector is sign(den) eps / |eps|
    const double coef = (Ge - (bp.w + bc.w) * (den > 0.0 ? elen : -elen)) / den;
    return make_double2(coef * hx, coef * hy);
}
// Outer vertex at hc of the cell in slot cs (record mec), between the entries hp and hn: the contact edge with j (bj: its
// displacement and prefactors) starts (entry) or ends (!entry) here  (finite_voronoi.py:653-746)
__device__ __forceinline__ double2 force_outer_term(const ForceArgs &a, const int cs, const double4 mec, const bool entry, const int j, const double4 bj,
                                                    const double2 hp, const double2 hc, const double2 hn, int &misses) {
    const double r = a.ph.r, r2 = r * r;
    const double cAc = mec.z, cPc = mec.w, cLc = mec.w + a.ph.Lambda;
    const double hx = hc.x, hy = hc.y;
    const double fxe = entry ? hn.x : hp.x, fye = entry ? hn.y : hp.y;     // far end of the contact edge
    const double oxe = entry ? hp.x : hn.x, oye = entry ? hp.y : hn.y;     // other end of i's arc
    const double pjx = bj.x, pjy = bj.y, cAj = bj.z, cPj = bj.w;
    const double il = force_inv_len(hx, hy, fxe, fye);
    double qx, qy;                            // other end of j's arc at h, relative to j
    misses += !arc_far_end(a, j, cs, entry, qx, qy);
    const double ujx = hx - pjx, ujy = hy - pjy;                           // h relative to j
    const double Qx = qx + pjx, Qy = qy + pjy;                             // ... relative to i
    // unit vector from the far end of the contact edge towards h: dP/dh of both cells along the edge
    const double tx = (hx - fxe) * il, ty = (hy - fye) * il;
    // polygon parts (cell_geom.pyx:478-518): dA/dv1 = 0.5 (y0 - y2, x2 - x0) with (v0, v2) the ring neighbours of
    // h -- for i: (hp, hn); for j: (far end, Q) when the edge arrives at h in j's ring (entry), (Q, far end) else
    const double sgn = entry ? 1.0 : -1.0;
    const double Gx = cAc * 0.5 * (hp.y - hn.y) + cAj * 0.5 * sgn * (fye - Qy) + (cPc + cPj) * tx;
    const double Gy = cAc * 0.5 * (hn.x - hp.x) + cAj * 0.5 * sgn * (Qx - fxe) + (cPc + cPj) * ty;
    // arc-end weights (cell_geom.pyx:521-539): +-(cA/2 (r^2 - v1 . v2) + (cP + Lambda) r), + where the arc starts
    const double wi = cAc * 0.5 * (r2 - (hx * oxe + hy * oye)) + cLc * r;
    const double wj = cAj * 0.5 * (r2 - (ujx * qx + ujy * qy)) + (cPj + a.ph.Lambda) * r;
    const double Wi = entry ? -wi : wi, Wj = entry ? wj : -wj;
    const double rho2 = pjx * pjx + pjy * pjy, irho = rsqrt(rho2);
    const double root = sqrt(fmax(4.0 * r2 - rho2, 0.0));
    const double cr = pjx * hy - pjy * hx;
    const double sg = (cr > 0.0) ? 1.0 : ((cr < 0.0) ? -1.0 : 0.0);
    const double k1 = 2.0 * r2 * irho * irho * irho / fmax(root, a.ph.delta);
    const double k2 = 0.5 * root * irho;
    // T_x = dx_terms, T_y = dy_terms of the reference with r_IJ = -p_j
    const double Txx = -k1 * pjx * pjy,       Txy = k1 * pjx * pjx - k2;
    const double Tyx = -k1 * pjy * pjy + k2,  Tyy = k1 * pjx * pjy;
    // dh/dx_i = (1/2 + s Txx, s Txy), dh/dy_i = (s Tyx, 1/2 + s Tyy); dh/dr_j = e - dh/dr_i
    const double hxi_x = 0.5 + sg * Txx, hxi_y = sg * Txy, hyi_x = sg * Tyx, hyi_y = 0.5 + sg * Tyy;
    const double hxj_x = 0.5 - sg * Txx, hxj_y = -sg * Txy, hyj_x = -sg * Tyx, hyj_y = 0.5 - sg * Tyy;
    // arc angles: theta_i of h about i, theta_j of h about j
    const double iui = 1.0 / (hx * hx + hy * hy), iuj = 1.0 / (ujx * ujx + ujy * ujy);
    const double upix = -hy * iui, upiy = hx * iui, upjx = -ujy * iuj, upjy = ujx * iuj;
    const double dti_x = -(hxj_x * upix + hxj_y * upiy), dti_y = -(hyj_x * upix + hyj_y * upiy);
    const double dtj_x = hxi_x * upjx + hxi_y * upjy, dtj_y = hyi_x * upjx + hyi_y * upjy;
    return make_double2((Gx * hxi_x + Gy * hxi_y) + (Wi * dti_x + Wj * dtj_x), (Gx * hyi_x + Gy * hyi_y) + (Wi * dti_y + Wj * dtj_y));
}
// the term of ring entry e of the cell in slot cs, whatever its kind (pooled rings, and the outer vertices of the queue)
__device__ __forceinline__ double2 force_entry_term(const ForceArgs &a, const int cs, const double4 mec, const RingRef R, const int e, int &misses) {
    const int e0 = (e == 0) ? R.E - 1 : e - 1, e1 = (e + 1 == R.E) ? 0 : e + 1;
    const double2 hp = R.h[(size_t)e0 * R.st], hc = R.h[(size_t)e * R.st], hn = R.h[(size_t)e1 * R.st];
    const int np_ = R.nb[(size_t)e0 * R.st], nc = R.nb[(size_t)e * R.st];
    if (np_ != ARC && nc != ARC) return force_inner_term(mec, force_fetch(a, mec, np_), force_fetch(a, mec, nc), hp, hc, hn);
    if ((np_ != ARC) != (nc != ARC)) {
        const bool entry = (nc != ARC);
        const int j = entry ? nc : np_;
        return force_outer_term(a, cs, mec, entry, j, force_fetch(a, mec, j), hp, hc, hn, misses);
    }
    return make_double2(0.0, 0.0);
}

__global__ void __launch_bounds__(FB_THREADS, FORCE_MIN_BLOCKS) k_force(ForceArgs a) {
    __shared__ double2 s_term[RS][FB_CELLS];          // [entry][cell of the block]
    __shared__ unsigned short s_queue[FB_CELLS * RS]; // outer vertices: cell << 4 | entry
    __shared__ unsigned char s_len[FB_CELLS];
    __shared__ int s_nq;
    const int c = threadIdx.x & (FB_CELLS - 1), row0 = threadIdx.x / FB_CELLS;
    const int s = blockIdx.x * FB_CELLS + c;
    // (both loads at once: the ring length is not held back until the ownership mark has arrived)
    const int own_ = s < a.n ? a.owned[s] : CELL_HALO, len_ = s < a.n ? a.ring_len[s] : 0;
    const bool live = own_ == CELL_OWNED;
    const int Eraw = live ? len_ : 0;
    const bool pooled = Eraw == RING_IN_POOL;
    const int E = (pooled || Eraw < 2) ? 0 : Eraw;    // entries of the fast-path ring done in parallel (a pooled ring: by its cell, below)
    if (threadIdx.x == 0) s_nq = 0;
    if (row0 == 0) s_len[c] = (unsigned char)E;
    double4 mec = make_double4(0.0, 0.0, 0.0, 0.0);
    if (live) mec = *reinterpret_cast<const double4 *>(a.rec + s);
    // what the update at the end needs is requested now: its round trip overlaps everything in between
    double th = 0.0, xi_ = 0.0;
    double2 pos = make_double2(0.0, 0.0);
    if (row0 == 0 && live && a.do_step) {
        th = a.theta[s]; pos = a.xy[s];
        if (a.noise_amp != 0.0) {
            const int64_t id = a.gid[s];
            xi_ = a.noise ? a.noise[id] : philox_normal(a.seed, id, a.step_dev ? (int64_t)*a.step_dev : a.step);
        }
    }
    __syncthreads();
    int misses = 0;
    {
        const double2 *const rh = a.ring_h + (live ? s : 0);
        const int *const rn = a.ring_nbr + (live ? s : 0);
        const size_t st = a.stride;
        // the thread's rows of the ring are requested up front: the loop below then finds them in L1 instead of paying one DRAM
        // round trip per entry (0.239 -> 0.229 ms; requesting the neighbours' records as well: 0.244)
        for (int e = row0; e < E; e += FB_ROWS) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(rh + (size_t)e * st));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(rn + (size_t)e * st));
        }
        for (int e = row0; e < E; e += FB_ROWS) {
            const int e0 = (e == 0) ? E - 1 : e - 1, e1 = (e + 1 == E) ? 0 : e + 1;
            const int np_ = rn[(size_t)e0 * st], nc = rn[(size_t)e * st];
            const double2 hp = rh[(size_t)e0 * st], hc = rh[(size_t)e * st], hn = rh[(size_t)e1 * st];
            double2 t = make_double2(0.0, 0.0);
            if (np_ != ARC && nc != ARC) t = force_inner_term(mec, force_fetch(a, mec, np_), force_fetch(a, mec, nc), hp, hc, hn);
            else if ((np_ != ARC) != (nc != ARC)) s_queue[atomicAdd(&s_nq, 1)] = (unsigned short)(c << 4 | e);
            s_term[e][c] = t;
        }
    }
    __syncthreads();
    {
        // the queue, one outer vertex per thread (their order in the queue is arbitrary: each result has its own place)
        const int nq = s_nq;
        for (int i = threadIdx.x; i < nq; i += FB_THREADS) {
            const int qc = s_queue[i] >> 4, qe = s_queue[i] & 15, qs = blockIdx.x * FB_CELLS + qc;
            const RingRef R = {a.ring_h + qs, a.ring_nbr + qs, a.stride, (int)s_len[qc]};
            s_term[qe][qc] = force_entry_term(a, qs, *reinterpret_cast<const double4 *>(a.rec + qs), R, qe, misses);
        }
    }
    __syncthreads();
    if (misses) atomicAdd(&a.flags[1], misses);
    if (row0 != 0 || s >= a.n) return;
    if (!live) { a.fx[s] = 0.0; a.fy[s] = 0.0; return; }
    double dEx = 0.0, dEy = 0.0;
    for (int e = 0; e < E; ++e) { const double2 t = s_term[e][c]; dEx += t.x; dEy += t.y; }
    if (pooled) {   // a ring in the overflow pool (a handful of cells per step): walked by its own thread
        const int pi = a.ovf_index[s];
        const RingRef R = {a.ovf_ring_h + (size_t)pi * RS_SLOW, a.ovf_ring_nbr + (size_t)pi * RS_SLOW, 1, a.ovf_len[pi]};
        int m2 = 0;
        if (R.E >= 2) for (int e = 0; e < R.E; ++e) { const double2 t = force_entry_term(a, s, mec, R, e, m2); dEx += t.x; dEy += t.y; }
        if (m2) atomicAdd(&a.flags[1], m2);
    }
    const double Fx = -dEx, Fy = -dEy;
    a.fx[s] = Fx; a.fy[s] = Fy;
    // a step whose geometry exceeded every capacity does not move the cells: the state stays the one the error is about
    if (a.do_step && a.flags[3] == 0) {
        double sn, cs_;
        sincos(th, &sn, &cs_);
        double xn = pos.x + (a.mu * Fx + a.v0 * cs_) * a.dt;
        double yn = pos.y + (a.mu * Fy + a.v0 * sn) * a.dt;
        if (a.Lx > 0.0) { xn -= a.Lx * floor(xn / a.Lx); if (xn >= a.Lx) xn = 0.0; }
        if (a.Ly > 0.0) { yn -= a.Ly * floor(yn / a.Ly); if (yn >= a.Ly) yn = 0.0; }
        a.xy[s] = make_double2(xn, yn); a.theta[s] = th + a.noise_amp * xi_;
    }
}

// ------------------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------------------
__global__ void k_iota(int *p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }

__global__ void k_init_cells(int n, int64_t *__restrict__ gid, unsigned char *__restrict__ owned) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gid[i] = i; owned[i] = 1;
}
// afv_set_positions on a handle that keeps its cells: one pass brings theta and A0 back to the caller's order (scatter by id
// into the other buffers), resets ids and ownership marks, and wraps the uploaded positions into the box (L > 0)
__global__ void k_reset_order(int n, const double *__restrict__ th0, const double *__restrict__ a00, const int64_t *__restrict__ gid0,
                              double *__restrict__ th1, double *__restrict__ a01, int64_t *__restrict__ gid1, unsigned char *__restrict__ owned1,
                              double2 *__restrict__ xy1, double L) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t g = gid0[i];
    th1[g] = th0[i]; a01[g] = a00[i];
    gid1[i] = i; owned1[i] = 1;
    if (L > 0.0) {
        double2 q = xy1[i];
        q.x -= L * floor(q.x / L); if (q.x >= L) q.x = 0.0;
        q.y -= L * floor(q.y / L); if (q.y >= L) q.y = 0.0;
        xy1[i] = q;
    }
}
__global__ void k_unsort_xy(const double2 *__restrict__ src, const int64_t *__restrict__ gid, int n, double2 *__restrict__ out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) out[gid[s]] = src[s];
}

// positions of a periodic box into [0, L): the caller's host loop may leave them outside (pts += ...; update_positions(pts))
__global__ void k_wrap_box(double2 *__restrict__ xy, int n, double L) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 q = xy[i];
    q.x -= L * floor(q.x / L); if (q.x >= L) q.x = 0.0;
    q.y -= L * floor(q.y / L); if (q.y >= L) q.y = 0.0;
    xy[i] = q;
}

__global__ void k_fill(double *p, int n, double v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

// out[gid[s]] = src[s]   (back to the caller's order)
__global__ void k_unsort_f64(const double *__restrict__ src, const int64_t *__restrict__ gid, int n, double *__restrict__ out,
                             int stride, int comp) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) out[gid[s] * stride + comp] = src[s];
}
__global__ void k_unsort_i32_to_i64(const int *__restrict__ src, const int64_t *__restrict__ gid, int n, int64_t *__restrict__ out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) out[gid[s]] = (int64_t)src[s];
}
// dst[s] = src[gid[s]]   (caller's order -> current sorted order)
__global__ void k_gather_by_gid(const double *__restrict__ src, const int64_t *__restrict__ gid, int n, double *__restrict__ dst) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) dst[s] = src[gid[s]];
}

// read-only view of all rings (regular fixed-stride storage + overflow pool)
struct RingView {
    const double2 *ring_h;
    const int *ring_nbr;
    const unsigned char *ring_len;
    const int *ovf_index, *ovf_len;
    const double2 *ovf_ring_h;
    const int *ovf_ring_nbr;
    size_t stride;
    // entry e of the ring: vertex at h[e * ros], neighbour slot (ARC = arc) at nb[e * ros]
    __device__ __forceinline__ int get(int s, const double2 *&h, const int *&nb, size_t &ros) const {
        int E = ring_len[s];
        h = ring_h + s; nb = ring_nbr + s; ros = stride;
        if (E == RING_IN_POOL) { const int pi = ovf_index[s]; E = ovf_len[pi]; h = ovf_ring_h + (size_t)pi * RS_SLOW; nb = ovf_ring_nbr + (size_t)pi * RS_SLOW; ros = 1; }
        return E;
    }
};

// connections: count and emit pairs (gid_i < gid_j) from the rings (contact <=> straight ring edge)
__global__ void k_conn_count(RingView rv, const int64_t *__restrict__ gid, const unsigned char *__restrict__ owned, int n, int *__restrict__ cnt) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int c = 0;
    if (owned[s] == CELL_OWNED) {
        const double2 *h; const int *nb; size_t ros;
        const int E = rv.get(s, h, nb, ros);
        const int64_t gi = gid[s];
        for (int e = 0; e < E; ++e) {
            const int j = nb[e * ros];
            if (j != ARC && gi < gid[j]) ++c;
        }
    }
    cnt[s] = c;
}
__global__ void k_conn_emit(RingView rv, const int64_t *__restrict__ gid, const unsigned char *__restrict__ owned, int n,
                            const int *__restrict__ off, int64_t *__restrict__ pairs) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || owned[s] != CELL_OWNED) return;
    const double2 *h; const int *nb; size_t ros;
    const int E = rv.get(s, h, nb, ros);
    const int64_t gi = gid[s];
    int o = off[s];
    for (int e = 0; e < E; ++e) {
        const int j = nb[e * ros];
        if (j != ARC) {
            const int64_t gj = gid[j];
            if (gi < gj) { pairs[2 * (size_t)o] = gi; pairs[2 * (size_t)o + 1] = gj; ++o; }
        }
    }
}

// ring export: lengths in the caller's order, then entries
__global__ void k_ring_len_by_gid(RingView rv, const int64_t *__restrict__ gid, int n, int *__restrict__ len_out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double2 *h; const int *nb; size_t ros;
    len_out[gid[s]] = rv.get(s, h, nb, ros);
}
__global__ void k_ring_emit(RingView rv, const int64_t *__restrict__ gid, const double2 *__restrict__ xy, int n,
                            const int *__restrict__ off_by_gid, signed char *__restrict__ types, double *__restrict__ vxy) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double2 *h; const int *nb; size_t ros;
    const int E = rv.get(s, h, nb, ros);
    const int o = off_by_gid[gid[s]];
    const double xi = xy[s].x, yi = xy[s].y;
    for (int e = 0; e < E; ++e) {
        types[o + e] = (signed char)(nb[e * ros] != ARC ? 1 : 0);
        const double2 q = h[e * ros];
        vxy[2 * (size_t)(o + e)] = xi + q.x; vxy[2 * (size_t)(o + e) + 1] = yi + q.y;
    }
}

#include "afv_cluster.cuh"
#include "afv_relax.cuh"

// ------------------------------------------------------------------------------------------------------
// Exchange over peer memory (NVLink): the pack kernel of a rank writes its two messages straight into the inboxes of
// its ring neighbours (device memory of the other GPU, mapped through CUDA IPC), and the two ends synchronise with a
// sequence number per inbox instead of a send / receive pair: k_flag_signal, launched behind the pack, publishes the
// number (the writes of the kernels before it on the stream are complete; the fences order the flag behind them at
// system scope), k_flag_wait holds the receiving stream until the number has arrived.  No library call, no staging copy,
// no host in between.  A wait that outlives FLAG_WAIT_NS gives up and raises the error flag instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------------
constexpr int FLAG_TIMEOUT_MARK = 1 << 20;                  // added to flags[3]: told apart from capacity overflows
constexpr unsigned long long FLAG_WAIT_NS = 20000000000ull; // 20 s
// (two flag words per launch: a rank has two ring neighbours; flag_b may be null)
__global__ void k_flag_signal(unsigned *flag_a, unsigned *flag_b, unsigned value) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned *>(flag_a) = value;
    if (flag_b) *reinterpret_cast<volatile unsigned *>(flag_b) = value;
    __threadfence_system();
}
__global__ void k_flag_wait(const unsigned *flag_a, const unsigned *flag_b, unsigned value, int *flags) {
    const unsigned *const flag = (threadIdx.x == 0) ? flag_a : flag_b;      // one thread per flag word
    if (flag) {
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            unsigned cur;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(flag) : "memory");
            if ((int)(cur - value) >= 0) break;                 // (sequence numbers: wrap-safe)
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > FLAG_WAIT_NS) { atomicAdd(&flags[3], FLAG_TIMEOUT_MARK); break; }
        }
    }
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------------------
// slab sharding helpers.  Rows are (n,5) float64: x, y, theta, A0, gid
// ------------------------------------------------------------------------------------------------------
__global__ void k_unpack_rows(const double *__restrict__ rows, int n, int offset, int n_owned_in_rows, double2 *__restrict__ xy,
                              double *__restrict__ th, double *__restrict__ a0, int64_t *__restrict__ gid,
                              unsigned char *__restrict__ owned) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *q = rows + 5 * (size_t)i;
    const int d = offset + i;
    xy[d] = make_double2(q[0], q[1]); th[d] = q[2]; a0[d] = q[3]; gid[d] = (int64_t)q[4];
    owned[d] = (i < n_owned_in_rows) ? 1 : 0;
}
__global__ void k_pack_rows(const int *__restrict__ idx, int m, double x_shift, const double2 *__restrict__ xy,
                            const double *__restrict__ th, const double *__restrict__ a0, const int64_t *__restrict__ gid,
                            double *__restrict__ rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int s = idx[i];
    double *q = rows + 5 * (size_t)i;
    q[0] = xy[s].x + x_shift; q[1] = xy[s].y; q[2] = th[s]; q[3] = a0[s]; q[4] = (double)gid[s];
}
// message = (cap+1, 5) rows: row 0 is the header [count, 0, 0, 0, 0], rows 1..count the packed cells.  The count is
// read on the device (no host round trip between the selection and the packing); rows beyond cap are dropped and
// the header still carries the true count, so both ends detect the overflow.
__global__ void k_pack_message(const int *__restrict__ idx, const int *__restrict__ count, int cap, double x_shift,
                               const double2 *__restrict__ xy, const double *__restrict__ th, const double *__restrict__ a0,
                               const int64_t *__restrict__ gid, double *__restrict__ msg) {
    const int m = *count;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { msg[0] = (double)m; msg[1] = msg[2] = msg[3] = msg[4] = 0.0; }
    if (i >= m || i >= cap) return;
    const int s = idx[i];
    double *q = msg + 5 * (size_t)(i + 1);
    q[0] = xy[s].x + x_shift; q[1] = xy[s].y; q[2] = th[s]; q[3] = a0[s]; q[4] = (double)gid[s];
}
// ---- the per-step exchange: ONE launch packs both messages ------------------------------------------------------
// exchange message = (cap+1, 5) rows: header [n_transfer, n_halo, n_deep, 0, 0]; the transfer rows (cells that change
// owner) fill rows 1 .. n_transfer, the halo rows fill the buffer from the BACK (halo row i at row cap - i), so that both
// kinds can be appended concurrently without knowing each other's count.  Rows are claimed with warp-aggregated atomics:
// their order is arbitrary, and nothing downstream depends on it.
// Owned cells fall into up to four lists (xch_lists above): q = 0 transfers to the left rank, 1 halo rows for the left
// rank, 2 transfers to the right rank, 3 halo rows for the right rank.
constexpr int XCH_TPB = 256;
// row k of the cells a message carries (k < n_own: transfers, then the halo rows)
__device__ __forceinline__ const double *msg_row(const double *__restrict__ msg, int k, int n_own, int cap) {
    return msg + 5 * (size_t)(k < n_own ? k + 1 : cap - (k - n_own));
}
// work: int[8], zero on entry and zero again on exit: [0..3] rows claimed per list, [4] / [5] deep transfers left / right,
// [6] blocks done.  The last block writes the headers and the totals (pair_cnt, read back with the counts of the received
// messages).  Also updates the ownership marks: halo copies of the previous step become CELL_DEAD (the next cell-list pass
// drops them), cells that left the slab stay as halo copies of their new owner's cell.
__global__ void __launch_bounds__(XCH_TPB) k_exchange_pack(const double2 *__restrict__ xy, const double *__restrict__ th,
                                                           const double *__restrict__ a0, const int64_t *__restrict__ gid,
                                                           unsigned char *__restrict__ owned, int n,
                                                           double lo, double hi, double halo, double L, double shift_left, double shift_right,
                                                           double deep_lo, double deep_hi, int *__restrict__ work, int *__restrict__ totals,
                                                           int cap, double *__restrict__ msg_left, double *__restrict__ msg_right) {
    const int t = threadIdx.x, lane = t & 31;
    const int i = blockIdx.x * XCH_TPB + t;
    double2 p = make_double2(0.0, 0.0);
    unsigned char own = CELL_DEAD;
    unsigned m = 0u;
    double xr = 0.0;
    if (i < n) { p = xy[i]; own = owned[i]; xr = slab_offset(p.x, lo, hi - lo, L); m = xch_lists(xr, own, hi - lo, halo); }
    if (i < n) {
        if (own == CELL_HALO) owned[i] = CELL_DEAD;
        else if (m & 5u) owned[i] = CELL_HALO;
    }
    // transfers that land beyond the receiver's face columns (header slot 2): its face pass then covers every cell
    const unsigned deepl = __ballot_sync(FULL_MASK, (m & 1u) && xr < deep_lo - lo), deepr = __ballot_sync(FULL_MASK, (m & 4u) && xr >= deep_hi - lo);
    if (lane == 0) { if (deepl) atomicAdd(&work[4], __popc(deepl)); if (deepr) atomicAdd(&work[5], __popc(deepr)); }
    double tv = 0.0, av = 0.0, gv = 0.0;
    if (m) { tv = th[i]; av = a0[i]; gv = (double)gid[i]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const unsigned bal = __ballot_sync(FULL_MASK, (m >> q) & 1u);
        if (!bal) continue;
        int base = 0;
        if (lane == __ffs(bal) - 1) base = atomicAdd(&work[q], __popc(bal));
        base = __shfl_sync(FULL_MASK, base, __ffs(bal) - 1);
        if (!((m >> q) & 1u)) continue;
        const int r = base + __popc(bal & ((1u << lane) - 1u));
        if (r >= cap) continue;                                  // dropped; the header keeps the true count, both ends see the overflow
        double *row = ((q < 2) ? msg_left : msg_right) + 5 * (size_t)((q & 1) ? cap - r : r + 1);
        row[0] = p.x + ((q < 2) ? shift_left : shift_right); row[1] = p.y; row[2] = tv; row[3] = av; row[4] = gv;
    }
    // the last block to finish publishes the counts (only warps that claimed rows have anything to order before the block count)
    __shared__ int s_last;
    if (__any_sync(FULL_MASK, m != 0u)) __threadfence();
    __syncthreads();
    if (t == 0) s_last = (atomicAdd(&work[6], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last && t == 0) {
        __threadfence();
        volatile int *w = work;
        const int c0 = w[0], c1 = w[1], c2 = w[2], c3 = w[3], dl = w[4], dr = w[5];
        totals[0] = c0; totals[1] = c1; totals[2] = c2; totals[3] = c3;
        msg_left[0] = (double)c0; msg_left[1] = (double)c1; msg_left[2] = (double)dl; msg_left[3] = msg_left[4] = 0.0;
        msg_right[0] = (double)c2; msg_right[1] = (double)c3; msg_right[2] = (double)dr; msg_right[3] = msg_right[4] = 0.0;
        for (int k = 0; k < 7; ++k) work[k] = 0;
    }
}
// received messages -> cells appended behind slot `base`: [left transfers][left halo][right transfers][right halo];
// with `perm` (late cells, sorted by bin among themselves) row perm[i] of that sequence goes to slot base + i
// nl < 0: the four counts are read from the message headers on the device (clamped to the capacity; the host has not seen
// them yet) and the new cell count base + nl + nr goes to n_all_dev
__global__ void k_exchange_append(const double *__restrict__ in_left, const double *__restrict__ in_right, int nl_own, int nl, int nr_own,
                                  int nr, int base, int cap, const int *__restrict__ perm, double2 *__restrict__ xy, double *__restrict__ th,
                                  double *__restrict__ a0, int64_t *__restrict__ gid, unsigned char *__restrict__ owned, int *__restrict__ n_all_dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (nl < 0) {
        nl_own = min((int)in_left[0], cap); nl = nl_own + min((int)in_left[1], cap - nl_own);
        nr_own = min((int)in_right[0], cap); nr = nr_own + min((int)in_right[1], cap - nr_own);
        if (i == 0) *n_all_dev = base + nl + nr;
    }
    if (i >= nl + nr) return;
    const int src = perm ? perm[i] : i;
    const bool left = src < nl;
    const int k = left ? src : src - nl;
    const double *q = msg_row(left ? in_left : in_right, k, left ? nl_own : nr_own, cap);
    const int d = base + i;
    xy[d] = make_double2(q[0], q[1]); th[d] = q[2]; a0[d] = q[3]; gid[d] = (int64_t)q[4];
    owned[d] = (k < (left ? nl_own : nr_own)) ? CELL_OWNED : CELL_HALO;
}
// bin table of the (few, clustered) late cells: one thread per bin, lower bound by bisection -- the boundary walk of
// k_bin_start would leave the whole interior of the grid to a single thread
__global__ void k_bin_start_search(const unsigned *__restrict__ keys_sorted, int n, const GridDesc *__restrict__ gp, int *__restrict__ bin_start) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > gp->nbx * gp->nby) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys_sorted[mid] < (unsigned)b) lo = mid + 1; else hi = mid;
    }
    bin_start[b] = lo;
}
// bin keys of the received rows, in the order of k_exchange_append's unpermuted sequence
__global__ void k_late_keys(const double *__restrict__ in_left, const double *__restrict__ in_right, int nl_own, int nl, int nr_own, int nr,
                            int cap, const GridDesc *__restrict__ gp, unsigned *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nl + nr) return;
    const double *q = (i < nl) ? msg_row(in_left, i, nl_own, cap) : msg_row(in_right, i - nl, nr_own, cap);
    const GridDesc g = *gp;
    const int bx = bin_col(g, q[0]);
    int by = (int)floor((q[1] - g.y0) * g.inv_by);
    by = min(max(by, 0), g.nby - 1);
    keys[i] = bin_key(g, bx, by);
}
__global__ void k_gather_counts8(const int *__restrict__ c, const double *__restrict__ in0, const double *__restrict__ in1,
                                 const int *__restrict__ flags, const int *__restrict__ bin_start, int bin_a, int bin_b,
                                 double *__restrict__ out) {
    out[0] = (double)c[0]; out[1] = (double)c[1]; out[2] = (double)c[2]; out[3] = (double)c[3];
    out[4] = in0 ? in0[0] : 0.0; out[5] = in0 ? in0[1] : 0.0; out[6] = in1 ? in1[0] : 0.0; out[7] = in1 ? in1[1] : 0.0;
    out[8] = (double)flags[0]; out[9] = (double)flags[1]; out[10] = (double)flags[2]; out[11] = (double)flags[3];
    out[12] = out[13] = 0.0;
    out[14] = in0 ? in0[2] : 0.0; out[15] = in1 ? in1[2] : 0.0;
    out[16] = bin_start ? (double)bin_start[bin_a] : 0.0; out[17] = bin_start ? (double)bin_start[bin_b] : 0.0;   // ends of the face columns
}
// keep = !(owned && x in either range)
__global__ void k_flag_keep2(const double2 *__restrict__ xy, const unsigned char *__restrict__ owned, int n, double lo0, double hi0,
                             double lo1, double hi1, int *__restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xv = xy[i].x;
    const bool sel = owned[i] == CELL_OWNED && ((xv >= lo0 && xv < hi0) || (xv >= lo1 && xv < hi1));
    flag[i] = sel ? 0 : 1;
}
__global__ void k_gather_counts(const int *__restrict__ c0, const int *__restrict__ c1, const double *__restrict__ in0,
                                const double *__restrict__ in1, double *__restrict__ out) {
    out[0] = (double)*c0; out[1] = (double)*c1; out[2] = in0 ? in0[0] : 0.0; out[3] = in1 ? in1[0] : 0.0;
}
// rows (m,11): x, y, theta, A0, gid, fx, fy, area, perimeter, arclen, coord_num of the selected slots
__global__ void k_pack_results(const int *__restrict__ idx, int m, const double2 *__restrict__ xy, const double *__restrict__ th,
                               const double *__restrict__ a0, const int64_t *__restrict__ gid, const double *__restrict__ fx,
                               const double *__restrict__ fy, const double *__restrict__ area, const double *__restrict__ perim,
                               const double *__restrict__ arcl, const int *__restrict__ coord, double *__restrict__ rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int s = idx[i];
    double *q = rows + 11 * (size_t)i;
    q[0] = xy[s].x; q[1] = xy[s].y; q[2] = th[s]; q[3] = a0[s]; q[4] = (double)gid[s];
    q[5] = fx[s]; q[6] = fy[s]; q[7] = area[s]; q[8] = perim[s]; q[9] = arcl[s]; q[10] = (double)coord[s];
}
// flag = owned && x in [lo,hi)   (invert != 0: the complement, i.e. the cells to keep)
__global__ void k_flag_slab(const double2 *__restrict__ xy, const unsigned char *__restrict__ owned, int n, double lo, double hi, int invert,
                            int *__restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xv = xy[i].x;
    const int sel = (owned[i] == CELL_OWNED && xv >= lo && xv < hi) ? 1 : 0;
    flag[i] = invert ? 1 - sel : sel;
}
__global__ void k_flag_owned(const unsigned char *__restrict__ owned, int n, int *__restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (owned[i] == CELL_OWNED) ? 1 : 0;
}

// DFMA throughput micro-benchmark: 8 independent chains per thread
__global__ void k_dfma_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace afv
