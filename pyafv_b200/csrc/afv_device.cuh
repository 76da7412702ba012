// Device-side building blocks of the AFV hot path (sm_100a).  See DESIGN.md for the data layout.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace afv {

constexpr int RS = 16;      // ring stride: ring entries stored per cell on the fast path
constexpr int ARC = -1;     // neighbour label of an arc edge
constexpr int RS_SLOW = 96;       // ring capacity of the slow path (pooled rings)
constexpr int KPOLY_SLOW = 96;    // clip-polygon capacity of the slow path
constexpr int OVF_MAX = 8192;     // cells per build that may take the slow path
constexpr int RING_IN_POOL = 255; // ring_len sentinel: the ring lives in the overflow pool
// ownership mark of a cell (slab sharding): owned by this rank / halo copy of a neighbour rank's cell / stale halo copy
// waiting to be dropped by the next sort
constexpr unsigned char CELL_HALO = 0, CELL_OWNED = 1, CELL_DEAD = 2;

// Uniform cell list description.  Lives in device memory so that the open-boundary path can refresh it on the
// device (bounding box -> grid) without a host round trip.
struct GridDesc {
    double x0, y0;        // origin of bin (0,0)
    double inv_bx, inv_by; // 1 / bin side
    double Lx, halfLx;    // periodic period in x (used when perx)
    double Ly, halfLy;    // periodic period in y (used when pery)
    int nbx, nby;
    int perx, pery;       // periodic wrap of the bin grid per direction (and minimum image for the cells in its outermost bins)
    int swap;             // 0: key = by * nbx + bx (rows contiguous); 1: key = bx * nby + by (columns contiguous: x-slabs, whose face
                          // columns then are the two ends of the sorted order)
    int mix;              // x-slab of a periodic box: the bin grid is open in x, but positions are the box's own (wrapped into
                          // [0, Lx)), so bins are taken from the wrapped offset to x0 and displacements get the minimum image in x
    int seam_lo, seam_hi; // mix: the bin columns around the box seam x = 0 = Lx, if it runs through this grid (else lo > hi).
                          // Only their cells can have a neighbour whose x differs by a period; for all others the minimum
                          // image is the identity and is skipped (same bits, 20 % less work in k_nbrs)
};

__device__ __forceinline__ unsigned bin_key(const GridDesc &g, int bx, int by) {
    return g.swap ? (unsigned)(bx * g.nby + by) : (unsigned)(by * g.nbx + bx);
}
// key -> (position along the line, line); bin column / row: see bin_xy
__device__ __forceinline__ void bin_xy(const GridDesc &g, unsigned key, int &bx, int &by) {
    const unsigned w = (unsigned)(g.swap ? g.nby : g.nbx);
    const int line = (int)(key / w), pos = (int)(key - (unsigned)line * w);
    bx = g.swap ? line : pos; by = g.swap ? pos : line;
}

// What a neighbour needs from a cell during the force gather: its position and the two energy prefactors
// cA = 2 KA (A - A0), cP = 2 KP (P - P0).  32 B = one DRAM sector, so one gather per ring edge.
// With these two scalars cell i can rebuild, from positions it already holds, every term of dE_j/dr_i that passes
// through an inner vertex (circumcentre) -- see k_force.  Only outer vertices (circle-circle intersections, whose
// Jacobian the reference regularises with delta) need one more datum from the neighbour: the far end of the
// neighbour's arc that starts / ends there (ArcTable below).
struct __align__(16) CellRec {
    double x, y;    // copy of the cell centre (same bits as xy[])
    double cA, cP;  // 2 KA (A - A0), 2 KP (P - P0)
};

// Ring storage, slot-major (entry e of cell s at [e * stride + s], stride = allocated cells), so that the lanes of a
// warp -- consecutive cells working on the same entry -- touch consecutive words:
//   ring_h[e]   vertex position relative to the cell centre (double2)
//   ring_nbr[e] sorted slot of the neighbour across the edge LEAVING this vertex (clockwise), ARC for an arc
// Arcs, CELL-major (arc q of cell s at [s * ARC_MAX + q]: the first four keys of a cell are one 32-byte sector, which is
// what a neighbour gathers): arc_key = (slot of the neighbour whose contact edge ENDS where the arc starts, slot of the
// neighbour whose contact edge STARTS where the arc ends), arc_pts = (start x, y, end x, y) relative to the cell centre.
// A ring of E entries has at most E/2 arcs, and the fast path stores that many: at rho l^2 = 1 ten cells in a million
// have five or more, and with room for four they were the whole population of the slow path (0.019 ms per step for a
// dozen cells).
constexpr int ARC_MAX = RS / 2;
constexpr int ARC_SLOW = RS_SLOW / 2;

// bin column of x (clamped)
__device__ __forceinline__ int bin_col(const GridDesc &g, double x) {
    double xr = x - g.x0;
    if (g.mix) xr -= g.Lx * floor(xr / g.Lx);
    return min(max((int)floor(xr * g.inv_bx), 0), g.nbx - 1);
}

__device__ __forceinline__ double min_image(double d, double L, double halfL) {
    if (d > halfL) d -= L;
    else if (d < -halfL) d += L;
    return d;
}

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al., SC'11).  counter = (cell id, step), key = seed, so the
// stream does not depend on the sort order or on how cells are sharded over GPUs.
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)M0 * c[0];
        uint64_t p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

// one standard normal per (seed, cell id, step): Box-Muller on two 53-bit uniforms in (0,1]
__host__ __device__ inline double philox_normal(uint64_t seed, int64_t cell, int64_t step) {
    uint32_t c[4] = {(uint32_t)(uint64_t)cell, (uint32_t)((uint64_t)cell >> 32), (uint32_t)(uint64_t)step,
                     (uint32_t)((uint64_t)step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double two53 = 1.0 / 9007199254740992.0;
    uint64_t a = (((uint64_t)c[0] << 32) | c[1]) >> 11;
    uint64_t b = (((uint64_t)c[2] << 32) | c[3]) >> 11;
    double u1 = ((double)a + 1.0) * two53;   // (0,1]
    double u2 = (double)b * two53;           // [0,1)
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}

// ---------------------------------------------------------------------------------------------------------
// Convex polygon (counter-clockwise, relative to the cell centre) clipped in place by the half-plane
// n.x <= |n|^2/2 (the bisector with a neighbour at displacement n).  Every edge k is itself such a bisector with
// (px[k], py[k]); the four initial sides of the bounding square are bisectors with virtual neighbours.  New
// vertices are computed as the intersection of the two generating bisectors, so a vertex is a function of its
// generating triple only -- independent of the order in which neighbours were clipped.
// Returns false when the result would exceed the capacity K.
// ---------------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ bool clip_polygon(double (&vx)[K], double (&vy)[K], double (&px)[K], double (&py)[K],
                                             int (&lab)[K], int &m, double nx, double ny, double c, int label) {
    int a = -1, b = -1;
    bool prev_in = (nx * vx[m - 1] + ny * vy[m - 1] <= c);
    for (int k = 0; k < m; ++k) {
        bool cur_in = (nx * vx[k] + ny * vy[k] <= c);
        int km = (k == 0) ? m - 1 : k - 1;
        if (prev_in && !cur_in) a = km;   // edge a leaves the half-plane
        if (!prev_in && cur_in) b = km;   // edge b re-enters it
        prev_in = cur_in;
    }
    if (a < 0 || b < 0) return true;      // not cut (the centre itself always satisfies the inequality)

    // P = edge a  ^  new bisector,  Q = edge b ^ new bisector
    double ax = px[a], ay = py[a], ca = 0.5 * (ax * ax + ay * ay);
    double bx = px[b], by = py[b], cb = 0.5 * (bx * bx + by * by);
    int blab = lab[b];
    double da = ax * ny - ay * nx, db = bx * ny - by * nx;
    double Px = (ca * ny - c * ay) / da, Py = (ax * c - nx * ca) / da;
    double Qx = (cb * ny - c * by) / db, Qy = (bx * c - nx * cb) / db;

    if (a < b) {
        // outside run a+1..b; keep 0..a, then P, Q, then b+1..m-1
        int shift = 2 - (b - a);
        int mnew = m + shift;
        if (mnew > K) return false;
        if (shift > 0) {
            for (int k = m - 1; k > b; --k) { vx[k + 1] = vx[k]; vy[k + 1] = vy[k]; px[k + 1] = px[k]; py[k + 1] = py[k]; lab[k + 1] = lab[k]; }
        } else if (shift < 0) {
            for (int k = b + 1; k < m; ++k) { vx[k + shift] = vx[k]; vy[k + shift] = vy[k]; px[k + shift] = px[k]; py[k + shift] = py[k]; lab[k + shift] = lab[k]; }
        }
        vx[a + 1] = Px; vy[a + 1] = Py; px[a + 1] = nx; py[a + 1] = ny; lab[a + 1] = label;
        vx[a + 2] = Qx; vy[a + 2] = Qy; px[a + 2] = bx; py[a + 2] = by; lab[a + 2] = blab;
        m = mnew;
    } else {
        // outside run wraps: a+1..m-1 and 0..b; keep b+1..a (moved to the front), then P, Q
        int sh = b + 1, cnt = a - b;
        if (cnt + 2 > K) return false;
        for (int k = 0; k < cnt; ++k) { vx[k] = vx[k + sh]; vy[k] = vy[k + sh]; px[k] = px[k + sh]; py[k] = py[k + sh]; lab[k] = lab[k + sh]; }
        vx[cnt] = Px; vy[cnt] = Py; px[cnt] = nx; py[cnt] = ny; lab[cnt] = label;
        vx[cnt + 1] = Qx; vy[cnt + 1] = Qy; px[cnt + 1] = bx; py[cnt + 1] = by; lab[cnt + 1] = blab;
        m = cnt + 2;
    }
    return true;
}

}  // namespace afv
