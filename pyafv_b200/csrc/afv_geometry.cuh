// K2: geometry of the finite-Voronoi cells.  Included by afv_kernels.cuh inside namespace afv.
//
// k_nbrs: neighbour lists (all cells within 2r, from the 3 x 3 bins).  k_clip (fast path): one thread per cell; the
// Voronoi polygon is a bounding square clipped by the neighbours' bisector half-planes (edges in shared-memory slots,
// their order in a register), and the ring is walked right away, exactly as the reference defines it
// (finite_voronoi.py:363-459), with area / perimeter / arc length (cell_geom.pyx:432-475) accumulated on the fly.
// What it leaves behind for the force gather is small: the ring (vertex + neighbour slot, 20 B per entry), the arcs,
// and ONE 32-byte record per cell with the energy prefactors 2KA(A-A0), 2KP(P-P0) -- no per-vertex gradient records.
//
// k_geometry_slow: cells whose neighbourhood / polygon / ring exceeds the fast-path capacity, one thread each:
// in-place clipping of a bounding square (afv_device.cuh: clip_polygon) with large local arrays gives the polygon, a
// ring walk that rebuilds the vertices from consecutive edges then writes into the overflow pool.

struct GeomArgs {
    const double2 *xy;
    const double *a0;
    const unsigned *keys;      // sorted
    const int *bin_start;
    const GridDesc *grid;
    int n;
    PhysDev ph;
    size_t stride;             // allocated cells
    double2 *ring_h;           // [RS * stride] slot-major ring vertices (relative to the cell centre)
    int *ring_nbr;             // [RS * stride] slot-major: neighbour slot across the edge leaving the vertex, ARC for an arc
    unsigned char *ring_len;   // [n] ring entries; RING_IN_POOL when the ring lives in the overflow pool
    int2 *arc_key;             // [stride * ARC_MAX] cell-major, see afv_device.cuh
    double4 *arc_pts;          // [stride * ARC_MAX]
    unsigned char *n_arc;      // [n]
    CellRec *rec;              // [n] what the neighbours gather in k_force
    int *nb_list;              // [cmax * stride] slot-major neighbour lists (k_nbrs -> k_clip)
    unsigned char *nb_cnt;     // [n] entries; NB_OVERFLOW when the list was too short
    int *nb_min;               // [n] sorted slot of the nearest neighbour (-1: none within 2r)
    double *area, *perim, *arcl;
    int *coord;
    int *flags;                // [0] cells sent to the slow path, [1] lookup misses, [2] max ring length, [3] hard overflows
    // slow path (cells whose neighbourhood / polygon / ring exceeds the fast-path capacity)
    int *ovf_cells;            // [OVF_MAX] sorted slots queued by the fast kernels
    int *ovf_index;            // [n] slot -> pool index, valid where ring_len == RING_IN_POOL
    int *ovf_len;              // [OVF_MAX] ring length of pooled rings
    int *ovf_hull;             // [OVF_MAX * KPOLY_SLOW] polygon of a queued cell (neighbour slots counter-clockwise, -2: edge at infinity)
    int *ovf_m;                // [OVF_MAX] its edge count when k_geom handed the polygon over, 0: clip from scratch
    double2 *ovf_ring_h;       // [OVF_MAX * RS_SLOW]
    int *ovf_ring_nbr;         // [OVF_MAX * RS_SLOW]
    int2 *ovf_arc_key;         // [OVF_MAX * ARC_SLOW]
    double4 *ovf_arc_pts;      // [OVF_MAX * ARC_SLOW]
    // late cells (slab sharding with the exchange overlapped): cells appended behind the sorted ones AFTER the sort,
    // slots n_main .. n-1, with a bin table of their own (entries relative to n_main)
    int n_main;
    const int *bin_start_late; // nullptr: no late cells to look at
    int filter;                // 0 every cell; 1 interior cells only; 2 the others (face columns and late cells)
    int face_cols;             // bin columns at either end of the slab grid whose cells can see a late cell
    // filter == 2: the launch is dense over the face cells -- thread t handles slot t (t < face_len0: the first face_cols
    // columns), face_s1 + (t - face_len0) (the last face_cols columns, face_len1 cells), or a late cell
    int face_len0, face_len1, face_s1;
    // minimum image per direction (period, half period; 0 in an open system).  By value: the hot loops read them from the
    // constant bank instead of holding the grid description's copies in registers
    double mi_Lx, mi_hx, mi_Ly, mi_hy;
};

// sorted slot handled by thread t of a geometry launch (n = number of threads with work)
__device__ __forceinline__ int slot_of_thread(const GeomArgs &a, int t) {
    if (a.filter != 2) return t;
    if (t < a.face_len0) return t;
    t -= a.face_len0;
    if (t < a.face_len1) return a.face_s1 + t;
    return a.n_main + (t - a.face_len1);
}

// the launch handles cell s (bin column bx)?  Interior cells: placed by the sort, at least face_cols columns from both
// ends of the grid, so that no late cell lies in their 3 x 3 bin neighbourhood.
__device__ __forceinline__ bool pass_selects(const GeomArgs &a, const GridDesc &g, int s, int bx) {
    if (a.filter == 0) return true;
    const bool interior = s < a.n_main && bx >= a.face_cols && bx < g.nbx - a.face_cols;
    return (a.filter == 1) == interior;
}

constexpr unsigned FULL_MASK = 0xffffffffu;

// ------------------------------------------------------------------------------------------------------
// candidate runs of one line of bins (a row; a column when the grid is swapped): positions pos-1..pos+1 of the line as
// (up to) two contiguous runs of sorted slots.  npos bins per line, periodic along the line or not.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int row_runs(const int npos, const bool per, const int *__restrict__ bin_start, int line, int pos, int (&c0)[2], int (&c1)[2]) {
    const int base = line * npos;
    if (!per || npos < 3) {
        const int lo = per ? 0 : max(pos - 1, 0), hi = per ? npos - 1 : min(pos + 1, npos - 1);
        c0[0] = bin_start[base + lo]; c1[0] = bin_start[base + hi + 1];
        return 1;
    }
    if (pos == 0) {
        c0[0] = bin_start[base]; c1[0] = bin_start[base + 2];
        c0[1] = bin_start[base + npos - 1]; c1[1] = bin_start[base + npos];
        return 2;
    }
    if (pos == npos - 1) {
        c0[0] = bin_start[base + pos - 1]; c1[0] = bin_start[base + pos + 1];
        c0[1] = bin_start[base]; c1[1] = bin_start[base + 1];
        return 2;
    }
    c0[0] = bin_start[base + pos - 1]; c1[0] = bin_start[base + pos + 2];
    return 1;
}

// lines to visit for a cell in line `line` (of nline, periodic or not): returns the count, fills rows[]
__device__ __forceinline__ int cell_rows(const int nline, const bool per, int line, int (&rows)[3]) {
    if (per) {
        if (nline < 3) { for (int i = 0; i < nline; ++i) rows[i] = i; return nline; }
        rows[0] = (line == 0) ? nline - 1 : line - 1; rows[1] = line; rows[2] = (line == nline - 1) ? 0 : line + 1;
        return 3;
    }
    int n = 0;
    if (line > 0) rows[n++] = line - 1;
    rows[n++] = line;
    if (line < nline - 1) rows[n++] = line + 1;
    return n;
}

// ------------------------------------------------------------------------------------------------------
// fast path
// ------------------------------------------------------------------------------------------------------
#ifndef GEOM_TPB_
#define GEOM_TPB_ 128
#endif
constexpr int GEOM_TPB = GEOM_TPB_;
#ifndef NBRS_MIN_BLOCKS
#define NBRS_MIN_BLOCKS 8
#endif
constexpr int KP = 14;           // polygon edges on the fast path (<= 16: the order of the slots is a sequence of 4-bit ids)
constexpr int CMAX_HARD = 128;   // upper bound of the per-cell neighbour list (the launch picks cmax <= this)
constexpr int NB_OVERFLOW = 255; // nb_cnt: more neighbours within 2r than the list holds
constexpr int LAB_SIDE = -2;     // polygon-edge labels LAB_SIDE - q: side q of the bounding square (labels >= 0: sorted slots)

// One contiguous run of sorted slots [c_begin, c_end) against cell s at (xi, yi): the slots within 2r are appended at wp
// (stride apart, up to wend), the nearest is tracked.  The body is short and branch-free on purpose: two of three
// candidates fail the distance test, and a divergent accept branch would run for nearly every candidate with a third of
// the lanes.  WRAP: minimum image (cells in the outermost bins of a periodic direction; every cell of a slab grid).
template <bool WRAP>
__device__ __forceinline__ void scan_run(const double2 *__restrict__ xy, const int c_begin, const int c_end, const int s, const double xi,
                                         const double yi, const double cut2, const bool wrapx, const bool wrapy, const GridDesc &g,
                                         const size_t stride, int *const wend, int *&wp, double &dmin, int &smin, bool &tie, bool &ok) {
    const double Lx = wrapx ? g.Lx : 0.0, hx = wrapx ? g.halfLx : 1e300, Ly = wrapy ? g.Ly : 0.0, hy = wrapy ? g.halfLy : 1e300;
#pragma unroll 4
    for (int c = c_begin; c < c_end; ++c) {
        const double2 pc = __ldg(xy + c);
        double dx = pc.x - xi, dy = pc.y - yi;
        if (WRAP) { dx = min_image(dx, Lx, hx); dy = min_image(dy, Ly, hy); }   // (a direction that does not wrap: h = 1e300)
        const double d2 = dx * dx + dy * dy;
        const bool acc = d2 < cut2 && d2 > 0.0;               // (the cell itself: d2 == 0)
        const bool lt = acc && d2 < dmin, eq = acc && d2 == dmin;
        tie = eq || (tie && !lt);             // several cells at exactly the smallest distance so far
        dmin = lt ? d2 : dmin;
        smin = lt ? c : smin;
        const bool put = acc && wp != wend;
        if (put) *wp = c;
        wp += put ? stride : 0;
        ok = ok && (put || !acc);
    }
}

// ------------------------------------------------------------------------------------------------------
// K2a: neighbour lists.  One thread per cell: the sorted slots of all cells within 2r, from the 3 bin rows (<= 2
// contiguous runs each), slot-major (entry e of cell s at nb_list[e * stride + s]) so that k_clip, whose lanes walk their
// lists in step, reads them coalesced.  Also the nearest neighbour (where the ring of the cell will start).
// SLAB = false: one box, every cell, rows of bins contiguous (the single-GPU path, nothing of the slab machinery is
// compiled in).  SLAB = true: an x-slab -- columns contiguous, pass filters, the late cells' bin table.
// ------------------------------------------------------------------------------------------------------
template <bool SLAB>
__global__ void __launch_bounds__(128, NBRS_MIN_BLOCKS) k_nbrs(GeomArgs a, const int cmax) {
    const int t_ = blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= a.n) return;
    const int s = SLAB ? slot_of_thread(a, t_) : t_;
    GridDesc g = *a.grid;
    if (!SLAB) g.swap = 0;
    int bx, by;
    bin_xy(g, a.keys[s], bx, by);
    if (SLAB && !pass_selects(a, g, s, bx)) return;
    const double r = a.ph.r, cut2 = 4.0 * r * r;
    const double2 pi_ = a.xy[s];
    const double xi = pi_.x, yi = pi_.y;
    const bool wrapx = (g.mix && bx >= g.seam_lo && bx <= g.seam_hi) || (g.perx && (bx == 0 || bx == g.nbx - 1 || g.nbx < 3));   // only cells in the outermost bins can see a periodic image
    const bool wrapy = g.pery && (by == 0 || by == g.nby - 1 || g.nby < 3);
    // lines of bins: rows, or columns in a swapped (slab) grid
    const int npos = g.swap ? g.nby : g.nbx, nline = g.swap ? g.nbx : g.nby, pos = g.swap ? by : bx;
    const bool per_pos = g.swap ? g.pery : g.perx, per_line = g.swap ? g.perx : g.pery;
    int rows[3];
    const int nrow = cell_rows(nline, per_line, g.swap ? bx : by, rows);
    int *const list = a.nb_list + s;
    int *wp = list;                                       // next free entry of the list
    int *const wend = list + (size_t)cmax * a.stride;
    int smin = -1;
    bool ok = true, tie = false;
    double dmin = 1e300;
    for (int tab = 0; tab < (SLAB ? 2 : 1); ++tab) {     // the sorted cells, then (face pass of a slab) the late cells
        const int *const bin_start = tab ? a.bin_start_late : a.bin_start;
        if (!bin_start) break;
        const int first = tab ? a.n_main : 0;
        for (int ir = 0; ir < nrow; ++ir) {
            int c0[2], c1[2];
            const int nrun = row_runs(npos, per_pos, bin_start, rows[ir], pos, c0, c1);
            for (int q = 0; q < nrun; ++q) {
                if (wrapx || wrapy) scan_run<true>(a.xy, first + c0[q], first + c1[q], s, xi, yi, cut2, wrapx, wrapy, g, a.stride, wend, wp, dmin, smin, tie, ok);
                else scan_run<false>(a.xy, first + c0[q], first + c1[q], s, xi, yi, cut2, false, false, g, a.stride, wend, wp, dmin, smin, tie, ok);
            }
        }
    }
    const int n = (int)((size_t)(wp - list) / a.stride);
    if (tie) {
        // nearest neighbour among equals (lattices): the smallest displacement -- the ring start must not depend on the
        // order the candidates are met in.  Rare: scan the runs once more.
        double mdx = 0.0, mdy = 0.0;
        bool have = false;
        for (int tab = 0; tab < (SLAB ? 2 : 1); ++tab) {
            const int *const bin_start = tab ? a.bin_start_late : a.bin_start;
            if (!bin_start) break;
            const int first = tab ? a.n_main : 0;
            for (int ir = 0; ir < nrow; ++ir) {
                int c0[2], c1[2];
                const int nrun = row_runs(npos, per_pos, bin_start, rows[ir], pos, c0, c1);
                for (int q = 0; q < nrun; ++q)
                    for (int c = first + c0[q]; c < first + c1[q]; ++c) {
                        const double2 pc = a.xy[c];
                        double dx = pc.x - xi, dy = pc.y - yi;
                        if (wrapx) dx = min_image(dx, g.Lx, g.halfLx);
                        if (wrapy) dy = min_image(dy, g.Ly, g.halfLy);
                        if (c != s && dx * dx + dy * dy == dmin && (!have || dx < mdx || (dx == mdx && dy < mdy))) { have = true; smin = c; mdx = dx; mdy = dy; }
                    }
            }
        }
    }
    a.nb_cnt[s] = ok ? (unsigned char)n : (unsigned char)NB_OVERFLOW;
    a.nb_min[s] = smin;
}

// intersection of the bisectors of two neighbours at displacements a and b, (a, b) in counter-clockwise order of the
// polygon edges they carry (a x b > 0)
__device__ __forceinline__ double2 bisector_junction(double ax, double ay, double bx, double by) {
    // explicit roundings: the compiler must not contract the two inlined copies of this expression differently
    const double ca = __dmul_rn(0.5, __fma_rn(ax, ax, __dmul_rn(ay, ay))), cb = __dmul_rn(0.5, __fma_rn(bx, bx, __dmul_rn(by, by)));
#ifdef AFV_FAST_RCP
    const double det = __fma_rn(ax, by, -__dmul_rn(ay, bx));
    double idet;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(idet) : "d"(det));          // ~20 good bits, then two Newton steps
    idet = __fma_rn(idet, __fma_rn(-det, idet, 1.0), idet);
    idet = __fma_rn(idet, __fma_rn(-det, idet, 1.0), idet);
#else
    const double idet = __drcp_rn(__fma_rn(ax, by, -__dmul_rn(ay, bx)));
#endif
    return make_double2(__dmul_rn(__fma_rn(ca, by, -__dmul_rn(cb, ay)), idet), __dmul_rn(__fma_rn(ax, cb, -__dmul_rn(bx, ca)), idet));
}

inline size_t clip_smem_bytes() { return (size_t)GEOM_TPB * ((size_t)KP * (sizeof(double2) + sizeof(int)) + sizeof(int)) + (CMAX_HARD + 2) * sizeof(int); }

#ifndef CLIP_MIN_BLOCKS
#define CLIP_MIN_BLOCKS 6
#endif
#ifndef CLIP_TEST_CHUNK
#define CLIP_TEST_CHUNK 4        // vertices tested per batch of the half-plane test (1: 0.353 ms, 2: 0.321, 4: 0.318 at N = 1e6)
#endif
// ------------------------------------------------------------------------------------------------------
// K2b: polygon, ring and measures.  One thread per cell, everything between the neighbour list and the ring on chip:
// (1) polygon: a bounding square (half side 2r) clipped by the bisector half-plane of every listed neighbour.  The
//     edges live in fixed SLOTS of shared memory (start vertex + label, [slot][thread], conflict free); their counter-
//     clockwise ORDER is a sequence of 4-bit slot ids in one 64-bit register.  Per neighbour one unrolled pass
//     evaluates every vertex against the half-plane (one 16-byte shared load + 2 FMAs each) into a bit mask; the
//     edges that leave / re-enter the half-plane follow from two bit operations, the two new vertices are the
//     intersections of the generating bisectors (a vertex is a function of its generating triple only, whatever the
//     order of the cuts), and the cut itself is a rotate-mask-insert of the order register: no data moves.
// (2) ring + measures: clockwise from the edge of the nearest neighbour (so that every sum is taken in an order the
//     geometry defines, not the sort), with the vertices at hand; exactly the reference's ring
//     (finite_voronoi.py:363-459) and measures (cell_geom.pyx:432-475).
// Cells with more than cmax neighbours, KP polygon edges, RS ring entries or ARC_MAX arcs are queued for
// k_geometry_slow.
// ------------------------------------------------------------------------------------------------------
template <bool SLAB>
__global__ void __launch_bounds__(GEOM_TPB, CLIP_MIN_BLOCKS) k_clip(GeomArgs a) {
    extern __shared__ double2 s_mem[];
    const int tid = threadIdx.x;
    constexpr int T = GEOM_TPB;
    double2 *const sv = s_mem + tid;                                       // [KP][T] start vertex of the edge in slot k
    int *const sl = reinterpret_cast<int *>(s_mem + KP * T) + tid;         // [KP][T] its label: sorted slot of the neighbour, or LAB_SIDE - q
    // The cells of the block are dealt to its threads by decreasing list length, so that the lanes of a warp run (nearly)
    // the same number of clips: the loop bounds below are warp maxima.  All of the block's cells share their
    // neighbourhood, so the locality of the loads does not change.
    // (A counting sort on the list length; which thread takes which cell of a tie is arbitrary and changes no result.)
    int *const s_deal = reinterpret_cast<int *>(s_mem + KP * T) + KP * T;  // [T] rank -> thread index in slot order
    int *const s_hist = s_deal + T;                                        // [CMAX_HARD + 2] cells per list length, then start ranks
    int t_;
    {
        const int t0 = blockIdx.x * GEOM_TPB + tid;
        int key = CMAX_HARD + 1;                                           // threads without a cell rank last
        // (cells this pass does not select may carry a stale count: any value 0 .. 255)
        if (t0 < a.n) { const int c = a.nb_cnt[SLAB ? slot_of_thread(a, t0) : t0]; key = (c > CMAX_HARD) ? CMAX_HARD : CMAX_HARD - c; }
        for (int j = tid; j < CMAX_HARD + 2; j += T) s_hist[j] = 0;
        __syncthreads();
        atomicAdd(&s_hist[key], 1);
        __syncthreads();
        if (tid < 32) {   // exclusive scan of the CMAX_HARD + 2 counters by one warp
            int carry = 0;
            for (int base = 0; base < CMAX_HARD + 2; base += 32) {
                const int j = base + tid;
                const int v = (j < CMAX_HARD + 2) ? s_hist[j] : 0;
                int x = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(FULL_MASK, x, d); if (tid >= d) x += y; }
                if (j < CMAX_HARD + 2) s_hist[j] = carry + x - v;
                carry += __shfl_sync(FULL_MASK, x, 31);
            }
        }
        __syncthreads();
        s_deal[atomicAdd(&s_hist[key], 1)] = tid;
        __syncthreads();
        t_ = blockIdx.x * GEOM_TPB + s_deal[tid];
    }
    bool live = t_ < a.n;
    const int s = SLAB ? slot_of_thread(a, t_) : t_;
    const double r = a.ph.r, r2 = r * r;
    GridDesc g = *a.grid;
    if (!SLAB) g.swap = 0;
    int n = 0, smin = -1;
    bool ok = true;
    double xi = 0.0, yi = 0.0, a0s = 0.0;
    bool wrapx = false, wrapy = false;
    if (live) {
        int bx, by;
        bin_xy(g, a.keys[s], bx, by);
        if (SLAB) live = pass_selects(a, g, s, bx);
        if (live) {
            const double2 pi_ = a.xy[s];
            xi = pi_.x; yi = pi_.y;
            wrapx = (g.mix && bx >= g.seam_lo && bx <= g.seam_hi) || (g.perx && (bx == 0 || bx == g.nbx - 1 || g.nbx < 3));
            wrapy = g.pery && (by == 0 || by == g.nby - 1 || g.nby < 3);
            n = a.nb_cnt[s]; smin = a.nb_min[s];
            a0s = a.a0[s];                         // (needed at the very end; requested here, its round trip is free)
            if (n == NB_OVERFLOW) { ok = false; n = 0; }
        }
    }
    // displacement of the cell in sorted slot `slot`; normal of a polygon edge (2 x the foot point of its bisector)
    auto disp_of = [&](int slot, double &dx, double &dy) {
        const double2 pc = a.xy[slot];
        dx = pc.x - xi; dy = pc.y - yi;
        if (wrapx) dx = min_image(dx, a.mi_Lx, a.mi_hx);
        if (wrapy) dy = min_image(dy, a.mi_Ly, a.mi_hy);
    };
    const double B = 2.0 * r;                     // half side of the bounding square: its sides are bisectors of virtual neighbours at 4r
    // (no branch: a side label reads the cell's own position -- a zero displacement -- and the normal of the side is
    // selected afterwards; the sides occur in the first cuts of every cell, next to real neighbours in the same warp)
    auto normal_of = [&](int lab, double &nx, double &ny) {
        const bool side = lab < 0;
        disp_of(side ? s : lab, nx, ny);
        const int q = LAB_SIDE - lab;             // bottom, right, top, left
        const double sx = (q & 1) ? ((q & 2) ? -2.0 * B : 2.0 * B) : 0.0;
        const double sy = (q & 1) ? 0.0 : ((q & 2) ? 2.0 * B : -2.0 * B);
        nx = side ? sx : nx; ny = side ? sy : ny;
    };

    // ---- (1) the polygon.  Edge at position p of the order (counter-clockwise) runs from its start vertex to the start
    //      vertex of the edge at position p+1. ----
    // `ord` is a permutation of all KP slot ids, 4 bits each: positions 0 .. m-1 are the edges of the polygon in order,
    // the positions behind them the free slots.
    unsigned long long ord = 0xfedcba9876543210ull;
    int m = 4;
#define NIB(w, p) ((int)(((w) >> (4 * (p))) & 15ull))
    const bool work = live && ok && n > 0;
    if (work) {
        sv[0 * T] = make_double2(-B, -B); sv[1 * T] = make_double2(B, -B); sv[2 * T] = make_double2(B, B); sv[3 * T] = make_double2(-B, B);
        sl[0 * T] = LAB_SIDE; sl[1 * T] = LAB_SIDE - 1; sl[2 * T] = LAB_SIDE - 2; sl[3 * T] = LAB_SIDE - 3;
    }
    {
        const int nloop = __reduce_max_sync(FULL_MASK, work ? n : 0);
        const int *const list = a.nb_list + (live ? s : 0);
        // neighbour slots are requested two iterations ahead, their positions one iteration ahead
        int c1 = 0, c2 = 0;
        double2 p1 = make_double2(0.0, 0.0);
        if (work) { c1 = list[0]; if (n > 1) c2 = list[a.stride]; p1 = a.xy[c1]; }
        for (int ci = 0; ci < nloop; ++ci) {
            const bool act = work && ok && ci < n;
            const int c = c1;
            const double2 pc = p1;
            c1 = c2;
            if (work && ci + 1 < n) p1 = a.xy[c1];
            if (work && ci + 2 < n) c2 = list[(size_t)(ci + 2) * a.stride];
            double nx = pc.x - xi, ny = pc.y - yi;
            if (wrapx) nx = min_image(nx, a.mi_Lx, a.mi_hx);
            if (wrapy) ny = min_image(ny, a.mi_Ly, a.mi_hy);
            const double cc = 0.5 * (nx * nx + ny * ny);
            // which vertices lie inside the new half-plane?  (bit p; the loop bound is the warp's largest polygon)
            unsigned in = 0u;
            const int mloop = __reduce_max_sync(FULL_MASK, act ? m : 0);
            // in chunks: the loads of a chunk are issued together and its tests are independent (positions >= m hold free
            // slots; their bits are masked off below)
#pragma unroll
            for (int p0 = 0; p0 < KP; p0 += CLIP_TEST_CHUNK) {
                if (p0 >= mloop) break;
                double2 v[CLIP_TEST_CHUNK];
#pragma unroll
                for (int q = 0; q < CLIP_TEST_CHUNK; ++q) if (p0 + q < KP) v[q] = sv[NIB(ord, p0 + q) * T];
#pragma unroll
                for (int q = 0; q < CLIP_TEST_CHUNK; ++q) if (p0 + q < KP) { if (fma(nx, v[q].x, ny * v[q].y) <= cc) in |= 1u << (p0 + q); }
            }
            const unsigned full = (1u << m) - 1u;
            in &= full;
            if (act && in != full && in != 0u) {
                // ea: the edge that leaves the half-plane (its start inside, its end outside); eb: the edge that comes back
                const unsigned rot = ((in >> 1) | ((in & 1u) << (m - 1))) & full;      // bit p = vertex p+1 inside
                const int ea = __ffs(in & ~rot) - 1, eb = __ffs(~in & rot & full) - 1;
                const int sa = NIB(ord, ea), sb = NIB(ord, eb);
                // P = edge ea ^ new bisector, Q = edge eb ^ new bisector (intersections of the two generating bisectors)
                // Both with the same expression, arguments in counter-clockwise order of the two edges: the bits of a vertex
                // depend on its generating pair only, not on which of the two was clipped first.
                double ax, ay, bx_, by_;
                normal_of(sl[sa * T], ax, ay);
                normal_of(sl[sb * T], bx_, by_);
                const double2 Pv = bisector_junction(ax, ay, nx, ny);
                const double2 Qv = bisector_junction(nx, ny, bx_, by_);
                // Rotate the first m positions so that eb comes first.  Then the edges at 1 .. t stay (t = where ea went), those
                // at t+1 .. m-1 lie entirely outside, the new edge follows t, and edge eb (position 0) starts at Q.  The
                // slot at position t+1 -- the first one that fell free, or, when only one vertex was cut off (t+1 = m), the
                // first slot of the free tail -- takes the new edge, and everything behind it is free: the rotation IS the cut.
                int t = ea - eb; if (t < 0) t += m;
                if (t + 2 > KP) ok = false;       // more than KP edges
                else {
                    const unsigned long long mmask = (1ull << (4 * m)) - 1ull, low = ord & mmask;
                    ord = (ord & ~mmask) | (((low >> (4 * eb)) | (low << (4 * (m - eb)))) & mmask);
                    const int ns = NIB(ord, t + 1);
                    m = t + 2;
                    sv[ns * T] = Pv; sl[ns * T] = c; sv[sb * T] = Qv;
                }
            }
        }
    }
    const bool poly_ok = work && ok;

    // ---- (2) ring, clockwise from the edge of the nearest neighbour: an edge is walked from the start of its
    //      counter-clockwise successor (V1) to its own start (V2)  (finite_voronoi.py:363-459) ----
    const double PI = 3.14159265358979323846, TWO_PI = 6.2831853071795864769;
    double2 *const rh = a.ring_h + (live ? s : 0);
    int *const rn = a.ring_nbr + (live ? s : 0);
    int2 *const ak = a.arc_key + (size_t)(live ? s : 0) * ARC_MAX;       // cell-major: what a neighbour looks up is one sector
    double4 *const ap = a.arc_pts + (size_t)(live ? s : 0) * ARC_MAX;
    const size_t ros = a.stride;
    int E = 0, z = 0, na = 0;
    double A = 0.0, P = 0.0;
    double x0 = 0.0, y0 = 0.0;                   // first ring entry
    int nbr0 = ARC;
    double xp = 0.0, yp = 0.0;                   // previous ring entry
    int nbr_p = ARC, arc_from = ARC;             // its leaving edge; the contact edge that ended where the pending arc starts
    // connect the previous entry to (x, y), the start of the edge with neighbour nb
    auto close_prev = [&](double x, double y, int nb) {
        if (nbr_p != ARC) {
            const double sx = xp - x, sy = yp - y;
            const double l2 = sx * sx + sy * sy;
            P += l2 > 0.0 ? l2 * rsqrt(l2) : 0.0;
            A += -0.5 * (xp * y - yp * x);
            ++z;
        } else if (na < ARC_MAX) {               // the arc that started at (xp, yp) ends here
            ak[na] = make_int2(arc_from, nb);
            ap[na] = make_double4(xp, yp, x, y);
            ++na;
        } else ok = false;                       // more arcs than the fast path stores
    };
    // new ring entry at (x, y); nb = neighbour across the edge leaving it (ARC: an arc starts here, at the end of the
    // contact edge with neighbour `edge`)
    auto emit = [&](double x, double y, int nb, int edge) {
        if (E >= RS) { ok = false; return; }
        if (E > 0) close_prev(x, y, nb); else { x0 = x; y0 = y; nbr0 = nb; }
        rh[E * ros] = make_double2(x, y);
        rn[E * ros] = nb;
        if (nb == ARC) arc_from = edge;
        xp = x; yp = y; nbr_p = nb; ++E;
    };
    {
        const int mw = poly_ok ? m : 0;
        const int mloop = __reduce_max_sync(FULL_MASK, mw);
        // position of the nearest neighbour's edge (it is always a Voronoi neighbour)
        int p = 0;
#pragma unroll
        for (int q = 0; q < KP; ++q) {
            if (q >= mloop) break;
            if (q < mw && sl[NIB(ord, q) * T] == smin) p = q;
        }
        double2 V1 = make_double2(0.0, 0.0);
        // label and neighbour position of the edge at position p are requested one iteration ahead
        int jn = LAB_SIDE;
        double2 pn = make_double2(0.0, 0.0);
        if (mw) {
            V1 = sv[NIB(ord, (p + 1 == mw) ? 0 : p + 1) * T];
            jn = sl[NIB(ord, p) * T];
            if (jn >= 0) pn = a.xy[jn];
        }
        for (int st = 0; st < mloop; ++st) {
            if (st < mw) {
                const int k = NIB(ord, p);
                const double2 V2 = sv[k * T];
                const int j = jn;
                const double2 pj = pn;
                p = (p == 0) ? mw - 1 : p - 1;
                if (st + 1 < mw) { jn = sl[NIB(ord, p) * T]; if (jn >= 0) pn = a.xy[jn]; }
                if (j >= 0 && ok) {                        // a real neighbour (the sides of the square emit nothing)
                    double pcx = pj.x - xi, pcy = pj.y - yi;
                    if (wrapx) pcx = min_image(pcx, a.mi_Lx, a.mi_hx);
                    if (wrapy) pcy = min_image(pcy, a.mi_Ly, a.mi_hy);
                    const bool in1 = V1.x * V1.x + V1.y * V1.y <= r2;
                    // first entry of the edge: its inner start vertex (d1 <= r) or the point where it enters the disk (never
                    // both); second: the point where it leaves the disk.  An edge between two inner vertices has neither.
                    bool has1 = in1, has2 = false;
                    double x1 = V1.x, y1 = V1.y, x2 = 0.0, y2 = 0.0;
                    const double d2c = 0.25 * (pcx * pcx + pcy * pcy);               // |C|^2, C = p/2
                    if (d2c < r2 && !(in1 && V2.x * V2.x + V2.y * V2.y <= r2)) {      // the bisector crosses the circle
                        const double inv = 0.5 * rsqrt(d2c);
                        const double ux = pcy * inv, uy = -pcx * inv;                  // clockwise unit vector along the edge
                        const double t1 = V1.x * ux + V1.y * uy, t2 = V2.x * ux + V2.y * uy;
                        const double tr = sqrt(r2 - d2c);
                        if (-tr < t2 && -tr > t1) { has1 = true; x1 = 0.5 * pcx - tr * ux; y1 = 0.5 * pcy - tr * uy; }    // entry point
                        if (tr < t2 && tr > t1) { has2 = true; x2 = 0.5 * pcx + tr * ux; y2 = 0.5 * pcy + tr * uy; }      // exit point
                    }
                    if (has1) emit(x1, y1, j, j);
                    if (has2) emit(x2, y2, ARC, j);
                }
                V1 = V2;
            }
        }
    }
    if (live && ok && E >= 2) close_prev(x0, y0, nbr0);   // closing edge: last entry -> first entry
    const bool ringed = live && ok && E >= 2;
    if (live && !ok) {   // too many neighbours / polygon edges / ring entries / arcs for the fast path: queue for the large-capacity kernel
        const int idx = atomicAdd(&a.flags[0], 1);
        if (idx < OVF_MAX) {
            a.ovf_cells[idx] = s;
            // only the ring outgrew the fast path: hand the polygon over, the slow path need not clip again
            const int mm = poly_ok ? m : 0;
            a.ovf_m[idx] = mm;
            for (int q = 0; q < mm; ++q) { const int lab = sl[NIB(ord, q) * T]; a.ovf_hull[(size_t)idx * KPOLY_SLOW + q] = lab >= 0 ? lab : -2; }
        } else atomicAdd(&a.flags[3], 1);
        a.ring_len[s] = 0; a.n_arc[s] = 0;
    }
#undef NIB

    // ---- arcs: the lanes of a warp evaluate their q-th arc together (cell_geom.pyx:454-469) ----
    double Larc = 0.0;
    {
        const int nq = ringed ? na : 0;
        const int aloop = __reduce_max_sync(FULL_MASK, nq);
        for (int q = 0; q < aloop; ++q) {
            if (q < nq) {
                const double4 c = ap[q];                  // (v1, v2) = (start, end)
                double ang = atan2(c.z * c.y - c.w * c.x, c.x * c.z + c.y * c.w);   // angle(v1) - angle(v2)
                if (ang < 0.0) ang += TWO_PI;
                Larc += r * ang;
                A += 0.5 * r2 * ang;
            }
        }
    }
    P += Larc;
    if (live && ok) {
        int zz = z;
        if (!ringed) {                           // isolated cell (cell_geom.pyx:402-406)
            A = PI * r2; P = TWO_PI * r; Larc = TWO_PI * r; zz = 0;
            a.ring_len[s] = 0; a.n_arc[s] = 0;
        } else {
            a.ring_len[s] = (unsigned char)E;
            a.n_arc[s] = (unsigned char)na;
            if (E > 12) atomicMax(&a.flags[2], E);
        }
        a.area[s] = A; a.perim[s] = P; a.arcl[s] = Larc; a.coord[s] = zz;
        CellRec q;
        q.x = xi; q.y = yi;
        q.cA = 2.0 * a.ph.KA * (A - a0s);
        q.cP = 2.0 * a.ph.KP * (P - a.ph.P0);
        a.rec[s] = q;
    }
}

// ------------------------------------------------------------------------------------------------------
// ring and measures of a queued cell from its polygon (hull[k * hstride], k < m_pool: neighbour slots counter-clockwise,
// < 0 for edges at infinity), streamed into slot `pool_idx` of the overflow pool; at most RS_SLOW entries, returns false
// when even that is too short.  Vertices are rebuilt as the intersections of consecutive bisectors.
// ------------------------------------------------------------------------------------------------------
template <bool POOL, bool SLAB>
__device__ __forceinline__ bool ring_cell(const GeomArgs &a, const int s, const bool live_in, const int pool_idx,
                                          const int *hull, const size_t hstride, const int m_pool) {
    bool live = live_in;
    const double r = a.ph.r, B4 = 4.0 * r;
    GridDesc g = *a.grid;
    if (!SLAB) g.swap = 0;
    double xi = 0.0, yi = 0.0;
    bool wrapx = false, wrapy = false;
    int m = 0;
    bool ok = true;
    const int *hs = hull;                                            // edge k at hs[k * hst]
    const size_t hst = hstride;
    if (live) {
        int bx, by;
        bin_xy(g, a.keys[s], bx, by);
        if (!POOL && SLAB) live = pass_selects(a, g, s, bx);
        if (live) {
            const double2 pi_ = a.xy[s];
            xi = pi_.x; yi = pi_.y;
            wrapx = (g.mix && bx >= g.seam_lo && bx <= g.seam_hi) || (g.perx && (bx == 0 || bx == g.nbx - 1 || g.nbx < 3));
            wrapy = g.pery && (by == 0 || by == g.nby - 1 || g.nby < 3);
            m = m_pool;
        }
    }
    const bool queued = false;
    // the walk starts at the edge of the nearest neighbour, like k_clip's: edge index k of the walk is entry k + rot of the
    // polygon (so that the sums are taken in the same order whichever path a cell takes)
    int rot = 0;
    if (live && m > 0) {
        const int smin = a.nb_min[s];
        for (int k = 0; k < m; ++k) if (hs[k * hst] == smin) rot = (k + 1 == m) ? 0 : k + 1;
    }
    auto hidx = [&](int k) { int q = k + rot; if (q >= m) q -= m; return hs[q * hst]; };
    // displacement of polygon edge k's neighbour
    auto disp = [&](int k, double &dx, double &dy, int &slot) {
        slot = hidx(k);
        if (slot >= 0) {
            const double2 pc = a.xy[slot];
            dx = pc.x - xi; dy = pc.y - yi;
            if (wrapx) dx = min_image(dx, g.Lx, g.halfLx);
            if (wrapy) dy = min_image(dy, g.Ly, g.halfLy);
        } else { dx = 0.0; dy = 0.0; }
    };

    // ---- ring, streamed.  The polygon edges are e_0..e_{m-1} (counter-clockwise); the ring is clockwise, so edge k
    //      is walked from vertex k+1 to vertex k for k = m-1 .. 0 (finite_voronoi.py:363-459). ----
    const double r2 = r * r;
    const double PI = 3.14159265358979323846, TWO_PI = 6.2831853071795864769;
    double2 *const rh = POOL ? a.ovf_ring_h + (size_t)pool_idx * RS_SLOW : a.ring_h + s;
    int *const rn = POOL ? a.ovf_ring_nbr + (size_t)pool_idx * RS_SLOW : a.ring_nbr + s;
    int2 *const ak = POOL ? a.ovf_arc_key + (size_t)pool_idx * ARC_SLOW : a.arc_key + (size_t)s * ARC_MAX;
    double4 *const ap = POOL ? a.ovf_arc_pts + (size_t)pool_idx * ARC_SLOW : a.arc_pts + (size_t)s * ARC_MAX;
    const size_t ros = POOL ? 1 : a.stride;
    constexpr int ECAP = POOL ? RS_SLOW : RS, ACAP = POOL ? ARC_SLOW : ARC_MAX;
    const int mm = m;
    const int mloop = POOL ? mm : __reduce_max_sync(FULL_MASK, mm);
    // Junction between the counter-clockwise consecutive polygon edges A (slot ja, displacement a) and B (jb, b):
    // the END of A and the START of B.  Normally one point, the intersection of the two bisectors (a proper left
    // turn, a x b > 0).  Next to the edge at infinity -- or for a degenerate turn -- the bisector runs off to
    // infinity instead: a point 4r along it stands in (it only has to lie outside the disk).
    auto junction = [&](int ja, double ax, double ay, int jb, double bx_, double by_, double &eax, double &eay, double &sbx, double &sby) {
        const double det = ax * by_ - ay * bx_;
        if (ja >= 0 && jb >= 0 && det > 0.0) {
            const double2 v = bisector_junction(ax, ay, bx_, by_);
            eax = sbx = v.x; eay = sby = v.y;
        } else {
            eax = eay = sbx = sby = 0.0;
            if (ja >= 0) { const double f = B4 * rsqrt(ax * ax + ay * ay); eax = 0.5 * ax - f * ay; eay = 0.5 * ay + f * ax; }      // C_A + 4r ccw(A)
            if (jb >= 0) { const double f = B4 * rsqrt(bx_ * bx_ + by_ * by_); sbx = 0.5 * bx_ + f * by_; sby = 0.5 * by_ - f * bx_; } // C_B - 4r ccw(B)
        }
    };
    int E = 0, z = 0, na = 0;
    double A = 0.0, P = 0.0;
    double x0 = 0.0, y0 = 0.0;                   // first ring entry
    int nbr0 = ARC;
    double xp = 0.0, yp = 0.0;                   // previous ring entry
    int nbr_p = ARC, arc_from = ARC;             // its leaving edge; the contact edge that ended where the pending arc starts
    // connect the previous entry to (x, y), the start of the edge with neighbour nb
    auto close_prev = [&](double x, double y, int nb) {
        if (nbr_p != ARC) {
            const double sx = xp - x, sy = yp - y;
            const double l2 = sx * sx + sy * sy;
            P += l2 > 0.0 ? l2 * rsqrt(l2) : 0.0;
            A += -0.5 * (xp * y - yp * x);
            ++z;
        } else if (na < ACAP) {                  // the arc that started at (xp, yp) ends here
            ak[na] = make_int2(arc_from, nb);
            ap[na] = make_double4(xp, yp, x, y);
            ++na;
        } else ok = false;                       // more arcs than the fast path stores: the cell goes to the pool
    };
    // new ring entry at (x, y); nb = neighbour across the edge leaving it (ARC: an arc starts here, at the end of the
    // contact edge with neighbour `edge`)
    auto emit = [&](double x, double y, int nb, int edge) {
        if (E >= ECAP) { ok = false; return; }
        if (E > 0) close_prev(x, y, nb); else { x0 = x; y0 = y; nbr0 = nb; }
        rh[E * ros] = make_double2(x, y);
        rn[E * ros] = nb;
        if (nb == ARC) arc_from = edge;
        xp = x; yp = y; nbr_p = nb; ++E;
    };
    {
        // the walk visits the edges m-1, then m-2, ..., 0 and needs the PREVIOUS polygon edge of each: slot indices are
        // requested two iterations ahead, positions one iteration ahead, so neither load sits on the critical path
        auto load_slot = [&](int k) { return hidx(k); };
        auto load_pos = [&](int slot) { return slot >= 0 ? a.xy[slot] : make_double2(0.0, 0.0); };
        auto rel = [&](int slot, double2 pc, double &dx, double &dy) {
            if (slot >= 0) {
                dx = pc.x - xi; dy = pc.y - yi;
                if (wrapx) dx = min_image(dx, g.Lx, g.halfLx);
                if (wrapy) dy = min_image(dy, g.Ly, g.halfLy);
            } else { dx = 0.0; dy = 0.0; }
        };
        double pcx = 0.0, pcy = 0.0, pnx = 0.0, pny = 0.0, V1x = 0.0, V1y = 0.0;
        int jc = -2, jn = -2;
        int slA = -2, slB = -2;                   // previous edge of iteration kk / kk + 1
        double2 poA = make_double2(0.0, 0.0);
        if (mm > 0) {
            double tx, ty;
            disp(mm - 1, pcx, pcy, jc);
            disp(0, pnx, pny, jn);
            slA = load_slot(mm > 1 ? mm - 2 : 0);
            slB = load_slot(mm > 2 ? mm - 3 : mm - 1);
            poA = load_pos(slA);
            junction(jc, pcx, pcy, jn, pnx, pny, V1x, V1y, tx, ty);   // end of edge m-1 (junction with edge 0)
        }
        for (int kk = 0; kk < mloop; ++kk) {
            if (kk < mm) {
                double ppx, ppy, V2x, V2y, Enx, Eny;
                const int jp = slA;
                rel(slA, poA, ppx, ppy);
                {   // edge index of the previous edge two iterations on: k - 1 with k = mm - 1 - (kk + 2)
                    int kn = mm - 4 - kk; if (kn < 0) kn += mm; if (kn < 0) kn += mm; if (kn < 0) kn += mm;
                    const int slC = (kk + 2 < mm) ? load_slot(kn) : -2;
                    poA = load_pos(slB);
                    slA = slB; slB = slC;
                }
                junction(jp, ppx, ppy, jc, pcx, pcy, Enx, Eny, V2x, V2y);   // end of edge k-1, start of edge k
                if (jc >= 0 && ok) {                       // a real neighbour (virtual edges emit nothing)
                    const int j = jc;
                    const bool in1 = V1x * V1x + V1y * V1y <= r2;
                    // first entry of the edge: its inner start vertex (d1 <= r) or the point where it enters the disk (never
                    // both); second: the point where it leaves the disk.  An edge between two inner vertices has neither.
                    bool has1 = in1, has2 = false;
                    double x1 = V1x, y1 = V1y, x2 = 0.0, y2 = 0.0;
                    const double d2c = 0.25 * (pcx * pcx + pcy * pcy);               // |C|^2, C = p/2
                    if (d2c < r2 && !(in1 && V2x * V2x + V2y * V2y <= r2)) {          // the bisector crosses the circle
                        const double inv = 0.5 * rsqrt(d2c);
                        const double ux = pcy * inv, uy = -pcx * inv;                  // clockwise unit vector along the edge
                        const double t1 = V1x * ux + V1y * uy, t2 = V2x * ux + V2y * uy;
                        const double tr = sqrt(r2 - d2c);
                        if (-tr < t2 && -tr > t1) { has1 = true; x1 = 0.5 * pcx - tr * ux; y1 = 0.5 * pcy - tr * uy; }    // entry point
                        if (tr < t2 && tr > t1) { has2 = true; x2 = 0.5 * pcx + tr * ux; y2 = 0.5 * pcy + tr * uy; }      // exit point
                    }
                    if (has1) emit(x1, y1, j, j);
                    if (has2) emit(x2, y2, ARC, j);
                }
                pcx = ppx; pcy = ppy; jc = jp; V1x = Enx; V1y = Eny;
            }
        }
    }
    if (live && ok && !queued && E >= 2) close_prev(x0, y0, nbr0);   // closing edge: last entry -> first entry
    const bool ringed = live && ok && !queued && E >= 2;

    if (POOL && !ok) return false;               // even the pool is too short: the caller reports it
    if (!POOL && live && !ok) {                  // ring longer than the stride: queue the cell for the large-capacity kernel
        const int idx = atomicAdd(&a.flags[0], 1);
        if (idx < OVF_MAX) a.ovf_cells[idx] = s; else atomicAdd(&a.flags[3], 1);
        a.ring_len[s] = 0; a.n_arc[s] = 0;
    }
    if (live && queued) { a.ring_len[s] = 0; a.n_arc[s] = 0; }   // k_geometry_slow fills everything in

    // ---- arcs: the lanes of a warp evaluate their q-th arc together (cell_geom.pyx:454-469) ----
    double Larc = 0.0;
    {
        const int nq = ringed ? na : 0;
        const int aloop = POOL ? nq : __reduce_max_sync(FULL_MASK, nq);
        for (int q = 0; q < aloop; ++q) {
            if (q < nq) {
                const double4 c = ap[q];                  // (v1, v2) = (start, end)
                double ang = atan2(c.z * c.y - c.w * c.x, c.x * c.z + c.y * c.w);   // angle(v1) - angle(v2)
                if (ang < 0.0) ang += TWO_PI;
                Larc += r * ang;
                A += 0.5 * r2 * ang;
            }
        }
    }
    P += Larc;
    if (live && ok && !queued) {
        int zz = z;
        if (!ringed) {                           // isolated cell (cell_geom.pyx:402-406)
            A = PI * r2; P = TWO_PI * r; Larc = TWO_PI * r; zz = 0;
            a.ring_len[s] = 0; a.n_arc[s] = 0;
        } else {
            if (POOL) { a.ring_len[s] = RING_IN_POOL; a.ovf_index[s] = pool_idx; a.ovf_len[pool_idx] = E; }
            else a.ring_len[s] = (unsigned char)E;
            a.n_arc[s] = (unsigned char)na;
            if (E > 12) atomicMax(&a.flags[2], E);
        }
        a.area[s] = A; a.perim[s] = P; a.arcl[s] = Larc; a.coord[s] = zz;
        CellRec q;
        q.x = xi; q.y = yi;
        q.cA = 2.0 * a.ph.KA * (A - a.a0[s]);
        q.cP = 2.0 * a.ph.KP * (P - a.ph.P0);
        a.rec[s] = q;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------
// slow path: the (few) queued cells, one thread each
// ------------------------------------------------------------------------------------------------------
// polygon of cell s by in-place clipping of a bounding square with large local arrays -> ovf_hull of pool slot
// `pool_idx` (neighbour slots counter-clockwise, -2 for the sides of the square); returns the edge count, -1 when the
// polygon outgrows KPOLY_SLOW
__device__ __noinline__ int clip_cell_slow(const GeomArgs &a, const int s, const int pool_idx) {
    constexpr int K = KPOLY_SLOW;
    const double r = a.ph.r, cut2 = 4.0 * r * r;
    const double xi = a.xy[s].x, yi = a.xy[s].y;
    const GridDesc g = *a.grid;
    int bx, by;
    bin_xy(g, a.keys[s], bx, by);
    const int npos = g.swap ? g.nby : g.nbx, nline = g.swap ? g.nbx : g.nby, pos = g.swap ? by : bx;
    const bool per_pos = g.swap ? g.pery : g.perx, per_line = g.swap ? g.perx : g.pery;
    // bounding square of half side B = 2r as four virtual bisectors (virtual neighbours at distance 4r)
    double vx[K], vy[K], px[K], py[K];
    int lab[K];
    const double B = 2.0 * r;
    vx[0] = -B; vy[0] = -B; px[0] = 0.0;      py[0] = -2.0 * B; lab[0] = -2;
    vx[1] = B;  vy[1] = -B; px[1] = 2.0 * B;  py[1] = 0.0;      lab[1] = -2;
    vx[2] = B;  vy[2] = B;  px[2] = 0.0;      py[2] = 2.0 * B;  lab[2] = -2;
    vx[3] = -B; vy[3] = B;  px[3] = -2.0 * B; py[3] = 0.0;      lab[3] = -2;
    int m = 4;
    bool ok = true;
    int rows[3];
    const int nrow = cell_rows(nline, per_line, g.swap ? bx : by, rows);
    for (int tab = 0; tab < 2; ++tab) {
        const int *const bin_start = tab ? a.bin_start_late : a.bin_start;
        if (!bin_start) break;
        const int first = tab ? a.n_main : 0;
        for (int ir = 0; ir < nrow; ++ir) {
            int c0[2], c1[2];
            const int nrun = row_runs(npos, per_pos, bin_start, rows[ir], pos, c0, c1);
            for (int q = 0; q < nrun; ++q) {
                for (int c = first + c0[q]; c < first + c1[q]; ++c) {
                    if (c == s) continue;
                    const double2 pc = a.xy[c];
                    double dx = pc.x - xi, dy = pc.y - yi;
                    if (g.perx || g.mix) dx = min_image(dx, g.Lx, g.halfLx);
                    if (g.pery) dy = min_image(dy, g.Ly, g.halfLy);
                    const double d2 = dx * dx + dy * dy;
                    if (d2 >= cut2 || d2 == 0.0) continue;
                    ok = ok && clip_polygon<K>(vx, vy, px, py, lab, m, dx, dy, 0.5 * d2, c);
                }
            }
        }
    }
    if (!ok) return -1;
    int *out = a.ovf_hull + (size_t)pool_idx * KPOLY_SLOW;
    for (int k = 0; k < m; ++k) out[k] = lab[k];
    return m;
}

// The queue is usually a handful of cells and each is a long serial job: they are spread over the launch, one cell per
// warp (and one warp per block first) as long as there are warps to spare, so that the kernel lasts as long as its
// longest cell and not as long as a warp's worth of divergent ones.
__global__ void __launch_bounds__(64) k_geometry_slow(GeomArgs a) {
    const int cnt = min(a.flags[0], OVF_MAX);
    const int T = gridDim.x * blockDim.x;
    const bool sparse = cnt * 32 <= T;                     // a warp for every queued cell: its lane 0 works
    if (sparse && (threadIdx.x & 31) != 0) return;
    const int job0 = sparse ? (threadIdx.x >> 5) * gridDim.x + blockIdx.x : blockIdx.x * blockDim.x + threadIdx.x;
    const int njob = sparse ? T / 32 : T;
    for (int idx = job0; idx < cnt; idx += njob) {
        const int s = a.ovf_cells[idx];
        const int m0 = a.ovf_m[idx];
        const int m = m0 > 0 ? m0 : clip_cell_slow(a, s, idx);
        const bool done = m >= 0 && ring_cell<true, true>(a, s, true, idx, a.ovf_hull + (size_t)idx * KPOLY_SLOW, 1, m);
        if (!done) {   // exceeded even the slow-path capacity: the host reports AFV_ERR_CAPACITY; the cell counts as isolated
            atomicAdd(&a.flags[3], 1);
            const double r = a.ph.r, PI = 3.14159265358979323846;
            const double A = PI * r * r, P = 2.0 * PI * r;
            a.ring_len[s] = 0; a.n_arc[s] = 0;
            a.area[s] = A; a.perim[s] = P; a.arcl[s] = P; a.coord[s] = 0;
            CellRec q; q.x = a.xy[s].x; q.y = a.xy[s].y; q.cA = 2.0 * a.ph.KA * (A - a.a0[s]); q.cP = 2.0 * a.ph.KP * (P - a.ph.P0);
            a.rec[s] = q;
        }
    }
}
