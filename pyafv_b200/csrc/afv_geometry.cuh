// K2: geometry of the finite-Voronoi cells.  Included by afv_kernels.cuh inside namespace afv.
//
// Fast path (k_geometry_hull): one thread per cell; the in-range neighbours of each cell are gathered once into
// shared memory ([entry][lane] layout, conflict free) and the Voronoi polygon is found by gift wrapping in the
// dual plane (half-plane p.x <= |p|^2/2  <->  dual point 2p/|p|^2; the cell's edges are the hull vertices, in
// counter-clockwise order).  The loops are regular (hull steps x candidates), warp-uniform in their bounds, hold no
// per-thread arrays while wrapping, and need no division; a vertex is computed once per final edge as the
// intersection of its two bisectors.  Cells with too many in-range neighbours / hull vertices / ring entries are
// queued for the slow path.
//
// Slow path (k_geometry_slow): the queued cells, one thread each, by in-place clipping of a bounding square
// (afv_device.cuh: clip_polygon) with large local arrays; long rings go to the overflow pool.
//
// Both paths then share finish_cell(): ring emission exactly as the reference defines it
// (finite_voronoi.py:363-459), area / perimeter / arc length (cell_geom.pyx:432-475), per-vertex energy gradients
// of this cell (cell_geom.pyx:478-539), written as ring records.

struct GeomArgs {
    const double2 *xy;
    const double *a0;
    const unsigned *keys;      // sorted
    const int *bin_start;
    const GridDesc *grid;
    int n;
    PhysDev ph;
    RingOwn *ro;               // slot-major, stride `stride`
    RingSelf *self;            // slot-major, stride `stride`: gradient records, gathered by the neighbours in k_force
    size_t stride;             // allocated cells
    unsigned char *ring_len;
    int *hull_slot;            // [KHULL_STRIDE * stride], slot-major: polygon edges found by k_hull (neighbour slot, or -2 for the edge at infinity)
    unsigned char *hull_len;   // [n] number of polygon edges; HULL_QUEUED when the cell went to the slow path
    int *ring_key;             // [n * RS] neighbour slot of every ring entry (ARC for arcs; stale beyond ring_len): what k_force scans
    double *area, *perim, *arcl;
    int *coord;
    int *flags;                // [0] cells sent to the slow path, [1] lookup misses, [2] max ring length, [3] hard overflows
    // slow path (cells whose neighbourhood / polygon / ring exceeds the fast-path capacity)
    int *ovf_cells;            // [OVF_MAX] sorted slots queued by the fast kernel
    int *ovf_index;            // [n] slot -> pool index, valid where ring_len == RING_IN_POOL
    int *ovf_len;              // [OVF_MAX] ring length of pooled rings
    RingShared *ovf_rs;        // [OVF_MAX * RS_SLOW]
    RingOwn *ovf_ro;
    // late cells (slab sharding with the exchange overlapped): cells appended behind the sorted ones AFTER the sort,
    // slots n_main .. n-1, with a bin table of their own (entries relative to n_main)
    int n_main;
    const int *bin_start_late; // nullptr: no late cells to look at
    int filter;                // 0 every cell; 1 interior cells only; 2 the others (face columns and late cells)
    int face_cols;             // bin columns at either end of the slab grid whose cells can see a late cell
    // filter == 2: the launch is dense over the face cells -- thread t handles slot t (t < face_len0: the first face_cols
    // columns), face_s1 + (t - face_len0) (the last face_cols columns, face_len1 cells), or a late cell
    int face_len0, face_len1, face_s1;
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
constexpr int KHULL_STRIDE = 16;
constexpr int HULL_QUEUED = 255;

// ------------------------------------------------------------------------------------------------------
// Shared second half.  Input: the convex polygon of the cell, counter-clockwise, relative to the centre: vertex k,
// and for the edge from vertex k to vertex k+1 the displacement (px,py) of the neighbour whose bisector carries
// it and that neighbour's sorted slot `lab` (< 0: side of the bounding square).
// LOCKSTEP: every lane of the warp calls this (invalid lanes with m = 0); loop bounds are warp maxima so that the
// expensive bodies run converged.
// ------------------------------------------------------------------------------------------------------
template <int K, int R, bool SLOW>
__device__ __forceinline__ void finish_cell(const GeomArgs &a, const int s, const int pool_idx, const bool live, bool ok,
                                            const double (&vx)[K], const double (&vy)[K], const double (&px)[K],
                                            const double (&py)[K], const int (&lab)[K], const int m) {
    constexpr bool LOCKSTEP = !SLOW;
    const double r = a.ph.r, r2 = r * r;
    const double PI = 3.14159265358979323846, TWO_PI = 6.2831853071795864769;

    // ---- ring, clockwise: edge k of the ccw polygon is walked from vertex k+1 to vertex k ----
    double hx[R], hy[R], qx[R], qy[R];
    int typ[R], nbr[R];
    int E = 0;
    const int mm = (live && ok) ? m : 0;
    const int mloop = LOCKSTEP ? __reduce_max_sync(FULL_MASK, mm) : mm;
    for (int kk = 0; kk < mloop; ++kk) {
        if (kk < mm && ok) {
            const int k = mm - 1 - kk;
            const int j = lab[k];
            if (j >= 0) {                                  // j < 0: side of the bounding square (the reference's ray-ray pair)
                if (E + 2 > R) ok = false;                 // an edge emits at most 2 entries
                else {
                    const int k1 = (k + 1 == mm) ? 0 : k + 1;
                    const double V1x = vx[k1], V1y = vy[k1], V2x = vx[k], V2y = vy[k];
                    const double ex = px[k], ey = py[k];
                    if (V1x * V1x + V1y * V1y <= r2) {     // inner start vertex: d1 <= r  (finite_voronoi.py:375-377)
                        hx[E] = V1x; hy[E] = V1y; qx[E] = ex; qy[E] = ey; typ[E] = 1; nbr[E] = j; ++E;
                    }
                    const double d2c = 0.25 * (ex * ex + ey * ey);   // |C|^2, C = p/2 the foot of the bisector
                    if (d2c < r2) {                        // bisector crosses the disk (finite_voronoi.py:395-407)
                        // clockwise unit vector along the bisector: u = (ey, -ex)/|p|
                        const double inv = 0.5 / sqrt(d2c);
                        const double ux = ey * inv, uy = -ex * inv;
                        const double t1 = V1x * ux + V1y * uy, t2 = V2x * ux + V2y * uy;
                        const double tr = sqrt(r2 - d2c);
                        const double Cx = 0.5 * ex, Cy = 0.5 * ey;
                        if (-tr < t2 && -tr > t1) {        // entry point: the straight edge starts on the circle
                            hx[E] = Cx - tr * ux; hy[E] = Cy - tr * uy; qx[E] = ex; qy[E] = ey; typ[E] = 1; nbr[E] = j; ++E;
                        }
                        if (tr < t2 && tr > t1) {          // exit point: an arc starts here
                            hx[E] = Cx + tr * ux; hy[E] = Cy + tr * uy; qx[E] = ex; qy[E] = ey; typ[E] = 0; nbr[E] = ARC; ++E;
                        }
                    }
                }
            }
        }
    }

    if (live && !ok) {
        if (!SLOW) {  // queue the cell for the large-capacity kernel
            const int idx = atomicAdd(&a.flags[0], 1);
            if (idx < OVF_MAX) a.ovf_cells[idx] = s; else atomicAdd(&a.flags[3], 1);
            a.ring_len[s] = 0;
        } else {
            atomicAdd(&a.flags[3], 1);  // exceeded even the slow-path capacity: the host reports AFV_ERR_CAPACITY
        }
    }
    const bool act = live && (ok || SLOW);     // this lane writes results
    if (!ok) E = 0;

    // ---- measures (cell_geom.pyx:402-475) ----
    const bool ringed = act && E >= 2;
    if (act && E < 2) {
        a.area[s] = PI * r2; a.perim[s] = TWO_PI * r; a.arcl[s] = TWO_PI * r; a.coord[s] = 0;
        a.ring_len[s] = 0;
    }
    const int EE = ringed ? E : 0;
    const int Eloop = LOCKSTEP ? __reduce_max_sync(FULL_MASK, EE) : EE;
    double ux_[R], uy_[R];   // unit vector (v1 - v2)/|v1 - v2| of straight edge e, 0 for arcs
    int arc_e[R / 2 + 1];
    int n_arc = 0;
    double A = 0.0, P = 0.0, Larc = 0.0;
    int z = 0;
    for (int e = 0; e < Eloop; ++e) {
        if (e < EE) {
            const int e2 = (e + 1 == EE) ? 0 : e + 1;
            if (typ[e] == 1) {
                const double v1x = hx[e], v1y = hy[e], v2x = hx[e2], v2y = hy[e2];
                const double sx = v1x - v2x, sy = v1y - v2y;
                const double l2 = sx * sx + sy * sy;
                const double l = sqrt(l2);
                P += l;
                A += -0.5 * (v1x * v2y - v1y * v2x);
                const double il = l > 0.0 ? 1.0 / l : 0.0;
                ux_[e] = sx * il; uy_[e] = sy * il;
                ++z;
            } else {
                ux_[e] = 0.0; uy_[e] = 0.0;
                if (n_arc < R / 2 + 1) arc_e[n_arc++] = e;
            }
        }
    }
    // arcs, compacted so that the lanes of a warp evaluate their atan2 together
    const int aloop = LOCKSTEP ? __reduce_max_sync(FULL_MASK, n_arc) : n_arc;
    for (int q = 0; q < aloop; ++q) {
        if (q < n_arc) {
            const int e = arc_e[q];
            const int e2 = (e + 1 == EE) ? 0 : e + 1;
            const double v1x = hx[e], v1y = hy[e], v2x = hx[e2], v2y = hy[e2];
            double ang = atan2(v2x * v1y - v2y * v1x, v1x * v2x + v1y * v2y);  // angle(v1) - angle(v2)
            if (ang < 0.0) ang += TWO_PI;
            Larc += r * ang;
            A += 0.5 * r2 * ang;
        }
    }
    P += Larc;

    RingShared *rs = nullptr;       // only pooled rings have per-cell records
    RingOwn *ro = a.ro + s;
    size_t ros = a.stride;          // RingOwn / RingSelf element stride (slot-major), 1 inside the overflow pool
    bool pooled = false;
    double cA = 0.0, cP = 0.0, cL = 0.0;
    if (ringed) {
        a.area[s] = A; a.perim[s] = P; a.arcl[s] = Larc; a.coord[s] = z;
        if (SLOW && E > RS) {
            a.ring_len[s] = RING_IN_POOL; a.ovf_index[s] = pool_idx; a.ovf_len[pool_idx] = E;
            rs = a.ovf_rs + (size_t)pool_idx * RS_SLOW; ro = a.ovf_ro + (size_t)pool_idx * RS_SLOW; ros = 1; pooled = true;
        } else {
            a.ring_len[s] = (unsigned char)E;
        }
        if (E > 12) atomicMax(&a.flags[2], E);
        cA = 2.0 * a.ph.KA * (A - a.a0[s]);
        cP = 2.0 * a.ph.KP * (P - a.ph.P0);
        cL = cP + a.ph.Lambda;
    }

    // ---- per-vertex gradients of this cell (cell_geom.pyx:478-539) ----
    for (int e = 0; e < Eloop; ++e) {
        if (e < EE) {
            const int e2 = (e + 1 == EE) ? 0 : e + 1, e0 = (e == 0) ? EE - 1 : e - 1;
            // dA/dv1 = 0.5 (perp(v0) - perp(v2)), perp(u) = (uy, -ux);  dP/dv1 = u(e) - u(e-1)
            const double gx = cA * 0.5 * (hy[e0] - hy[e2]) + cP * (ux_[e] - ux_[e0]);
            const double gy = cA * 0.5 * (hx[e2] - hx[e0]) + cP * (uy_[e] - uy_[e0]);
            double w = 0.0;
            if (typ[e] == 0)        // an arc starts here: da = +0.5 r^2 (1 - cos D), dl = +r
                w = cA * 0.5 * (r2 - (hx[e] * hx[e2] + hy[e] * hy[e2])) + cL * r;
            else if (typ[e0] == 0)  // an arc ends here
                w = -(cA * 0.5 * (r2 - (hx[e0] * hx[e] + hy[e0] * hy[e])) + cL * r);
            RingOwn q; q.hx = hx[e]; q.hy = hy[e]; q.px = qx[e]; q.py = qy[e];
            ro[e * ros] = q;
            if (pooled) { RingShared o; o.nbr = nbr[e]; o.type = typ[e]; o.gx = gx; o.gy = gy; o.w = w; rs[e] = o; }
            else { RingSelf f; f.gx = gx; f.gy = gy; f.w = w; f.nbr = nbr[e]; f.type = typ[e]; a.self[e * a.stride + s] = f; }
        }
    }
    if (ringed && E <= RS) {   // the compact key row that k_force scans (entries beyond ring_len are stale)
        int *key = a.ring_key + (size_t)s * RS;
        for (int e = 0; e < EE; ++e) key[e] = nbr[e];
    }
}

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
constexpr int HULL_TPB = 128;
constexpr int KHULL = 16;        // hull vertices (= polygon edges) on the fast path
constexpr int CMAX_HARD = 128;   // upper bound of the per-cell candidate list (the launch picks cmax <= this); < 255 (8-bit hull indices)

#ifndef HULL_MIN_BLOCKS
#define HULL_MIN_BLOCKS 10
#endif
// SLAB = false: one box, every cell, rows of bins contiguous (the single-GPU path, nothing of the slab machinery is
// compiled in).  SLAB = true: an x-slab -- columns contiguous, pass filters, the late cells' bin table.
template <bool SLAB>
__global__ void __launch_bounds__(HULL_TPB, SLAB ? HULL_MIN_BLOCKS - 1 : HULL_MIN_BLOCKS) k_hull(GeomArgs a, const int cmax) {
    extern __shared__ int s_idx[];               // [cmax][HULL_TPB]: sorted slots of the in-range neighbours of each lane's cell
    const int tid = threadIdx.x;
    const int t_ = blockIdx.x * HULL_TPB + tid;
    bool live = t_ < a.n;
    const int s = SLAB ? slot_of_thread(a, t_) : t_;
    const double r = a.ph.r, cut2 = 4.0 * r * r;
    GridDesc g = *a.grid;
    if (!SLAB) g.swap = 0;
    int n = 0, imin = 0;
    bool ok = true;
    double xi = 0.0, yi = 0.0;
    bool wrapx = false, wrapy = false;            // only cells in the outermost bins can see a periodic image
    int bx = 0, by = 0;
    if (live) {
        bin_xy(g, a.keys[s], bx, by);
        if (SLAB) live = pass_selects(a, g, s, bx);
    }

    // ---- gather the neighbours within 2r: the sorted cells, then (face pass of a slab) the late cells ----
    if (live) {
        const double2 pi_ = a.xy[s];
        xi = pi_.x; yi = pi_.y;
        wrapx = g.perx && (bx == 0 || bx == g.nbx - 1 || g.nbx < 3);
        wrapy = g.pery && (by == 0 || by == g.nby - 1 || g.nby < 3);
        // lines of bins: rows, or columns in a swapped (slab) grid
        const int npos = g.swap ? g.nby : g.nbx, nline = g.swap ? g.nbx : g.nby, pos = g.swap ? by : bx;
        const bool per_pos = g.swap ? g.pery : g.perx, per_line = g.swap ? g.perx : g.pery;
        int rows[3];
        const int nrow = cell_rows(nline, per_line, g.swap ? bx : by, rows);
        double dmin = 1e300;
        for (int tab = 0; tab < (SLAB ? 2 : 1); ++tab) {
            const int *const bin_start = tab ? a.bin_start_late : a.bin_start;
            if (!bin_start) break;
            const int first = tab ? a.n_main : 0;
            for (int ir = 0; ir < nrow; ++ir) {
                int c0[2], c1[2];
                const int nrun = row_runs(npos, per_pos, bin_start, rows[ir], pos, c0, c1);
                for (int q = 0; q < nrun; ++q) {
#pragma unroll 4
                    for (int c = first + c0[q]; c < first + c1[q]; ++c) {
                        const double2 pc = a.xy[c];
                        double dx = pc.x - xi, dy = pc.y - yi;
                        if (wrapx) dx = min_image(dx, g.Lx, g.halfLx);
                        if (wrapy) dy = min_image(dy, g.Ly, g.halfLy);
                        const double d2 = dx * dx + dy * dy;
                        if (d2 < cut2 && c != s && d2 > 0.0) {
                            if (n < cmax) {
                                s_idx[n * HULL_TPB + tid] = c;
                                if (d2 < dmin) { dmin = d2; imin = n; }
                                ++n;
                            } else ok = false;
                        }
                    }
                }
            }
        }
    }
    // ---- gift wrapping in the dual plane.  List entry c <-> homogeneous dual point (p_c, k_c) with k_c = |p_c|^2
    //      (twice the true k; the factor cancels in every sign).  Entry n is the ORIGIN of the dual plane, (p, k) =
    //      ((0,0), 2): the "half-plane at infinity".  It becomes a hull vertex exactly when the neighbours do not
    //      surround the cell, i.e. when the Voronoi polygon is unbounded; the two polygon edges next to it then run
    //      off to infinity.
    //      orient(a, b, c) = sign((p_b k_a - p_a k_b) x (p_c k_a - p_a k_c));  the nearest neighbour is on the hull.
    //      Two independent running-best chains (even / odd list entries) halve the dependent latency.  The loop is
    //      compiled twice: warps none of whose cells touches a periodic face skip the minimum-image code. ----
    unsigned long long hull_lo = 0ull, hull_hi = 0ull;   // 16 x 8-bit list indices (n = the origin)
    int m = 0;
    const bool wrap = live && ok && n > 0;
    const int npair_max = (__reduce_max_sync(FULL_MASK, wrap ? n : 0) + 1) >> 1;
    auto hull_steps = [&](auto wrap_tag) {
        constexpr bool W = decltype(wrap_tag)::value;
        auto disp = [&](int c, double &dx, double &dy) {       // displacement of list entry c (< n)
            const double2 pc = a.xy[s_idx[c * HULL_TPB + tid]];
            dx = pc.x - xi; dy = pc.y - yi;
            if (W) {
                if (wrapx) dx = min_image(dx, g.Lx, g.halfLx);
                if (wrapy) dy = min_image(dy, g.Ly, g.halfLy);
            }
        };
        bool done = !wrap;
        int cur = imin;
        while (__any_sync(FULL_MASK, !done)) {
            double pax = 0.0, pay = 0.0, ka = 2.0;
            if (!done) {
                if (m < 8) hull_lo |= (unsigned long long)cur << (8 * m); else hull_hi |= (unsigned long long)cur << (8 * (m - 8));
                ++m;
                if (cur < n) { disp(cur, pax, pay); ka = pax * pax + pay * pay; }
            }
            int best0 = -1, best1 = -1;
            double u0x = 0.0, u0y = 0.0, u1x = 0.0, u1y = 0.0;
            for (int cp = 0; cp < npair_max; ++cp) {
                const int ca = 2 * cp, cb = 2 * cp + 1;
                if (!done && ca < n && ca != cur) {
                    double qx_, qy_; disp(ca, qx_, qy_);
                    const double kc = qx_ * qx_ + qy_ * qy_;
                    const double wx = qx_ * ka - pax * kc, wy = qy_ * ka - pay * kc;
                    if (best0 < 0 || u0x * wy - u0y * wx < 0.0) { best0 = ca; u0x = wx; u0y = wy; }
                }
                if (!done && cb < n && cb != cur) {
                    double qx_, qy_; disp(cb, qx_, qy_);
                    const double kc = qx_ * qx_ + qy_ * qy_;
                    const double wx = qx_ * ka - pax * kc, wy = qy_ * ka - pay * kc;
                    if (best1 < 0 || u1x * wy - u1y * wx < 0.0) { best1 = cb; u1x = wx; u1y = wy; }
                }
            }
            if (!done) {
                // merge the chains: the overall next hull vertex is the one with every other point on its left
                int best = best0;
                double ux = u0x, uy = u0y;
                if (best0 < 0 || (best1 >= 0 && u0x * u1y - u0y * u1x < 0.0)) { best = best1; ux = u1x; uy = u1y; }
                // the origin: w = 0 * k_a - p_a * k_O, direction -p_a.  An exact tie means q_a, the origin and the best
                // neighbour are collinear (e.g. collinear cells): the farther point wins, as in any gift wrapping;
                // |q_c - q_a| is proportional to |w_c| / k_c.
                if (cur != n) {
                    const double wx = -pax, wy = -pay;
                    bool take = best < 0;
                    if (!take) {
                        const double cr = ux * wy - uy * wx;
                        if (cr < 0.0) take = true;
                        else if (cr == 0.0) {
                            double bx_, by_;
                            disp(best, bx_, by_);
                            const double kb = bx_ * bx_ + by_ * by_;
                            take = (wx * wx + wy * wy) * kb * kb > (ux * ux + uy * uy);
                        }
                    }
                    if (take) best = n;
                }
                cur = best;
                if (cur == imin) done = true;
                else if (m >= KHULL) { ok = false; done = true; }
            }
        }
    };
    if (__any_sync(FULL_MASK, wrapx || wrapy)) hull_steps(std::true_type{}); else hull_steps(std::false_type{});

    // ---- hand the polygon to k_ring ----
    if (live) {
        if (!ok) {   // too many neighbours / hull vertices for the fast path: queue for the large-capacity kernel
            const int idx = atomicAdd(&a.flags[0], 1);
            if (idx < OVF_MAX) a.ovf_cells[idx] = s; else atomicAdd(&a.flags[3], 1);
            a.hull_len[s] = HULL_QUEUED;
        } else {
            const int mm = wrap ? m : 0;
            a.hull_len[s] = (unsigned char)mm;
            int *hs = a.hull_slot + s;                       // slot-major: edge k at hs[k * stride]
            for (int k = 0; k < mm; ++k) {
                const int h = (int)(((k < 8) ? (hull_lo >> (8 * k)) : (hull_hi >> (8 * (k - 8)))) & 0xffull);
                hs[k * a.stride] = (h < n) ? s_idx[h * HULL_TPB + tid] : -2;   // -2: the edge at infinity
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K2b: ring, measures and per-vertex gradients from the polygon found by k_hull.  One thread per cell, no shared
// memory, no per-thread arrays.
// ------------------------------------------------------------------------------------------------------
#ifndef RING_MIN_BLOCKS
#define RING_MIN_BLOCKS 6
#endif
// POOL = false: the fast path, every lane of the warp calls it (warp-uniform loop bounds), records slot-major, at most
// RS entries; a longer ring queues the cell.  POOL = true: one queued cell per thread (k_geometry_slow), records into
// slot `pool_idx` of the overflow pool, at most RING_POOL_FAST entries; returns false when even that is too short.
constexpr int RING_POOL_FAST = 32;               // bounded by the 32-bit arc mask
template <bool POOL, bool SLAB>
__device__ __forceinline__ bool ring_cell(const GeomArgs &a, const int s, const bool live_in, const int pool_idx) {
    bool live = live_in;
    const double r = a.ph.r, B4 = 4.0 * r;
    GridDesc g = *a.grid;
    if (!SLAB) g.swap = 0;
    double xi = 0.0, yi = 0.0;
    bool wrapx = false, wrapy = false;
    int m = 0;
    bool ok = true;
    const int *hs = a.hull_slot + (live ? s : 0);      // slot-major: edge k at hs[k * a.stride]
    if (live) {
        int bx, by;
        bin_xy(g, a.keys[s], bx, by);
        if (!POOL && SLAB) live = pass_selects(a, g, s, bx);
        if (live) {
            const double2 pi_ = a.xy[s];
            xi = pi_.x; yi = pi_.y;
            wrapx = g.perx && (bx == 0 || bx == g.nbx - 1 || g.nbx < 3);
            wrapy = g.pery && (by == 0 || by == g.nby - 1 || g.nby < 3);
            m = a.hull_len[s];
        }
    }
    const bool queued = (m == HULL_QUEUED);
    if (queued) m = 0;
    // displacement of polygon edge k's neighbour
    auto disp = [&](int k, double &dx, double &dy, int &slot) {
        slot = hs[k * a.stride];
        if (slot >= 0) {
            const double2 pc = a.xy[slot];
            dx = pc.x - xi; dy = pc.y - yi;
            if (wrapx) dx = min_image(dx, g.Lx, g.halfLx);
            if (wrapy) dy = min_image(dy, g.Ly, g.halfLy);
        } else { dx = 0.0; dy = 0.0; }
    };

    // ---- ring, streamed.  The polygon edges are the hull vertices h_0..h_{m-1} (counter-clockwise); the ring is
    //      clockwise, so edge k is walked from vertex k+1 to vertex k for k = m-1 .. 0 (finite_voronoi.py:363-459).
    //      Ring records are written as they are produced; their gx/gy slots double as scratch (1/length of a
    //      straight edge, or the (cross, dot) of an arc's end points) until the last pass overwrites them, so the
    //      kernel keeps no per-thread arrays at all. ----
    const double r2 = r * r;
    const double PI = 3.14159265358979323846, TWO_PI = 6.2831853071795864769;
    // slot-major: entry e at ro[e * ros]; inside the pool the records of a cell are contiguous (RingShared and RingSelf
    // have the same layout)
    RingOwn *const ro = POOL ? a.ovf_ro + (size_t)pool_idx * RS_SLOW : a.ro + s;
    RingSelf *const self = POOL ? reinterpret_cast<RingSelf *>(a.ovf_rs + (size_t)pool_idx * RS_SLOW) : a.self + s;
    const size_t ros = POOL ? 1 : a.stride;
    int *const key = a.ring_key + (size_t)s * RS;   // the compact key row k_force scans (fast path only)
    constexpr int ECAP = POOL ? RING_POOL_FAST : RS;
    const int mm = m;
    const int mloop = POOL ? mm : __reduce_max_sync(FULL_MASK, mm);
    // Junction between the counter-clockwise consecutive polygon edges A (slot ja, displacement a) and B (jb, b):
    // the END of A and the START of B.  Normally one point, the intersection of the two bisectors (a proper left
    // turn, a x b > 0).  Next to the edge at infinity -- or for a degenerate turn -- the bisector runs off to
    // infinity instead: a point 4r along it stands in (it only has to lie outside the disk).
    auto junction = [&](int ja, double ax, double ay, int jb, double bx_, double by_, double &eax, double &eay, double &sbx, double &sby) {
        const double det = ax * by_ - ay * bx_;
        if (ja >= 0 && jb >= 0 && det > 0.0) {
            const double ca = 0.5 * (ax * ax + ay * ay), cb = 0.5 * (bx_ * bx_ + by_ * by_);
            const double idet = __drcp_rn(det);
            eax = sbx = (ca * by_ - cb * ay) * idet;
            eay = sby = (ax * cb - bx_ * ca) * idet;
        } else {
            eax = eay = sbx = sby = 0.0;
            if (ja >= 0) { const double f = B4 * rsqrt(ax * ax + ay * ay); eax = 0.5 * ax - f * ay; eay = 0.5 * ay + f * ax; }      // C_A + 4r ccw(A)
            if (jb >= 0) { const double f = B4 * rsqrt(bx_ * bx_ + by_ * by_); sbx = 0.5 * bx_ + f * by_; sby = 0.5 * by_ - f * bx_; } // C_B - 4r ccw(B)
        }
    };
    int E = 0, z = 0;
    unsigned arcmask = 0u;                       // bit e set: ring entry e starts an arc
    double A = 0.0, P = 0.0;
    double x0 = 0.0, y0 = 0.0;                   // first ring entry
    double xp = 0.0, yp = 0.0;                   // previous ring entry (its record is written when the next one arrives)
    int nbr_p = ARC;
    // connect the previous entry to (x, y): measure the edge, then write the previous entry's shared record
    auto close_prev = [&](double x, double y) {
        RingSelf o; o.nbr = nbr_p; o.type = (nbr_p != ARC) ? 1 : 0; o.w = 0.0;
        if (nbr_p != ARC) {
            const double sx = xp - x, sy = yp - y;
            const double l2 = sx * sx + sy * sy;
            const double il = l2 > 0.0 ? rsqrt(l2) : 0.0;
            P += l2 * il;
            A += -0.5 * (xp * y - yp * x);
            o.gx = il; o.gy = 0.0;
            ++z;
        } else {
            o.gx = x * yp - y * xp;              // v2 x v1
            o.gy = xp * x + yp * y;              // v1 . v2
        }
        self[(E - 1) * ros] = o;
        if (!POOL) key[E - 1] = nbr_p;
    };
    auto emit = [&](double x, double y, double ex, double ey, int nb) {
        if (E >= ECAP) { ok = false; return; }
        if (E > 0) close_prev(x, y); else { x0 = x; y0 = y; }
        RingOwn q; q.hx = x; q.hy = y; q.px = ex; q.py = ey;
        ro[E * ros] = q;
        if (nb == ARC) arcmask |= 1u << E;
        xp = x; yp = y; nbr_p = nb; ++E;
    };
    for (int k = 0; k < mm; ++k) {               // pull the neighbours' positions towards L1 before the dependent walk
        const int sl = hs[k * a.stride];
        if (sl >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.xy + sl));
    }
    {
        double pcx = 0.0, pcy = 0.0, pnx = 0.0, pny = 0.0, V1x = 0.0, V1y = 0.0;
        int jc = -2, jn = -2;
        if (mm > 0) {
            double tx, ty;
            disp(mm - 1, pcx, pcy, jc);
            disp(0, pnx, pny, jn);
            junction(jc, pcx, pcy, jn, pnx, pny, V1x, V1y, tx, ty);   // end of edge m-1 (junction with edge 0)
        }
        for (int kk = 0; kk < mloop; ++kk) {
            if (kk < mm) {
                const int k = mm - 1 - kk;
                double ppx, ppy, V2x, V2y, Enx, Eny;
                int jp;
                disp(k == 0 ? mm - 1 : k - 1, ppx, ppy, jp);
                junction(jp, ppx, ppy, jc, pcx, pcy, Enx, Eny, V2x, V2y);   // end of edge k-1, start of edge k
                if (jc >= 0 && ok) {                       // a real neighbour (virtual edges emit nothing)
                    const int j = jc;
                    if (V1x * V1x + V1y * V1y <= r2) emit(V1x, V1y, pcx, pcy, j);      // inner start vertex, d1 <= r
                    const double d2c = 0.25 * (pcx * pcx + pcy * pcy);               // |C|^2, C = p/2
                    if (d2c < r2) {                        // the bisector crosses the disk
                        const double inv = 0.5 * rsqrt(d2c);
                        const double ux = pcy * inv, uy = -pcx * inv;                  // clockwise unit vector along the edge
                        const double t1 = V1x * ux + V1y * uy, t2 = V2x * ux + V2y * uy;
                        const double tr = sqrt(r2 - d2c);
                        if (-tr < t2 && -tr > t1) emit(0.5 * pcx - tr * ux, 0.5 * pcy - tr * uy, pcx, pcy, j);   // entry point
                        if (tr < t2 && tr > t1) emit(0.5 * pcx + tr * ux, 0.5 * pcy + tr * uy, pcx, pcy, ARC);  // exit point
                    }
                }
                pcx = ppx; pcy = ppy; jc = jp; V1x = Enx; V1y = Eny;
            }
        }
    }
    const bool ringed = live && ok && !queued && E >= 2;
    if (ringed) close_prev(x0, y0);              // closing edge: last entry -> first entry

    if (POOL && !ok) return false;               // the caller falls back to the clipping path
    if (!POOL && live && !ok) {                  // ring longer than the stride: queue the cell for the large-capacity kernel
        const int idx = atomicAdd(&a.flags[0], 1);
        if (idx < OVF_MAX) a.ovf_cells[idx] = s; else atomicAdd(&a.flags[3], 1);
        a.ring_len[s] = 0;
    }
    if (live && queued) a.ring_len[s] = 0;       // k_geometry_slow fills everything in
    if (live && ok && !queued && E < 2) {                   // isolated cell (cell_geom.pyx:402-406)
        a.area[s] = PI * r2; a.perim[s] = TWO_PI * r; a.arcl[s] = TWO_PI * r; a.coord[s] = 0;
        a.ring_len[s] = 0;
    }

    // ---- arcs: the lanes of a warp evaluate their q-th arc together (cell_geom.pyx:454-469) ----
    double Larc = 0.0;
    {
        unsigned rem = ringed ? arcmask : 0u;
        const int aloop = POOL ? __popc(rem) : __reduce_max_sync(FULL_MASK, __popc(rem));
        for (int q = 0; q < aloop; ++q) {
            if (rem) {
                const int e = __ffs(rem) - 1;
                rem &= rem - 1;
                const RingSelf sc = self[e * ros];
                double ang = atan2(sc.gx, sc.gy);         // angle(v1) - angle(v2)
                if (ang < 0.0) ang += TWO_PI;
                Larc += r * ang;
                A += 0.5 * r2 * ang;
            }
        }
    }
    P += Larc;
    double cA = 0.0, cP = 0.0, cL = 0.0;
    if (ringed) {
        a.area[s] = A; a.perim[s] = P; a.arcl[s] = Larc; a.coord[s] = z;
        if (POOL) { a.ring_len[s] = RING_IN_POOL; a.ovf_index[s] = pool_idx; a.ovf_len[pool_idx] = E; }
        else a.ring_len[s] = (unsigned char)E;
        if (E > 12) atomicMax(&a.flags[2], E);
        cA = 2.0 * a.ph.KA * (A - a.a0[s]);
        cP = 2.0 * a.ph.KP * (P - a.ph.P0);
        cL = cP + a.ph.Lambda;
    }

    // ---- per-vertex gradients of this cell (cell_geom.pyx:478-539), streamed over the ring records ----
    {
        const int EE = ringed ? E : 0;
        const int Eloop = POOL ? EE : __reduce_max_sync(FULL_MASK, EE);
        // software pipeline: the record of entry e+2 (vertex position, 1/length scratch) is requested while vertex e is
        // being processed; entry e's scratch is only overwritten after its last use
        double hpx = 0.0, hpy = 0.0, hcx = 0.0, hcy = 0.0, hnx = 0.0, hny = 0.0, ilp = 0.0, ilc = 0.0, iln = 0.0;
        auto load_h = [&](int e, double &x, double &y) { const double2 v = *reinterpret_cast<const double2 *>(&ro[e * ros]); x = v.x; y = v.y; };
        auto load_il = [&](int e) { return ((arcmask >> e) & 1u) ? 0.0 : self[e * ros].gx; };
        if (EE > 0) {
            const int e1 = (1 == EE) ? 0 : 1;
            load_h(EE - 1, hpx, hpy); load_h(0, hcx, hcy); load_h(e1, hnx, hny);
            ilp = load_il(EE - 1); ilc = load_il(0); iln = load_il(e1);
        }
        for (int e = 0; e < Eloop; ++e) {
            if (e < EE) {
                const int e0 = (e == 0) ? EE - 1 : e - 1;
                int e3 = e + 2; if (e3 >= EE) e3 -= EE; if (e3 >= EE) e3 -= EE;
                double hqx, hqy;
                load_h(e3, hqx, hqy);
                const double ilq = load_il(e3);
                const bool arc_c = (arcmask >> e) & 1u, arc_p = (arcmask >> e0) & 1u;
                // u(e) = (v1 - v2)/|v1 - v2| for straight edges, 0 for arcs;  dP/dv1 = u(e) - u(e-1);  dA/dv1 = 0.5 (perp(v0) - perp(v2))
                const double uex = (hcx - hnx) * ilc, uey = (hcy - hny) * ilc;
                const double upx = (hpx - hcx) * ilp, upy = (hpy - hcy) * ilp;
                const double gx = cA * 0.5 * (hpy - hny) + cP * (uex - upx);
                const double gy = cA * 0.5 * (hnx - hpx) + cP * (uey - upy);
                double w = 0.0;
                if (arc_c)       // an arc starts here: da = +0.5 r^2 (1 - cos D), dl = +r
                    w = cA * 0.5 * (r2 - (hcx * hnx + hcy * hny)) + cL * r;
                else if (arc_p)  // an arc ends here
                    w = -(cA * 0.5 * (r2 - (hpx * hcx + hpy * hcy)) + cL * r);
                // final record (nbr / type were written when the entry was closed)
                double *dst = reinterpret_cast<double *>(&self[e * ros]);
                dst[0] = gx; dst[1] = gy; dst[2] = w;
                hpx = hcx; hpy = hcy; hcx = hnx; hcy = hny; hnx = hqx; hny = hqy; ilp = ilc; ilc = iln; iln = ilq;
            }
        }
    }
    return true;
}

template <bool SLAB>
__global__ void __launch_bounds__(HULL_TPB, RING_MIN_BLOCKS) k_ring(GeomArgs a) {
    const int t = blockIdx.x * HULL_TPB + threadIdx.x;
    ring_cell<false, SLAB>(a, SLAB ? slot_of_thread(a, t) : t, t < a.n, 0);
}

// ------------------------------------------------------------------------------------------------------
// slow path: the (few) queued cells, one thread each, in-place clipping with large local arrays
// ------------------------------------------------------------------------------------------------------
__device__ __noinline__ void geom_cell_slow(const GeomArgs &a, const int s, const int pool_idx) {
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
                    if (g.perx) dx = min_image(dx, g.Lx, g.halfLx);
                    if (g.pery) dy = min_image(dy, g.Ly, g.halfLy);
                    const double d2 = dx * dx + dy * dy;
                    if (d2 >= cut2 || d2 == 0.0) continue;
                    ok = ok && clip_polygon<K>(vx, vy, px, py, lab, m, dx, dy, 0.5 * d2, c);
                }
            }
        }
    }
    finish_cell<K, RS_SLOW, true>(a, s, pool_idx, true, ok, vx, vy, px, py, lab, m);
}

__global__ void __launch_bounds__(64) k_geometry_slow(GeomArgs a) {
    const int cnt = min(a.flags[0], OVF_MAX);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < cnt; idx += gridDim.x * blockDim.x) {
        const int s = a.ovf_cells[idx];
        // queued by k_ring (its polygon is fine, only the ring outgrew the stride): emit the ring again, into the pool;
        // queued by k_hull, or a ring beyond RING_POOL_FAST: clip from scratch
        if (a.hull_len[s] == HULL_QUEUED || !ring_cell<true, true>(a, s, true, idx)) geom_cell_slow(a, s, idx);
    }
}
