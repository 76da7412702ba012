// Energy and FIRE relaxation on the device.  Included by afv_kernels.cuh inside namespace afv.
//
// E = sum_i KA (A_i - A0_i)^2 + KP (P_i - P0)^2 + Lambda L_arc,i is the energy whose gradient the reference's force
// code implements (finite_voronoi.py:529,746; the doublet closed form at physical_params.py:160-171); the reference
// never evaluates it.  All global sums are taken in a fixed order (grid-stride partial sums, shared-memory trees, one
// final block), so they are bit-reproducible run to run.

constexpr int SUMS_TPB = 256;
constexpr int SUMS_MAX_BLOCKS = 1024;
constexpr int N_SUMS = 5;      // F.v, v.v, F.F, max |F|^2, E

__device__ __forceinline__ double cell_energy(const PhysDev &ph, double A, double P, double Larc, double A0) {
    const double dA = A - A0, dP = P - ph.P0;
    return ph.KA * dA * dA + ph.KP * dP * dP + ph.Lambda * Larc;
}

__global__ void k_cell_energy(const double *__restrict__ area, const double *__restrict__ perim, const double *__restrict__ arcl,
                              const double *__restrict__ a0, const unsigned char *__restrict__ owned, int n, PhysDev ph, double *__restrict__ e) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) e[s] = owned[s] == CELL_OWNED ? cell_energy(ph, area[s], perim[s], arcl[s], a0[s]) : 0.0;
}

// partial[b * N_SUMS + q]: the five sums over the cells block b visits (grid-stride); vel (indexed by gid) may be null.
// slot_of_gid (may be null): visit the cells in the CALLER's order -- the order inside a bin of the cell list is not
// reproducible from run to run, the caller's order is, and with it every bit of the sums.
__global__ void __launch_bounds__(SUMS_TPB) k_sums_partial(const double *__restrict__ fx, const double *__restrict__ fy, const double2 *__restrict__ vel,
                                                           const int64_t *__restrict__ gid, const unsigned char *__restrict__ owned,
                                                           const double *__restrict__ area, const double *__restrict__ perim,
                                                           const double *__restrict__ arcl, const double *__restrict__ a0, int n, PhysDev ph,
                                                           const int *__restrict__ slot_of_gid, double *__restrict__ partial) {
    __shared__ double sm[N_SUMS][SUMS_TPB];
    double acc[N_SUMS] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = blockIdx.x * SUMS_TPB + threadIdx.x; i < n; i += gridDim.x * SUMS_TPB) {
        const int s = slot_of_gid ? slot_of_gid[i] : i;
        if (owned[s] != CELL_OWNED) continue;
        const double Fx = fx[s], Fy = fy[s];
        if (vel) { const double2 v = vel[gid[s]]; acc[0] += Fx * v.x + Fy * v.y; acc[1] += v.x * v.x + v.y * v.y; }
        const double f2 = Fx * Fx + Fy * Fy;
        acc[2] += f2;
        acc[3] = fmax(acc[3], f2);
        acc[4] += cell_energy(ph, area[s], perim[s], arcl[s], a0[s]);
    }
#pragma unroll
    for (int q = 0; q < N_SUMS; ++q) sm[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int w = SUMS_TPB / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
#pragma unroll
            for (int q = 0; q < N_SUMS; ++q)
                sm[q][threadIdx.x] = (q == 3) ? fmax(sm[q][threadIdx.x], sm[q][threadIdx.x + w]) : sm[q][threadIdx.x] + sm[q][threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x < N_SUMS) partial[blockIdx.x * N_SUMS + threadIdx.x] = sm[threadIdx.x][0];
}

// state of the FIRE integrator, resident on the device
struct FireState {
    double dt, alpha;          // current time step and mixing parameter
    double mix_a, mix_b;       // this iteration: v <- mix_a v + mix_b F   (then v += dt F, x += dt v)
    double E, fmax, P;         // energy, max |F|, power F.v of the configuration the forces belong to
    double dt_max;
    int npos, zero_v, iter, pad_;
};

// one block: final sums (fixed order) and the FIRE update (Bitzek et al., PRL 97, 170201: N_min = 5, f_inc = 1.1,
// f_dec = 0.5, alpha_start = 0.1, f_alpha = 0.99).  update == 0: only E, fmax, P are refreshed.
__global__ void __launch_bounds__(SUMS_TPB) k_fire_state(const double *__restrict__ partial, int nblocks, FireState *__restrict__ st, int update) {
    __shared__ double sm[N_SUMS][SUMS_TPB];
    double acc[N_SUMS] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < nblocks; b += SUMS_TPB) {
#pragma unroll
        for (int q = 0; q < N_SUMS; ++q) acc[q] = (q == 3) ? fmax(acc[q], partial[b * N_SUMS + q]) : acc[q] + partial[b * N_SUMS + q];
    }
#pragma unroll
    for (int q = 0; q < N_SUMS; ++q) sm[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int w = SUMS_TPB / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
#pragma unroll
            for (int q = 0; q < N_SUMS; ++q)
                sm[q][threadIdx.x] = (q == 3) ? fmax(sm[q][threadIdx.x], sm[q][threadIdx.x + w]) : sm[q][threadIdx.x] + sm[q][threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const double P = sm[0][0], vv = sm[1][0], ff = sm[2][0];
    st->E = sm[4][0]; st->fmax = sqrt(sm[3][0]); st->P = P;
    if (!update) return;
    if (P > 0.0) {
        st->mix_a = 1.0 - st->alpha;
        st->mix_b = ff > 0.0 ? st->alpha * sqrt(vv / ff) : 0.0;
        st->zero_v = 0;
        if (st->npos > 5) { st->dt = fmin(st->dt * 1.1, st->dt_max); st->alpha *= 0.99; }
        st->npos += 1;
    } else {
        st->mix_a = 0.0; st->mix_b = 0.0; st->zero_v = 1;
        if (st->iter > 0) st->dt *= 0.5;       // (the very first iteration starts from rest: P = 0 is no uphill move)
        st->alpha = 0.1; st->npos = 0;
    }
    st->iter += 1;
}

// v <- mix_a v + mix_b F (or 0 after an uphill move); v += dt F; x += dt v, wrapped into the box
__global__ void k_fire_move(double2 *__restrict__ xy, double2 *__restrict__ vel, const double *__restrict__ fx, const double *__restrict__ fy,
                            const int64_t *__restrict__ gid, const unsigned char *__restrict__ owned, int n, const FireState *__restrict__ st,
                            double Lx, double Ly) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || owned[s] != CELL_OWNED) return;
    const double dt = st->dt, ma = st->mix_a, mb = st->mix_b;
    const int64_t g = gid[s];
    const double Fx = fx[s], Fy = fy[s];
    double2 v = vel[g];
    if (st->zero_v) { v.x = 0.0; v.y = 0.0; }
    else { v.x = ma * v.x + mb * Fx; v.y = ma * v.y + mb * Fy; }
    v.x += dt * Fx; v.y += dt * Fy;
    vel[g] = v;
    double2 p = xy[s];
    p.x += dt * v.x; p.y += dt * v.y;
    if (Lx > 0.0) { p.x -= Lx * floor(p.x / Lx); if (p.x >= Lx) p.x = 0.0; }
    if (Ly > 0.0) { p.y -= Ly * floor(p.y / Ly); if (p.y >= Ly) p.y = 0.0; }
    xy[s] = p;
}
