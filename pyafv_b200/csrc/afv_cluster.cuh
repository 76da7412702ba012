// Connected components of the contact graph on the device (the reference does this on the host with
// scipy.sparse.csgraph on the (K,2) connections array, pyafv/_connectutil.py:101-172).  Included by afv_kernels.cuh inside
// namespace afv.
//
// Union by minimum with full path compression, in the index space of the caller (gid): every pass hooks, for every
// contact, the larger of the two current roots onto the smaller one (atomicMin), then compresses every path; passes
// repeat until one changes nothing.  O(log N) passes; the final root of a component is its smallest cell index, so
// numbering the roots in ascending order reproduces scipy's component labels exactly.

__device__ __forceinline__ void cc_hook(int *__restrict__ parent, int a, int b, int *__restrict__ changed) {
    const int ra = parent[a], rb = parent[b];
    if (ra == rb) return;
    const int hi = max(ra, rb), lo = min(ra, rb);
    atomicMin(&parent[hi], lo);
    *changed = 1;
}

__global__ void k_cc_init(int *__restrict__ parent, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = i;
}

// contacts straight from the rings (no (K,2) list is ever built): ring entry with a neighbour <=> contact
__global__ void k_cc_hook_rings(RingView rv, const int64_t *__restrict__ gid, const unsigned char *__restrict__ owned, int n,
                                int *__restrict__ parent, int *__restrict__ changed) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || owned[s] != CELL_OWNED) return;
    const double2 *h; const int *nb; size_t ros;
    const int E = rv.get(s, h, nb, ros);
    const int g = (int)gid[s];
    for (int e = 0; e < E; ++e) {
        const int j = nb[e * ros];
        if (j == ARC) continue;
        const int gj = (int)gid[j];
        if (g < gj) cc_hook(parent, g, gj, changed);          // every contact once
    }
}

// contacts from an explicit (K,2) int64 list
__global__ void k_cc_hook_edges(const int64_t *__restrict__ pairs, int64_t K, int n, int *__restrict__ parent, int *__restrict__ changed,
                                int *__restrict__ bad) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int64_t a = pairs[2 * k], b = pairs[2 * k + 1];
    if (a < 0 || b < 0 || a >= n || b >= n) { *bad = 1; return; }
    if (a != b) cc_hook(parent, (int)a, (int)b, changed);
}

__global__ void k_cc_compress(int *__restrict__ parent, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = parent[i];
    while (true) { const int q = parent[p]; if (q == p) break; p = q; }
    parent[i] = p;
}

__global__ void k_cc_root_flags(const int *__restrict__ parent, int n, int *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (parent[i] == i) ? 1 : 0;
}

// label = rank of the component's root among all roots (ascending); sizes by integer atomics (order independent)
__global__ void k_cc_labels(const int *__restrict__ parent, const int *__restrict__ root_rank, int n, int64_t *__restrict__ labels,
                            int *__restrict__ sizes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = root_rank[parent[i]];
    labels[i] = c;
    atomicAdd(&sizes[c], 1);
}
