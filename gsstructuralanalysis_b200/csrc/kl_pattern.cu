// kl_pattern.cu — one-time symbolic build ON THE GPU of
//   (1) the compressed pattern (outer/inner) of the structurally symmetric stiffness matrix,
//       bit-exact with what gsExprAssembler::initSystem + coeffRef + makeCompressed leave behind
//       when no local entry is an exact zero, including eliminated and
//       matched (clamped / collapsed) DoFs of gsDofMapper (Appendix A.6);
//   (2) the scatter table  pos[(J,d)][stencil(I-J)][c] -> index into the value array.
// Replaces: gsExprAssembler::initSystem / gsSparseMatrix::reservePerColumn + makeCompressed
// (reference consumers: gsSparseMatrix in src/gsStructuralAnalysisTools/gsStructuralAnalysisTypes.h:88).
#include <cub/cub.cuh>
#include "kl_internal.h"

struct PatArgs {
    int p, n1, n2, ncp, nfree, nst;
    const int* map;
    const int *lo1, *hi1, *lo2, *hi2;   // element range of every 1-D basis function
};

__device__ __forceinline__ bool coupled(const PatArgs& a, int I1, int I2, int J1, int J2) {
    if (J1 < 0 || J1 >= a.n1 || J2 < 0 || J2 >= a.n2) return false;
    return !(a.hi1[J1] < a.lo1[I1] || a.lo1[J1] > a.hi1[I1] || a.hi2[J2] < a.lo2[I2] || a.lo2[J2] > a.hi2[I2]);
}

// one key per (column control point J, stencil slot, d, c); invalid slots get the sentinel
__global__ void k_gen_keys(PatArgs a, unsigned long long* __restrict__ keys, long long total) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const int c = (int)(k % 3);
    const int d = (int)((k / 3) % 3);
    const int st = (int)((k / 9) % a.nst);
    const int J = (int)(k / (9LL * a.nst));
    const int W = 2 * a.p + 1;
    const int J1 = J % a.n1, J2 = J / a.n1;
    const int I1 = J1 + (st % W) - a.p, I2 = J2 + (st / W) - a.p;
    unsigned long long key = ~0ULL;
    if (coupled(a, J1, J2, I1, I2)) {
        const int col = a.map[d * a.ncp + J];
        const int row = a.map[c * a.ncp + (I1 + a.n1 * I2)];
        if (col < a.nfree && row < a.nfree) key = ((unsigned long long)col << 32) | (unsigned int)row;
    }
    keys[k] = key;
}

__global__ void k_outer_inner(const unsigned long long* __restrict__ ukeys, long long nnz, int nfree, int* __restrict__ outer,
                              int* __restrict__ inner) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) inner[k] = (int)(ukeys[k] & 0xffffffffULL);
    if (k <= nfree) {
        // lower_bound of (k << 32)
        const unsigned long long target = (unsigned long long)k << 32;
        long long lo = 0, hi = nnz;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (ukeys[mid] < target) lo = mid + 1; else hi = mid;
        }
        outer[k] = (int)lo;
    }
}

__global__ void k_pos_table(PatArgs a, const int* __restrict__ outer, const int* __restrict__ inner, int* __restrict__ pos, long long total) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    // layout: pos[(J*3 + d) * (nst*3) + st*3 + c]
    const int c = (int)(k % 3);
    const int st = (int)((k / 3) % a.nst);
    const int d = (int)((k / (3LL * a.nst)) % 3);
    const int J = (int)(k / (9LL * a.nst));
    const int W = 2 * a.p + 1;
    const int J1 = J % a.n1, J2 = J / a.n1;
    const int I1 = J1 + (st % W) - a.p, I2 = J2 + (st / W) - a.p;
    int res = -1;
    if (coupled(a, J1, J2, I1, I2)) {
        const int col = a.map[d * a.ncp + J];
        const int row = a.map[c * a.ncp + (I1 + a.n1 * I2)];
        if (col < a.nfree && row < a.nfree) {
            int lo = outer[col], hi = outer[col + 1] - 1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1;
                const int v = inner[mid];
                if (v == row) { res = mid; break; }
                if (v < row) lo = mid + 1; else hi = mid - 1;
            }
        }
    }
    pos[k] = res;
}

// colbase[J][d] = outer[map[d][J]] (or -1), colbase[J][3] = 1 when every column (J,d) holds exactly the full
// (2p+1)^2 x 3 stencil in canonical order, i.e. pos[(J,d)][st][c] == outer[col] + c*nst + st for all st,c,d.
__global__ void k_colbase(PatArgs a, const int* __restrict__ outer, const int* __restrict__ pos, int* __restrict__ colbase) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    if (J >= a.ncp) return;
    int regular = 1;
    for (int d = 0; d < 3; ++d) {
        const int col = a.map[d * a.ncp + J];
        const int base = col < a.nfree ? outer[col] : -1;
        colbase[4 * J + d] = base;
        if (base < 0) { regular = 0; continue; }
        const int* pp = pos + (size_t)(J * 3 + d) * (a.nst * 3);
        for (int st = 0; st < a.nst && regular; ++st)
            for (int c = 0; c < 3; ++c)
                if (pp[st * 3 + c] != base + c * a.nst + st) { regular = 0; break; }
    }
    colbase[4 * J + 3] = regular;
}

template <class T>
static int dev_alloc(kl_ctx* ctx, T** p, size_t n) {
    KL_CUDA(cudaMalloc((void**)p, sizeof(T) * (n ? n : 1)));
    ctx->owned.push_back((void*)*p);
    return 0;
}

static int pat_args(kl_ctx* ctx, PatArgs* out) {
    KLDev& d = ctx->d;
    PatArgs a{};
    a.p = d.p; a.n1 = d.n1; a.n2 = d.n2; a.ncp = d.ncp; a.nfree = d.nfree; a.nst = d.nst;
    a.map = d.map;
    if (!ctx->d_flohi[0]) {
        const std::vector<int>* src[4] = {&ctx->flo[0], &ctx->fhi[0], &ctx->flo[1], &ctx->fhi[1]};
        for (int k = 0; k < 4; ++k) {
            if (int rc = dev_alloc(ctx, &ctx->d_flohi[k], src[k]->size())) return rc;
            KL_CUDA(cudaMemcpy(ctx->d_flohi[k], src[k]->data(), sizeof(int) * src[k]->size(), cudaMemcpyHostToDevice));
        }
    }
    a.lo1 = ctx->d_flohi[0]; a.hi1 = ctx->d_flohi[1]; a.lo2 = ctx->d_flohi[2]; a.hi2 = ctx->d_flohi[3];
    *out = a;
    return 0;
}

long long kl_pattern_key_count(const kl_ctx* ctx) { return 9LL * ctx->d.nst * ctx->d.ncp; }

int kl_pattern_gen_keys(kl_ctx* ctx, unsigned long long* keys) {
    PatArgs a;
    if (int rc = pat_args(ctx, &a)) return rc;
    const long long total = kl_pattern_key_count(ctx);
    if (total >= (1LL << 31)) { kl_set_error("kl_create: 9 * (2p+1)^2 * n_cp exceeds int32 (index_t); the mesh is too large for one context"); return KL_E_ARG; }
    const int T = 256;
    k_gen_keys<<<(unsigned)((total + T - 1) / T), T>>>(a, keys, total);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    return 0;
}

// keys (possibly of several patches) -> compressed pattern; keys / keys_alt are scratch of `total` entries each
int kl_pattern_compress(kl_ctx* owner, unsigned long long* keys, unsigned long long* keys_alt, long long total, int nfree,
                        int** outer_out, int** inner_out, long long* nnz_out) {
    const int T = 256;
    long long* d_num = nullptr;
    KL_CUDA(cudaMalloc(&d_num, sizeof(long long)));
    // radix sort (all 64 bits: the sentinel must end up last)
    cub::DoubleBuffer<unsigned long long> db(keys, keys_alt);
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    KL_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, db, total));
    KL_CUDA(cudaMalloc(&tmp, tmp_bytes));
    KL_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, db, total));
    KL_CUDA(cudaFree(tmp));
    unsigned long long* sorted = db.Current();
    unsigned long long* ukeys = (sorted == keys) ? keys_alt : keys;
    tmp = nullptr; tmp_bytes = 0;
    KL_CUDA(cub::DeviceSelect::Unique(tmp, tmp_bytes, sorted, ukeys, d_num, total));
    KL_CUDA(cudaMalloc(&tmp, tmp_bytes));
    KL_CUDA(cub::DeviceSelect::Unique(tmp, tmp_bytes, sorted, ukeys, d_num, total));
    KL_CUDA(cudaFree(tmp));
    long long nuniq = 0;
    KL_CUDA(cudaMemcpy(&nuniq, d_num, sizeof(long long), cudaMemcpyDeviceToHost));
    KL_CUDA(cudaFree(d_num));
    // drop the sentinel if present (it is the largest key)
    if (nuniq > 0) {
        unsigned long long last = 0;
        KL_CUDA(cudaMemcpy(&last, ukeys + (nuniq - 1), sizeof(last), cudaMemcpyDeviceToHost));
        if (last == ~0ULL) --nuniq;
    }
    if (nuniq >= (1LL << 31)) { kl_set_error("nnz exceeds int32 (index_t)"); return KL_E_ARG; }
    int *outer, *inner;
    if (int rc = dev_alloc(owner, &outer, (size_t)nfree + 1)) return rc;
    if (int rc = dev_alloc(owner, &inner, (size_t)nuniq)) return rc;
    {
        const long long n = std::max<long long>(nuniq, (long long)nfree + 1);
        k_outer_inner<<<(unsigned)((n + T - 1) / T), T>>>(ukeys, nuniq, nfree, outer, inner);
        owner->launches++;
        KL_CUDA(cudaGetLastError());
    }
    KL_CUDA(cudaDeviceSynchronize());
    *outer_out = outer; *inner_out = inner; *nnz_out = nuniq;
    return 0;
}

int kl_pattern_tables(kl_ctx* ctx) {
    KLDev& d = ctx->d;
    PatArgs a;
    if (int rc = pat_args(ctx, &a)) return rc;
    const long long total = kl_pattern_key_count(ctx);
    const int T = 256;
    int* pos;
    if (int rc = dev_alloc(ctx, &pos, (size_t)total)) return rc;
    k_pos_table<<<(unsigned)((total + T - 1) / T), T>>>(a, d.outer, d.inner, pos, total);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    int* colbase;
    if (int rc = dev_alloc(ctx, &colbase, (size_t)4 * d.ncp)) return rc;
    k_colbase<<<(d.ncp + T - 1) / T, T>>>(a, d.outer, pos, colbase);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    KL_CUDA(cudaDeviceSynchronize());
    d.colbase = colbase;
    d.pos = pos;
    return 0;
}

int kl_build_pattern(kl_ctx* ctx) {
    KLDev& d = ctx->d;
    const long long total = kl_pattern_key_count(ctx);
    if (total >= (1LL << 31)) { kl_set_error("kl_create: 9 * (2p+1)^2 * n_cp exceeds int32 (index_t); the mesh is too large for one context"); return KL_E_ARG; }
    unsigned long long *keys = nullptr, *keys_alt = nullptr;
    KL_CUDA(cudaMalloc(&keys, sizeof(unsigned long long) * total));
    KL_CUDA(cudaMalloc(&keys_alt, sizeof(unsigned long long) * total));
    int rc = kl_pattern_gen_keys(ctx, keys);
    int *outer = nullptr, *inner = nullptr;
    long long nnz = 0;
    if (!rc) rc = kl_pattern_compress(ctx, keys, keys_alt, total, d.nfree, &outer, &inner, &nnz);
    cudaFree(keys);
    cudaFree(keys_alt);
    if (rc) return rc;
    ctx->nnz = nnz;
    d.outer = outer; d.inner = inner;
    if ((rc = kl_pattern_tables(ctx))) return rc;
    double* values;
    if ((rc = dev_alloc(ctx, &values, (size_t)nnz))) return rc;
    KL_CUDA(cudaMemset(values, 0, sizeof(double) * (nnz ? nnz : 1)));
    KL_CUDA(cudaDeviceSynchronize());
    d.values = values;
    return 0;
}


// ------------------------------------------------------------------------------------------------
// Lower-triangular view (row >= col) of the symmetric stiffness matrix for consumers that only read one triangle
// (gsSparseSolver<>::SimplicialLDLT of benchmarks/benchmark_Roof.cpp:359-360 factorises selfadjointView<Lower>): the same
// compressed column layout with the upper entries dropped, i.e. the tail of every column (row indices ascend).  Half the
// bytes cross PCIe.
__global__ void k_lower_count(const int* __restrict__ outer, const int* __restrict__ inner, int n, int* __restrict__ cnt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int lo = outer[j], hi = outer[j + 1];
    const int end = hi;
    while (lo < hi) {                       // first entry with row >= j
        const int mid = (lo + hi) >> 1;
        if (inner[mid] < j) lo = mid + 1; else hi = mid;
    }
    cnt[j] = end - lo;
}
// one warp per column: copy the tail (inner indices once, values every call)
template <bool INNER>
__global__ void k_lower_pack(const int* __restrict__ outer, const int* __restrict__ outer_lo, const int* __restrict__ inner,
                             const double* __restrict__ val, int col_begin, int col_end, int* __restrict__ inner_lo, double* __restrict__ val_lo) {
    const int lane = threadIdx.x & 31;
    const int j = col_begin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= col_end) return;
    const int a = outer_lo[j], b = outer_lo[j + 1];
    const int src = outer[j + 1] - (b - a);
    for (int k = lane; k < b - a; k += 32) {
        if (INNER) inner_lo[a + k] = inner[src + k];
        else val_lo[a + k] = val[src + k];
    }
}

int kl_lower_tables(kl_ctx* ctx) {
    if (ctx->d_outer_lower) return 0;
    const KLDev& d = ctx->d;
    const int n = d.nfree, T = 256;
    int* cnt;
    if (int rc = dev_alloc(ctx, &cnt, (size_t)n + 1)) return rc;
    KL_CUDA(cudaMemset(cnt, 0, sizeof(int) * ((size_t)n + 1)));
    k_lower_count<<<(n + T - 1) / T, T>>>(d.outer, d.inner, n, cnt);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    int* outer_lo;
    if (int rc = dev_alloc(ctx, &outer_lo, (size_t)n + 1)) return rc;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    KL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, outer_lo, n + 1));
    KL_CUDA(cudaMalloc(&tmp, tmp_bytes));
    KL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, outer_lo, n + 1));
    KL_CUDA(cudaDeviceSynchronize());
    KL_CUDA(cudaFree(tmp));
    ctx->h_outer_lower.resize((size_t)n + 1);
    KL_CUDA(cudaMemcpy(ctx->h_outer_lower.data(), outer_lo, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost));
    ctx->nnz_lower = ctx->h_outer_lower[n];
    if (int rc = dev_alloc(ctx, &ctx->d_inner_lower, (size_t)ctx->nnz_lower)) return rc;
    if (int rc = dev_alloc(ctx, &ctx->d_values_lower, (size_t)ctx->nnz_lower)) return rc;
    if (n > 0) {
        k_lower_pack<true><<<(n + 7) / 8, T>>>(d.outer, outer_lo, d.inner, nullptr, 0, n, ctx->d_inner_lower, nullptr);
        ctx->launches++;
        KL_CUDA(cudaGetLastError());
    }
    KL_CUDA(cudaDeviceSynchronize());
    ctx->d_outer_lower = outer_lo;
    return 0;
}

int kl_launch_pack_lower(kl_ctx* ctx, int col_begin, int col_end, cudaStream_t s) {
    if (col_end <= col_begin) return 0;
    k_lower_pack<false><<<(col_end - col_begin + 7) / 8, 256, 0, s>>>(ctx->d.outer, ctx->d_outer_lower, nullptr, ctx->d.values, col_begin, col_end,
                                                                     nullptr, ctx->d_values_lower);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    return 0;
}
