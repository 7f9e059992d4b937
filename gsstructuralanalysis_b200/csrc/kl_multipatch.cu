// kl_multipatch.cu — several C0-coupled tensor-product patches behind ONE gsDofMapper numbering and ONE sparse matrix.
//
// Reference: a gsMultiPatch with computeTopology() / addInterface() handed to gsThinShellAssembler(mp, dbasis, bc, force, materialMatrix)
// (benchmarks/benchmark_Wrinkling.cpp:446-522, benchmarks/benchmark_cylinder_DC.cpp:146, benchmarks/benchmark_Pillow.cpp); the
// conforming interfaces are glued by gsDofMapper::matchDofs inside gsFeSpace::setupMapper, after which the
// element loop of gsExprAssembler simply runs over all patches and pushes into one matrix.
//
// Here every patch keeps its own context (tables, control net, scatter map, per-point records) whose dof_map holds GLOBAL indices;
// the pattern is the union of the patch patterns (all keys sorted once on the GPU), every patch gets a scatter table against that
// union, and the unchanged single-patch kernels assemble into the shared value array: a column that belongs to an interface DoF
// is an "irregular" column (table look-up instead of arithmetic addressing), everything else runs on the fast path.
// The kl_mp owns a geometry-free "matrix context" that all matrix / vector level entry points of kl_shell.h accept
// (kl_jacobian, kl_residual, kl_al_residual, kl_force, kl_mass, kl_cg_solve, kl_newton_solve, kl_alm_step, kl_fetch_values, ...).
// Patch -> GPU partition: kl_mp_set_active restricts the assembly to the patches this process owns; the interface
// columns are completed by the halo reduce of the host layer (gsstructuralanalysis_b200/parallel.py).
#include <algorithm>
#include <cstring>
#include "kl_internal.h"

static int mp_rebuild_loads(kl_mp* mp) {
    kl_ctx* g = mp->g;
    const int n = g->d.nfree;
    const size_t vb = sizeof(double) * (size_t)std::max(n, 1);
    KL_CUDA(cudaMemset(g->d_fext, 0, vb));
    KL_CUDA(cudaMemset(g->d_force, 0, vb));
    for (size_t q = 0; q < mp->patch.size(); ++q) {
        if (!mp->active[q]) continue;
        if (int rc = kl_launch_axpby(g, g->d_fext, mp->patch[q]->d_fext, 1.0, 1.0, n, 0)) return rc;
        if (int rc = kl_launch_axpby(g, g->d_force, mp->patch[q]->d_force, 1.0, 1.0, n, 0)) return rc;
    }
    KL_CUDA(cudaDeviceSynchronize());
    return KL_OK;
}

extern "C" void kl_mp_destroy(kl_mp* mp) {
    if (!mp) return;
    cudaSetDevice(mp->device);
    cudaDeviceSynchronize();
    for (kl_ctx* c : mp->patch) kl_destroy(c);
    for (cudaStream_t st : mp->pstream) cudaStreamDestroy(st);
    for (cudaEvent_t e : mp->pdone) cudaEventDestroy(e);
    if (mp->fork) cudaEventDestroy(mp->fork);
    if (mp->g) { mp->g->mp = nullptr; kl_destroy(mp->g); }
    delete mp;
}

extern "C" int kl_mp_create(int32_t n_patches, const kl_problem* probs, int device, kl_mp** out) {
    if (!probs || !out || n_patches < 1) { kl_set_error("kl_mp_create: bad argument"); return KL_E_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        kl_set_error("no CUDA device: libkl_shell has no CPU fallback");
        return KL_E_NOGPU;
    }
    for (int q = 1; q < n_patches; ++q)
        if (probs[q].n_free != probs[0].n_free || probs[q].n_fixed != probs[0].n_fixed) {
            kl_set_error("kl_mp_create: every patch must carry the GLOBAL n_free / n_fixed of the common DoF mapper");
            return KL_E_ARG;
        }
    if (device >= 0) KL_CUDA(cudaSetDevice(device));
    kl_mp* mp = new kl_mp();
    KL_CUDA(cudaGetDevice(&mp->device));
    const int n = probs[0].n_free;
    int rc = KL_OK;
#define MP_FAIL(code) do { rc = (code); if (rc) { kl_mp_destroy(mp); return rc; } } while (0)
#define MP_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { kl_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); kl_mp_destroy(mp); return KL_E_CUDA; } } while (0)
    // ---- the matrix context: no geometry, only the global vectors, streams and (below) the pattern
    kl_ctx* g = new kl_ctx();
    mp->g = g;
    g->mp = mp;
    g->device = mp->device;
    g->prob = probs[0];
    g->d.p = probs[0].degree[0];
    g->d.nst = (2 * g->d.p + 1) * (2 * g->d.p + 1);
    g->d.nfree = n;
    g->nfixed = probs[0].n_fixed;
    g->spec_allowed = 0;
    const size_t vb = sizeof(double) * (size_t)std::max(n, 1);
    double** vecs[] = {&g->d_x, &g->d_r, &g->d_fext, &g->d_force, &g->d_xstate};
    for (double** v : vecs) { MP_CUDA(cudaMalloc((void**)v, vb)); g->owned.push_back(*v); MP_CUDA(cudaMemset(*v, 0, vb)); }
    MP_CUDA(cudaMalloc((void**)&g->d.flag, sizeof(int))); g->owned.push_back(g->d.flag);
    MP_CUDA(cudaMemset(g->d.flag, 0, sizeof(int)));
    MP_CUDA(cudaMalloc((void**)&g->d_same, sizeof(int))); g->owned.push_back(g->d_same);
    MP_CUDA(cudaMallocHost((void**)&g->h_pinned_x, vb));
    MP_CUDA(cudaMallocHost((void**)&g->h_pinned_r, vb));
    MP_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    MP_CUDA(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
    for (auto& e : g->ev) MP_CUDA(cudaEventCreate(&e));
    g->strip_ev.resize(1); g->copy_ev.resize(1);
    MP_CUDA(cudaEventCreateWithFlags(&g->strip_ev[0], cudaEventDisableTiming));
    MP_CUDA(cudaEventCreateWithFlags(&g->copy_ev[0], cudaEventDisableTiming));
    MP_CUDA(cudaDeviceGetAttribute(&g->n_sm, cudaDevAttrMultiProcessorCount, g->device));
    // ---- patch contexts up to the pattern
    long long total = 0;
    for (int q = 0; q < n_patches; ++q) {
        kl_ctx* c = nullptr;
        MP_FAIL(kl_ctx_create_base(&probs[q], mp->device, &c));
        mp->patch.push_back(c);
        c->mp_member = 1;
        c->zero_ranges.assign(1, std::make_pair((size_t)0, (size_t)0));   // the kl_mp zeroes the shared value array once per assembly
        c->strip_ev.resize(1); c->copy_ev.resize(1);
        MP_CUDA(cudaEventCreateWithFlags(&c->strip_ev[0], cudaEventDisableTiming));
        MP_CUDA(cudaEventCreateWithFlags(&c->copy_ev[0], cudaEventDisableTiming));
        total += kl_pattern_key_count(c);
        mp->n_elements += (int64_t)c->d.nel1 * c->d.nel2;
        mp->n_qp += (int64_t)c->d.nel1 * c->d.nel2 * c->d.nq * c->d.nq;
        if (probs[q].pressure != 0.0) g->d.mat.pressure = probs[q].pressure;   // only read as "the tangent is unsymmetric"
    }
    mp->active.assign(n_patches, 1);
    // ---- one pattern for all patches: the keys of every patch sorted together
    {
        unsigned long long *keys = nullptr, *keys_alt = nullptr;
        MP_CUDA(cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)std::max<long long>(total, 1)));
        cudaError_t e2 = cudaMalloc(&keys_alt, sizeof(unsigned long long) * (size_t)std::max<long long>(total, 1));
        if (e2 != cudaSuccess) { cudaFree(keys); kl_set_error("kl_mp_create: out of device memory for the pattern keys"); kl_mp_destroy(mp); return KL_E_CUDA; }
        long long off = 0;
        for (kl_ctx* c : mp->patch) {
            if (!rc) rc = kl_pattern_gen_keys(c, keys + off);
            off += kl_pattern_key_count(c);
        }
        int *outer = nullptr, *inner = nullptr;
        long long nnz = 0;
        if (!rc) rc = kl_pattern_compress(g, keys, keys_alt, total, n, &outer, &inner, &nnz);
        cudaFree(keys);
        cudaFree(keys_alt);
        MP_FAIL(rc);
        g->d.outer = outer; g->d.inner = inner; g->nnz = nnz;
        double* values = nullptr;
        MP_CUDA(cudaMalloc((void**)&values, sizeof(double) * (size_t)std::max<long long>(nnz, 1)));
        g->owned.push_back(values);
        MP_CUDA(cudaMemset(values, 0, sizeof(double) * (size_t)std::max<long long>(nnz, 1)));
        g->d.values = values;
        g->h_outer.resize((size_t)n + 1);
        MP_CUDA(cudaMemcpy(g->h_outer.data(), outer, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost));
    }
    // ---- per patch: scatter tables against the union pattern, load vectors
    for (int q = 0; q < n_patches; ++q) {
        kl_ctx* c = mp->patch[q];
        c->d.outer = g->d.outer; c->d.inner = g->d.inner; c->d.values = g->d.values; c->nnz = g->nnz;
        MP_FAIL(kl_pattern_tables(c));
        MP_FAIL(kl_ctx_finish(c, &probs[q]));
    }
    MP_FAIL(mp_rebuild_loads(mp));
    // interface DoFs: free DoFs that more than one patch maps to
    {
        std::vector<int> last(n, -1);
        std::vector<char> shared(n, 0);
        for (int q = 0; q < n_patches; ++q) {
            const kl_ctx* c = mp->patch[q];
            for (size_t k = 0; k < c->h_map.size(); ++k) {
                const int gd = c->h_map[k];
                if (gd >= n) continue;
                if (last[gd] >= 0 && last[gd] != q) shared[gd] = 1;
                last[gd] = q;
            }
        }
        for (int i = 0; i < n; ++i) if (shared[i]) mp->coupled_cols.push_back(i);
    }
    *out = mp;
    return KL_OK;
#undef MP_FAIL
#undef MP_CUDA
}

extern "C" kl_ctx* kl_mp_context(kl_mp* mp) { return mp ? mp->g : nullptr; }
extern "C" kl_ctx* kl_mp_patch(kl_mp* mp, int32_t q) { return (mp && q >= 0 && q < (int)mp->patch.size()) ? mp->patch[q] : nullptr; }
extern "C" int32_t kl_mp_num_patches(const kl_mp* mp) { return mp ? (int32_t)mp->patch.size() : 0; }

extern "C" int kl_mp_set_active(kl_mp* mp, const int32_t* active) {
    if (!mp) return KL_E_ARG;
    KL_CUDA(cudaSetDevice(mp->device));
    for (size_t q = 0; q < mp->patch.size(); ++q) mp->active[q] = active ? (active[q] != 0) : 1;
    return mp_rebuild_loads(mp);
}

// interface columns (free DoFs shared by several patches): count, then the list; what a patch partition has to reduce
extern "C" int kl_mp_interface_dofs(const kl_mp* mp, int32_t* count, int32_t* dofs) {
    if (!mp) return KL_E_ARG;
    if (count) *count = (int32_t)mp->coupled_cols.size();
    if (dofs) std::memcpy(dofs, mp->coupled_cols.data(), sizeof(int) * mp->coupled_cols.size());
    return KL_OK;
}

// ---- device-resident assembly: what the matrix context's kl_jacobian_device / kl_residual_device / kl_check dispatch to -------------
// Run f(patch, stream) for every active patch on the patch's own stream, forked from and joined into the caller's stream s: the grids
// of a patch are a few waves long, so one after the other they leave the tail of every launch idle (8 patches of 576 x 72 elements:
// 6.16 ms in sequence).  KL_MP_STREAMS=0 keeps everything on s (A/B).
template <class F>
static int mp_for_patches(kl_mp* mp, cudaStream_t s, F&& f) {
    static const bool sequential = [] { const char* e = getenv("KL_MP_STREAMS"); return e && e[0] == '0'; }();
    int nact = 0;
    for (size_t q = 0; q < mp->patch.size(); ++q) nact += mp->active[q] ? 1 : 0;
    if (sequential || nact < 2) {
        for (size_t q = 0; q < mp->patch.size(); ++q)
            if (mp->active[q])
                if (int rc = f(mp->patch[q], s)) return rc;
        return KL_OK;
    }
    if (mp->pstream.empty()) {
        mp->pstream.resize(mp->patch.size(), nullptr);
        mp->pdone.resize(mp->patch.size(), nullptr);
        for (auto& st : mp->pstream) KL_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (auto& e : mp->pdone) KL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        KL_CUDA(cudaEventCreateWithFlags(&mp->fork, cudaEventDisableTiming));
    }
    KL_CUDA(cudaEventRecord(mp->fork, s));
    int rc = KL_OK;
    for (size_t q = 0; q < mp->patch.size(); ++q) {
        if (!mp->active[q]) continue;
        KL_CUDA(cudaStreamWaitEvent(mp->pstream[q], mp->fork, 0));
        if (!rc) rc = f(mp->patch[q], mp->pstream[q]);
        KL_CUDA(cudaEventRecord(mp->pdone[q], mp->pstream[q]));      // joined even after an error: s never runs ahead of a patch stream
        KL_CUDA(cudaStreamWaitEvent(s, mp->pdone[q], 0));
    }
    return rc;
}

int kl_mp_jacobian_device(kl_mp* mp, const double* x_dev, cudaStream_t s) {
    kl_ctx* g = mp->g;
    KL_CUDA(cudaMemsetAsync(g->d.values, 0, sizeof(double) * (size_t)g->nnz, s));
    return mp_for_patches(mp, s, [&](kl_ctx* c, cudaStream_t st) { return kl_jacobian_device(c, x_dev, st); });
}

int kl_mp_residual_device(kl_mp* mp, const double* x_dev, double lam_fext, double sign_fint, double* r_dev, cudaStream_t s) {
    kl_ctx* g = mp->g;
    KL_CUDA(cudaMemsetAsync(r_dev, 0, sizeof(double) * g->d.nfree, s));
    if (int rc = mp_for_patches(mp, s, [&](kl_ctx* c, cudaStream_t st) { return kl_residual_accumulate(c, x_dev, r_dev, st); })) return rc;
    return kl_launch_axpby(g, r_dev, g->d_fext, sign_fint, lam_fext, g->d.nfree, s);
}

int kl_mp_mass_device(kl_mp* mp, double density, double* values, double* lumped, cudaStream_t s) {
    for (size_t q = 0; q < mp->patch.size(); ++q)
        if (mp->active[q])
            if (int rc = kl_launch_mass(mp->patch[q], density * mp->patch[q]->d.mat.t, values, lumped, s)) return rc;
    return KL_OK;
}

int kl_mp_check(kl_mp* mp, cudaStream_t s) {
    int first = KL_OK;
    for (size_t q = 0; q < mp->patch.size(); ++q) {
        if (!mp->active[q]) continue;
        const int rc = kl_check(mp->patch[q], s);
        if (rc && !first) first = rc;
    }
    return first;
}

int kl_mp_launches(const kl_mp* mp) {
    int n = 0;
    for (const kl_ctx* c : mp->patch) n += c->launches;
    return n;
}
