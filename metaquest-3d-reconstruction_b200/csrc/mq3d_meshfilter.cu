// N1: filter_mesh_components on the device.
//
// Replaces the legacy-CPU round trip of processing/reconstruction/utils/o3d_utils.py:241-321, which the reference
// runs between marching cubes and the colour-view raycast (reconstruct_scene.py:115-118,192-195):
//   mesh.to_legacy() -> cluster_connected_triangles -> keep clusters with >= min_triangle_count triangles (or the
//   largest one) -> remove_triangles_by_mask + remove_unreferenced_vertices (only if something was removed) ->
//   remove_degenerate_triangles -> remove_duplicated_triangles -> remove_duplicated_vertices ->
//   remove_non_manifold_edges -> from_legacy.
// Same semantics, step for step, without the mesh ever leaving HBM:
//   * edge adjacency without sorting: every triangle inserts its three undirected edges (min, max vertex index) into
//     an open-addressing hash set whose value is the lowest triangle index seen so far for that edge (atomicMin);
//     the previous holder of the edge is united with the triangle in a lock-free union-find (roots hook towards the
//     smaller index, so a component's label is its lowest triangle index), followed by a flattening pass;
//   * component sizes by atomics on the roots; validity / largest-component fallback decided from one small
//     read-back;
//   * degenerate triangles: a repeated vertex index;
//   * duplicated triangles (equal up to rotation) and duplicated vertices (identical coordinates): hash sets that
//     keep the LOWEST index of every key, i.e. Open3D's "first occurrence wins";
//   * compaction of vertices / attributes / triangles by exclusive scans (order preserving, as the CPU code);
//   * non-manifold edges (more than two triangles on an edge) are counted on the device; a marching-cubes mesh has
//     none.  If some exist the caller finishes that one step with the host routine (the CPU loop removes smallest-
//     area triangles in unordered_map order, which no parallel formulation reproduces bit for bit anyway).
// Welding identical vertices is also what turns the concatenation of per-rank meshes of a multi-GPU run (vertices on
// ghost edges are emitted by every rank that references them) into the single-GPU mesh.
#include "mq3d_common.cuh"

#define MF_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define MF_EMPTY_ID 0x7FFFFFFF

// ------------------------------------------------------------------------------------------------
// generic exclusive scan of int flags (two kernels, deterministic, no inter-CTA waiting)
// ------------------------------------------------------------------------------------------------
#define MF_SCAN_THREADS 256
#define MF_SCAN_PER 8
#define MF_SCAN_CHUNK (MF_SCAN_THREADS * MF_SCAN_PER)

__global__ void __launch_bounds__(MF_SCAN_THREADS) k_mf_scan_totals(const int *__restrict__ flags, int64_t n, long long *__restrict__ totals) {
    __shared__ long long s[MF_SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long a = 0;
    for (int q = 0; q < MF_SCAN_PER; ++q) {
        const int64_t i = (int64_t)blockIdx.x * MF_SCAN_CHUNK + q * MF_SCAN_THREADS + tid;
        if (i < n) a += flags[i];
    }
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
    if (lane == 0) s[warp] = a;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int w = 0; w < MF_SCAN_THREADS / 32; ++w) t += s[w];
        totals[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(MF_SCAN_THREADS) k_mf_scan(const int *__restrict__ flags, int64_t n, const long long *__restrict__ totals,
                                                             int *__restrict__ excl, long long *__restrict__ total_out) {
    __shared__ long long s[MF_SCAN_THREADS / 32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        long long a = 0;
        for (int i = tid; i < (int)blockIdx.x; i += MF_SCAN_THREADS) a += totals[i];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        if (lane == 0) s[warp] = a;
        __syncthreads();
        if (tid == 0) {
            long long t = 0;
            for (int w = 0; w < MF_SCAN_THREADS / 32; ++w) t += s[w];
            s_carry = t;
        }
        __syncthreads();
    }
    const int64_t i0 = (int64_t)blockIdx.x * MF_SCAN_CHUNK + (int64_t)tid * MF_SCAN_PER;
    long long a = 0, la[MF_SCAN_PER];
#pragma unroll
    for (int q = 0; q < MF_SCAN_PER; ++q) {
        la[q] = a;
        if (i0 + q < n) a += flags[i0 + q];
    }
    long long ia = a;
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xFFFFFFFFu, ia, o);
        if (lane >= o) ia += t;
    }
    __syncthreads();
    if (lane == 31) s[warp] = ia;
    __syncthreads();
    long long wa = 0, ta = 0;
    for (int w = 0; w < MF_SCAN_THREADS / 32; ++w) {
        if (w < warp) wa += s[w];
        ta += s[w];
    }
    const long long p = s_carry + wa + ia - a;
#pragma unroll
    for (int q = 0; q < MF_SCAN_PER; ++q)
        if (i0 + q < n) excl[i0 + q] = (int)(p + la[q]);
    if (blockIdx.x == gridDim.x - 1 && tid == 0) *total_out = s_carry + ta;
}

static int mf_scan(const int *flags, int64_t n, int *excl, long long *totals, long long *total_out, cudaStream_t st) {
    if (n == 0) {
        MQ3D_CUDA(cudaMemsetAsync(total_out, 0, sizeof(long long), st));
        return MQ3D_OK;
    }
    const unsigned chunks = (unsigned)((n + MF_SCAN_CHUNK - 1) / MF_SCAN_CHUNK);
    k_mf_scan_totals<<<chunks, MF_SCAN_THREADS, 0, st>>>(flags, n, totals);
    k_mf_scan<<<chunks, MF_SCAN_THREADS, 0, st>>>(flags, n, totals, excl, total_out);
    MQ3D_CUDA(cudaGetLastError());
    return MQ3D_OK;
}

// ------------------------------------------------------------------------------------------------
// union-find over triangles
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mf_find(int *parent, int x) {
    int p = parent[x];
    while (p != x) {
        const int gp = parent[p];
        if (gp != p) parent[x] = gp;      // path halving (benign race: only ever points further up the same tree)
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void mf_unite(int *parent, int a, int b) {
    for (;;) {
        a = mf_find(parent, a);
        b = mf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }          // hook the larger root under the smaller one
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}

__global__ void k_mf_iota(int *p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}

__global__ void k_mf_fill64(unsigned long long *p, unsigned long long v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void k_mf_fill32(int *p, int v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// edge (min, max) -> lowest triangle index that carries it; every later arrival is united with the holder
__global__ void k_mf_edges(const int32_t *__restrict__ tris, int64_t T, unsigned long long *__restrict__ keys, int *__restrict__ vals,
                           uint32_t mask, int *__restrict__ parent) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int v[3] = {tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]};
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const unsigned a = (unsigned)v[e], b = (unsigned)v[(e + 1) % 3];
        const unsigned long long key = ((unsigned long long)(a < b ? a : b) << 32) | (unsigned long long)(a < b ? b : a);
        uint32_t slot = mq3d_hash64(key) & mask;
        for (;;) {
            unsigned long long cur = keys[slot];
            if (cur == MF_EMPTY_KEY) {
                cur = atomicCAS(&keys[slot], MF_EMPTY_KEY, key);
                if (cur == MF_EMPTY_KEY) cur = key;
            }
            if (cur == key) {
                const int old = atomicMin(&vals[slot], (int)t);
                if (old != MF_EMPTY_ID) mf_unite(parent, (int)t, old);
                break;
            }
            slot = (slot + 1) & mask;
        }
    }
}

__global__ void k_mf_flatten_count(int *__restrict__ parent, int64_t T, int *__restrict__ count) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int r = (int)t;
    while (parent[r] != r) r = parent[r];
    parent[t] = r;          // (another thread may still be walking through t: r is an ancestor of t, walks stay correct)
    atomicAdd(&count[r], 1);
}

// stats[0] components, [1] components with >= min triangles, [2] (max count << 32) | ~root of the first largest one
__global__ void k_mf_stats(const int *__restrict__ parent, const int *__restrict__ count, int64_t T, int min_count,
                           unsigned long long *__restrict__ stats) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || parent[t] != (int)t) return;
    atomicAdd(&stats[0], 1ull);
    if (count[t] >= min_count) atomicAdd(&stats[1], 1ull);
    atomicMax(&stats[2], ((unsigned long long)(unsigned)count[t] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)t));
}

// keep1: component large enough (or the chosen largest one); marks the vertices it references
__global__ void k_mf_keep_component(const int32_t *__restrict__ tris, const int *__restrict__ parent, const int *__restrict__ count,
                                    int64_t T, int min_count, int only_root, int *__restrict__ keep, int *__restrict__ vused) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int r = parent[t];
    const int k = only_root >= 0 ? (r == only_root) : (count[r] >= min_count);
    keep[t] = k;
    if (k && vused) {
        vused[tris[3 * t]] = 1;
        vused[tris[3 * t + 1]] = 1;
        vused[tris[3 * t + 2]] = 1;
    }
}

// keep2 = keep1 && no repeated vertex index
__global__ void k_mf_degenerate(const int32_t *__restrict__ tris, int64_t T, int *__restrict__ keep) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || !keep[t]) return;
    const int a = tris[3 * t], b = tris[3 * t + 1], c = tris[3 * t + 2];
    if (a == b || b == c || a == c) keep[t] = 0;
}

__device__ __forceinline__ void mf_canonical(const int32_t *__restrict__ tris, int64_t t, int &a, int &b, int &c) {
    const int v0 = tris[3 * t], v1 = tris[3 * t + 1], v2 = tris[3 * t + 2];
    if (v0 <= v1 && v0 <= v2) { a = v0; b = v1; c = v2; }          // rotate the smallest index to the front
    else if (v1 <= v0 && v1 <= v2) { a = v1; b = v2; c = v0; }
    else { a = v2; b = v0; c = v1; }
}

__device__ __forceinline__ uint32_t mf_hash3(unsigned a, unsigned b, unsigned c) {
    return mq3d_hash64(((unsigned long long)a << 32 | b) ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(c + 1u)));
}

// set of canonical triples; slot value = lowest triangle index with that triple
template <bool LOOKUP>
__global__ void k_mf_dup_tris(const int32_t *__restrict__ tris, int64_t T, int *__restrict__ slots, uint32_t mask, int *__restrict__ keep) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || !keep[t]) return;
    int a, b, c;
    mf_canonical(tris, t, a, b, c);
    uint32_t slot = mf_hash3((unsigned)a, (unsigned)b, (unsigned)c) & mask;
    for (;;) {
        int cur = slots[slot];
        if (!LOOKUP && cur == MF_EMPTY_ID) {
            cur = atomicCAS(&slots[slot], MF_EMPTY_ID, (int)t);
            if (cur == MF_EMPTY_ID) return;                        // first of its kind so far
        }
        int x, y, z;
        mf_canonical(tris, cur, x, y, z);
        if (x == a && y == b && z == c) {
            if (LOOKUP) {
                if (cur != (int)t) keep[t] = 0;                    // a lower-indexed twin exists
            } else {
                atomicMin(&slots[slot], (int)t);
            }
            return;
        }
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ void mf_coords(const float *__restrict__ v, int64_t i, unsigned &x, unsigned &y, unsigned &z) {
    // value equality as np.unique / Eigen ==: -0.0 and +0.0 are the same coordinate
    const float fx = v[3 * i] + 0.0f, fy = v[3 * i + 1] + 0.0f, fz = v[3 * i + 2] + 0.0f;
    x = __float_as_uint(fx);
    y = __float_as_uint(fy);
    z = __float_as_uint(fz);
}

// set of coordinates; slot value = lowest vertex index with those coordinates.  LOOKUP writes rep[i].
template <bool LOOKUP>
__global__ void k_mf_dup_verts(const float *__restrict__ verts, int64_t V, const int *__restrict__ vused, int *__restrict__ slots, uint32_t mask,
                               int *__restrict__ rep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V || !vused[i]) return;
    unsigned a, b, c;
    mf_coords(verts, i, a, b, c);
    uint32_t slot = mf_hash3(a, b, c) & mask;
    for (;;) {
        int cur = slots[slot];
        if (!LOOKUP && cur == MF_EMPTY_ID) {
            cur = atomicCAS(&slots[slot], MF_EMPTY_ID, (int)i);
            if (cur == MF_EMPTY_ID) return;
        }
        unsigned x, y, z;
        mf_coords(verts, cur, x, y, z);
        if (x == a && y == b && z == c) {
            if (LOOKUP) rep[i] = cur;
            else atomicMin(&slots[slot], (int)i);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

// vkeep[i] = used && own representative
__global__ void k_mf_vkeep(const int *__restrict__ vused, const int *__restrict__ rep, int64_t V, int *__restrict__ vkeep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V) vkeep[i] = vused[i] && rep[i] == (int)i;
}

__global__ void k_mf_gather_verts(const float *__restrict__ v, const float *__restrict__ n, const float *__restrict__ c, const int *__restrict__ vkeep,
                                  const int *__restrict__ vnew, int64_t V, float *__restrict__ ov, float *__restrict__ on, float *__restrict__ oc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V || !vkeep[i]) return;
    const int64_t j = vnew[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ov[3 * j + k] = v[3 * i + k];
        if (n) on[3 * j + k] = n[3 * i + k];
        if (c) oc[3 * j + k] = c[3 * i + k];
    }
}

__global__ void k_mf_gather_tris(const int32_t *__restrict__ tris, const int *__restrict__ keep, const int *__restrict__ tnew, const int *__restrict__ rep,
                                 const int *__restrict__ vnew, int64_t T, int32_t *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || !keep[t]) return;
    const int64_t j = tnew[t];
#pragma unroll
    for (int k = 0; k < 3; ++k) out[3 * j + k] = vnew[rep[tris[3 * t + k]]];
}

// triangles per undirected edge of the final mesh; counts the edges carried by more than two
__global__ void k_mf_edge_count(const int32_t *__restrict__ tris, int64_t T, unsigned long long *__restrict__ keys, int *__restrict__ vals,
                                uint32_t mask, unsigned long long *__restrict__ n_bad) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int v[3] = {tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]};
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const unsigned a = (unsigned)v[e], b = (unsigned)v[(e + 1) % 3];
        const unsigned long long key = ((unsigned long long)(a < b ? a : b) << 32) | (unsigned long long)(a < b ? b : a);
        uint32_t slot = mq3d_hash64(key) & mask;
        for (;;) {
            unsigned long long cur = keys[slot];
            if (cur == MF_EMPTY_KEY) {
                cur = atomicCAS(&keys[slot], MF_EMPTY_KEY, key);
                if (cur == MF_EMPTY_KEY) cur = key;
            }
            if (cur == key) {
                if (atomicAdd(&vals[slot], 1) == 2) atomicAdd(n_bad, 1ull);      // third triangle on this edge
                break;
            }
            slot = (slot + 1) & mask;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct MfScratch {          // frees everything it owns on scope exit
    void *p[16];
    int n;
    MfScratch() : n(0) {}
    ~MfScratch() { for (int i = 0; i < n; ++i) cudaFree(p[i]); }
    template <class Tp> cudaError_t alloc(Tp **out, size_t count) {
        *out = nullptr;
        cudaError_t e = cudaMalloc((void **)out, sizeof(Tp) * (count ? count : 1));
        if (e == cudaSuccess) p[n++] = *out;
        return e;
    }
};

static uint32_t mf_table_size(int64_t entries) {
    int64_t s = 1024;
    while (s < 2 * entries) s <<= 1;
    return (uint32_t)s;
}

extern "C" int mq3d_mesh_filter(const float *vertices_dev, const float *normals_dev, const float *colors_dev, int64_t n_vertices,
                                const int32_t *triangles_dev, int64_t n_triangles, int64_t min_triangle_count,
                                float *out_vertices_dev, float *out_normals_dev, float *out_colors_dev,
                                int32_t *out_triangles_dev, int64_t *out_n_vertices, int64_t *out_n_triangles,
                                mq3d_mesh_filter_info *info, int device, void *stream) {
    MQ3D_REQUIRE(vertices_dev && triangles_dev && out_vertices_dev && out_triangles_dev && out_n_vertices && out_n_triangles,
                 "null argument");
    MQ3D_REQUIRE(n_vertices > 0 && n_triangles > 0, "empty mesh (the caller returns it unchanged)");
    MQ3D_REQUIRE(n_vertices < 0x7FFFFFFF && n_triangles < 0x7FFFFFFF, "mesh too large for int32 indices");
    MQ3D_REQUIRE((normals_dev == nullptr) == (out_normals_dev == nullptr) && (colors_dev == nullptr) == (out_colors_dev == nullptr),
                 "attribute inputs and outputs must be given together");
    MQ3D_TRY(mq3d_set_device(device));
    cudaStream_t st = as_stream(stream);
    const int64_t V = n_vertices, T = n_triangles;
    const int min_count = (int)(min_triangle_count > 0x7FFFFFFF ? 0x7FFFFFFF : (min_triangle_count < 0 ? 0 : min_triangle_count));
    mq3d_mesh_filter_info inf;
    memset(&inf, 0, sizeof(inf));
    MfScratch m;
    const uint32_t esize = mf_table_size(3 * T), tsize = mf_table_size(T), vsize = mf_table_size(V);
    unsigned long long *ekeys, *stats;
    int *evals, *parent, *count, *keep, *tnew, *vused, *rep, *vkeep, *vnew, *slots;
    long long *totals, *total_dev;
    long long *h_total = nullptr;
    MQ3D_CUDA(m.alloc(&ekeys, esize));
    MQ3D_CUDA(m.alloc(&evals, esize));
    MQ3D_CUDA(m.alloc(&parent, (size_t)T));
    MQ3D_CUDA(m.alloc(&count, (size_t)T));
    MQ3D_CUDA(m.alloc(&keep, (size_t)T));
    MQ3D_CUDA(m.alloc(&tnew, (size_t)T));
    MQ3D_CUDA(m.alloc(&vused, (size_t)V));
    MQ3D_CUDA(m.alloc(&rep, (size_t)V));
    MQ3D_CUDA(m.alloc(&vkeep, (size_t)V));
    MQ3D_CUDA(m.alloc(&vnew, (size_t)V));
    MQ3D_CUDA(m.alloc(&slots, (size_t)(tsize > vsize ? tsize : vsize)));
    MQ3D_CUDA(m.alloc(&totals, (size_t)((T > V ? T : V) / MF_SCAN_CHUNK + 2)));
    MQ3D_CUDA(m.alloc(&stats, 8));
    MQ3D_CUDA(m.alloc(&total_dev, 2));
    struct Pinned {
        long long *p;
        ~Pinned() { if (p) cudaFreeHost(p); }
    } pin = {nullptr};
    MQ3D_CUDA(cudaMallocHost(&pin.p, sizeof(long long) * 8));
    h_total = pin.p;
    const unsigned gT = (unsigned)((T + 255) / 256), gV = (unsigned)((V + 255) / 256);

    // ---- connected components of the triangle-edge graph ----
    k_mf_fill64<<<148 * 8, 256, 0, st>>>(ekeys, MF_EMPTY_KEY, esize);
    k_mf_fill32<<<148 * 8, 256, 0, st>>>(evals, MF_EMPTY_ID, esize);
    k_mf_iota<<<gT, 256, 0, st>>>(parent, T);
    MQ3D_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * T, st));
    MQ3D_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 8, st));
    k_mf_edges<<<gT, 256, 0, st>>>(triangles_dev, T, ekeys, evals, esize - 1, parent);
    k_mf_flatten_count<<<gT, 256, 0, st>>>(parent, T, count);
    k_mf_stats<<<gT, 256, 0, st>>>(parent, count, T, min_count, stats);
    MQ3D_CUDA(cudaGetLastError());
    MQ3D_CUDA(cudaMemcpyAsync(h_total, stats, sizeof(unsigned long long) * 3, cudaMemcpyDeviceToHost, st));
    MQ3D_CUDA(cudaStreamSynchronize(st));
    inf.components = h_total[0];
    inf.components_kept = h_total[1];
    inf.largest_component = (int64_t)((unsigned long long)h_total[2] >> 32);
    int only_root = -1;
    if (inf.components_kept == 0) {           // no component is large enough: keep the (first) largest one
        inf.fallback_largest = 1;
        inf.components_kept = 1;
        only_root = (int)(0xFFFFFFFFu - (unsigned)((unsigned long long)h_total[2] & 0xFFFFFFFFull));
    }

    // ---- triangle and vertex selection ----
    MQ3D_CUDA(cudaMemsetAsync(vused, 0, sizeof(int) * V, st));
    k_mf_keep_component<<<gT, 256, 0, st>>>(triangles_dev, parent, count, T, min_count, only_root, keep, vused);
    MQ3D_TRY(mf_scan(keep, T, tnew, totals, total_dev, st));
    MQ3D_CUDA(cudaMemcpyAsync(h_total, total_dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
    MQ3D_CUDA(cudaStreamSynchronize(st));
    inf.input_triangles = T;
    inf.removed_triangles = T - h_total[0];
    if (inf.removed_triangles == 0)           // remove_unreferenced_vertices only runs when triangles were removed
        k_mf_fill32<<<148 * 8, 256, 0, st>>>(vused, 1, V);
    k_mf_degenerate<<<gT, 256, 0, st>>>(triangles_dev, T, keep);
    k_mf_fill32<<<148 * 8, 256, 0, st>>>(slots, MF_EMPTY_ID, tsize);
    k_mf_dup_tris<false><<<gT, 256, 0, st>>>(triangles_dev, T, slots, tsize - 1, keep);
    k_mf_dup_tris<true><<<gT, 256, 0, st>>>(triangles_dev, T, slots, tsize - 1, keep);
    k_mf_fill32<<<148 * 8, 256, 0, st>>>(slots, MF_EMPTY_ID, vsize);
    k_mf_iota<<<gV, 256, 0, st>>>(rep, V);
    k_mf_dup_verts<false><<<gV, 256, 0, st>>>(vertices_dev, V, vused, slots, vsize - 1, rep);
    k_mf_dup_verts<true><<<gV, 256, 0, st>>>(vertices_dev, V, vused, slots, vsize - 1, rep);
    k_mf_vkeep<<<gV, 256, 0, st>>>(vused, rep, V, vkeep);
    MQ3D_CUDA(cudaGetLastError());
    MQ3D_TRY(mf_scan(vkeep, V, vnew, totals, total_dev, st));
    MQ3D_TRY(mf_scan(keep, T, tnew, totals, total_dev + 1, st));
    MQ3D_CUDA(cudaMemcpyAsync(h_total, total_dev, sizeof(long long) * 2, cudaMemcpyDeviceToHost, st));

    // ---- compaction ----
    k_mf_gather_verts<<<gV, 256, 0, st>>>(vertices_dev, normals_dev, colors_dev, vkeep, vnew, V, out_vertices_dev, out_normals_dev,
                                          out_colors_dev);
    k_mf_gather_tris<<<gT, 256, 0, st>>>(triangles_dev, keep, tnew, rep, vnew, T, out_triangles_dev);
    MQ3D_CUDA(cudaGetLastError());
    MQ3D_CUDA(cudaStreamSynchronize(st));
    const int64_t Vo = h_total[0], To = h_total[1];

    // ---- non-manifold edges of the result (counted; removal is the caller's host step if there are any) ----
    if (To > 0) {
        k_mf_fill64<<<148 * 8, 256, 0, st>>>(ekeys, MF_EMPTY_KEY, esize);
        MQ3D_CUDA(cudaMemsetAsync(evals, 0, sizeof(int) * (size_t)esize, st));
        MQ3D_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned long long), st));
        k_mf_edge_count<<<(unsigned)((To + 255) / 256), 256, 0, st>>>(out_triangles_dev, To, ekeys, evals, esize - 1, stats);
        MQ3D_CUDA(cudaGetLastError());
        MQ3D_CUDA(cudaMemcpyAsync(h_total, stats, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        MQ3D_CUDA(cudaStreamSynchronize(st));
        inf.non_manifold_edges = h_total[0];
    }
    *out_n_vertices = Vo;
    *out_n_triangles = To;
    if (info) *info = inf;
    return MQ3D_OK;
}
