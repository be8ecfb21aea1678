// K6: colour-aligned depth raycast -- RaycastingScene replacement.
//
// Reference call sites: scene = o3d.t.geometry.RaycastingScene(); scene.add_triangles(mesh.cpu());
// rays = scene.create_rays_pinhole(K, E, width_px, height_px); scene.cast_rays(rays)['t_hit']
// (processing/reconstruction/reconstruct_scene.py:197-198; utils/o3d_utils.py:324-342).
// Open3D delegates to Embree on the CPU; here the scene is an LBVH (Morton sort + Karras 2012
// hierarchy + bottom-up refit) over the marching-cubes triangles, traversed by one thread per ray
// with a short per-thread stack and a float32 Moeller-Trumbore closest-hit test.  The 64-byte node
// holds both children's boxes so that one 4 x float4 fetch decides both descents.
// The Morton sort uses cub::DeviceRadixSort (CUDA toolkit library; build-time plumbing, not the
// traversal hot loop).
#include <cub/device/device_radix_sort.cuh>

#include "mq3d_common.cuh"

struct __align__(16) BvhNode {
    float4 a;  // l.lo.xyz, l.hi.x
    float4 b;  // l.hi.yz, r.lo.xy
    float4 c;  // r.lo.z, r.hi.xyz
    int left, right;  // >= 0 internal node, < 0 leaf ~index (into the sorted triangle array)
    int pad0, pad1;
};

struct mq3d_scene {
    int device;
    int64_t n_tris;
    float4 *tri;     // [n][3] vertices of the Morton-sorted triangles (w unused)
    BvhNode *nodes;  // [n-1]
    int64_t alloc_tris;
};

extern "C" int mq3d_scene_create(int device, mq3d_scene **out) {
    MQ3D_REQUIRE(out != nullptr, "null output handle");
    int n_dev = 0;
    MQ3D_CUDA(cudaGetDeviceCount(&n_dev));
    MQ3D_REQUIRE(device >= 0 && device < n_dev, "CUDA device not available (no CPU fallback)");
    mq3d_scene *s = new mq3d_scene();
    memset(s, 0, sizeof(*s));
    s->device = device;
    *out = s;
    return MQ3D_OK;
}

extern "C" int mq3d_scene_destroy(mq3d_scene *s) {
    if (!s) return MQ3D_OK;
    cudaSetDevice(s->device);
    cudaFree(s->tri);
    cudaFree(s->nodes);
    delete s;
    return MQ3D_OK;
}

// ------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void k_scene_bounds(const float *__restrict__ v, const int32_t *__restrict__ t, int64_t n, unsigned *__restrict__ bounds) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (i < n) {
        for (int k = 0; k < 3; ++k) {
            const float *p = v + 3 * (int64_t)t[3 * i + k];
            for (int a = 0; a < 3; ++a) {
                lo[a] = fminf(lo[a], p[a]);
                hi[a] = fmaxf(hi[a], p[a]);
            }
        }
    }
    for (int a = 0; a < 3; ++a) {
        unsigned l = f2ord(lo[a]), h = f2ord(hi[a]);
        l = __reduce_min_sync(0xFFFFFFFFu, l);
        h = __reduce_max_sync(0xFFFFFFFFu, h);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&bounds[a], l);
            atomicMax(&bounds[3 + a], h);
        }
    }
}

__device__ __forceinline__ unsigned expand10(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_morton(const float *__restrict__ v, const int32_t *__restrict__ t, int64_t n, const unsigned *__restrict__ bounds,
                         unsigned long long *__restrict__ keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float c[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        const float *p = v + 3 * (int64_t)t[3 * i + k];
        for (int a = 0; a < 3; ++a) c[a] += p[a];
    }
    unsigned q[3];
    for (int a = 0; a < 3; ++a) {
        float lo = ord2f(bounds[a]), hi = ord2f(bounds[3 + a]);
        float ext = hi - lo;
        float f = ext > 0.0f ? (c[a] * (1.0f / 3.0f) - lo) / ext : 0.0f;
        q[a] = (unsigned)fminf(fmaxf(f * 1024.0f, 0.0f), 1023.0f);
    }
    unsigned m = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    keys[i] = ((unsigned long long)m << 32) | (unsigned long long)(unsigned)i;  // unique keys
}

__global__ void k_gather_tris(const float *__restrict__ v, const int32_t *__restrict__ t, const unsigned long long *__restrict__ keys,
                              int64_t n, float4 *__restrict__ tri, float *__restrict__ leaf_box) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t src = (int64_t)(unsigned)(keys[i] & 0xFFFFFFFFull);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = 0; k < 3; ++k) {
        const float *p = v + 3 * (int64_t)t[3 * src + k];
        tri[3 * i + k] = make_float4(p[0], p[1], p[2], 0.0f);
        for (int a = 0; a < 3; ++a) {
            lo[a] = fminf(lo[a], p[a]);
            hi[a] = fmaxf(hi[a], p[a]);
        }
    }
    for (int a = 0; a < 3; ++a) {
        leaf_box[6 * i + a] = lo[a];
        leaf_box[6 * i + 3 + a] = hi[a];
    }
}

__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);
}

// Karras 2012: one thread per internal node
__global__ void k_build_hierarchy(const unsigned long long *__restrict__ keys, int n, int *__restrict__ left, int *__restrict__ right,
                                  int *__restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? ~gamma : gamma;            // leaf encoded as ~index
    int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lc;
    right[i] = rc;
    // parent links: internal nodes [0,n-1), leaves stored at (n-1)+index
    parent[lc < 0 ? (n - 1) + ~lc : lc] = i;
    parent[rc < 0 ? (n - 1) + ~rc : rc] = i;
    if (i == 0) parent[0] = -1;
}

// bottom-up refit: second arrival at a node merges its children's boxes
__global__ void k_refit(int n, const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ parent,
                        const float *__restrict__ leaf_box, float *__restrict__ node_box, int *__restrict__ flags,
                        BvhNode *__restrict__ nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cur = parent[(n - 1) + i];
    while (cur >= 0) {
        if (atomicAdd(&flags[cur], 1) == 0) return;  // first arrival: sibling not ready
        __threadfence();
        int lc = left[cur], rc = right[cur];
        const float *lb = lc < 0 ? leaf_box + 6 * (int64_t)(~lc) : node_box + 6 * (int64_t)lc;
        const float *rb = rc < 0 ? leaf_box + 6 * (int64_t)(~rc) : node_box + 6 * (int64_t)rc;
        float l[6], r[6];
        for (int a = 0; a < 6; ++a) {
            l[a] = __ldcg(lb + a);
            r[a] = __ldcg(rb + a);
        }
        for (int a = 0; a < 3; ++a) {
            node_box[6 * (int64_t)cur + a] = fminf(l[a], r[a]);
            node_box[6 * (int64_t)cur + 3 + a] = fmaxf(l[3 + a], r[3 + a]);
        }
        BvhNode nd;
        nd.a = make_float4(l[0], l[1], l[2], l[3]);
        nd.b = make_float4(l[4], l[5], r[0], r[1]);
        nd.c = make_float4(r[2], r[3], r[4], r[5]);
        nd.left = lc;
        nd.right = rc;
        nd.pad0 = nd.pad1 = 0;
        nodes[cur] = nd;
        __threadfence();
        cur = parent[cur];
    }
}

extern "C" int mq3d_scene_add_triangles(mq3d_scene *s, const float *vertices_dev, int64_t n_vertices,
                                        const int32_t *triangles_dev, int64_t n_triangles, void *stream) {
    MQ3D_REQUIRE(s != nullptr, "null scene");
    MQ3D_REQUIRE(n_triangles >= 0 && n_vertices >= 0, "negative sizes");
    MQ3D_REQUIRE(n_triangles < 2147483647LL / 4, "too many triangles");
    MQ3D_REQUIRE(s->n_tris == 0, "only one add_triangles call per scene is supported (as the reference uses it)");
    MQ3D_TRY(mq3d_set_device(s->device));
    if (n_triangles == 0) return MQ3D_OK;
    MQ3D_REQUIRE(vertices_dev && triangles_dev, "null geometry");
    cudaStream_t st = as_stream(stream);
    const int n = (int)n_triangles;
    // one scratch arena for the build (a single allocation / free instead of ten)
    unsigned *bounds = nullptr;
    unsigned long long *keys = nullptr, *keys_sorted = nullptr;
    float *leaf_box = nullptr, *node_box = nullptr;
    int *left = nullptr, *right = nullptr, *parent = nullptr, *flags = nullptr;
    void *tmp = nullptr;
    char *arena = nullptr;
    size_t tmp_bytes = 0;
    int rc = [&]() -> int {
        MQ3D_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, keys_sorted, n, 0, 64, st));
        const size_t nn = (size_t)n, ni = (size_t)(n > 1 ? n - 1 : 1);
        const size_t sizes[10] = {sizeof(unsigned) * 6, sizeof(unsigned long long) * nn, sizeof(unsigned long long) * nn,
                                  sizeof(float) * 6 * nn, sizeof(float) * 6 * ni, sizeof(int) * nn, sizeof(int) * nn,
                                  sizeof(int) * 2 * nn, sizeof(int) * nn, tmp_bytes};
        size_t offs[10], total = 0;
        for (int i = 0; i < 10; ++i) {
            offs[i] = total;
            total += (sizes[i] + 255) & ~(size_t)255;
        }
        MQ3D_CUDA(cudaMalloc(&arena, total));
        bounds = reinterpret_cast<unsigned *>(arena + offs[0]);
        keys = reinterpret_cast<unsigned long long *>(arena + offs[1]);
        keys_sorted = reinterpret_cast<unsigned long long *>(arena + offs[2]);
        leaf_box = reinterpret_cast<float *>(arena + offs[3]);
        node_box = reinterpret_cast<float *>(arena + offs[4]);
        left = reinterpret_cast<int *>(arena + offs[5]);
        right = reinterpret_cast<int *>(arena + offs[6]);
        parent = reinterpret_cast<int *>(arena + offs[7]);
        flags = reinterpret_cast<int *>(arena + offs[8]);
        tmp = arena + offs[9];
        MQ3D_CUDA(cudaMalloc(&s->tri, sizeof(float4) * 3 * n));
        MQ3D_CUDA(cudaMalloc(&s->nodes, sizeof(BvhNode) * ni));
        unsigned init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
        MQ3D_CUDA(cudaMemcpyAsync(bounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
        MQ3D_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * n, st));
        MQ3D_CUDA(cudaMemsetAsync(parent, 0xFF, sizeof(int) * 2 * n, st));
        unsigned grid = (unsigned)((n + 255) / 256);
        k_scene_bounds<<<grid, 256, 0, st>>>(vertices_dev, triangles_dev, n, bounds);
        k_morton<<<grid, 256, 0, st>>>(vertices_dev, triangles_dev, n, bounds, keys);
        MQ3D_CUDA(cudaGetLastError());
        MQ3D_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, keys_sorted, n, 0, 64, st));
        k_gather_tris<<<grid, 256, 0, st>>>(vertices_dev, triangles_dev, keys_sorted, n, s->tri, leaf_box);
        if (n > 1) {
            k_build_hierarchy<<<grid, 256, 0, st>>>(keys_sorted, n, left, right, parent);
            k_refit<<<grid, 256, 0, st>>>(n, left, right, parent, leaf_box, node_box, flags, s->nodes);
        }
        MQ3D_CUDA(cudaGetLastError());
        MQ3D_CUDA(cudaStreamSynchronize(st));
        return MQ3D_OK;
    }();
    cudaFree(arena);
    if (rc != MQ3D_OK) return rc;
    s->n_tris = n;
    return MQ3D_OK;
}

// ------------------------------------------------------------------------------------------------
// rays
// ------------------------------------------------------------------------------------------------
struct PinholeParams {
    float c[3];
    float m[9];  // (R^T K^-1) cast to float32
};

__global__ void k_rays_pinhole(PinholeParams p, int W, int H, float *__restrict__ rays) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)W * H) return;
    int x = (int)(i % W), y = (int)(i / W);
    float px = __fadd_rn((float)x, 0.5f), py = __fadd_rn((float)y, 0.5f);
    float *r = rays + 6 * i;
    r[0] = p.c[0];
    r[1] = p.c[1];
    r[2] = p.c[2];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        r[3 + k] = __fadd_rn(__fadd_rn(__fmul_rn(p.m[3 * k], px), __fmul_rn(p.m[3 * k + 1], py)), __fmul_rn(p.m[3 * k + 2], 1.0f));
}

static int inv3(const double m[9], double o[9]) {
    double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (det == 0.0) return 1;
    double id = 1.0 / det;
    o[0] = (m[4] * m[8] - m[5] * m[7]) * id;
    o[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * id;
    o[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = (m[3] * m[7] - m[4] * m[6]) * id;
    o[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return 0;
}

extern "C" int mq3d_scene_create_rays_pinhole(const double K[9], const double E[16], int width, int height,
                                              float *rays_dev, void *stream) {
    MQ3D_REQUIRE(K && E && rays_dev, "null argument");
    MQ3D_REQUIRE(width > 0 && height > 0, "empty image");
    double invK[9];
    MQ3D_REQUIRE(inv3(K, invK) == 0, "singular intrinsic matrix");
    // C = -R^T t ; M = (R^T K^-1) in float64, cast to float32 (RaycastingScene::CreateRaysPinhole)
    PinholeParams p;
    double RT[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) RT[i * 3 + j] = E[j * 4 + i];
    for (int i = 0; i < 3; ++i) p.c[i] = (float)(-(RT[i * 3] * E[3] + RT[i * 3 + 1] * E[7] + RT[i * 3 + 2] * E[11]));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc += RT[i * 3 + k] * invK[k * 3 + j];
            p.m[i * 3 + j] = (float)acc;
        }
    int64_t n = (int64_t)width * height;
    k_rays_pinhole<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(p, width, height, rays_dev);
    MQ3D_CUDA(cudaGetLastError());
    return MQ3D_OK;
}

// ------------------------------------------------------------------------------------------------
// traversal
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool hit_box(const float lo[3], const float hi[3], const float o[3], const float inv[3], float tmax,
                                        float &tnear) {
    float t0 = 0.0f, t1 = tmax;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float ta = (lo[a] - o[a]) * inv[a], tb = (hi[a] - o[a]) * inv[a];
        t0 = fmaxf(t0, fminf(ta, tb));   // fminf/fmaxf drop NaNs (0 * inf on a slab face)
        t1 = fminf(t1, fmaxf(ta, tb));
    }
    tnear = t0;
    return t0 <= t1 * 1.0000004f;
}

__device__ __forceinline__ void hit_tri(const float4 *__restrict__ tri, int idx, const float o[3], const float d[3], float &best) {
    float4 v0 = __ldg(tri + 3 * (int64_t)idx), v1 = __ldg(tri + 3 * (int64_t)idx + 1), v2 = __ldg(tri + 3 * (int64_t)idx + 2);
    float e1x = v1.x - v0.x, e1y = v1.y - v0.y, e1z = v1.z - v0.z;
    float e2x = v2.x - v0.x, e2y = v2.y - v0.y, e2z = v2.z - v0.z;
    float px = d[1] * e2z - d[2] * e2y, py = d[2] * e2x - d[0] * e2z, pz = d[0] * e2y - d[1] * e2x;
    float det = e1x * px + e1y * py + e1z * pz;
    if (det == 0.0f) return;
    float inv = 1.0f / det;
    float tx = o[0] - v0.x, ty = o[1] - v0.y, tz = o[2] - v0.z;
    float u = (tx * px + ty * py + tz * pz) * inv;
    const float eps = 1e-6f;
    if (u < -eps || u > 1.0f + eps) return;
    float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
    float v = (d[0] * qx + d[1] * qy + d[2] * qz) * inv;
    if (v < -eps || u + v > 1.0f + eps) return;
    float t = (e2x * qx + e2y * qy + e2z * qz) * inv;
    if (t >= 0.0f && t < best) best = t;
}

__global__ void __launch_bounds__(128)
k_cast_rays(const BvhNode *__restrict__ nodes, const float4 *__restrict__ tri, int n_tris, const float *__restrict__ rays,
            int64_t n_rays, float *__restrict__ t_hit) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    float o[3], d[3], inv[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        o[a] = rays[6 * i + a];
        d[a] = rays[6 * i + 3 + a];
        inv[a] = 1.0f / d[a];
    }
    float best = INFINITY;
    if (n_tris == 1) {
        hit_tri(tri, 0, o, d, best);
    } else if (n_tris > 1) {
        int stack[64];
        int sp = 0;
        int cur = 0;
        for (;;) {
            const BvhNode *nd = nodes + cur;
            float4 A = __ldg(&nd->a), B = __ldg(&nd->b), C = __ldg(&nd->c);
            int lc = __ldg(&nd->left), rc = __ldg(&nd->right);
            float llo[3] = {A.x, A.y, A.z}, lhi[3] = {A.w, B.x, B.y};
            float rlo[3] = {B.z, B.w, C.x}, rhi[3] = {C.y, C.z, C.w};
            float tl, tr;
            bool hl = hit_box(llo, lhi, o, inv, best, tl);
            bool hr = hit_box(rlo, rhi, o, inv, best, tr);
            if (hl && lc < 0) { hit_tri(tri, ~lc, o, d, best); hl = false; }
            if (hr && rc < 0) { hit_tri(tri, ~rc, o, d, best); hr = false; }
            if (hl && hr) {
                // descend into the nearer child first
                int nearc = tl <= tr ? lc : rc, farc = tl <= tr ? rc : lc;
                if (sp < 64) stack[sp++] = farc;
                cur = nearc;
            } else if (hl) {
                cur = lc;
            } else if (hr) {
                cur = rc;
            } else {
                if (sp == 0) break;
                cur = stack[--sp];
            }
        }
    }
    t_hit[i] = best;
}

extern "C" int mq3d_scene_cast_rays(mq3d_scene *s, const float *rays_dev, int64_t n_rays, float *t_hit_dev,
                                    void *stream) {
    MQ3D_REQUIRE(s && (n_rays == 0 || (rays_dev && t_hit_dev)), "null argument");
    if (n_rays == 0) return MQ3D_OK;
    MQ3D_TRY(mq3d_set_device(s->device));
    k_cast_rays<<<(unsigned)((n_rays + 127) / 128), 128, 0, as_stream(stream)>>>(s->nodes, s->tri, (int)s->n_tris, rays_dev,
                                                                                  n_rays, t_hit_dev);
    MQ3D_CUDA(cudaGetLastError());
    return MQ3D_OK;
}
