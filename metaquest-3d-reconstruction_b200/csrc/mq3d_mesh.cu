// K5: marching cubes / point-cloud extraction over the hashed voxel-block grid.
//
// Semantics: Open3D 0.19 ExtractTriangleMesh / ExtractPointCloud as reached from
// vbg.extract_triangle_mesh(weight_threshold, -1) and vbg.extract_point_cloud()
// (reference: processing/reconstruction/reconstruct_scene.py:90,105-108,186-189).
//
// GPU formulation (no global 16 B/voxel "mesh_structure" volume, no global atomics):
//   classify (one CTA per block): stage validity/sign bits of the (-1..16)^3 neighbourhood in shared
//     memory, derive the cube case of every cube in (-1..15)^3, mark the lattice edges the block owns
//     with warp ballots (3 bit-planes x 128 words), popcount them, and keep per-word prefix counts so
//     that  id(voxel, axis) = vertex_offset[block] + prefix[word] + popc(mask[word] & lower_lanes)
//     is computable by any neighbour without a lookup table;
//   scan: exclusive prefix of per-block vertex / triangle counts;
//   emit (one CTA per non-empty block): stage the (-1..17)^3 tsdf neighbourhood, write vertices and
//     normals at their scanned positions and triangles with ids resolved through the bit-planes.
// Output order is deterministic.  A cube counts on this rank only if its block is owned (multi-GPU
// partition); single-GPU grids own every block.
#include "mc_tables.h"
#include "mq3d_common.cuh"

#define CODE_R 18   // validity/sign region: coords -1..16
#define CUBE_R 17   // cube region: coords -1..15
#define TS_R 19     // tsdf region: coords -1..17
#define EWORDS 384  // 3 planes x 128 words

__device__ __forceinline__ int nb_of(int r) { return r < 0 ? 0 : (r > 15 ? 2 : 1); }  // -> d+1

// ------------------------------------------------------------------------------------------------
__global__ void k_mc_neighbors(HashView h, const int32_t *__restrict__ block_keys, int64_t n, int32_t *__restrict__ nb) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 27) return;
    int64_t b = i / 27;
    int k = (int)(i % 27);
    int dx = k % 3 - 1, dy = (k / 3) % 3 - 1, dz = k / 9 - 1;
    int x = block_keys[3 * b] + dx, y = block_keys[3 * b + 1] + dy, z = block_keys[3 * b + 2] + dz;
    int32_t r = -1;
    if (k == 13) {
        r = (int32_t)b;
    } else if (mq3d_key_in_range(x, y, z)) {
        uint32_t s = hash_find(h, mq3d_pack_key(x, y, z));
        if (s != 0xFFFFFFFFu) r = h.vals[s];
    }
    nb[i] = r;
}

// exclusive scan of 128 or 384 small counts held in shared memory, by warp 0 (n % 32 == 0)
__device__ __forceinline__ int warp0_exclusive_scan(int *s, int n, int lane) {
    const int per = n / 32;
    int sum = 0;
    for (int i = 0; i < per; ++i) sum += s[lane * per + i];
    int incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    int run = incl - sum;
    for (int i = 0; i < per; ++i) {
        int c = s[lane * per + i];
        s[lane * per + i] = run;
        run += c;
    }
    return __shfl_sync(0xFFFFFFFFu, incl, 31);  // total
}

// ------------------------------------------------------------------------------------------------
// classify
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mc_classify(const float *__restrict__ tsdf, const float *__restrict__ weight, const int32_t *__restrict__ block_keys,
              const int32_t *__restrict__ nb, float weight_thr, Partition part, uint32_t *__restrict__ emask,
              uint16_t *__restrict__ eprefix, uint8_t *__restrict__ cubes, int32_t *__restrict__ counts) {
    __shared__ uint8_t s_code[CODE_R * CODE_R * CODE_R];
    __shared__ uint8_t s_cube[CUBE_R * CUBE_R * CUBE_R];
    __shared__ int s_nb[27];
    __shared__ uint8_t s_owned[27];
    __shared__ uint8_t s_tric[256];
    __shared__ int s_cnt[EWORDS];
    __shared__ int s_tri[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x;
    if (tid < 27) {
        s_nb[tid] = nb[b * 27 + tid];
        int dx = tid % 3 - 1, dy = (tid / 3) % 3 - 1, dz = tid / 9 - 1;
        s_owned[tid] = mq3d_block_owned(block_keys[3 * b] + dx, block_keys[3 * b + 1] + dy, block_keys[3 * b + 2] + dz, part);
    }
    s_tric[tid] = MC_TRI_COUNT[tid];
    __syncthreads();
    // stage validity + sign bits of the (-1..16)^3 neighbourhood
    for (int r = tid; r < CODE_R * CODE_R * CODE_R; r += 256) {
        int rx = r % CODE_R - 1, ry = (r / CODE_R) % CODE_R - 1, rz = r / (CODE_R * CODE_R) - 1;
        int k = nb_of(rx) + 3 * nb_of(ry) + 9 * nb_of(rz);
        int bi = s_nb[k];
        uint8_t code = 0;
        if (bi >= 0) {
            int64_t li = (int64_t)bi * MQ3D_RES3 + (((rz & 15) * 16 + (ry & 15)) * 16 + (rx & 15));
            float t = __ldg(tsdf + li), w = __ldg(weight + li);
            code = (w > weight_thr ? 1 : 0) | (t < 0.0f ? 2 : 0);   // reject is `w <= thr`
        }
        s_code[r] = code;
    }
    __syncthreads();
    // cube cases for cubes in (-1..15)^3
    for (int c = tid; c < CUBE_R * CUBE_R * CUBE_R; c += 256) {
        int cx = c % CUBE_R, cy = (c / CUBE_R) % CUBE_R, cz = c / (CUBE_R * CUBE_R);  // region coords (cube -1 -> 0)
        int valid = 1, table = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int code = s_code[((cz + MC_VTX_SHIFTS[i][2]) * CODE_R + cy + MC_VTX_SHIFTS[i][1]) * CODE_R + cx + MC_VTX_SHIFTS[i][0]];
            valid &= code;
            table |= ((code >> 1) & 1) << i;
        }
        int k = nb_of(cx - 1) + 3 * nb_of(cy - 1) + 9 * nb_of(cz - 1);
        s_cube[c] = (valid & 1) && s_owned[k] ? (uint8_t)table : (uint8_t)0;
    }
    __syncthreads();
    // edge marks of the voxels this block owns + own-cube triangle counts
    int tri_local = 0;
    for (int it = 0; it < 16; ++it) {
        int v = it * 256 + tid;
        int x = v & 15, y = (v >> 4) & 15, z = v >> 8;
        int so = (s_code[((z + 1) * CODE_R + y + 1) * CODE_R + x + 1] >> 1) & 1;
        int sx = (s_code[((z + 1) * CODE_R + y + 1) * CODE_R + x + 2] >> 1) & 1;
        int sy = (s_code[((z + 1) * CODE_R + y + 2) * CODE_R + x + 1] >> 1) & 1;
        int sz = (s_code[((z + 2) * CODE_R + y + 1) * CODE_R + x + 1] >> 1) & 1;
        // cubes sharing each edge (cube region index: coord + 1)
#define CUBE(ix, iy, iz) s_cube[(((iz) + 1) * CUBE_R + (iy) + 1) * CUBE_R + (ix) + 1]
        int c000 = CUBE(x, y, z);
        int mx = (so != sx) && (c000 | CUBE(x, y - 1, z) | CUBE(x, y, z - 1) | CUBE(x, y - 1, z - 1));
        int my = (so != sy) && (c000 | CUBE(x - 1, y, z) | CUBE(x, y, z - 1) | CUBE(x - 1, y, z - 1));
        int mz = (so != sz) && (c000 | CUBE(x - 1, y, z) | CUBE(x, y - 1, z) | CUBE(x - 1, y - 1, z));
#undef CUBE
        unsigned bx_ = __ballot_sync(0xFFFFFFFFu, mx), by_ = __ballot_sync(0xFFFFFFFFu, my),
                 bz_ = __ballot_sync(0xFFFFFFFFu, mz);
        int word = it * 8 + warp;
        if (lane == 0) {
            emask[b * EWORDS + word] = bx_;
            emask[b * EWORDS + 128 + word] = by_;
            emask[b * EWORDS + 256 + word] = bz_;
            s_cnt[word] = __popc(bx_);
            s_cnt[128 + word] = __popc(by_);
            s_cnt[256 + word] = __popc(bz_);
        }
        cubes[b * MQ3D_RES3 + v] = (uint8_t)c000;
        tri_local += s_tric[c000];
    }
    for (int o = 16; o > 0; o >>= 1) tri_local += __shfl_xor_sync(0xFFFFFFFFu, tri_local, o);
    if (lane == 0) s_tri[warp] = tri_local;
    __syncthreads();
    if (warp == 0) {
        int total = warp0_exclusive_scan(s_cnt, EWORDS, lane);
        if (lane == 0) {
            int t = 0;
            for (int i = 0; i < 8; ++i) t += s_tri[i];
            counts[2 * b] = total;
            counts[2 * b + 1] = t;
        }
    }
    __syncthreads();
    for (int i = tid; i < EWORDS; i += 256) eprefix[b * EWORDS + i] = (uint16_t)s_cnt[i];
}

// ------------------------------------------------------------------------------------------------
// scan of per-block (a, b) counts -> int64 exclusive offsets [n+1][2]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan_counts(const int32_t *__restrict__ counts, int64_t n, int64_t *__restrict__ offsets) {
    __shared__ long long s_a[32], s_b[32];
    __shared__ long long s_carry[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry[0] = s_carry[1] = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        int64_t i = base + tid;
        long long a = i < n ? counts[2 * i] : 0, c = i < n ? counts[2 * i + 1] : 0;
        long long ia = a, ic = c;
        for (int o = 1; o < 32; o <<= 1) {
            long long ta = __shfl_up_sync(0xFFFFFFFFu, ia, o), tc = __shfl_up_sync(0xFFFFFFFFu, ic, o);
            if (lane >= o) { ia += ta; ic += tc; }
        }
        if (lane == 31) { s_a[warp] = ia; s_b[warp] = ic; }
        __syncthreads();
        if (warp == 0) {
            long long wa = s_a[lane], wb = s_b[lane];
            long long xa = wa, xb = wb;
            for (int o = 1; o < 32; o <<= 1) {
                long long ta = __shfl_up_sync(0xFFFFFFFFu, xa, o), tb = __shfl_up_sync(0xFFFFFFFFu, xb, o);
                if (lane >= o) { xa += ta; xb += tb; }
            }
            s_a[lane] = xa - wa;
            s_b[lane] = xb - wb;
        }
        __syncthreads();
        long long ca = s_carry[0], cb = s_carry[1];
        if (i < n) {
            offsets[2 * i] = ca + s_a[warp] + ia - a;
            offsets[2 * i + 1] = cb + s_b[warp] + ic - c;
        }
        __syncthreads();
        if (tid == 1023) { s_carry[0] = ca + s_a[warp] + ia; s_carry[1] = cb + s_b[warp] + ic; }
        __syncthreads();
    }
    if (tid == 0) { offsets[2 * n] = s_carry[0]; offsets[2 * n + 1] = s_carry[1]; }
}

// ------------------------------------------------------------------------------------------------
// shared staging of the (-1..17)^3 tsdf neighbourhood
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_tsdf(const float *__restrict__ tsdf, const int *s_nb, float *s_t, int tid) {
    for (int r = tid; r < TS_R * TS_R * TS_R; r += 256) {
        int rx = r % TS_R - 1, ry = (r / TS_R) % TS_R - 1, rz = r / (TS_R * TS_R) - 1;
        int bi = s_nb[nb_of(rx) + 3 * nb_of(ry) + 9 * nb_of(rz)];
        float t = 0.0f;
        if (bi >= 0) t = __ldg(tsdf + (int64_t)bi * MQ3D_RES3 + (((rz & 15) * 16 + (ry & 15)) * 16 + (rx & 15)));
        s_t[r] = t;
    }
}
#define TS(ix, iy, iz) s_t[(((iz) + 1) * TS_R + (iy) + 1) * TS_R + (ix) + 1]
#define EXISTS(ix, iy, iz) (s_nb[nb_of(ix) + 3 * nb_of(iy) + 9 * nb_of(iz)] >= 0)

// DeviceGetNormal: central differences where both neighbours exist; other components untouched
__device__ __forceinline__ void get_normal(const float *s_t, const int *s_nb, int x, int y, int z, float *n) {
    if (EXISTS(x + 1, y, z) && EXISTS(x - 1, y, z)) n[0] = __fsub_rn(TS(x + 1, y, z), TS(x - 1, y, z));
    if (EXISTS(x, y + 1, z) && EXISTS(x, y - 1, z)) n[1] = __fsub_rn(TS(x, y + 1, z), TS(x, y - 1, z));
    if (EXISTS(x, y, z + 1) && EXISTS(x, y, z - 1)) n[2] = __fsub_rn(TS(x, y, z + 1), TS(x, y, z - 1));
}

__device__ __forceinline__ void write_vertex(float *__restrict__ verts, float *__restrict__ normals, int32_t *__restrict__ vkeys,
                                             int64_t id, float vs, int gx, int gy, int gz, int e, float ratio,
                                             const float *no, const float *ne) {
    float rx = __fmul_rn(ratio, e == 0 ? 1.0f : 0.0f), ry = __fmul_rn(ratio, e == 1 ? 1.0f : 0.0f),
          rz = __fmul_rn(ratio, e == 2 ? 1.0f : 0.0f);
    verts[3 * id + 0] = __fmul_rn(vs, __fadd_rn((float)gx, rx));
    verts[3 * id + 1] = __fmul_rn(vs, __fadd_rn((float)gy, ry));
    verts[3 * id + 2] = __fmul_rn(vs, __fadd_rn((float)gz, rz));
    if (vkeys) {
        vkeys[4 * id] = gx; vkeys[4 * id + 1] = gy; vkeys[4 * id + 2] = gz; vkeys[4 * id + 3] = e;
    }
    if (normals) {
        float om = __fsub_rn(1.0f, ratio);
        float nx = __fadd_rn(__fmul_rn(om, no[0]), __fmul_rn(ratio, ne[0]));
        float ny = __fadd_rn(__fmul_rn(om, no[1]), __fmul_rn(ratio, ne[1]));
        float nz = __fadd_rn(__fmul_rn(om, no[2]), __fmul_rn(ratio, ne[2]));
        float s = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
        float norm = (float)((double)__fsqrt_rn(s) + 1e-5);
        normals[3 * id + 0] = __fdiv_rn(nx, norm);
        normals[3 * id + 1] = __fdiv_rn(ny, norm);
        normals[3 * id + 2] = __fdiv_rn(nz, norm);
    }
}

// ------------------------------------------------------------------------------------------------
// emit
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mc_emit(const float *__restrict__ tsdf, const int32_t *__restrict__ block_keys, const int32_t *__restrict__ nb,
          const uint32_t *__restrict__ emask, const uint16_t *__restrict__ eprefix, const uint8_t *__restrict__ cubes,
          const int32_t *__restrict__ counts, const int64_t *__restrict__ offsets, float vs,
          float *__restrict__ verts, float *__restrict__ normals, int32_t *__restrict__ tris, int32_t *__restrict__ vkeys) {
    __shared__ float s_t[TS_R * TS_R * TS_R];
    __shared__ int s_nb[27];
    __shared__ uint32_t s_mask[EWORDS];
    __shared__ uint16_t s_pref[EWORDS];
    __shared__ int s_tsum[128];
    __shared__ signed char s_tt[256 * 16];
    __shared__ uint8_t s_tric[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x;
    const int nv = counts[2 * b], nt = counts[2 * b + 1];
    if (nv == 0 && nt == 0) return;
    if (tid < 27) s_nb[tid] = nb[b * 27 + tid];
    for (int i = tid; i < EWORDS; i += 256) {
        s_mask[i] = emask[b * EWORDS + i];
        s_pref[i] = eprefix[b * EWORDS + i];
    }
    for (int i = tid; i < 256 * 16; i += 256) s_tt[i] = MC_TRI_TABLE[i >> 4][i & 15];
    s_tric[tid] = MC_TRI_COUNT[tid];
    __syncthreads();
    const int kx = block_keys[3 * b], ky = block_keys[3 * b + 1], kz = block_keys[3 * b + 2];
    const int64_t voff = offsets[2 * b], toff = offsets[2 * b + 1];
    if (nv > 0) {
        stage_tsdf(tsdf, s_nb, s_t, tid);
        __syncthreads();
        for (int it = 0; it < 16; ++it) {
            int v = it * 256 + tid;
            int word = v >> 5;
            unsigned low = (1u << (v & 31)) - 1u, bit = 1u << (v & 31);
            unsigned m0 = s_mask[word], m1 = s_mask[128 + word], m2 = s_mask[256 + word];
            if (!((m0 | m1 | m2) & bit)) continue;
            int x = v & 15, y = (v >> 4) & 15, z = v >> 8;
            float to = TS(x, y, z);
            float no[3] = {0.0f, 0.0f, 0.0f}, ne[3] = {0.0f, 0.0f, 0.0f};
            get_normal(s_t, s_nb, x, y, z, no);
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                unsigned m = e == 0 ? m0 : (e == 1 ? m1 : m2);
                if (!(m & bit)) continue;
                int ex = x + (e == 0), ey = y + (e == 1), ez = z + (e == 2);
                float te = TS(ex, ey, ez);
                float ratio = __fdiv_rn(__fsub_rn(0.0f, to), __fsub_rn(te, to));
                int64_t id = voff + s_pref[e * 128 + word] + __popc(m & low);
                get_normal(s_t, s_nb, ex, ey, ez, ne);   // stale components persist across e (Open3D)
                write_vertex(verts, normals, vkeys, id, vs, kx * 16 + x, ky * 16 + y, kz * 16 + z, e, ratio, no, ne);
            }
        }
    }
    if (nt > 0 && tris) {
        // exclusive scan of per-cube triangle counts in voxel order
        int cube[16];
        for (int it = 0; it < 16; ++it) {
            cube[it] = cubes[b * MQ3D_RES3 + it * 256 + tid];
            int n = s_tric[cube[it]];
            for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
            if (lane == 0) s_tsum[it * 8 + warp] = n;
        }
        __syncthreads();
        if (warp == 0) warp0_exclusive_scan(s_tsum, 128, lane);
        __syncthreads();
        for (int it = 0; it < 16; ++it) {
            int c = cube[it];
            int n = s_tric[c];
            int incl = n;
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            if (n == 0) continue;
            int64_t tbase = toff + s_tsum[it * 8 + warp] + incl - n;
            int v = it * 256 + tid;
            int x = v & 15, y = (v >> 4) & 15, z = v >> 8;
            for (int k = 0; k < n; ++k) {
#pragma unroll
                for (int vtx = 0; vtx < 3; ++vtx) {
                    int edge = s_tt[c * 16 + 3 * k + vtx];
                    int ox = x + MC_EDGE_SHIFTS[edge][0], oy = y + MC_EDGE_SHIFTS[edge][1], oz = z + MC_EDGE_SHIFTS[edge][2];
                    int ax = MC_EDGE_SHIFTS[edge][3];
                    int nbk = nb_of(ox) + 3 * nb_of(oy) + 9 * nb_of(oz);
                    int lv = ((oz & 15) * 16 + (oy & 15)) * 16 + (ox & 15);
                    int w = ax * 128 + (lv >> 5);
                    unsigned low = (1u << (lv & 31)) - 1u;
                    int64_t id;
                    if (nbk == 13) {
                        id = voff + s_pref[w] + __popc(s_mask[w] & low);
                    } else {
                        int64_t bn = s_nb[nbk];
                        id = offsets[2 * bn] + eprefix[bn * EWORDS + w] + __popc(emask[bn * EWORDS + w] & low);
                    }
                    tris[3 * (tbase + k) + (2 - vtx)] = (int32_t)id;   // winding reversed
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// point cloud: count + emit (weights read directly; tsdf neighbourhood staged for normals)
// ------------------------------------------------------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(256)
k_points(const float *__restrict__ tsdf, const float *__restrict__ weight, const int32_t *__restrict__ block_keys,
         const int32_t *__restrict__ nb, float weight_thr, Partition part, int32_t *__restrict__ counts,
         const int64_t *__restrict__ offsets, float vs, float *__restrict__ points, float *__restrict__ normals,
         int32_t *__restrict__ pkeys) {
    __shared__ float s_t[TS_R * TS_R * TS_R];
    __shared__ int s_nb[27];
    __shared__ int s_sum[128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x;
    if (EMIT && counts[2 * b] == 0) return;
    if (tid < 27) s_nb[tid] = nb[b * 27 + tid];
    __syncthreads();
    const int kx = block_keys[3 * b], ky = block_keys[3 * b + 1], kz = block_keys[3 * b + 2];
    const bool owned = mq3d_block_owned(kx, ky, kz, part);
    stage_tsdf(tsdf, s_nb, s_t, tid);
    __syncthreads();
    unsigned flags[16];  // 3 bits per voxel iteration
    for (int it = 0; it < 16; ++it) {
        int v = it * 256 + tid;
        int x = v & 15, y = (v >> 4) & 15, z = v >> 8;
        unsigned fl = 0;
        float wo = __ldg(weight + b * MQ3D_RES3 + v);
        if (owned && wo > weight_thr) {
            float to = TS(x, y, z);
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                int ex = x + (e == 0), ey = y + (e == 1), ez = z + (e == 2);
                int bi = s_nb[nb_of(ex) + 3 * nb_of(ey) + 9 * nb_of(ez)];
                if (bi < 0) continue;
                float wi = __ldg(weight + (int64_t)bi * MQ3D_RES3 + (((ez & 15) * 16 + (ey & 15)) * 16 + (ex & 15)));
                float ti = TS(ex, ey, ez);
                if (wi > weight_thr && __fmul_rn(ti, to) < 0.0f) fl |= 1u << e;
            }
        }
        flags[it] = fl;
        int n = __popc(fl);
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
        if (lane == 0) s_sum[it * 8 + warp] = n;
    }
    __syncthreads();
    int total = 0;
    if (warp == 0) total = warp0_exclusive_scan(s_sum, 128, lane);
    if (!EMIT) {
        if (tid == 0) { counts[2 * b] = total; counts[2 * b + 1] = 0; }
        return;
    }
    __syncthreads();
    const int64_t off = offsets[2 * b];
    for (int it = 0; it < 16; ++it) {
        unsigned fl = flags[it];
        int n = __popc(fl);
        int incl = n;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (!fl) continue;
        int64_t id = off + s_sum[it * 8 + warp] + incl - n;
        int v = it * 256 + tid;
        int x = v & 15, y = (v >> 4) & 15, z = v >> 8;
        float to = TS(x, y, z);
        float no[3] = {0.0f, 0.0f, 0.0f}, ni[3] = {0.0f, 0.0f, 0.0f};
        get_normal(s_t, s_nb, x, y, z, no);
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            if (!(fl & (1u << e))) continue;
            int ex = x + (e == 0), ey = y + (e == 1), ez = z + (e == 2);
            float ti = TS(ex, ey, ez);
            float ratio = __fdiv_rn(__fsub_rn(0.0f, to), __fsub_rn(ti, to));
            get_normal(s_t, s_nb, ex, ey, ez, ni);
            write_vertex(points, normals, pkeys, id, vs, kx * 16 + x, ky * 16 + y, kz * 16 + z, e, ratio, no, ni);
            ++id;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int mc_prepare(mq3d_grid *g, cudaStream_t st) {
    MQ3D_TRY(mq3d_grid_sync_count(g, st));
    int64_t n = g->n_blocks_host;
    MQ3D_REQUIRE(n <= g->capacity, "internal: block count exceeds pool capacity");
    if (n > g->mc_alloc_blocks || g->mc_offsets == nullptr) {
        cudaFree(g->mc_nb); cudaFree(g->mc_emask); cudaFree(g->mc_eprefix); cudaFree(g->mc_cubes);
        cudaFree(g->mc_counts); cudaFree(g->mc_offsets);
        g->mc_nb = nullptr; g->mc_emask = nullptr; g->mc_eprefix = nullptr; g->mc_cubes = nullptr;
        g->mc_counts = nullptr; g->mc_offsets = nullptr;
        g->mc_alloc_blocks = 0;
        int64_t a = n + n / 4 + 16;
        MQ3D_CUDA(cudaMalloc(&g->mc_nb, sizeof(int32_t) * 27 * a));
        MQ3D_CUDA(cudaMalloc(&g->mc_emask, sizeof(uint32_t) * EWORDS * a));
        MQ3D_CUDA(cudaMalloc(&g->mc_eprefix, sizeof(uint16_t) * EWORDS * a));
        MQ3D_CUDA(cudaMalloc(&g->mc_cubes, sizeof(uint8_t) * MQ3D_RES3 * a));
        MQ3D_CUDA(cudaMalloc(&g->mc_counts, sizeof(int32_t) * 2 * a));
        MQ3D_CUDA(cudaMalloc(&g->mc_offsets, sizeof(int64_t) * 2 * (a + 1)));
        g->mc_alloc_blocks = a;
    }
    g->mc_blocks = n;
    if (n > 0) {
        k_mc_neighbors<<<(unsigned)((n * 27 + 255) / 256), 256, 0, st>>>(g->hash, g->block_keys, n, g->mc_nb);
        MQ3D_CUDA(cudaGetLastError());
    }
    return MQ3D_OK;
}

static int mc_finish_count(mq3d_grid *g, cudaStream_t st, int64_t *a, int64_t *b) {
    int64_t n = g->mc_blocks;
    k_scan_counts<<<1, 1024, 0, st>>>(g->mc_counts, n, g->mc_offsets);
    MQ3D_CUDA(cudaGetLastError());
    int64_t tot[2];
    MQ3D_CUDA(cudaMemcpyAsync(tot, g->mc_offsets + 2 * n, sizeof(tot), cudaMemcpyDeviceToHost, st));
    MQ3D_CUDA(cudaStreamSynchronize(st));
    *a = tot[0];
    *b = tot[1];
    return MQ3D_OK;
}

extern "C" int mq3d_extract_mesh_count(mq3d_grid *g, float weight_threshold, int64_t *n_vertices,
                                       int64_t *n_triangles, void *stream) {
    MQ3D_REQUIRE(g && n_vertices && n_triangles, "null argument");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    g->mc_state = 0;
    MQ3D_TRY(mc_prepare(g, st));
    int64_t n = g->mc_blocks;
    if (n > 0) {
        k_mc_classify<<<(unsigned)n, 256, 0, st>>>(g->tsdf, g->weight, g->block_keys, g->mc_nb, weight_threshold, g->part,
                                                   g->mc_emask, g->mc_eprefix, g->mc_cubes, g->mc_counts);
        MQ3D_CUDA(cudaGetLastError());
    }
    MQ3D_TRY(mc_finish_count(g, st, &g->mc_V, &g->mc_T));
    MQ3D_REQUIRE(g->mc_V < 2147483647LL && g->mc_T < 2147483647LL, "mesh too large for int32 indices");
    *n_vertices = g->mc_V;
    *n_triangles = g->mc_T;
    g->mc_weight_thr = weight_threshold;
    g->mc_state = 1;
    return MQ3D_OK;
}

extern "C" int mq3d_extract_mesh_fill(mq3d_grid *g, float *vertices_dev, float *normals_dev, int32_t *triangles_dev,
                                      int32_t *vertex_keys_dev, void *stream) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    if (g->mc_state != 1) {
        mq3d_set_error("extract_mesh_fill called without a preceding extract_mesh_count on an unchanged grid");
        return MQ3D_ERR_STATE;
    }
    MQ3D_REQUIRE(g->mc_V == 0 || vertices_dev, "null vertex buffer");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    if (g->mc_blocks > 0 && (g->mc_V > 0 || g->mc_T > 0)) {
        k_mc_emit<<<(unsigned)g->mc_blocks, 256, 0, st>>>(g->tsdf, g->block_keys, g->mc_nb, g->mc_emask, g->mc_eprefix,
                                                          g->mc_cubes, g->mc_counts, g->mc_offsets, g->voxel_size,
                                                          vertices_dev, normals_dev, triangles_dev, vertex_keys_dev);
        MQ3D_CUDA(cudaGetLastError());
    }
    return MQ3D_OK;
}

extern "C" int mq3d_extract_points_count(mq3d_grid *g, float weight_threshold, int64_t *n_points, void *stream) {
    MQ3D_REQUIRE(g && n_points, "null argument");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    g->mc_state = 0;
    MQ3D_TRY(mc_prepare(g, st));
    int64_t n = g->mc_blocks;
    if (n > 0) {
        k_points<false><<<(unsigned)n, 256, 0, st>>>(g->tsdf, g->weight, g->block_keys, g->mc_nb, weight_threshold, g->part,
                                                     g->mc_counts, nullptr, g->voxel_size, nullptr, nullptr, nullptr);
        MQ3D_CUDA(cudaGetLastError());
    }
    int64_t dummy;
    MQ3D_TRY(mc_finish_count(g, st, &g->mc_V, &dummy));
    *n_points = g->mc_V;
    g->mc_weight_thr = weight_threshold;
    g->mc_state = 2;
    return MQ3D_OK;
}

extern "C" int mq3d_extract_points_fill(mq3d_grid *g, float *points_dev, float *normals_dev, int32_t *point_keys_dev,
                                        void *stream) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    if (g->mc_state != 2) {
        mq3d_set_error("extract_points_fill called without a preceding extract_points_count on an unchanged grid");
        return MQ3D_ERR_STATE;
    }
    MQ3D_REQUIRE(g->mc_V == 0 || points_dev, "null point buffer");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    if (g->mc_blocks > 0 && g->mc_V > 0) {
        k_points<true><<<(unsigned)g->mc_blocks, 256, 0, st>>>(g->tsdf, g->weight, g->block_keys, g->mc_nb, g->mc_weight_thr,
                                                               g->part, g->mc_counts, g->mc_offsets, g->voxel_size,
                                                               points_dev, normals_dev, point_keys_dev);
        MQ3D_CUDA(cudaGetLastError());
    }
    return MQ3D_OK;
}
