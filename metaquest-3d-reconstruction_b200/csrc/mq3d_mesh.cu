// K5: marching cubes / point-cloud extraction over the hashed voxel-block grid.
//
// Semantics: Open3D 0.19 ExtractTriangleMesh / ExtractPointCloud as reached from
// vbg.extract_triangle_mesh(weight_threshold, -1) and vbg.extract_point_cloud()
// (reference: processing/reconstruction/reconstruct_scene.py:90,105-108,186-189).
//
// GPU formulation (no global 16 B/voxel "mesh_structure" volume, no global atomics):
//   rows (streaming, one thread per 16-voxel x-row): tsdf and weight of every resident block are read exactly
//     once, fully coalesced, and reduced to a 16-bit validity row (w > thr) and a 16-bit sign row (tsdf < 0) --
//     1 KiB per block instead of 32 KiB; this is the only pass over the voxel data;
//   classify (one CTA per block): the 18-bit rows of the (-1..16)^3 neighbourhood are assembled from the block's
//     and its neighbours' 16-bit rows (L2-resident).  Everything else is row-wise bit arithmetic: cube validity
//     = AND of four rows and their shift, "surface" cubes = valid cubes whose 8 signs differ, edge
//     marks = sign difference AND (OR of the four cubes sharing the edge).  Per block it stores
//     16-bit edge-mark rows for the three axes, their exclusive prefix counts, the surface-cube rows,
//     triangle prefix counts and the sign rows (5.4 KB), so that
//       id(voxel, axis) = vertex_offset[block] + prefix[axis][z,y] + popc(mask[axis][z,y] & below(x))
//     is computable by any neighbour without a lookup table;
//   scan: exclusive prefix of per-block vertex / triangle counts;
//   emit (one CTA per non-empty block): one thread per (axis, row) walks its marked edges and writes
//     vertices + normals (tsdf read straight from global/L2: ~14 values per vertex), one thread per
//     cube row walks its surface cubes and writes triangles with ids resolved through the bit rows.
// Output order is deterministic.  A cube counts on this rank only if its block is owned (multi-GPU
// partition); single-GPU grids own every block.
#include <stdlib.h>

#include "mc_tables.h"
#include "mq3d_common.cuh"

#define TS_R 19      // tsdf region of the point-cloud kernel: coords -1..17
#define ROW_R 18     // validity / sign rows: (y,z) in -1..16, bit i <-> x = i - 1
#define SROW_WORDS (ROW_R * ROW_R)
// per-block uint16 table
#define R16_EMASK 0      // [3][256] edge marks, bit x
#define R16_EPREF 768    // [3][256] exclusive prefix of popc(edge marks) in (axis, z, y) order
#define R16_SURF 1536    // [256]    surface cubes of the own rows, bit x
#define R16_TPREF 1792   // [256]    exclusive prefix of per-row triangle counts
#define R16_SPREF 2048   // [256]    exclusive prefix of per-row surface-cube counts
#define R16_TPLANE 2304  // [3][256] bit planes of the per-cube triangle counts: bit x of plane k = bit k of the count of cube x
#define R16_WORDS 3072

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

// exclusive scan of n (multiple of 32) small counts held in shared memory, by warp 0; returns total
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
    return __shfl_sync(0xFFFFFFFFu, incl, 31);
}

// cube case (Bourke corner order) of the cube whose low corner is bit i of rows (y,z):
// s00 = (y,z), s10 = (y+1,z), s01 = (y,z+1), s11 = (y+1,z+1)
__device__ __forceinline__ int cube_case(unsigned s00, unsigned s10, unsigned s01, unsigned s11, int i) {
    unsigned a = (s00 >> i) & 3u, b = (s10 >> i) & 3u, c = (s01 >> i) & 3u, d = (s11 >> i) & 3u;
    // corners 0:(0,0,0) 1:(1,0,0) 2:(1,1,0) 3:(0,1,0) 4:(0,0,1) 5:(1,0,1) 6:(1,1,1) 7:(0,1,1)
    return (int)((a & 1u) | (a & 2u) | ((b & 2u) << 1) | ((b & 1u) << 3) | ((c & 1u) << 4) | ((c & 2u) << 4) |
                 ((d & 2u) << 5) | ((d & 1u) << 7));
}

// ------------------------------------------------------------------------------------------------
// classify
// ------------------------------------------------------------------------------------------------
// One thread per x-row of a block: 16 tsdf + 16 weight floats (four float4 each, a warp reads 2 KiB contiguous of
// both arrays) -> (validity << 16) | sign, bit x.  No neighbour table, no barriers: the pass over the voxel data
// runs at memory speed.  Open3D rejects `w <= thr`.
__global__ void __launch_bounds__(256)
k_mc_rows(const float *__restrict__ tsdf, const float *__restrict__ weight, int64_t n_blocks, float weight_thr,
          uint32_t *__restrict__ rows_vs) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;    // global row index: block * 256 + z * 16 + y
    if (r >= n_blocks * 256) return;
    const float4 *t4 = reinterpret_cast<const float4 *>(tsdf) + r * 4;
    const float4 *w4 = reinterpret_cast<const float4 *>(weight) + r * 4;
    float4 tq[4], wq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        tq[q] = __ldcs(t4 + q);
        wq[q] = __ldcs(w4 + q);
    }
    unsigned vrow = 0, srow = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        vrow |= ((wq[q].x > weight_thr ? 1u : 0u) | (wq[q].y > weight_thr ? 2u : 0u) | (wq[q].z > weight_thr ? 4u : 0u) |
                 (wq[q].w > weight_thr ? 8u : 0u)) << (4 * q);
        srow |= ((tq[q].x < 0.0f ? 1u : 0u) | (tq[q].y < 0.0f ? 2u : 0u) | (tq[q].z < 0.0f ? 4u : 0u) |
                 (tq[q].w < 0.0f ? 8u : 0u)) << (4 * q);
    }
    rows_vs[r] = (vrow << 16) | srow;
}

#define MC_CLASSIFY_THREADS 256
// One CTA per block, working on the 16-bit rows of k_mc_rows only (a few KiB per block, L2-resident): neighbour
// indices are read by the threads that need them (no staging barrier), cube validity is recomputed by the rows
// that need it (no second staging barrier), the five per-row counts are scanned in registers with warp shuffles
// (no serial warp-0 scan): two barriers per block, 6 CTAs per SM to hide what latency remains.
__global__ void __launch_bounds__(MC_CLASSIFY_THREADS, 6)
k_mc_classify(const uint32_t *__restrict__ rows_vs, const int32_t *__restrict__ block_keys,
              const int32_t *__restrict__ nb, int64_t n_blocks, Partition part,
              uint32_t *__restrict__ srow_out, uint16_t *__restrict__ rows16, int32_t *__restrict__ counts) {
    __shared__ unsigned s_valid[SROW_WORDS], s_sign[SROW_WORDS];
    __shared__ unsigned s_own[4];            // ownership masks of the cube rows with (cube y == -1, cube z == -1) = bits of the index
    __shared__ int s_wtot[8][3];
    __shared__ unsigned char s_tc[256];      // triangles per cube case (constant-bank reads with a per-lane index serialise)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x;
    s_tc[tid] = (unsigned char)(__ldg(&MC_TRI_PACKED[tid]) >> 60);
    {
        // ---- phase A: one thread per (y,z) row of the (-1..16)^2 neighbourhood: the row's 16 interior bits from the
        // block that holds it, the x = -1 / x = 16 bits from that block's -x / +x neighbours (bit i <-> x = i - 1);
        // missing blocks contribute zeros
        unsigned sign_mix = 0;                 // bit 0: a valid voxel with tsdf < 0, bit 1: a valid voxel with tsdf >= 0
        for (int r = tid; r < SROW_WORDS; r += MC_CLASSIFY_THREADS) {
            const int ry = r % ROW_R - 1, rz = r / ROW_R - 1;
            const int k0 = 3 * nb_of(ry) + 9 * nb_of(rz);
            const int bm = __ldg(nb + b * 27 + k0), b0 = __ldg(nb + b * 27 + k0 + 1), bp = __ldg(nb + b * 27 + k0 + 2);
            const int ri = (rz & 15) * 16 + (ry & 15);
            const unsigned w0 = b0 >= 0 ? __ldg(rows_vs + (int64_t)b0 * 256 + ri) : 0u;
            const unsigned wm = bm >= 0 ? __ldg(rows_vs + (int64_t)bm * 256 + ri) : 0u;
            const unsigned wp = bp >= 0 ? __ldg(rows_vs + (int64_t)bp * 256 + ri) : 0u;
            const unsigned vrow = ((w0 >> 16) << 1) | (wm >> 31) | (((wp >> 16) & 1u) << 17);
            const unsigned srow = ((w0 & 0xFFFFu) << 1) | ((wm >> 15) & 1u) | ((wp & 1u) << 17);
            s_valid[r] = vrow;
            s_sign[r] = srow;
            srow_out[b * SROW_WORDS + r] = srow;
            sign_mix |= ((srow & vrow) != 0u ? 1u : 0u) | ((~srow & vrow) != 0u ? 2u : 0u);
        }
        if (tid < 4) {
            // cubes count only in owned blocks (multi-GPU partition); cube x = -1 lies in block x - 1, cube y / z = -1 in
            // block y - 1 / z - 1.  Bit 0 <-> cube x = -1, bits 1..16 <-> cubes of this block's x range.
            unsigned own = 0x1FFFFu;
            if (part.world > 1) {
                const int kx = block_keys[3 * b], ky = block_keys[3 * b + 1] - (tid & 1), kz = block_keys[3 * b + 2] - (tid >> 1);
                own = (mq3d_block_owned(kx - 1, ky, kz, part) ? 1u : 0u) | (mq3d_block_owned(kx, ky, kz, part) ? 0x1FFFEu : 0u);
            }
            s_own[tid] = own;
        }
        // Most blocks of a truncation band hold no surface: a surface cube (and a marked edge) needs eight valid corners
        // with differing signs, so when the valid voxels of the (-1..16)^3 neighbourhood all have the same sign there is
        // nothing to emit.  Such a block writes zero counts and none of its tables -- nobody reads them: emit skips the
        // block, and a neighbour resolves vertex ids only on marked edges, which this neighbourhood does not have.
        // (__syncthreads_or is a vote, not a bitwise OR: one barrier per bit; the first one replaces the staging barrier)
        const int any_set = __syncthreads_or((int)(sign_mix & 1u)), any_clear = __syncthreads_or((int)(sign_mix & 2u));
        if (!(any_set && any_clear)) {
            if (tid == 0) {
                counts[2 * b] = 0;
                counts[2 * b + 1] = 0;
            }
            return;
        }
    // ---- phase C: one thread per own row (y,z): edge marks, surface cubes, triangle count ----
    int cnt[5] = {0, 0, 0, 0, 0};   // popc(x marks), popc(y marks), popc(z marks), triangles, surface cubes
    {
        const int y = tid & 15, z = tid >> 4;
        // valid (and owned) cubes of the cube rows (y,z) (y-1,z) (y,z-1) (y-1,z-1), bit i <-> cube x = i-1: AND of the
        // four neighbourhood rows around the cube row and of their shift by one (region row index = coordinate + 1)
        unsigned v[3][3];
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) v[dz][dy] = s_valid[(z + dz) * ROW_R + y + dy];
        unsigned cok[2][2];
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const unsigned m = v[dz][dy] & v[dz][dy + 1] & v[dz + 1][dy] & v[dz + 1][dy + 1];
                // cube row (y - 1 + dy, z - 1 + dz)
                cok[dz][dy] = m & (m >> 1) & s_own[((y + dy == 0) ? 1 : 0) | ((z + dz == 0) ? 2 : 0)];
            }
        const unsigned c00 = cok[1][1], cm0 = cok[1][0], c0m = cok[0][1], cmm = cok[0][0];
        const unsigned s00 = s_sign[(z + 1) * ROW_R + y + 1], s10 = s_sign[(z + 1) * ROW_R + y + 2];
        const unsigned s01 = s_sign[(z + 2) * ROW_R + y + 1], s11 = s_sign[(z + 2) * ROW_R + y + 2];
        const unsigned ex = (s00 ^ (s00 >> 1)) & (c00 | cm0 | c0m | cmm);
        const unsigned cy_ = c00 | c0m, cz_ = c00 | cm0;
        const unsigned ey = (s00 ^ s10) & (cy_ | (cy_ << 1));
        const unsigned ez = (s00 ^ s01) & (cz_ | (cz_ << 1));
        const unsigned mx = (ex >> 1) & 0xFFFFu, my = (ey >> 1) & 0xFFFFu, mz = (ez >> 1) & 0xFFFFu;
        // surface cubes of this row: valid cubes whose 8 corner signs are not all equal
        const unsigned same4 = ~((s00 ^ s10) | (s00 ^ s01) | (s00 ^ s11));
        const unsigned flat = same4 & (same4 >> 1) & ~(s00 ^ (s00 >> 1));
        const unsigned surf = (c00 & ~flat) >> 1 & 0xFFFFu;
        // triangles per surface cube, kept as three bit planes so that any cube's offset inside its row is a popcount
        unsigned p0 = 0, p1 = 0, p2 = 0;
        for (unsigned m = surf; m; m &= m - 1) {
            const int i = __ffs(m);                           // cube x = i - 1
            const unsigned t = s_tc[cube_case(s00, s10, s01, s11, i)];
            p0 |= (t & 1u) << (i - 1);
            p1 |= ((t >> 1) & 1u) << (i - 1);
            p2 |= (t >> 2) << (i - 1);
        }
        const int ntri = __popc(p0) + 2 * __popc(p1) + 4 * __popc(p2);
        uint16_t *r16 = rows16 + b * R16_WORDS;
        r16[R16_EMASK + tid] = (uint16_t)mx;
        r16[R16_EMASK + 256 + tid] = (uint16_t)my;
        r16[R16_EMASK + 512 + tid] = (uint16_t)mz;
        r16[R16_SURF + tid] = (uint16_t)surf;
        r16[R16_TPLANE + tid] = (uint16_t)p0;
        r16[R16_TPLANE + 256 + tid] = (uint16_t)p1;
        r16[R16_TPLANE + 512 + tid] = (uint16_t)p2;
        cnt[0] = __popc(mx); cnt[1] = __popc(my); cnt[2] = __popc(mz); cnt[3] = ntri; cnt[4] = __popc(surf);
    }
    // block-wide exclusive scans of the five counts over the 256 rows, packed into three words (every prefix stays
    // below 2^16: at most 4096 edges per axis, 4096 cubes and 5 * 4096 triangles per block): warp shuffles, then the
    // 8 warp totals are scanned by 8 lanes of every warp
    unsigned pk[3] = {(unsigned)cnt[0] | ((unsigned)cnt[1] << 16), (unsigned)cnt[2] | ((unsigned)cnt[3] << 16), (unsigned)cnt[4]};
    unsigned incl[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned v = pk[k];
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if (lane >= o) v += t;
        }
        incl[k] = v;
        if (lane == 31) s_wtot[warp][k] = (int)v;
    }
    __syncthreads();
    {
        unsigned base[3], tot[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const unsigned own = lane < 8 ? (unsigned)s_wtot[lane][k] : 0u;
            unsigned v = own;
            for (int o = 1; o < 8; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xFFFFFFFFu, v, o);
                if (lane >= o) v += t;
            }
            base[k] = __shfl_sync(0xFFFFFFFFu, v - own, warp);     // totals of the warps before this one
            tot[k] = __shfl_sync(0xFFFFFFFFu, v, 7);
        }
        const unsigned ex0 = base[0] + incl[0] - pk[0], ex1 = base[1] + incl[1] - pk[1], ex2 = base[2] + incl[2] - pk[2];
        const unsigned tx = tot[0] & 0xFFFFu, ty = tot[0] >> 16, tz = tot[1] & 0xFFFFu;
        uint16_t *r16 = rows16 + b * R16_WORDS;
        // vertex numbering is axis-major: x edges, then y edges, then z edges
        r16[R16_EPREF + tid] = (uint16_t)(ex0 & 0xFFFFu);
        r16[R16_EPREF + 256 + tid] = (uint16_t)(tx + (ex0 >> 16));
        r16[R16_EPREF + 512 + tid] = (uint16_t)(tx + ty + (ex1 & 0xFFFFu));
        r16[R16_TPREF + tid] = (uint16_t)(ex1 >> 16);
        r16[R16_SPREF + tid] = (uint16_t)ex2;
        if (tid == 0) {
            counts[2 * b] = (int)(tx + ty + tz);
            counts[2 * b + 1] = (int)(tot[1] >> 16);
        }
    }
    }
}

// ------------------------------------------------------------------------------------------------
// scan of per-block (a, b) counts -> int64 exclusive offsets [n+1][2]: k_scan_totals sums chunks of
// SCAN_CHUNK blocks, k_scan_counts adds the totals of the preceding chunks (a few dozen values) to a local
// scan of its own chunk.  Deterministic, no inter-CTA waiting.
// ------------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_PER 8
#define SCAN_CHUNK (SCAN_THREADS * SCAN_PER)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_totals(const int32_t *__restrict__ counts, int64_t n, long long *__restrict__ totals) {
    __shared__ long long s_a[SCAN_THREADS / 32], s_b[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long a = 0, c = 0;
    for (int q = 0; q < SCAN_PER; ++q) {
        const int64_t i = (int64_t)blockIdx.x * SCAN_CHUNK + q * SCAN_THREADS + tid;
        if (i < n) {
            const int2 v = __ldg(reinterpret_cast<const int2 *>(counts) + i);
            a += v.x;
            c += v.y;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    }
    if (lane == 0) { s_a[warp] = a; s_b[warp] = c; }
    __syncthreads();
    if (tid == 0) {
        long long ta = 0, tb = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) { ta += s_a[w]; tb += s_b[w]; }
        totals[2 * blockIdx.x] = ta;
        totals[2 * blockIdx.x + 1] = tb;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_counts(const int32_t *__restrict__ counts, int64_t n, const long long *__restrict__ totals,
                                                              int64_t *__restrict__ offsets) {
    __shared__ long long s_a[SCAN_THREADS / 32], s_b[SCAN_THREADS / 32];
    __shared__ long long s_carry[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // carry-in: totals of the preceding chunks
    {
        long long a = 0, c = 0;
        for (int i = tid; i < (int)blockIdx.x; i += SCAN_THREADS) { a += totals[2 * i]; c += totals[2 * i + 1]; }
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
            c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        }
        if (lane == 0) { s_a[warp] = a; s_b[warp] = c; }
        __syncthreads();
        if (tid == 0) {
            long long ta = 0, tb = 0;
            for (int w = 0; w < SCAN_THREADS / 32; ++w) { ta += s_a[w]; tb += s_b[w]; }
            s_carry[0] = ta;
            s_carry[1] = tb;
        }
        __syncthreads();
    }
    // SCAN_PER consecutive blocks per thread
    const int64_t i0 = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)tid * SCAN_PER;
    long long a = 0, c = 0, la[SCAN_PER], lc[SCAN_PER];
#pragma unroll
    for (int q = 0; q < SCAN_PER; ++q) {
        la[q] = a;
        lc[q] = c;
        if (i0 + q < n) {
            const int2 v = __ldg(reinterpret_cast<const int2 *>(counts) + i0 + q);
            a += v.x;
            c += v.y;
        }
    }
    long long ia = a, ic = c;
    for (int o = 1; o < 32; o <<= 1) {
        const long long ta = __shfl_up_sync(0xFFFFFFFFu, ia, o), tc = __shfl_up_sync(0xFFFFFFFFu, ic, o);
        if (lane >= o) { ia += ta; ic += tc; }
    }
    __syncthreads();
    if (lane == 31) { s_a[warp] = ia; s_b[warp] = ic; }
    __syncthreads();
    long long wa = 0, wb = 0, ta = 0, tb = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) { wa += s_a[w]; wb += s_b[w]; }
        ta += s_a[w];
        tb += s_b[w];
    }
    const long long pa = s_carry[0] + wa + ia - a, pb = s_carry[1] + wb + ic - c;   // exclusive prefix of this thread
#pragma unroll
    for (int q = 0; q < SCAN_PER; ++q) {
        if (i0 + q < n) {
            offsets[2 * (i0 + q)] = pa + la[q];
            offsets[2 * (i0 + q) + 1] = pb + lc[q];
        }
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        offsets[2 * n] = s_carry[0] + ta;
        offsets[2 * n + 1] = s_carry[1] + tb;
    }
}

// ------------------------------------------------------------------------------------------------
// vertex writer shared by mesh and point cloud
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_vertex(float *__restrict__ verts, float *__restrict__ normals, int32_t *__restrict__ vkeys,
                                             int64_t id, float vs, int gx, int gy, int gz, int e, float ratio,
                                             const float *no, const float *ne) {
    float rx = __fmul_rn(ratio, e == 0 ? 1.0f : 0.0f), ry = __fmul_rn(ratio, e == 1 ? 1.0f : 0.0f),
          rz = __fmul_rn(ratio, e == 2 ? 1.0f : 0.0f);
    if (verts) {
        verts[3 * id + 0] = __fmul_rn(vs, __fadd_rn((float)gx, rx));
        verts[3 * id + 1] = __fmul_rn(vs, __fadd_rn((float)gy, ry));
        verts[3 * id + 2] = __fmul_rn(vs, __fadd_rn((float)gz, rz));
    }
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

// voxel (x,y,z) relative to the block (coords -1..17) -> linear pool index, -1 if its block is missing
__device__ __forceinline__ int64_t nb_lin(const int *s_nb, int x, int y, int z) {
    const int bi = s_nb[nb_of(x) + 3 * nb_of(y) + 9 * nb_of(z)];
    return bi < 0 ? -1 : (int64_t)bi * MQ3D_RES3 + (((z & 15) * 16 + (y & 15)) * 16 + (x & 15));
}

// DeviceGetNormal on global memory: central differences where both neighbours exist; other components
// keep their previous value (Open3D's caller-visible stale behaviour).  The six loads are issued together
// (a missing neighbour reads the voxel's own block start and is masked); returns the mask of the components
// that were computed.
__device__ __forceinline__ unsigned get_normal_g(const float *__restrict__ tsdf, const int *s_nb, int x, int y, int z, float *n) {
    // neighbours inside the voxel's own block sit at +-1, +-16, +-256 from it; only voxels on a block face go
    // through the neighbour table
    const int64_t c = nb_lin(s_nb, x, y, z);
    const int co[3] = {x & 15, y & 15, z & 15};
    const int stride[3] = {1, 16, 256};
    int64_t ip[3], im[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        ip[a] = (co[a] < 15 && c >= 0) ? c + stride[a] : nb_lin(s_nb, x + (a == 0), y + (a == 1), z + (a == 2));
        im[a] = (co[a] > 0 && c >= 0) ? c - stride[a] : nb_lin(s_nb, x - (a == 0), y - (a == 1), z - (a == 2));
    }
    float vp[3], vm[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        vp[a] = __ldg(tsdf + (ip[a] < 0 ? 0 : ip[a]));
        vm[a] = __ldg(tsdf + (im[a] < 0 ? 0 : im[a]));
    }
    unsigned have = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (ip[a] >= 0 && im[a] >= 0) {
            n[a] = __fsub_rn(vp[a], vm[a]);
            have |= 1u << a;
        }
    }
    return have;
}

// Open3D's colour branch of ExtractTriangleMesh / ExtractPointCloud: ((1-ratio)*c_o + ratio*c_e) / 255
// on the float32 colour attribute (voxel linear indices lo / le into the [n][4096][3] pool)
__device__ __forceinline__ void write_color(float *__restrict__ out, int64_t id, const float *__restrict__ color,
                                            int64_t lo, int64_t le, float ratio) {
    const float om = __fsub_rn(1.0f, ratio);
#pragma unroll
    for (int c = 0; c < 3; ++c)
        out[3 * id + c] = __fdiv_rn(__fadd_rn(__fmul_rn(om, __ldg(color + 3 * lo + c)), __fmul_rn(ratio, __ldg(color + 3 * le + c))),
                                    255.0f);
}

// A block carries ~100 vertices and ~100 surface cubes on real surfaces: the emit kernels use EMIT_THREADS = 64
// threads per block (two warps), so that most lanes have work and the per-warp instruction stream is paid twice
// rather than eight times per block.
#define EMIT_THREADS 64

// vertex colours of the mesh laid out by k_mc_classify / k_scan_counts (same vertex order as k_mc_emit).
// cap_v >= 0: do nothing if the mesh does not fit the caller's buffers (single-call extraction).
__global__ void __launch_bounds__(EMIT_THREADS)
k_mc_colors(const float *__restrict__ tsdf, const float *__restrict__ color, const int32_t *__restrict__ nb,
            const uint16_t *__restrict__ rows16, const int32_t *__restrict__ counts, const int64_t *__restrict__ offsets,
            float *__restrict__ vcolors, int64_t n_blocks, int64_t cap_v, int64_t cap_t) {
    __shared__ int s_nb[27];
    __shared__ uint16_t s_ep[768], s_em[768];
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    if (cap_v >= 0 && (offsets[2 * n_blocks] > cap_v || offsets[2 * n_blocks + 1] > cap_t)) return;
    const int nv = counts[2 * b];
    if (nv == 0) return;
    if (tid < 27) s_nb[tid] = nb[b * 27 + tid];
    for (int i = tid; i < 768; i += EMIT_THREADS) {
        s_ep[i] = rows16[b * R16_WORDS + R16_EPREF + i];
        s_em[i] = rows16[b * R16_WORDS + R16_EMASK + i];
    }
    __syncthreads();
    const int64_t voff = offsets[2 * b];
    for (int j = tid; j < nv; j += EMIT_THREADS) {
        int lo = 0, hi = 767;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((int)s_ep[mid] <= j) lo = mid; else hi = mid - 1;
        }
        const int item = lo, e = item >> 8, row = item & 255, y = row & 15, z = row >> 4;
        const int x = (int)__fns((unsigned)s_em[item], 0, j - (int)s_ep[item] + 1);
        const int64_t lin_o = b * MQ3D_RES3 + row * 16 + x;
        const int64_t lin_e = nb_lin(s_nb, x + (e == 0), y + (e == 1), z + (e == 2));
        const float to = __ldg(tsdf + lin_o), te = __ldg(tsdf + lin_e);
        const float ratio = __fdiv_rn(__fsub_rn(0.0f, to), __fsub_rn(te, to));
        write_color(vcolors, voff + j, color, lin_o, lin_e, ratio);
    }
}

// ------------------------------------------------------------------------------------------------
// emit
// ------------------------------------------------------------------------------------------------
// cap_v >= 0: do nothing if the mesh does not fit the caller's buffers (single-call extraction: the totals sit
// at the end of the scanned offsets, the host learns them afterwards and retries with exact sizes).
__global__ void __launch_bounds__(EMIT_THREADS)
k_mc_emit(const float *__restrict__ tsdf, const int32_t *__restrict__ block_keys, const int32_t *__restrict__ nb,
          const uint32_t *__restrict__ srow, const uint16_t *__restrict__ rows16, const int32_t *__restrict__ counts,
          const int64_t *__restrict__ offsets, float vs, float *__restrict__ verts, float *__restrict__ normals,
          int32_t *__restrict__ tris, int32_t *__restrict__ vkeys, int64_t n_blocks, int64_t cap_v, int64_t cap_t) {
    __shared__ int s_nb[27];
    __shared__ __align__(16) uint16_t s_r16[R16_WORDS];
    __shared__ unsigned s_sign[SROW_WORDS];
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    if (cap_v >= 0 && (offsets[2 * n_blocks] > cap_v || offsets[2 * n_blocks + 1] > cap_t)) return;
    const int nv = counts[2 * b], nt = counts[2 * b + 1];
    if (nv == 0 && nt == 0) return;
    if (tid < 27) s_nb[tid] = nb[b * 27 + tid];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(rows16 + b * R16_WORDS);
        for (int i = tid; i < R16_WORDS / 8; i += EMIT_THREADS) reinterpret_cast<uint4 *>(s_r16)[i] = __ldg(src + i);   // 6 KB
    }
    const bool do_tris = nt > 0 && tris != nullptr;
    if (do_tris)
        for (int i = tid; i < SROW_WORDS; i += EMIT_THREADS) s_sign[i] = srow[b * SROW_WORDS + i];
    __syncthreads();
    const int kx = block_keys[3 * b], ky = block_keys[3 * b + 1], kz = block_keys[3 * b + 2];
    const int64_t voff = offsets[2 * b], toff = offsets[2 * b + 1];
    // ---- vertices: one thread per vertex; (axis,row) by binary search in the prefix table, x = k-th set bit ----
    if (nv > 0) {
        for (int j = tid; j < nv; j += EMIT_THREADS) {
            int lo = 0, hi = 767;                      // largest item with prefix <= j and a non-empty mask
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if ((int)s_r16[R16_EPREF + mid] <= j) lo = mid; else hi = mid - 1;
            }
            const int item = lo;
            const unsigned m = s_r16[R16_EMASK + item];
            const int e = item >> 8, row = item & 255, y = row & 15, z = row >> 4;
            const int64_t id = voff + j;
            {
                const int x = (int)__fns(m, 0, j - (int)s_r16[R16_EPREF + item] + 1);
                const int ex = x + (e == 0), ey = y + (e == 1), ez = z + (e == 2);
                const float to = __ldg(tsdf + b * MQ3D_RES3 + row * 16 + x);
                const float te = __ldg(tsdf + nb_lin(s_nb, ex, ey, ez));
                const float ratio = __fdiv_rn(__fsub_rn(0.0f, to), __fsub_rn(te, to));
                // Open3D keeps `ne` across the three edges of a voxel: components that cannot be recomputed at a
                // later edge retain the value of the previous marked edge.  That only shows when the end voxel
                // lacks a neighbour (border of the allocated region): the common case needs this edge's gradient only.
                float no[3] = {0.0f, 0.0f, 0.0f}, ne[3] = {0.0f, 0.0f, 0.0f};
                get_normal_g(tsdf, s_nb, x, y, z, no);
                if (get_normal_g(tsdf, s_nb, ex, ey, ez, ne) != 7u) {
                    ne[0] = ne[1] = ne[2] = 0.0f;
                    for (int p = 0; p < e; ++p)
                        if ((s_r16[R16_EMASK + p * 256 + row] >> x) & 1u)
                            get_normal_g(tsdf, s_nb, x + (p == 0), y + (p == 1), z, ne);
                    get_normal_g(tsdf, s_nb, ex, ey, ez, ne);
                }
                write_vertex(verts, normals, vkeys, id, vs, kx * 16 + x, ky * 16 + y, kz * 16 + z, e, ratio, no, ne);
            }
        }
    }
    // ---- triangles: one thread per surface cube (row by binary search over the per-row cube counts) ----
    if (do_tris) {
        const uint16_t *s_sp = s_r16 + R16_SPREF;      // no barrier between the vertex and the triangle phase
        const int ncubes = (int)s_sp[255] + __popc((unsigned)s_r16[R16_SURF + 255]);
        for (int j = tid; j < ncubes; j += EMIT_THREADS) {
            int lo = 0, hi = 255;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if ((int)s_sp[mid] <= j) lo = mid; else hi = mid - 1;
            }
            const int row = lo, y = row & 15, z = row >> 4;
            const unsigned surf = s_r16[R16_SURF + row];
            const int kth = j - (int)s_sp[row];
            const unsigned s00 = s_sign[(z + 1) * ROW_R + y + 1], s10 = s_sign[(z + 1) * ROW_R + y + 2];
            const unsigned s01 = s_sign[(z + 2) * ROW_R + y + 1], s11 = s_sign[(z + 2) * ROW_R + y + 2];
            const int x = (int)__fns(surf, 0, kth + 1);                // x of the row's kth surface cube
            const unsigned below = (1u << x) - 1u;                     // the row's cubes before it: their triangle counts
            int64_t tbase = toff + s_r16[R16_TPREF + row] + __popc(s_r16[R16_TPLANE + row] & below) +
                            2 * __popc(s_r16[R16_TPLANE + 256 + row] & below) + 4 * __popc(s_r16[R16_TPLANE + 512 + row] & below);
            {
                // the case's triangle list: 4-bit edge ids, three per triangle (one 64-bit load, L1-resident table)
                unsigned long long tt = __ldg(&MC_TRI_PACKED[cube_case(s00, s10, s01, s11, x + 1)]);
                const int n_tri = (int)(tt >> 60);
                for (int k = 0; k < n_tri; ++k, ++tbase) {
#pragma unroll
                    for (int vtx = 0; vtx < 3; ++vtx, tt >>= 4) {
                        const int edge = (int)(tt & 15ull);
                        const unsigned sh = (unsigned)(MC_EDGE_SHIFT_BITS >> (5 * edge)) & 31u;   // dx | dy<<1 | dz<<2 | axis<<3
                        const int ox = x + (int)(sh & 1u), oy = y + (int)((sh >> 1) & 1u), oz = z + (int)((sh >> 2) & 1u);
                        const int ax = (int)(sh >> 3);
                        const int nbk = nb_of(ox) + 3 * nb_of(oy) + 9 * nb_of(oz);
                        const int item = ax * 256 + (oz & 15) * 16 + (oy & 15);
                        const unsigned low = (1u << (ox & 15)) - 1u;
                        int64_t id;
                        if (nbk == 13) {
                            id = voff + s_r16[R16_EPREF + item] + __popc(s_r16[R16_EMASK + item] & low);
                        } else {
                            const int64_t bn = s_nb[nbk];
                            const uint16_t *rn = rows16 + bn * R16_WORDS;
                            id = offsets[2 * bn] + __ldg(rn + R16_EPREF + item) + __popc(__ldg(rn + R16_EMASK + item) & low);
                        }
                        tris[3 * tbase + (2 - vtx)] = (int32_t)id;   // winding reversed
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// point cloud: count + emit.  The crossing test is Open3D's `tsdf_i * tsdf_o < 0` on the float product
// (not a sign comparison: a zero or an underflowing product does not count), so it stays per voxel.
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

__device__ __forceinline__ void get_normal(const float *s_t, const int *s_nb, int x, int y, int z, float *n) {
    if (EXISTS(x + 1, y, z) && EXISTS(x - 1, y, z)) n[0] = __fsub_rn(TS(x + 1, y, z), TS(x - 1, y, z));
    if (EXISTS(x, y + 1, z) && EXISTS(x, y - 1, z)) n[1] = __fsub_rn(TS(x, y + 1, z), TS(x, y - 1, z));
    if (EXISTS(x, y, z + 1) && EXISTS(x, y, z - 1)) n[2] = __fsub_rn(TS(x, y, z + 1), TS(x, y, z - 1));
}

template <bool EMIT>
__global__ void __launch_bounds__(256)
k_points(const float *__restrict__ tsdf, const float *__restrict__ weight, const int32_t *__restrict__ block_keys,
         const int32_t *__restrict__ nb, float weight_thr, Partition part, int32_t *__restrict__ counts,
         const int64_t *__restrict__ offsets, float vs, float *__restrict__ points, float *__restrict__ normals,
         int32_t *__restrict__ pkeys, const float *__restrict__ color, float *__restrict__ pcolors) {
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
            if (points) write_vertex(points, normals, pkeys, id, vs, kx * 16 + x, ky * 16 + y, kz * 16 + z, e, ratio, no, ni);
            if (pcolors) write_color(pcolors, id, color, b * MQ3D_RES3 + v, nb_lin(s_nb, ex, ey, ez), ratio);
            ++id;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int mc_prepare(mq3d_grid *g, cudaStream_t st) {
    MQ3D_TRY(mq3d_grid_fresh_count(g, st));
    int64_t n = g->n_blocks_host;
    MQ3D_REQUIRE(n <= g->capacity, "internal: block count exceeds pool capacity");
    if (n > g->mc_alloc_blocks || g->mc_offsets == nullptr) {
        cudaFree(g->mc_nb); cudaFree(g->mc_emask); cudaFree(g->mc_eprefix); cudaFree(g->mc_counts); cudaFree(g->mc_offsets);
        cudaFree(g->mc_totals); cudaFree(g->mc_rows);
        g->mc_rows = nullptr;
        g->mc_nb = nullptr; g->mc_emask = nullptr; g->mc_eprefix = nullptr; g->mc_counts = nullptr; g->mc_offsets = nullptr;
        g->mc_totals = nullptr;
        g->mc_alloc_blocks = 0;
        int64_t a = n + n / 4 + 16;
        MQ3D_CUDA(cudaMalloc(&g->mc_nb, sizeof(int32_t) * 27 * a));
        MQ3D_CUDA(cudaMalloc(&g->mc_rows, sizeof(uint32_t) * 256 * a));             // (validity << 16) | sign per x-row
        MQ3D_CUDA(cudaMalloc(&g->mc_emask, sizeof(uint32_t) * SROW_WORDS * a));     // 18-bit sign rows of the neighbourhood
        MQ3D_CUDA(cudaMalloc(&g->mc_eprefix, sizeof(uint16_t) * R16_WORDS * a));    // 16-bit row tables
        MQ3D_CUDA(cudaMalloc(&g->mc_counts, sizeof(int32_t) * 2 * a));
        MQ3D_CUDA(cudaMalloc(&g->mc_offsets, sizeof(int64_t) * 2 * (a + 1)));
        MQ3D_CUDA(cudaMalloc(&g->mc_totals, sizeof(long long) * 2 * (a / SCAN_CHUNK + 2)));
        g->mc_alloc_blocks = a;
    }
    g->mc_blocks = n;
    if (n > 0) {
        k_mc_neighbors<<<(unsigned)((n * 27 + 255) / 256), 256, 0, st>>>(g->hash, g->block_keys, n, g->mc_nb);
        MQ3D_CUDA(cudaGetLastError());
    }
    return MQ3D_OK;
}

static int mc_classify(mq3d_grid *g, float weight_threshold, cudaStream_t st, MqTrace *tr = nullptr) {
    const int64_t n = g->mc_blocks;
    k_mc_rows<<<(unsigned)n, 256, 0, st>>>(g->tsdf, g->weight, n, weight_threshold, g->mc_rows);
    if (tr) tr->mark("rows");
    k_mc_classify<<<(unsigned)n, MC_CLASSIFY_THREADS, 0, st>>>(g->mc_rows, g->block_keys, g->mc_nb, n, g->part, g->mc_emask,
                                                               g->mc_eprefix, g->mc_counts);
    if (tr) tr->mark("classify");
    MQ3D_CUDA(cudaGetLastError());
    return MQ3D_OK;
}

static int mc_scan(mq3d_grid *g, cudaStream_t st) {
    const int64_t n = g->mc_blocks;
    const unsigned chunks = (unsigned)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
    k_scan_totals<<<chunks, SCAN_THREADS, 0, st>>>(g->mc_counts, n, g->mc_totals);
    k_scan_counts<<<chunks, SCAN_THREADS, 0, st>>>(g->mc_counts, n, g->mc_totals, g->mc_offsets);
    MQ3D_CUDA(cudaGetLastError());
    return MQ3D_OK;
}

static int mc_finish_count(mq3d_grid *g, cudaStream_t st, int64_t *a, int64_t *b) {
    int64_t n = g->mc_blocks;
    if (n > 0) {
        MQ3D_TRY(mc_scan(g, st));
    } else {
        MQ3D_CUDA(cudaMemsetAsync(g->mc_offsets, 0, sizeof(int64_t) * 2, st));
    }
    MQ3D_CUDA(cudaMemcpyAsync(g->pinned_host64, g->mc_offsets + 2 * n, sizeof(int64_t) * 2, cudaMemcpyDeviceToHost, st));
    MQ3D_CUDA(cudaStreamSynchronize(st));
    *a = g->pinned_host64[0];
    *b = g->pinned_host64[1];
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
        MQ3D_TRY(mc_classify(g, weight_threshold, st));
    }
    MQ3D_TRY(mc_finish_count(g, st, &g->mc_V, &g->mc_T));
    MQ3D_REQUIRE(g->mc_V < 2147483647LL && g->mc_T < 2147483647LL, "mesh too large for int32 indices");
    *n_vertices = g->mc_V;
    *n_triangles = g->mc_T;
    g->mc_weight_thr = weight_threshold;
    g->mc_state = 1;
    return MQ3D_OK;
}

extern "C" int mq3d_extract_mesh(mq3d_grid *g, float weight_threshold, float *vertices_dev, float *normals_dev,
                                 int32_t *triangles_dev, int32_t *vertex_keys_dev, float *colors_dev, int64_t cap_vertices,
                                 int64_t cap_triangles, int64_t *n_vertices, int64_t *n_triangles, void *stream) {
    MQ3D_REQUIRE(g && n_vertices && n_triangles, "null argument");
    MQ3D_REQUIRE(cap_vertices >= 0 && cap_triangles >= 0, "negative capacity");
    MQ3D_REQUIRE(cap_vertices == 0 || vertices_dev, "null vertex buffer");
    MQ3D_REQUIRE(colors_dev == nullptr || g->color != nullptr, "grid has no colour attribute");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    g->mc_state = 0;
    MQ3D_TRY(mc_prepare(g, st));
    const int64_t n = g->mc_blocks;
    if (n > 0) {
        MqTrace tr(st);
        tr.mark("neighbors");
        MQ3D_TRY(mc_classify(g, weight_threshold, st, &tr));
        MQ3D_TRY(mc_scan(g, st));
        tr.mark("scan");
        // emission is enqueued at once: the kernels read the totals on the device and do nothing if the mesh does not
        // fit -- no host round trip between classification and emission
        k_mc_emit<<<(unsigned)n, EMIT_THREADS, 0, st>>>(g->tsdf, g->block_keys, g->mc_nb, g->mc_emask, g->mc_eprefix, g->mc_counts,
                                                        g->mc_offsets, g->voxel_size, vertices_dev, normals_dev, triangles_dev,
                                                        vertex_keys_dev, n, cap_vertices, cap_triangles);
        if (colors_dev)
            k_mc_colors<<<(unsigned)n, EMIT_THREADS, 0, st>>>(g->tsdf, g->color, g->mc_nb, g->mc_eprefix, g->mc_counts, g->mc_offsets,
                                                              colors_dev, n, cap_vertices, cap_triangles);
        MQ3D_CUDA(cudaGetLastError());
        tr.mark("emit");
        MQ3D_CUDA(cudaMemcpyAsync(g->pinned_host64, g->mc_offsets + 2 * n, sizeof(int64_t) * 2, cudaMemcpyDeviceToHost, st));
        MQ3D_CUDA(cudaStreamSynchronize(st));
        tr.report("extract_mesh");
        g->mc_V = g->pinned_host64[0];
        g->mc_T = g->pinned_host64[1];
    } else {
        g->mc_V = g->mc_T = 0;
    }
    MQ3D_REQUIRE(g->mc_V < 2147483647LL && g->mc_T < 2147483647LL, "mesh too large for int32 indices");
    *n_vertices = g->mc_V;
    *n_triangles = g->mc_T;
    g->mc_weight_thr = weight_threshold;
    g->mc_state = 1;       // classified: mq3d_extract_mesh_fill / _colors may follow (needed when the mesh did not fit)
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
        k_mc_emit<<<(unsigned)g->mc_blocks, EMIT_THREADS, 0, st>>>(g->tsdf, g->block_keys, g->mc_nb, g->mc_emask, g->mc_eprefix,
                                                                   g->mc_counts, g->mc_offsets, g->voxel_size, vertices_dev,
                                                                   normals_dev, triangles_dev, vertex_keys_dev, g->mc_blocks, -1, -1);
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
                                                     g->mc_counts, nullptr, g->voxel_size, nullptr, nullptr, nullptr, nullptr,
                                                     nullptr);
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
                                                               points_dev, normals_dev, point_keys_dev, nullptr, nullptr);
        MQ3D_CUDA(cudaGetLastError());
    }
    return MQ3D_OK;
}

extern "C" int mq3d_extract_mesh_colors(mq3d_grid *g, float *colors_dev, void *stream) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    if (g->mc_state != 1) {
        mq3d_set_error("extract_mesh_colors called without a preceding extract_mesh_count on an unchanged grid");
        return MQ3D_ERR_STATE;
    }
    MQ3D_REQUIRE(g->color != nullptr, "grid has no colour attribute");
    MQ3D_REQUIRE(g->mc_V == 0 || colors_dev, "null colour buffer");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    if (g->mc_blocks > 0 && g->mc_V > 0) {
        k_mc_colors<<<(unsigned)g->mc_blocks, EMIT_THREADS, 0, st>>>(g->tsdf, g->color, g->mc_nb, g->mc_eprefix, g->mc_counts,
                                                                     g->mc_offsets, colors_dev, g->mc_blocks, -1, -1);
        MQ3D_CUDA(cudaGetLastError());
    }
    return MQ3D_OK;
}

extern "C" int mq3d_extract_points_colors(mq3d_grid *g, float *colors_dev, void *stream) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    if (g->mc_state != 2) {
        mq3d_set_error("extract_points_colors called without a preceding extract_points_count on an unchanged grid");
        return MQ3D_ERR_STATE;
    }
    MQ3D_REQUIRE(g->color != nullptr, "grid has no colour attribute");
    MQ3D_REQUIRE(g->mc_V == 0 || colors_dev, "null colour buffer");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    if (g->mc_blocks > 0 && g->mc_V > 0) {
        k_points<true><<<(unsigned)g->mc_blocks, 256, 0, st>>>(g->tsdf, g->weight, g->block_keys, g->mc_nb, g->mc_weight_thr,
                                                               g->part, g->mc_counts, g->mc_offsets, g->voxel_size,
                                                               nullptr, nullptr, nullptr, g->color, colors_dev);
        MQ3D_CUDA(cudaGetLastError());
    }
    return MQ3D_OK;
}
