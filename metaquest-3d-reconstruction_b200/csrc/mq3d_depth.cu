// K1: raw NDC depth -> linear metres, fused with the per-frame validity reduction and the
// confidence / valid-count mask.  One coalesced float4 pass over all frames of a sequence.
// Reference: utils/depth_utils.py:21-46 (convert_depth_to_linear), dataio/depth_data_io.py:80-85
// (is_depth_map_valid), processing/reconstruction/utils/o3d_utils.py:131-142 (mask).
#include "mq3d_common.cuh"

struct NdcParams {
    double x, y;
    int y_is_f64;  // finite far: y is np.float64 and promotes the denominator to f64
    int pad;
};

__device__ __forceinline__ float ndc_to_linear(float raw, const NdcParams &p) {
    float ndc = __fadd_rn(__fmul_rn(raw, 2.0f), -1.0f);
    double den = p.y_is_f64 ? ((double)ndc + p.y) : (double)__fadd_rn(ndc, (float)p.y);
    return den != 0.0 ? (float)(p.x / den) : 0.0f;
}

__device__ __forceinline__ unsigned validity_bits(float v) {
    // bit0 any(!=0)  bit1 any(!=1)  bit2 any(isnan)  bit3 any(!(v>=0))
    return (v != 0.0f ? 1u : 0u) | (v != 1.0f ? 2u : 0u) | (isnan(v) ? 4u : 0u) | (!(v >= 0.0f) ? 8u : 0u);
}

// The per-frame conversion constants travel by value with the launch (constant bank), K1_GROUP frames per
// launch: no device-side parameter scratch, hence no state shared between callers, streams or threads.
#define K1_GROUP 512
struct NdcGroup {
    NdcParams p[K1_GROUP];   // 12 KB
};

// One CTA covers K1_UNROLL * 256 float4 of one frame: every thread has its loads in flight before the first
// float64 division starts (the divisions, ~1.1 ms of FP64 pipe time for 10^4 frames, then overlap the stream).
#define K1_UNROLL 4
template <bool MASK>
__global__ void __launch_bounds__(256)
k_depth_prepare(const float *__restrict__ raw, int64_t px_per_frame, const __grid_constant__ NdcGroup params, int frame0,
                const double *__restrict__ conf, const int32_t *__restrict__ count,
                const uint8_t *__restrict__ has_conf, double conf_thr, int32_t count_thr,
                float *__restrict__ out, int32_t *__restrict__ frame_bits) {
    const int f = frame0 + blockIdx.y;
    const NdcParams p = params.p[blockIdx.y];
    const int64_t base = (int64_t)f * px_per_frame;
    const int64_t n4 = px_per_frame >> 2;
    const bool mask = MASK && (has_conf == nullptr || has_conf[f]);
    unsigned bits = 0;
    const int64_t i0 = (int64_t)blockIdx.x * (256 * K1_UNROLL) + threadIdx.x;
    float4 r[K1_UNROLL];
#pragma unroll
    for (int q = 0; q < K1_UNROLL; ++q) {
        const int64_t i = i0 + 256 * q;
        r[q] = i < n4 ? __ldcs(reinterpret_cast<const float4 *>(raw + base) + i) : make_float4(0.5f, 0.5f, 0.5f, 0.5f);
    }
#pragma unroll
    for (int q = 0; q < K1_UNROLL; ++q) {
        const int64_t i = i0 + 256 * q;
        if (i >= n4) continue;
        bits |= validity_bits(r[q].x) | validity_bits(r[q].y) | validity_bits(r[q].z) | validity_bits(r[q].w);
        float4 o;
        o.x = ndc_to_linear(r[q].x, p);
        o.y = ndc_to_linear(r[q].y, p);
        o.z = ndc_to_linear(r[q].z, p);
        o.w = ndc_to_linear(r[q].w, p);
        if (mask) {
            const double2 *c2 = reinterpret_cast<const double2 *>(conf + base) + 2 * i;
            double2 c01 = __ldg(c2), c23 = __ldg(c2 + 1);
            int4 k = __ldg(reinterpret_cast<const int4 *>(count + base) + i);
            if (c01.x < conf_thr || k.x < count_thr) o.x = 0.0f;
            if (c01.y < conf_thr || k.y < count_thr) o.y = 0.0f;
            if (c23.x < conf_thr || k.z < count_thr) o.z = 0.0f;
            if (c23.y < conf_thr || k.w < count_thr) o.w = 0.0f;
        }
        reinterpret_cast<float4 *>(out + base)[i] = o;
    }
    // scalar tail (px_per_frame not a multiple of 4): first CTA of the frame
    if (blockIdx.x == 0) {
        for (int64_t i = (n4 << 2) + threadIdx.x; i < px_per_frame; i += blockDim.x) {
            float v = raw[base + i];
            bits |= validity_bits(v);
            float o = ndc_to_linear(v, p);
            if (mask && (conf[base + i] < conf_thr || count[base + i] < count_thr)) o = 0.0f;
            out[base + i] = o;
        }
    }
    bits = __reduce_or_sync(0xFFFFFFFFu, bits);
    if ((threadIdx.x & 31) == 0 && bits) atomicOr(&frame_bits[f], (int)bits);
}

__global__ void k_depth_finalize(int32_t *frame_bits, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = frame_bits[i];
    frame_bits[i] = ((b & 1) && (b & 2) && !(b & 4) && !(b & 8)) ? 1 : 0;
}

extern "C" int mq3d_depth_prepare(const float *raw_dev, int n_frames, int width, int height, const double *near_z,
                                  const double *far_z, const double *conf_dev, const int32_t *count_dev,
                                  const uint8_t *has_conf_dev, double conf_thr, int32_t count_thr, float *out_dev,
                                  int32_t *frame_valid_dev, void *stream) {
    MQ3D_REQUIRE(raw_dev && out_dev && frame_valid_dev && near_z && far_z, "null argument");
    MQ3D_REQUIRE(n_frames > 0 && width > 0 && height > 0, "empty input");
    MQ3D_REQUIRE((conf_dev == nullptr) == (count_dev == nullptr), "conf and count must be given together");
    int64_t px = (int64_t)width * height;
    cudaStream_t st = as_stream(stream);
    NdcParams *hp = (NdcParams *)malloc(sizeof(NdcParams) * n_frames);
    for (int i = 0; i < n_frames; ++i) {
        double n = near_z[i], f = far_z[i];
        if (isinf(f) || f < n) {  // depth_utils.py:22-24
            hp[i].x = -2.0 * n;
            hp[i].y = -1.0;
            hp[i].y_is_f64 = 0;
        } else {                  // depth_utils.py:25-27
            hp[i].x = -2.0 * f * n / (f - n);
            hp[i].y = -(f + n) / (f - n);
            hp[i].y_is_f64 = 1;
        }
        hp[i].pad = 0;
    }
    cudaError_t e = cudaMemsetAsync(frame_valid_dev, 0, sizeof(int32_t) * n_frames, st);
    if (e == cudaSuccess) {
        // alignment: float4 path needs 16-B aligned frame bases
        bool aligned = ((px & 3) == 0) && (((uintptr_t)raw_dev | (uintptr_t)out_dev) & 15) == 0 &&
                       (conf_dev == nullptr || ((((uintptr_t)conf_dev | (uintptr_t)count_dev) & 15) == 0));
        if (!aligned) {
            free(hp);
            mq3d_set_error("depth_prepare: buffers must be 16-byte aligned and W*H a multiple of 4");
            return MQ3D_ERR_INVALID;
        }
        int bx = (int)((px / 4 + 256 * K1_UNROLL - 1) / (256 * K1_UNROLL));
        if (bx < 1) bx = 1;
        NdcGroup *grp = (NdcGroup *)calloc(1, sizeof(NdcGroup));
        if (!grp) {
            free(hp);
            mq3d_set_error("out of host memory");
            return MQ3D_ERR_INVALID;
        }
        for (int f0 = 0; f0 < n_frames; f0 += K1_GROUP) {
            const int nf = n_frames - f0 < K1_GROUP ? n_frames - f0 : K1_GROUP;
            memcpy(grp->p, hp + f0, sizeof(NdcParams) * nf);
            dim3 grid(bx, nf);
            if (conf_dev)
                k_depth_prepare<true><<<grid, 256, 0, st>>>(raw_dev, px, *grp, f0, conf_dev, count_dev, has_conf_dev,
                                                            conf_thr, count_thr, out_dev, frame_valid_dev);
            else
                k_depth_prepare<false><<<grid, 256, 0, st>>>(raw_dev, px, *grp, f0, nullptr, nullptr, nullptr, 0.0, 0,
                                                             out_dev, frame_valid_dev);
        }
        free(grp);
        k_depth_finalize<<<(n_frames + 255) / 256, 256, 0, st>>>(frame_valid_dev, n_frames);
        e = cudaGetLastError();
    }
    free(hp);
    if (e != cudaSuccess) {
        mq3d_set_error("depth_prepare: %s", cudaGetErrorString(e));
        return MQ3D_ERR_CUDA;
    }
    return MQ3D_OK;
}
