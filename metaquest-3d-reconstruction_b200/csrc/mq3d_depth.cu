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

template <bool MASK>
__global__ void __launch_bounds__(256)
k_depth_prepare(const float *__restrict__ raw, int64_t px_per_frame, const NdcParams *__restrict__ params,
                const double *__restrict__ conf, const int32_t *__restrict__ count,
                const uint8_t *__restrict__ has_conf, double conf_thr, int32_t count_thr,
                float *__restrict__ out, int32_t *__restrict__ frame_bits) {
    int f = blockIdx.y;
    NdcParams p = params[f];
    int64_t base = (int64_t)f * px_per_frame;
    int64_t n4 = px_per_frame >> 2;
    bool mask = MASK && (has_conf == nullptr || has_conf[f]);
    unsigned bits = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 r = __ldg(reinterpret_cast<const float4 *>(raw + base) + i);
        bits |= validity_bits(r.x) | validity_bits(r.y) | validity_bits(r.z) | validity_bits(r.w);
        float4 o;
        o.x = ndc_to_linear(r.x, p);
        o.y = ndc_to_linear(r.y, p);
        o.z = ndc_to_linear(r.z, p);
        o.w = ndc_to_linear(r.w, p);
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
    // scalar tail (px_per_frame not a multiple of 4)
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px_per_frame;
         i += (int64_t)gridDim.x * blockDim.x) {
        float r = raw[base + i];
        bits |= validity_bits(r);
        float o = ndc_to_linear(r, p);
        if (mask && (conf[base + i] < conf_thr || count[base + i] < count_thr)) o = 0.0f;
        out[base + i] = o;
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
    MQ3D_REQUIRE(n_frames <= 65535, "at most 65535 frames per call");
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
    // per-device parameter scratch, kept across calls (one stream per device by contract)
    static NdcParams *s_dp[64] = {nullptr};
    static int s_dp_cap[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && dev < 64 && n_frames > s_dp_cap[dev]) {
        cudaFree(s_dp[dev]);
        s_dp[dev] = nullptr;
        s_dp_cap[dev] = 0;
        e = cudaMalloc(&s_dp[dev], sizeof(NdcParams) * n_frames);
        if (e == cudaSuccess) s_dp_cap[dev] = n_frames;
    }
    NdcParams *dp = (e == cudaSuccess && dev < 64) ? s_dp[dev] : nullptr;
    if (e == cudaSuccess && dp == nullptr) e = cudaErrorInvalidDevice;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dp, hp, sizeof(NdcParams) * n_frames, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(frame_valid_dev, 0, sizeof(int32_t) * n_frames, st);
    if (e == cudaSuccess) {
        // alignment: float4 path needs 16-B aligned frame bases
        bool aligned = ((px & 3) == 0) && (((uintptr_t)raw_dev | (uintptr_t)out_dev) & 15) == 0 &&
                       (conf_dev == nullptr || ((((uintptr_t)conf_dev | (uintptr_t)count_dev) & 15) == 0));
        if (!aligned) {
            free(hp);
            mq3d_set_error("depth_prepare: buffers must be 16-byte aligned and W*H a multiple of 4");
            return MQ3D_ERR_INVALID;
        }
        int bx = (int)((px / 4 + 255) / 256);
        if (bx > 148 * 4) bx = 148 * 4;
        dim3 grid(bx, n_frames);
        if (conf_dev)
            k_depth_prepare<true><<<grid, 256, 0, st>>>(raw_dev, px, dp, conf_dev, count_dev, has_conf_dev, conf_thr,
                                                        count_thr, out_dev, frame_valid_dev);
        else
            k_depth_prepare<false><<<grid, 256, 0, st>>>(raw_dev, px, dp, nullptr, nullptr, nullptr, 0.0, 0, out_dev,
                                                         frame_valid_dev);
        k_depth_finalize<<<(n_frames + 255) / 256, 256, 0, st>>>(frame_valid_dev, n_frames);
        e = cudaGetLastError();
    }
    free(hp);  // pageable H2D copies are staged before cudaMemcpyAsync returns
    if (e != cudaSuccess) {
        mq3d_set_error("depth_prepare: %s", cudaGetErrorString(e));
        return MQ3D_ERR_CUDA;
    }
    return MQ3D_OK;
}
