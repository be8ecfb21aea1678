// K4: multi-view depth confidence by reprojection against +-R neighbour frames.
//
// Restates build_confidence_map / compute_pixel_error_map / bilinear_interpolate_depth
// (reference: processing/reconstruction/confidence_estimation/estimate_depth_confidences.py:15-79,
//  compute_pixel_error_map.py:4-220) per reference pixel: float64 arithmetic on float32 inputs in
// the NumPy operation order (NEP-50 promotions: f32 images compared against Python floats in f32,
// int64 pixel coordinates minus np.float32 intrinsics promote to f64).  One launch per side: a thread
// owns one reference pixel and loops over the 2R targets, whose depth images stay L2-resident
// (sliding window of 2R+1 frames x 410 KB); outputs stay on the device for K1's mask.
#include "mq3d_common.cuh"

struct ConfFrame {
    double fx, fy, cx, cy;
    double ecw[12];   // camera -> world (rows 0..2)
    double einv[12];  // float32 inverse of the above (world -> camera)
};

__device__ __forceinline__ void xform(const double *e, double x, double y, double z, double *o) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
        o[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(e[4 * i], x), __dmul_rn(e[4 * i + 1], y)), __dmul_rn(e[4 * i + 2], z)),
                         __dmul_rn(e[4 * i + 3], 1.0));
}

__global__ void __launch_bounds__(256)
k_confidence(const float *__restrict__ depths, const int32_t *__restrict__ frame_valid, int N, int W, int H,
             const ConfFrame *__restrict__ frames, int range, double depth_max, float depth_max_f, float err_thr_f,
             double *__restrict__ conf, int32_t *__restrict__ count) {
    const int r = blockIdx.y;
    if (frame_valid && !frame_valid[r]) return;  // reference writes no map for an unreadable frame
    const int64_t npx = (int64_t)W * H;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    const int u = (int)(p % W), v = (int)(p / W);
    const float *ref = depths + (int64_t)r * npx;
    const ConfFrame &R = frames[r];
    const float d = ref[p];
    int valid_n = 0, cons_n = 0;
    if (d > 0.0f && d <= depth_max_f) {
        const double z = (double)d;
        const double x = __ddiv_rn(__dmul_rn(__dsub_rn((double)u, R.cx), z), R.fx);
        const double y = __ddiv_rn(__dmul_rn(__dsub_rn((double)v, R.cy), z), R.fy);
        double pw[3];
        xform(R.ecw, x, y, z, pw);
        const double max_coord = (double)((W > H ? W : H) * 10);
        int lo = r - range < 0 ? 0 : r - range;
        int hi = r + range + 1 > N ? N : r + range + 1;
        for (int t = lo; t < hi; ++t) {
            if (t == r) continue;
            if (frame_valid && !frame_valid[t]) continue;
            const ConfFrame &T = frames[t];
            double pt[3];
            xform(T.einv, pw[0], pw[1], pw[2], pt);
            if (!(pt[2] > 0.0 && isfinite(pt[2]) && pt[2] <= depth_max && isfinite(pt[0]) && isfinite(pt[1]))) continue;
            const double uu = __dadd_rn(__ddiv_rn(__dmul_rn(pt[0], T.fx), pt[2]), T.cx);
            const double vv = __dadd_rn(__ddiv_rn(__dmul_rn(pt[1], T.fy), pt[2]), T.cy);
            if (!(isfinite(uu) && isfinite(vv))) continue;
            if (!(uu >= -max_coord && uu < max_coord && vv >= -max_coord && vv < max_coord)) continue;
            const int u0 = (int)floor(uu), v0 = (int)floor(vv);
            const int u1 = u0 + 1, v1 = v0 + 1;
            if (!(u0 >= 0 && u1 < W && v0 >= 0 && v1 < H)) continue;
            const float *tg = depths + (int64_t)t * npx;
            const float Ia = __ldg(tg + (int64_t)v0 * W + u0), Ib = __ldg(tg + (int64_t)v0 * W + u1);
            const float Ic = __ldg(tg + (int64_t)v1 * W + u0), Id = __ldg(tg + (int64_t)v1 * W + u1);
            if (!(Ib > 0.0f && Ib <= depth_max_f && Ia > 0.0f && Ia <= depth_max_f && Ic > 0.0f && Ic <= depth_max_f &&
                  Id > 0.0f && Id <= depth_max_f))
                continue;
            const double du1 = __dsub_rn((double)u1, uu), du0 = __dsub_rn(uu, (double)u0);
            const double dv1 = __dsub_rn((double)v1, vv), dv0 = __dsub_rn(vv, (double)v0);
            const double wa = __dmul_rn(du1, dv1), wb = __dmul_rn(du0, dv1), wc = __dmul_rn(du1, dv0), wd = __dmul_rn(du0, dv0);
            const double zi = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(wa, (double)Ia), __dmul_rn(wb, (double)Ib)),
                                                  __dmul_rn(wc, (double)Ic)), __dmul_rn(wd, (double)Id));
            const float zt = (float)zi;
            if (!(zt > 0.0f && isfinite(zt))) continue;
            const double xt = __ddiv_rn(__dmul_rn(__dsub_rn(uu, T.cx), (double)zt), T.fx);
            const double yt = __ddiv_rn(__dmul_rn(__dsub_rn(vv, T.cy), (double)zt), T.fy);
            double tw[3];
            xform(T.ecw, xt, yt, (double)zt, tw);
            const double dx = __dsub_rn(pw[0], tw[0]), dy = __dsub_rn(pw[1], tw[1]), dz = __dsub_rn(pw[2], tw[2]);
            const float err = (float)__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            if (isnan(err)) continue;
            valid_n += 1;
            cons_n += (err <= err_thr_f) ? 1 : 0;
        }
    }
    count[(int64_t)r * npx + p] = valid_n;
    conf[(int64_t)r * npx + p] = valid_n == 0 ? 0.0 : __ddiv_rn((double)cons_n, (double)valid_n);
}

extern "C" int mq3d_confidence(const float *depths_dev, const int32_t *frame_valid_dev, int n_frames, int width,
                               int height, const float *K, const float *Ecw, const float *Ecw_inv,
                               int target_frame_range, double depth_max, double error_threshold, double *conf_dev,
                               int32_t *count_dev, void *stream) {
    MQ3D_REQUIRE(depths_dev && K && Ecw && Ecw_inv && conf_dev && count_dev, "null argument");
    MQ3D_REQUIRE(n_frames > 0 && n_frames <= 65535 && width > 0 && height > 0, "bad sequence geometry");
    MQ3D_REQUIRE(target_frame_range >= 0, "negative frame range");
    cudaStream_t st = as_stream(stream);
    ConfFrame *hf = (ConfFrame *)malloc(sizeof(ConfFrame) * n_frames);
    for (int i = 0; i < n_frames; ++i) {
        hf[i].fx = (double)K[9 * i + 0];
        hf[i].fy = (double)K[9 * i + 4];
        hf[i].cx = (double)K[9 * i + 2];
        hf[i].cy = (double)K[9 * i + 5];
        for (int j = 0; j < 12; ++j) {
            hf[i].ecw[j] = (double)Ecw[16 * i + j];
            hf[i].einv[j] = (double)Ecw_inv[16 * i + j];
        }
    }
    ConfFrame *df = nullptr;
    cudaError_t e = cudaMalloc(&df, sizeof(ConfFrame) * n_frames);
    if (e == cudaSuccess) e = cudaMemcpyAsync(df, hf, sizeof(ConfFrame) * n_frames, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        int64_t npx = (int64_t)width * height;
        dim3 grid((unsigned)((npx + 255) / 256), n_frames);
        k_confidence<<<grid, 256, 0, st>>>(depths_dev, frame_valid_dev, n_frames, width, height, df, target_frame_range,
                                           depth_max, (float)depth_max, (float)error_threshold, conf_dev, count_dev);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    free(hf);
    cudaFree(df);
    if (e != cudaSuccess) {
        mq3d_set_error("confidence: %s", cudaGetErrorString(e));
        return MQ3D_ERR_CUDA;
    }
    return MQ3D_OK;
}
