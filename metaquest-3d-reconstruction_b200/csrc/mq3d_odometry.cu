// N4: odometry information matrix between two depth frames.
//
// Replaces o3d.t.pipelines.odometry.compute_odometry_information_matrix(source_depth, target_depth, intrinsic,
// source_to_target, dist_threshold, depth_scale, depth_max) as called per frame pair (and per key-frame pair) by
// build_pose_graph_for_fragment (processing/reconstruction/depth_optimization/make_fragments.py:142-150,228-233);
// the pose-graph optimisation that consumes the edges stays out of scope.
//
// Open3D 0.19 semantics (t/pipelines/odometry/RGBDOdometry.cpp, kernel/RGBDOdometryJacobianImpl.h; recalled, the
// source is not in the reference tree -- parity unpinned, pinned to the oracle restatement only):
//   depth' = depth / depth_scale, NaN where depth' <= 0 or depth' >= depth_max     (ClipTransform(scale, 0, max, NaN))
//   vertex map: ((u - cx) d / fx, (v - cy) d / fy, d) in float32, NaN where depth' is NaN
//   for every source pixel with a valid vertex p:  q = R p + t (float32);  (u, v) = round(project(q));
//     skip if q.z < 0 or (u, v) outside the image (0 <= u <= W-1, 0 <= v <= H-1) or the target vertex at (u, v) is NaN
//     or |q - target|^2 > dist_threshold^2;
//     J_x = (0, q.z, -q.y, 1, 0, 0), J_y = (-q.z, 0, q.x, 0, 1, 0), J_z = (q.y, -q.x, 0, 0, 0, 1)
//     A += J_x^T J_x + J_y^T J_y + J_z^T J_z                                        (6 x 6, symmetric)
// Open3D sums A in float32 per thread in an unspecified order; here the 21 unique entries are summed in float64
// (warp shuffles, one atomicAdd per CTA and entry), so results agree with the CPU path to summation-order accuracy.
#include "mq3d_common.cuh"

struct OdoConsts {
    float fx, fy, cx, cy;
    float r[9], t[3];
    float scale, depth_max, dist2;
    int W, H;
};

__device__ __forceinline__ bool odo_vertex(const float *__restrict__ depth, int x, int y, const OdoConsts &k, float v[3]) {
    const float d = __fdiv_rn(depth[(int64_t)y * k.W + x], k.scale);
    if (!(d > 0.0f) || !(d < k.depth_max)) return false;      // clipped to NaN (also catches a NaN depth)
    v[0] = __fdiv_rn(__fmul_rn(__fsub_rn((float)x, k.cx), d), k.fx);
    v[1] = __fdiv_rn(__fmul_rn(__fsub_rn((float)y, k.cy), d), k.fy);
    v[2] = d;
    return true;
}

__global__ void __launch_bounds__(256)
k_odometry_information(const float *__restrict__ src, const float *__restrict__ tgt, OdoConsts k, double *__restrict__ acc /* [21] */) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double a[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) a[i] = 0.0;
    if (p < k.W * k.H) {
        const int x = p % k.W, y = p / k.W;
        float s[3];
        if (odo_vertex(src, x, y, k, s)) {
            float q[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                q[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(k.r[3 * i], s[0]), __fmul_rn(k.r[3 * i + 1], s[1])),
                                           __fmul_rn(k.r[3 * i + 2], s[2])), k.t[i]);
            const float inv_z = __fdiv_rn(1.0f, q[2]);
            const float u = roundf(__fadd_rn(__fmul_rn(__fmul_rn(k.fx, q[0]), inv_z), k.cx));
            const float v = roundf(__fadd_rn(__fmul_rn(__fmul_rn(k.fy, q[1]), inv_z), k.cy));
            if (!(q[2] < 0.0f) && u >= 0.0f && v >= 0.0f && u <= (float)k.W - 1.0f && v <= (float)k.H - 1.0f) {
                float tv[3];
                if (odo_vertex(tgt, (int)u, (int)v, k, tv)) {
                    const float rx = __fsub_rn(q[0], tv[0]), ry = __fsub_rn(q[1], tv[1]), rz = __fsub_rn(q[2], tv[2]);
                    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
                    if (r2 <= k.dist2) {
                        const float J[3][6] = {{0.0f, q[2], -q[1], 1.0f, 0.0f, 0.0f},
                                               {-q[2], 0.0f, q[0], 0.0f, 1.0f, 0.0f},
                                               {q[1], -q[0], 0.0f, 0.0f, 0.0f, 1.0f}};
                        int o = 0;
#pragma unroll
                        for (int i = 0; i < 6; ++i)
#pragma unroll
                            for (int j = 0; j <= i; ++j, ++o)
                                a[o] = (double)__fadd_rn(__fadd_rn(__fmul_rn(J[0][i], J[0][j]), __fmul_rn(J[1][i], J[1][j])),
                                                         __fmul_rn(J[2][i], J[2][j]));
                    }
                }
            }
        }
    }
    __shared__ double s_red[8][21];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 21; ++i) {
        double v = a[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 21) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += s_red[w][threadIdx.x];
        if (v != 0.0) atomicAdd(&acc[threadIdx.x], v);
    }
}

extern "C" int mq3d_odometry_information(const float *source_depth_dev, const float *target_depth_dev, int width, int height,
                                         const double K[9], const double source_to_target[16], float dist_threshold,
                                         float depth_scale, float depth_max, double info_out[36], int device, void *stream) {
    MQ3D_REQUIRE(source_depth_dev && target_depth_dev && K && source_to_target && info_out, "null argument");
    MQ3D_REQUIRE(width > 0 && height > 0, "empty depth image");
    MQ3D_TRY(mq3d_set_device(device));
    cudaStream_t st = as_stream(stream);
    OdoConsts k;
    k.fx = (float)K[0];
    k.fy = (float)K[4];
    k.cx = (float)K[2];
    k.cy = (float)K[5];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) k.r[3 * i + j] = (float)source_to_target[4 * i + j];
        k.t[i] = (float)source_to_target[4 * i + 3];
    }
    k.scale = depth_scale;
    k.depth_max = depth_max;
    k.dist2 = dist_threshold * dist_threshold;
    k.W = width;
    k.H = height;
    double *acc = nullptr;
    MQ3D_CUDA(cudaMalloc(&acc, sizeof(double) * 21));
    double h[21];
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * 21, st);
    if (e == cudaSuccess) {
        k_odometry_information<<<(unsigned)(((int64_t)width * height + 255) / 256), 256, 0, st>>>(source_depth_dev, target_depth_dev, k, acc);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, acc, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(acc);
    if (e != cudaSuccess) {
        mq3d_set_error("odometry_information: %s", cudaGetErrorString(e));
        return MQ3D_ERR_CUDA;
    }
    int o = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j <= i; ++j, ++o) info_out[6 * i + j] = info_out[6 * j + i] = h[o];
    return MQ3D_OK;
}
