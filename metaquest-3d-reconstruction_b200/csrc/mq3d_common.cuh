// Shared declarations of the sm_100a hot-path library (see include/mq3d.h for the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mq3d.h"

#define MQ3D_RES 16
#define MQ3D_RES3 4096
#define MQ3D_MAX_BATCH 256
#define MQ3D_COUNTER_WORDS 1024

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
void mq3d_set_error(const char *fmt, ...);

#define MQ3D_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            mq3d_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return MQ3D_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

#define MQ3D_REQUIRE(cond, msg)                              \
    do {                                                     \
        if (!(cond)) {                                       \
            mq3d_set_error("%s (%s)", msg, #cond);           \
            return MQ3D_ERR_INVALID;                         \
        }                                                    \
    } while (0)

#define MQ3D_TRY(call)              \
    do {                            \
        int rc__ = (call);          \
        if (rc__ != MQ3D_OK) return rc__; \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------------
// spatial hash: open addressing, 64-bit packed keys (3 x 21 bit, biased), value = block index
// ------------------------------------------------------------------------------------------------
#define MQ3D_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define MQ3D_KEY_BIAS (1 << 20)

__host__ __device__ __forceinline__ uint64_t mq3d_pack_key(int x, int y, int z) {
    return ((uint64_t)(uint32_t)(x + MQ3D_KEY_BIAS) << 42) | ((uint64_t)(uint32_t)(y + MQ3D_KEY_BIAS) << 21) |
           (uint64_t)(uint32_t)(z + MQ3D_KEY_BIAS);
}
__host__ __device__ __forceinline__ bool mq3d_key_in_range(int x, int y, int z) {
    return x >= -MQ3D_KEY_BIAS && x < MQ3D_KEY_BIAS && y >= -MQ3D_KEY_BIAS && y < MQ3D_KEY_BIAS &&
           z >= -MQ3D_KEY_BIAS && z < MQ3D_KEY_BIAS;
}
__host__ __device__ __forceinline__ void mq3d_unpack_key(uint64_t k, int &x, int &y, int &z) {
    x = (int)((k >> 42) & 0x1FFFFF) - MQ3D_KEY_BIAS;
    y = (int)((k >> 21) & 0x1FFFFF) - MQ3D_KEY_BIAS;
    z = (int)(k & 0x1FFFFF) - MQ3D_KEY_BIAS;
}
__host__ __device__ __forceinline__ uint32_t mq3d_hash64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}

struct HashView {
    unsigned long long *keys;  // [size], MQ3D_EMPTY_KEY when free
    int32_t *vals;             // [size] block index
    uint32_t mask;             // size - 1
};

#ifdef __CUDACC__
#define MQ3D_NO_SLOT 0xFFFFFFFFu
// insert-or-find; returns slot, or MQ3D_NO_SLOT when the table is full (the probe sequence wrapped
// around) -- the host then grows the table and repeats the launch.  `fresh` is set when this call
// claimed the slot.
__device__ __forceinline__ uint32_t hash_insert(const HashView &h, uint64_t key, bool &fresh) {
    uint32_t slot = mq3d_hash64(key) & h.mask;
    fresh = false;
    for (uint32_t probes = 0; probes <= h.mask; ++probes) {
        unsigned long long cur = h.keys[slot];
        if (cur == key) return slot;
        if (cur == MQ3D_EMPTY_KEY) {
            unsigned long long prev = atomicCAS(&h.keys[slot], MQ3D_EMPTY_KEY, (unsigned long long)key);
            if (prev == MQ3D_EMPTY_KEY) {
                fresh = true;
                return slot;
            }
            if (prev == key) return slot;
        }
        slot = (slot + 1) & h.mask;
    }
    return MQ3D_NO_SLOT;
}
// lookup only; returns slot or 0xFFFFFFFF
__device__ __forceinline__ uint32_t hash_find(const HashView &h, uint64_t key) {
    uint32_t slot = mq3d_hash64(key) & h.mask;
    for (uint32_t probes = 0; probes <= h.mask; ++probes) {
        unsigned long long cur = h.keys[slot];
        if (cur == key) return slot;
        if (cur == MQ3D_EMPTY_KEY) return 0xFFFFFFFFu;
        slot = (slot + 1) & h.mask;
    }
    return 0xFFFFFFFFu;
}
#endif

// ------------------------------------------------------------------------------------------------
// multi-GPU partition: owner(tile) = hash(tile) % world; ghosts = 1-block shell
// ------------------------------------------------------------------------------------------------
struct Partition {
    int rank, world, tile_shift;  // tile = 1 << tile_shift blocks per axis
    int integrate_ghosts;         // 1: frames are integrated into owned + ghost blocks (no exchange);
                                  // 0: owned blocks only, ghosts are fetched from their owners before extraction
};

__host__ __device__ __forceinline__ int mq3d_tile_owner(int bx, int by, int bz, const Partition &p) {
    int tx = bx >> p.tile_shift, ty = by >> p.tile_shift, tz = bz >> p.tile_shift;
    return (int)(mq3d_hash64(mq3d_pack_key(tx, ty, tz)) % (uint32_t)p.world);
}
__host__ __device__ __forceinline__ bool mq3d_block_owned(int bx, int by, int bz, const Partition &p) {
    return p.world <= 1 || mq3d_tile_owner(bx, by, bz, p) == p.rank;
}
// blocks a frame is integrated into on this rank
#define MQ3D_INTEGRATES(bx, by, bz, p) \
    ((p).integrate_ghosts ? mq3d_block_needed(bx, by, bz, p) : mq3d_block_owned(bx, by, bz, p))
// block is kept on this rank if it or any of its 26 neighbours lies in an owned tile
__host__ __device__ __forceinline__ bool mq3d_block_needed(int bx, int by, int bz, const Partition &p) {
    if (p.world <= 1) return true;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx)
                if (mq3d_tile_owner(bx + dx, by + dy, bz + dz, p) == p.rank) return true;
    return false;
}

// ------------------------------------------------------------------------------------------------
// camera (Open3D TransformIndexer semantics: float32 storage of float64 inputs)
// ------------------------------------------------------------------------------------------------
struct Camera {
    float fx, fy, cx, cy;
    float e[12];  // 3x4 row-major
};

static inline Camera make_camera(const double K[9], const double E[16]) {
    Camera c;
    c.fx = (float)K[0];
    c.fy = (float)K[4];
    c.cx = (float)K[2];
    c.cy = (float)K[5];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) c.e[i * 4 + j] = (float)E[i * 4 + j];
    return c;
}

// [R^T | -R^T t] in float64 (Open3D t::geometry::InverseTransformation)
static inline void inverse_transformation(const double E[16], double P[16]) {
    for (int i = 0; i < 16; ++i) P[i] = 0.0;
    P[15] = 1.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) P[i * 4 + j] = E[j * 4 + i];
    for (int i = 0; i < 3; ++i) {
        double s = P[i * 4 + 0] * E[3] + P[i * 4 + 1] * E[7] + P[i * 4 + 2] * E[11];
        P[i * 4 + 3] = -s;
    }
}

// device-resident state of one mq3d_integrate_sequence call: batches are enqueued back to back without
// host synchronisation; a batch whose touch overflows the pool / table (or meets a key outside the key
// range) records itself here, every later kernel of the call then returns at once, and the host -- which
// reads this once at the end -- grows the grid and resumes at that batch
struct SeqState {
    int fail_batch;           // -1, else the first batch that failed
    int fail_flags;           // 1 key out of range, 2 hash table full, 4 pool / load factor exceeded
    int first_empty_frame;    // INT_MAX, else lowest valid frame that touched no block on ANY rank's partition
    int frames_integrated;    // valid frames that touched at least one block
    int slow_div_batches;     // batches that held a depth in (0, 2^-75) and took the guarded division
    int pad;
    unsigned long long blocks_loaded;   // sum over batches of the listed slots (block residencies)
};

// per-frame parameters of the fused sequence path (device array)
struct FrameParams {
    Camera touch;   // K + inverse pose, scale 1   (DepthTouch)
    Camera integ;   // K + world->camera, scale = voxel_size (Integrate)
    float cfx, cfy, ccx, ccy;  // colour intrinsics
    int valid;      // host-side validity (frame_valid_dev is checked on device as well)
    int pad[3];
};

// ------------------------------------------------------------------------------------------------
// grid handle
// ------------------------------------------------------------------------------------------------
struct mq3d_grid {
    float voxel_size;
    int attr_mask;
    int device;
    Partition part;
    // hash
    HashView hash;
    int64_t table_size;
    // pool
    int64_t capacity;
    int32_t *block_keys;  // [capacity][3]
    float *tsdf, *weight, *color;
    int *n_blocks_dev;    // device counter (may exceed capacity transiently; host grows the pool)
    int64_t n_blocks_host;  // last synchronised value
    int count_dirty;        // blocks may have been added on the device since n_blocks_host was read
    // per-frame touch scratch ("frustum hashmap")
    HashView frustum;
    int64_t frustum_size;
    int *counter_dev;     // int scratch [MQ3D_COUNTER_WORDS]: counters of touch / list / sort / integrate, misc flags
    // fused-sequence scratch (indexed by hash slot)
    uint32_t *bitmap;     // [table_size][bitmap_words]
    int bitmap_words;
    int *slot_list;       // [table_size] slots touched in the current batch
    uint16_t *slot_cnt;   // [table_size] number of frames of the batch that touched the listed slot
    int *slot_sorted;     // [table_size] same, heavy-first (LPT order for the dynamic scheduler)
    // colour scratch of the fused path: a batch's colour frames resampled onto the depth pixel grid
    uint32_t *rgbx;
    int64_t rgbx_px;
    float *depth_scratch;   // frames / depth_scale when depth_scale != 1
    int64_t depth_scratch_size;
    float *dsan;            // one batch of depth frames with every pixel the integrator rejects replaced by -inf (k_touch)
    int64_t dsan_px;
    // validated fast division by the truncation constant
    float div_checked_trunc;
    int div_fast_ok;
    FrameParams *frame_params_dev;  // [frame_params_cap] (>= MQ3D_MAX_BATCH): the whole sequence of a call
    int64_t frame_params_cap;
    SeqState *seq_dev, *seq_host;   // device state of the running sequence call / pinned mirror
    int *frame_any_dev;             // [MQ3D_MAX_BATCH] frame touched some block (before the partition filter)
    int32_t *idx_scratch;  // per-frame integrate: block index per key
    int64_t idx_scratch_size;
    int *ghost_cnt_dev, *ghost_cnt_host;  // [64] per-destination ghost counts (device / pinned)
    // peer-memory ghost pull (mq3d_peer.cu): once a pool has been exported to other processes its
    // allocations are retired instead of freed on growth (peers may still have them mapped)
    int ipc_exported;
    void **retired;
    int n_retired, cap_retired;
    struct mq3d_peer_state *peer;
    int *pinned_host;    // pinned int[8] for async readbacks
    int64_t *pinned_host64;            // two pinned int64 (tail of pinned_host) for the MC totals
    int *frame_counts_dev;             // [MQ3D_MAX_BATCH]
    unsigned long long *stat_dev;      // [2]
    cudaEvent_t *events;               // persistent timing events of the sequence path
    int n_events;
    // batch gates of the NEXT sequence call (mq3d_grid_set_batch_gates): cudaEvent_t per batch, waited for on the
    // stream before that batch is enqueued; host copy, cleared by the call
    void **gates;
    int n_gates;
    // marching cubes scratch (valid between *_count and *_fill)
    int32_t *mc_nb;       // [n][27]
    uint32_t *mc_rows;    // [n][256] (validity << 16) | sign of every x-row
    uint32_t *mc_emask;   // [n][324] 18-bit sign rows of the (-1..16)^2 neighbourhood
    uint16_t *mc_eprefix; // [n][2304] 16-bit row tables (edge marks, prefixes, surface cubes)
    int32_t *mc_counts;   // [n][2] vertices, triangles (or points)
    int64_t *mc_offsets;  // [n+1][2]
    long long *mc_totals; // [chunks][2] partial sums of the count scan
    int64_t mc_blocks;    // n the scratch was built for
    int64_t mc_alloc_blocks;
    int mc_state;         // 0 none, 1 mesh counted, 2 points counted
    float mc_weight_thr;
    int64_t mc_V, mc_T;
};

#include <stdlib.h>
// MQ3D_TRACE: device time of the phases of a call (CUDA events on the stream), printed to stderr
struct MqTrace {
    cudaStream_t st;
    bool on;
    int n;
    cudaEvent_t ev[8];
    const char *name[8];
    explicit MqTrace(cudaStream_t s) : st(s), on(getenv("MQ3D_TRACE") != nullptr), n(0) { mark("start"); }
    void mark(const char *what) {
        if (!on || n >= 8) return;
        if (cudaEventCreate(&ev[n]) != cudaSuccess) { on = false; return; }
        cudaEventRecord(ev[n], st);
        name[n++] = what;
    }
    void report(const char *call) {     // after a stream synchronisation
        if (!on) return;
        fprintf(stderr, "[mq3d] %s:", call);
        for (int i = 1; i < n; ++i) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            fprintf(stderr, " %s %.3f ms", name[i], ms);
        }
        fprintf(stderr, "\n");
    }
    ~MqTrace() { for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]); }
};

int mq3d_grid_sync_count(mq3d_grid *g, cudaStream_t st);                 // refresh n_blocks_host (synchronises)
int mq3d_grid_fresh_count(mq3d_grid *g, cudaStream_t st);                // same, but only if blocks may have been added since
int mq3d_grid_ensure_capacity(mq3d_grid *g, int64_t need, cudaStream_t st, bool *rehashed);
int mq3d_set_device(int device);
// Activate + Find for an explicit key list; block indices land in g->idx_scratch.
int mq3d_grid_activate(mq3d_grid *g, const int32_t *keys_dev, int64_t n, bool integrating, cudaStream_t st,
                       const int *n_dev = nullptr);
void mq3d_peer_state_free(mq3d_grid *g);   // mq3d_peer.cu
