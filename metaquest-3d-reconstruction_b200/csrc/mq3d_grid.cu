// Voxel block grid state: GPU open-addressing spatial hash + dense block pool.
// Replaces o3d.t.geometry.VoxelBlockGrid construction / growth / save+load payload
// (reference: processing/reconstruction/utils/o3d_utils.py:171-179,
//  dataio/reconstruction_data_io.py:42-55).
#include <stdarg.h>

#include "mq3d_common.cuh"

static thread_local char g_err[512] = "";

void mq3d_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *mq3d_last_error(void) { return g_err; }
extern "C" int mq3d_version(void) { return 100; }

int mq3d_set_device(int device) {
    MQ3D_CUDA(cudaSetDevice(device));
    return MQ3D_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_fill_u64(unsigned long long *p, unsigned long long v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// re-insert every (key,val) of `src` into `dst` (dst pre-filled with EMPTY)
__global__ void k_rehash(HashView src, int64_t src_size, HashView dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= src_size) return;
    unsigned long long k = src.keys[i];
    if (k == MQ3D_EMPTY_KEY) return;
    bool fresh;
    uint32_t s = hash_insert(dst, k, fresh);
    if (s != MQ3D_NO_SLOT) dst.vals[s] = src.vals[i];   // (dst is at most half full: cannot fail)
}

// block_keys[val] = unpack(key) for every occupied slot (after pool growth)
__global__ void k_rebuild_block_keys(HashView h, int64_t size, int32_t *block_keys, int64_t capacity) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    unsigned long long k = h.keys[i];
    if (k == MQ3D_EMPTY_KEY) return;
    int v = h.vals[i];
    if (v < 0 || v >= capacity) return;
    int x, y, z;
    mq3d_unpack_key(k, x, y, z);
    block_keys[3 * (int64_t)v + 0] = x;
    block_keys[3 * (int64_t)v + 1] = y;
    block_keys[3 * (int64_t)v + 2] = z;
}

static int64_t next_pow2(int64_t v) {
    int64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static int alloc_table(HashView *h, int64_t size, cudaStream_t st) {
    MQ3D_CUDA(cudaMalloc(&h->keys, sizeof(unsigned long long) * size));
    MQ3D_CUDA(cudaMalloc(&h->vals, sizeof(int32_t) * size));
    h->mask = (uint32_t)(size - 1);
    k_fill_u64<<<1184, 256, 0, st>>>(h->keys, MQ3D_EMPTY_KEY, size);
    MQ3D_CUDA(cudaMemsetAsync(h->vals, 0xFF, sizeof(int32_t) * size, st));
    MQ3D_CUDA(cudaGetLastError());
    return MQ3D_OK;
}

static int alloc_slot_scratch(mq3d_grid *g, cudaStream_t st) {
    if (g->bitmap) cudaFree(g->bitmap);
    if (g->slot_cnt) cudaFree(g->slot_cnt);
    if (g->slot_list) cudaFree(g->slot_list);
    if (g->slot_sorted) cudaFree(g->slot_sorted);
    g->slot_sorted = nullptr;
    MQ3D_CUDA(cudaMalloc(&g->slot_sorted, sizeof(int) * g->table_size));
    g->bitmap_words = MQ3D_MAX_BATCH / 32;
    MQ3D_CUDA(cudaMalloc(&g->bitmap, sizeof(uint32_t) * g->table_size * g->bitmap_words));
    g->slot_cnt = nullptr;
    MQ3D_CUDA(cudaMalloc(&g->slot_cnt, sizeof(uint16_t) * g->table_size));
    MQ3D_CUDA(cudaMalloc(&g->slot_list, sizeof(int) * g->table_size));
    MQ3D_CUDA(cudaMemsetAsync(g->bitmap, 0, sizeof(uint32_t) * g->table_size * g->bitmap_words, st));
    return MQ3D_OK;
}

// free a pool array -- or park it until destroy when other processes may still have it mapped
static void release_pool_ptr(mq3d_grid *g, void *p) {
    if (!p) return;
    if (!g->ipc_exported) {
        cudaFree(p);
        return;
    }
    if (g->n_retired == g->cap_retired) {
        const int cap = g->cap_retired ? 2 * g->cap_retired : 16;
        void **r = (void **)realloc(g->retired, sizeof(void *) * cap);
        if (!r) return;   // out of host memory: the old array stays mapped for the peers and is leaked
        g->retired = r;
        g->cap_retired = cap;
    }
    g->retired[g->n_retired++] = p;
}

static int alloc_pool(mq3d_grid *g, int64_t cap, int32_t **keys, float **tsdf, float **weight, float **color,
                      cudaStream_t st) {
    // at least one 2 MB page so the key array never shares a driver block with other small allocations
    // (its IPC handle is exported by mq3d_grid_peer_descriptor)
    size_t key_bytes = sizeof(int32_t) * 3 * (size_t)cap;
    MQ3D_CUDA(cudaMalloc(keys, key_bytes < ((size_t)2 << 20) ? ((size_t)2 << 20) : key_bytes));
    MQ3D_CUDA(cudaMalloc(tsdf, sizeof(float) * MQ3D_RES3 * cap));
    MQ3D_CUDA(cudaMalloc(weight, sizeof(float) * MQ3D_RES3 * cap));
    *color = nullptr;
    if (g->attr_mask & MQ3D_ATTR_COLOR) MQ3D_CUDA(cudaMalloc(color, sizeof(float) * 3 * MQ3D_RES3 * cap));
    (void)st;
    return MQ3D_OK;
}

extern "C" int mq3d_grid_create(float voxel_size, int block_resolution, int64_t block_count, int attr_mask,
                                int device, mq3d_grid **out) {
    MQ3D_REQUIRE(out != nullptr, "null output handle");
    MQ3D_REQUIRE(block_resolution == MQ3D_RES, "only block_resolution == 16 is supported");
    MQ3D_REQUIRE(voxel_size > 0.0f, "voxel_size must be positive");
    MQ3D_REQUIRE(block_count > 0, "block_count must be positive");
    MQ3D_REQUIRE(attr_mask & MQ3D_ATTR_TSDF_WEIGHT, "tsdf/weight attributes are required");
    int n_dev = 0;
    MQ3D_CUDA(cudaGetDeviceCount(&n_dev));
    MQ3D_REQUIRE(device >= 0 && device < n_dev, "CUDA device not available (no CPU fallback)");
    MQ3D_TRY(mq3d_set_device(device));
    mq3d_grid *g = new mq3d_grid();
    memset(g, 0, sizeof(*g));
    g->voxel_size = voxel_size;
    g->attr_mask = attr_mask;
    g->device = device;
    g->part.rank = 0;
    g->part.world = 1;
    g->part.tile_shift = 3;
    g->part.integrate_ghosts = 1;
    g->capacity = block_count;
    g->table_size = next_pow2(block_count * 2 < 1024 ? 1024 : block_count * 2);
    cudaStream_t st = 0;
    int rc = alloc_table(&g->hash, g->table_size, st);
    if (rc == MQ3D_OK) rc = alloc_pool(g, g->capacity, &g->block_keys, &g->tsdf, &g->weight, &g->color, st);
    if (rc == MQ3D_OK) {
        rc = [&]() -> int {
            MQ3D_CUDA(cudaMemsetAsync(g->tsdf, 0, sizeof(float) * MQ3D_RES3 * g->capacity, st));
            MQ3D_CUDA(cudaMemsetAsync(g->weight, 0, sizeof(float) * MQ3D_RES3 * g->capacity, st));
            if (g->color) MQ3D_CUDA(cudaMemsetAsync(g->color, 0, sizeof(float) * 3 * MQ3D_RES3 * g->capacity, st));
            MQ3D_CUDA(cudaMalloc(&g->n_blocks_dev, sizeof(int)));
            MQ3D_CUDA(cudaMemsetAsync(g->n_blocks_dev, 0, sizeof(int), st));
            MQ3D_CUDA(cudaMalloc(&g->counter_dev, sizeof(int) * MQ3D_COUNTER_WORDS));
            MQ3D_CUDA(cudaMemsetAsync(g->counter_dev, 0, sizeof(int) * MQ3D_COUNTER_WORDS, st));
            MQ3D_CUDA(cudaMalloc(&g->frame_params_dev, sizeof(FrameParams) * MQ3D_MAX_BATCH));
            g->frame_params_cap = MQ3D_MAX_BATCH;
            MQ3D_CUDA(cudaMalloc(&g->seq_dev, sizeof(SeqState)));
            MQ3D_CUDA(cudaMallocHost(&g->seq_host, sizeof(SeqState)));
            MQ3D_CUDA(cudaMalloc(&g->frame_any_dev, sizeof(int) * MQ3D_MAX_BATCH));
            MQ3D_CUDA(cudaMallocHost(&g->pinned_host, sizeof(int) * 16));
            g->pinned_host64 = reinterpret_cast<int64_t *>(g->pinned_host + 8);
            MQ3D_CUDA(cudaMalloc(&g->frame_counts_dev, sizeof(int) * MQ3D_MAX_BATCH));
            MQ3D_CUDA(cudaMalloc(&g->stat_dev, sizeof(unsigned long long) * 2));
            return MQ3D_OK;
        }();
    }
    if (rc == MQ3D_OK) rc = alloc_slot_scratch(g, st);
    if (rc == MQ3D_OK) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            mq3d_set_error("grid_create sync: %s", cudaGetErrorString(e));
            rc = MQ3D_ERR_CUDA;
        }
    }
    if (rc != MQ3D_OK) {
        mq3d_grid_destroy(g);
        return rc;
    }
    *out = g;
    return MQ3D_OK;
}

static void free_mc(mq3d_grid *g) {
    cudaFree(g->mc_nb);
    cudaFree(g->mc_emask);
    cudaFree(g->mc_eprefix);
    cudaFree(g->mc_counts);
    cudaFree(g->mc_offsets);
    cudaFree(g->mc_totals);
    cudaFree(g->mc_rows);
    g->mc_nb = nullptr;
    g->mc_emask = nullptr;
    g->mc_eprefix = nullptr;
    g->mc_counts = nullptr;
    g->mc_offsets = nullptr;
    g->mc_totals = nullptr;
    g->mc_rows = nullptr;
    g->mc_alloc_blocks = 0;
    g->mc_state = 0;
}

extern "C" int mq3d_grid_destroy(mq3d_grid *g) {
    if (!g) return MQ3D_OK;
    cudaSetDevice(g->device);
    cudaFree(g->hash.keys);
    cudaFree(g->hash.vals);
    cudaFree(g->frustum.keys);
    cudaFree(g->frustum.vals);
    cudaFree(g->block_keys);
    cudaFree(g->tsdf);
    cudaFree(g->weight);
    cudaFree(g->color);
    cudaFree(g->n_blocks_dev);
    cudaFree(g->counter_dev);
    cudaFree(g->bitmap);
    cudaFree(g->slot_cnt);
    cudaFree(g->slot_list);
    cudaFree(g->slot_sorted);
    cudaFree(g->depth_scratch);
    cudaFree(g->dsan);
    cudaFree(g->rgbx);
    cudaFree(g->frame_params_dev);
    cudaFree(g->seq_dev);
    cudaFree(g->frame_any_dev);
    if (g->seq_host) cudaFreeHost(g->seq_host);
    cudaFree(g->idx_scratch);
    cudaFree(g->ghost_cnt_dev);
    mq3d_peer_state_free(g);
    for (int q = 0; q < g->n_retired; ++q) cudaFree(g->retired[q]);
    free(g->retired);
    if (g->ghost_cnt_host) cudaFreeHost(g->ghost_cnt_host);
    if (g->pinned_host) cudaFreeHost(g->pinned_host);
    cudaFree(g->frame_counts_dev);
    cudaFree(g->stat_dev);
    for (int q = 0; q < g->n_events; ++q) cudaEventDestroy(g->events[q]);
    free(g->events);
    free(g->gates);
    free_mc(g);
    delete g;
    return MQ3D_OK;
}

extern "C" int mq3d_grid_info(mq3d_grid *g, float *voxel_size, int *resolution, int64_t *capacity,
                              int *attr_mask, int *device) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    if (voxel_size) *voxel_size = g->voxel_size;
    if (resolution) *resolution = MQ3D_RES;
    if (capacity) *capacity = g->capacity;
    if (attr_mask) *attr_mask = g->attr_mask;
    if (device) *device = g->device;
    return MQ3D_OK;
}

extern "C" int mq3d_grid_set_partition(mq3d_grid *g, int rank, int world, int tile_blocks) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    MQ3D_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
    int shift = 0;
    while ((1 << shift) < tile_blocks) ++shift;
    MQ3D_REQUIRE((1 << shift) == tile_blocks && tile_blocks >= 1, "tile_blocks must be a power of two");
    MQ3D_REQUIRE(g->n_blocks_host == 0, "partition must be set on an empty grid");
    g->part.rank = rank;
    g->part.world = world;
    g->part.tile_shift = shift;
    return MQ3D_OK;
}

extern "C" int mq3d_grid_set_batch_gates(mq3d_grid *g, const void *const *events, int n_events) {
    MQ3D_REQUIRE(g && n_events >= 0 && (n_events == 0 || events), "bad gate list");
    free(g->gates);
    g->gates = nullptr;
    g->n_gates = 0;
    if (n_events == 0) return MQ3D_OK;
    g->gates = (void **)malloc(sizeof(void *) * (size_t)n_events);
    MQ3D_REQUIRE(g->gates != nullptr, "out of host memory");
    memcpy(g->gates, events, sizeof(void *) * (size_t)n_events);
    g->n_gates = n_events;
    return MQ3D_OK;
}

extern "C" int mq3d_grid_set_ghost_mode(mq3d_grid *g, int integrate_ghosts) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    MQ3D_REQUIRE(g->n_blocks_host == 0, "ghost mode must be set on an empty grid");
    g->part.integrate_ghosts = integrate_ghosts ? 1 : 0;
    return MQ3D_OK;
}

// flag the owned blocks that `dest` keeps as ghosts and compact their indices (order arbitrary)
__global__ void k_ghost_select(const int32_t *__restrict__ block_keys, int64_t n, Partition part, int dest,
                               int *__restrict__ count, int32_t *__restrict__ idx_out) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    int x = block_keys[3 * b], y = block_keys[3 * b + 1], z = block_keys[3 * b + 2];
    if (!mq3d_block_owned(x, y, z, part)) return;
    Partition pd = part;
    pd.rank = dest;
    if (!mq3d_block_needed(x, y, z, pd)) return;
    int i = atomicAdd(count, 1);
    if (idx_out) idx_out[i] = (int32_t)b;
}

__global__ void k_ghost_gather(const int32_t *__restrict__ idx, const int32_t *__restrict__ block_keys,
                               const float *__restrict__ tsdf, const float *__restrict__ weight,
                               const float *__restrict__ color, int32_t *__restrict__ keys_out, float *__restrict__ tsdf_out,
                               float *__restrict__ weight_out, float *__restrict__ color_out) {
    const int64_t i = blockIdx.x;
    const int64_t b = idx[i];
    if (threadIdx.x < 3) keys_out[3 * i + threadIdx.x] = block_keys[3 * b + threadIdx.x];
    const float4 *st = reinterpret_cast<const float4 *>(tsdf + b * MQ3D_RES3);
    const float4 *sw = reinterpret_cast<const float4 *>(weight + b * MQ3D_RES3);
    float4 *dt = reinterpret_cast<float4 *>(tsdf_out + i * MQ3D_RES3);
    float4 *dw = reinterpret_cast<float4 *>(weight_out + i * MQ3D_RES3);
    for (int k = threadIdx.x; k < MQ3D_RES3 / 4; k += blockDim.x) {
        dt[k] = st[k];
        dw[k] = sw[k];
    }
    if (color && color_out) {
        const float4 *sc = reinterpret_cast<const float4 *>(color + b * 3 * MQ3D_RES3);
        float4 *dc = reinterpret_cast<float4 *>(color_out + i * 3 * MQ3D_RES3);
        for (int k = threadIdx.x; k < 3 * MQ3D_RES3 / 4; k += blockDim.x) dc[k] = sc[k];
    }
}

// per-destination ghost counts in one pass (counts[d] = owned blocks that rank d keeps as ghosts)
__global__ void k_ghost_counts(const int32_t *__restrict__ block_keys, int64_t n, Partition part, int *__restrict__ counts) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    int x = block_keys[3 * b], y = block_keys[3 * b + 1], z = block_keys[3 * b + 2];
    if (!mq3d_block_owned(x, y, z, part)) return;
    unsigned long long seen = 0;   // destinations already counted for this block (world <= 64)
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                int d = mq3d_tile_owner(x + dx, y + dy, z + dz, part);
                if (d != part.rank && !((seen >> d) & 1ull)) {
                    seen |= 1ull << d;
                    atomicAdd(&counts[d], 1);
                }
            }
}

extern "C" int mq3d_grid_ghost_counts(mq3d_grid *g, int64_t *counts_out, void *stream) {
    MQ3D_REQUIRE(g && counts_out, "null argument");
    MQ3D_REQUIRE(g->part.world <= 64, "ghost exchange supports at most 64 ranks");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    MQ3D_TRY(mq3d_grid_sync_count(g, st));
    const int64_t n = g->n_blocks_host;
    for (int d = 0; d < g->part.world; ++d) counts_out[d] = 0;
    if (n == 0 || g->part.world <= 1) return MQ3D_OK;
    if (!g->ghost_cnt_dev) {
        MQ3D_CUDA(cudaMalloc(&g->ghost_cnt_dev, sizeof(int) * 64));
        MQ3D_CUDA(cudaMallocHost(&g->ghost_cnt_host, sizeof(int) * 64));
    }
    if (n > g->idx_scratch_size) {   // so the per-destination fills can run without further host syncs
        cudaFree(g->idx_scratch);
        g->idx_scratch = nullptr;
        g->idx_scratch_size = 0;
        MQ3D_CUDA(cudaMalloc(&g->idx_scratch, sizeof(int32_t) * next_pow2(n)));
        g->idx_scratch_size = next_pow2(n);
    }
    MQ3D_CUDA(cudaMemsetAsync(g->ghost_cnt_dev, 0, sizeof(int) * 64, st));
    k_ghost_counts<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g->block_keys, n, g->part, g->ghost_cnt_dev);
    MQ3D_CUDA(cudaGetLastError());
    MQ3D_CUDA(cudaMemcpyAsync(g->ghost_cnt_host, g->ghost_cnt_dev, sizeof(int) * 64, cudaMemcpyDeviceToHost, st));
    MQ3D_CUDA(cudaStreamSynchronize(st));
    for (int d = 0; d < g->part.world; ++d) counts_out[d] = g->ghost_cnt_host[d];
    return MQ3D_OK;
}

extern "C" int mq3d_grid_ghost_select(mq3d_grid *g, int dest_rank, int64_t *n_out, int32_t *keys_dev, float *tsdf_dev,
                                      float *weight_dev, float *color_dev, void *stream) {
    MQ3D_REQUIRE(g && n_out, "null argument");
    MQ3D_REQUIRE(dest_rank >= 0 && dest_rank < g->part.world, "bad destination rank");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    const int64_t expected = *n_out;   // > 0 with buffers: count known from mq3d_grid_ghost_counts -> no host sync
    if (keys_dev && expected > 0 && g->n_blocks_host > 0 && dest_rank != g->part.rank &&
        g->n_blocks_host <= g->idx_scratch_size) {
        MQ3D_REQUIRE(tsdf_dev && weight_dev, "null ghost payload buffers");
        const int64_t nb = g->n_blocks_host;
        MQ3D_CUDA(cudaMemsetAsync(g->counter_dev + 5, 0, sizeof(int), st));
        k_ghost_select<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(g->block_keys, nb, g->part, dest_rank,
                                                                     g->counter_dev + 5, g->idx_scratch);
        k_ghost_gather<<<(unsigned)expected, 256, 0, st>>>(g->idx_scratch, g->block_keys, g->tsdf, g->weight, g->color,
                                                           keys_dev, tsdf_dev, weight_dev, color_dev);
        MQ3D_CUDA(cudaGetLastError());
        return MQ3D_OK;
    }
    MQ3D_TRY(mq3d_grid_sync_count(g, st));
    const int64_t n = g->n_blocks_host;
    *n_out = 0;
    if (n == 0 || dest_rank == g->part.rank) return MQ3D_OK;
    if (n > g->idx_scratch_size) {
        cudaFree(g->idx_scratch);
        g->idx_scratch = nullptr;
        g->idx_scratch_size = 0;
        MQ3D_CUDA(cudaMalloc(&g->idx_scratch, sizeof(int32_t) * next_pow2(n)));
        g->idx_scratch_size = next_pow2(n);
    }
    MQ3D_CUDA(cudaMemsetAsync(g->counter_dev + 5, 0, sizeof(int), st));
    k_ghost_select<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g->block_keys, n, g->part, dest_rank, g->counter_dev + 5,
                                                                keys_dev ? g->idx_scratch : nullptr);
    MQ3D_CUDA(cudaGetLastError());
    MQ3D_CUDA(cudaMemcpyAsync(g->pinned_host + 5, g->counter_dev + 5, sizeof(int), cudaMemcpyDeviceToHost, st));
    MQ3D_CUDA(cudaStreamSynchronize(st));
    const int64_t m = g->pinned_host[5];
    *n_out = m;
    if (keys_dev && m > 0) {
        MQ3D_REQUIRE(tsdf_dev && weight_dev, "null ghost payload buffers");
        k_ghost_gather<<<(unsigned)m, 256, 0, st>>>(g->idx_scratch, g->block_keys, g->tsdf, g->weight, g->color, keys_dev,
                                                    tsdf_dev, weight_dev, color_dev);
        MQ3D_CUDA(cudaGetLastError());
    }
    return MQ3D_OK;
}

int mq3d_grid_sync_count(mq3d_grid *g, cudaStream_t st) {
    MQ3D_CUDA(cudaMemcpyAsync(g->pinned_host, g->n_blocks_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    MQ3D_CUDA(cudaStreamSynchronize(st));
    g->n_blocks_host = g->pinned_host[0];
    g->count_dirty = 0;
    return MQ3D_OK;
}

int mq3d_grid_fresh_count(mq3d_grid *g, cudaStream_t st) {
    return g->count_dirty ? mq3d_grid_sync_count(g, st) : MQ3D_OK;
}

extern "C" int mq3d_grid_num_blocks(mq3d_grid *g, int64_t *n, void *stream) {
    MQ3D_REQUIRE(g != nullptr && n != nullptr, "null argument");
    MQ3D_TRY(mq3d_set_device(g->device));
    MQ3D_TRY(mq3d_grid_sync_count(g, as_stream(stream)));
    *n = g->n_blocks_host;
    return MQ3D_OK;
}

// Grow pool (and table) so that `need` blocks fit.  n_blocks_host must be current.
int mq3d_grid_ensure_capacity(mq3d_grid *g, int64_t need, cudaStream_t st, bool *rehashed) {
    if (rehashed) *rehashed = false;
    if (need > g->capacity) {
        int64_t new_cap = g->capacity * 2;
        while (new_cap < need) new_cap *= 2;
        int32_t *nk;
        float *nt, *nw, *nc;
        MQ3D_TRY(alloc_pool(g, new_cap, &nk, &nt, &nw, &nc, st));
        int64_t live = g->n_blocks_host < g->capacity ? g->n_blocks_host : g->capacity;
        MQ3D_CUDA(cudaMemcpyAsync(nt, g->tsdf, sizeof(float) * MQ3D_RES3 * live, cudaMemcpyDeviceToDevice, st));
        MQ3D_CUDA(cudaMemcpyAsync(nw, g->weight, sizeof(float) * MQ3D_RES3 * live, cudaMemcpyDeviceToDevice, st));
        MQ3D_CUDA(cudaMemsetAsync(nt + MQ3D_RES3 * live, 0, sizeof(float) * MQ3D_RES3 * (new_cap - live), st));
        MQ3D_CUDA(cudaMemsetAsync(nw + MQ3D_RES3 * live, 0, sizeof(float) * MQ3D_RES3 * (new_cap - live), st));
        if (nc) {
            MQ3D_CUDA(cudaMemcpyAsync(nc, g->color, sizeof(float) * 3 * MQ3D_RES3 * live, cudaMemcpyDeviceToDevice, st));
            MQ3D_CUDA(cudaMemsetAsync(nc + 3 * MQ3D_RES3 * live, 0, sizeof(float) * 3 * MQ3D_RES3 * (new_cap - live), st));
        }
        MQ3D_CUDA(cudaStreamSynchronize(st));
        release_pool_ptr(g, g->block_keys);
        release_pool_ptr(g, g->tsdf);
        release_pool_ptr(g, g->weight);
        release_pool_ptr(g, g->color);
        g->block_keys = nk;
        g->tsdf = nt;
        g->weight = nw;
        g->color = nc;
        g->capacity = new_cap;
        int64_t grid = (g->table_size + 255) / 256;
        k_rebuild_block_keys<<<(unsigned)grid, 256, 0, st>>>(g->hash, g->table_size, g->block_keys, g->capacity);
        MQ3D_CUDA(cudaGetLastError());
        g->mc_state = 0;
    }
    if (need * 2 > g->table_size) {
        int64_t new_size = g->table_size;
        while (need * 2 > new_size) new_size *= 2;
        HashView nh;
        MQ3D_TRY(alloc_table(&nh, new_size, st));
        int64_t grid = (g->table_size + 255) / 256;
        k_rehash<<<(unsigned)grid, 256, 0, st>>>(g->hash, g->table_size, nh);
        MQ3D_CUDA(cudaGetLastError());
        MQ3D_CUDA(cudaStreamSynchronize(st));
        cudaFree(g->hash.keys);
        cudaFree(g->hash.vals);
        g->hash = nh;
        g->table_size = new_size;
        MQ3D_TRY(alloc_slot_scratch(g, st));
        if (rehashed) *rehashed = true;
    }
    return MQ3D_OK;
}

extern "C" int mq3d_grid_reserve(mq3d_grid *g, int64_t block_count, void *stream) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    MQ3D_TRY(mq3d_set_device(g->device));
    MQ3D_TRY(mq3d_grid_sync_count(g, as_stream(stream)));
    return mq3d_grid_ensure_capacity(g, block_count, as_stream(stream), nullptr);
}

extern "C" int mq3d_grid_reset(mq3d_grid *g, void *stream) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    MQ3D_TRY(mq3d_grid_fresh_count(g, st));
    int64_t live = g->n_blocks_host < g->capacity ? g->n_blocks_host : g->capacity;
    k_fill_u64<<<1184, 256, 0, st>>>(g->hash.keys, MQ3D_EMPTY_KEY, g->table_size);
    MQ3D_CUDA(cudaMemsetAsync(g->hash.vals, 0xFF, sizeof(int32_t) * g->table_size, st));
    MQ3D_CUDA(cudaMemsetAsync(g->tsdf, 0, sizeof(float) * MQ3D_RES3 * live, st));
    MQ3D_CUDA(cudaMemsetAsync(g->weight, 0, sizeof(float) * MQ3D_RES3 * live, st));
    if (g->color) MQ3D_CUDA(cudaMemsetAsync(g->color, 0, sizeof(float) * 3 * MQ3D_RES3 * live, st));
    MQ3D_CUDA(cudaMemsetAsync(g->n_blocks_dev, 0, sizeof(int), st));
    MQ3D_CUDA(cudaMemsetAsync(g->bitmap, 0, sizeof(uint32_t) * g->table_size * g->bitmap_words, st));
    g->n_blocks_host = 0;
    g->count_dirty = 0;
    g->mc_state = 0;
    return MQ3D_OK;
}

extern "C" int mq3d_grid_pool(mq3d_grid *g, int32_t **keys_dev, float **tsdf_dev, float **weight_dev,
                              float **color_dev) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    if (keys_dev) *keys_dev = g->block_keys;
    if (tsdf_dev) *tsdf_dev = g->tsdf;
    if (weight_dev) *weight_dev = g->weight;
    if (color_dev) *color_dev = g->color;
    return MQ3D_OK;
}

extern "C" int mq3d_grid_export(mq3d_grid *g, int32_t *keys_dev, float *tsdf_dev, float *weight_dev,
                                float *color_dev, void *stream) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    MQ3D_TRY(mq3d_grid_sync_count(g, st));
    int64_t n = g->n_blocks_host;
    if (n == 0) return MQ3D_OK;
    const cudaMemcpyKind d2d = cudaMemcpyDeviceToDevice;
    if (keys_dev) MQ3D_CUDA(cudaMemcpyAsync(keys_dev, g->block_keys, sizeof(int32_t) * 3 * n, d2d, st));
    if (tsdf_dev) MQ3D_CUDA(cudaMemcpyAsync(tsdf_dev, g->tsdf, sizeof(float) * MQ3D_RES3 * n, d2d, st));
    if (weight_dev) MQ3D_CUDA(cudaMemcpyAsync(weight_dev, g->weight, sizeof(float) * MQ3D_RES3 * n, d2d, st));
    if (color_dev && g->color) MQ3D_CUDA(cudaMemcpyAsync(color_dev, g->color, sizeof(float) * 3 * MQ3D_RES3 * n, d2d, st));
    MQ3D_CUDA(cudaStreamSynchronize(st));
    return MQ3D_OK;
}

// ------------------------------------------------------------------------------------------------
// activation of explicit key lists (per-frame integrate, import)
// ------------------------------------------------------------------------------------------------
// integrating = true : keys of a frame about to be integrated -> blocks this rank integrates
// integrating = false: import (VoxelBlockGrid.load, ghost exchange)  -> any block this rank keeps
__global__ void k_activate_keys(HashView h, const int32_t *keys, int64_t n, const int *__restrict__ n_dev, int *n_blocks,
                                int32_t *block_keys, int64_t capacity, Partition part, bool integrating, int *bad_key_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev && i >= *n_dev)) return;
    int x = keys[3 * i], y = keys[3 * i + 1], z = keys[3 * i + 2];
    if (!mq3d_key_in_range(x, y, z)) {
        *bad_key_flag = 1;
        return;
    }
    if (integrating ? !MQ3D_INTEGRATES(x, y, z, part) : !mq3d_block_needed(x, y, z, part)) return;
    bool fresh;
    uint32_t s = hash_insert(h, mq3d_pack_key(x, y, z), fresh);
    if (s == MQ3D_NO_SLOT) {   // cannot happen: the host reserved room for every key beforehand
        *bad_key_flag = 2;
        return;
    }
    if (fresh) {
        int b = atomicAdd(n_blocks, 1);
        h.vals[s] = b;
        if (b < capacity) {
            block_keys[3 * (int64_t)b] = x;
            block_keys[3 * (int64_t)b + 1] = y;
            block_keys[3 * (int64_t)b + 2] = z;
        }
    }
}

__global__ void k_find_keys(HashView h, const int32_t *keys, int64_t n, const int *__restrict__ n_dev, Partition part,
                            bool integrating, int32_t *idx_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev && i >= *n_dev)) return;
    int x = keys[3 * i], y = keys[3 * i + 1], z = keys[3 * i + 2];
    int32_t r = -1;
    if (mq3d_key_in_range(x, y, z) && (!integrating || MQ3D_INTEGRATES(x, y, z, part))) {
        uint32_t s = hash_find(h, mq3d_pack_key(x, y, z));
        if (s != 0xFFFFFFFFu) r = h.vals[s];
    }
    idx_out[i] = r;
}

// Activate + Find (Open3D Integrate preamble).  Leaves block indices in g->idx_scratch.  n_dev (optional): the
// number of keys lives on the device (at most n); nothing is read back.
int mq3d_grid_activate(mq3d_grid *g, const int32_t *keys_dev, int64_t n, bool integrating, cudaStream_t st, const int *n_dev) {
    if (n > g->idx_scratch_size) {
        cudaFree(g->idx_scratch);
        g->idx_scratch = nullptr;
        int64_t sz = next_pow2(n);
        MQ3D_CUDA(cudaMalloc(&g->idx_scratch, sizeof(int32_t) * sz));
        g->idx_scratch_size = sz;
    }
    // worst case every key is new: make room first so indices never exceed the pool
    MQ3D_TRY(mq3d_grid_fresh_count(g, st));
    bool rehashed;
    MQ3D_TRY(mq3d_grid_ensure_capacity(g, g->n_blocks_host + n, st, &rehashed));
    MQ3D_CUDA(cudaMemsetAsync(g->counter_dev, 0, sizeof(int), st));
    unsigned grid = (unsigned)((n + 255) / 256);
    k_activate_keys<<<grid, 256, 0, st>>>(g->hash, keys_dev, n, n_dev, g->n_blocks_dev, g->block_keys, g->capacity,
                                          g->part, integrating, g->counter_dev);
    k_find_keys<<<grid, 256, 0, st>>>(g->hash, keys_dev, n, n_dev, g->part, integrating, g->idx_scratch);
    MQ3D_CUDA(cudaGetLastError());
    g->count_dirty = 1;
    g->mc_state = 0;
    return MQ3D_OK;
}

__global__ void k_import_values(const int32_t *idx, int64_t n, const float *src_t, const float *src_w,
                                const float *src_c, float *tsdf, float *weight, float *color) {
    int64_t i = blockIdx.x;
    if (i >= n) return;
    int b = idx[i];
    if (b < 0) return;
    const float4 *st4 = reinterpret_cast<const float4 *>(src_t + i * MQ3D_RES3);
    const float4 *sw4 = reinterpret_cast<const float4 *>(src_w + i * MQ3D_RES3);
    float4 *dt4 = reinterpret_cast<float4 *>(tsdf + (int64_t)b * MQ3D_RES3);
    float4 *dw4 = reinterpret_cast<float4 *>(weight + (int64_t)b * MQ3D_RES3);
    for (int k = threadIdx.x; k < MQ3D_RES3 / 4; k += blockDim.x) {
        dt4[k] = st4[k];
        dw4[k] = sw4[k];
    }
    if (color && src_c) {
        const float4 *sc4 = reinterpret_cast<const float4 *>(src_c + i * 3 * MQ3D_RES3);
        float4 *dc4 = reinterpret_cast<float4 *>(color + (int64_t)b * 3 * MQ3D_RES3);
        for (int k = threadIdx.x; k < 3 * MQ3D_RES3 / 4; k += blockDim.x) dc4[k] = sc4[k];
    }
}

extern "C" int mq3d_grid_import(mq3d_grid *g, const int32_t *keys_dev, const float *tsdf_dev,
                                const float *weight_dev, const float *color_dev, int64_t n, void *stream) {
    MQ3D_REQUIRE(g != nullptr, "null grid");
    MQ3D_REQUIRE(n >= 0, "negative block count");
    if (n == 0) return MQ3D_OK;
    MQ3D_REQUIRE(keys_dev && tsdf_dev && weight_dev, "null block arrays");
    MQ3D_TRY(mq3d_set_device(g->device));
    cudaStream_t st = as_stream(stream);
    MQ3D_TRY(mq3d_grid_activate(g, keys_dev, n, /*integrating=*/false, st));
    k_import_values<<<(unsigned)n, 256, 0, st>>>(g->idx_scratch, n, tsdf_dev, weight_dev, color_dev, g->tsdf,
                                                  g->weight, g->color);
    MQ3D_CUDA(cudaGetLastError());
    MQ3D_TRY(mq3d_grid_sync_count(g, st));
    return MQ3D_OK;
}
