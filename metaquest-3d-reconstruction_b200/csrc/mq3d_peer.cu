// Multi-GPU ghost shell by peer memory: instead of packing ghost blocks, moving them with NCCL
// send/recv and importing them, every rank maps its peers' block pools (CUDA IPC over NVLink / NVSwitch) and
// PULLS the blocks it needs with one copy kernel -- owner pool -> own pool, no staging buffers.
//
//   mq3d_grid_peer_descriptor : 512-byte description of this rank's pool (IPC handles + block count); the
//                               host all-gathers the descriptors (that collective is also the "every rank
//                               has finished integrating" barrier, as it is stream-ordered after K3)
//   mq3d_grid_ghost_pull      : k_ghost_scan walks the peers' key arrays over NVLink and lists the blocks
//                               (owned by the peer, inside this rank's one-block shell); the list is
//                               activated in the local hash; k_ghost_copy moves tsdf | weight | colour
//                               straight from the owners' pools.
// The caller fences afterwards (any stream-ordered collective) before a pool may change again.
// Values are the owners', so the result is bit-identical to redundant ghost integration.
#include <unistd.h>

#include "mq3d_common.cuh"

#define MQ3D_PEER_MAGIC 0x5033514du   // "MQ3P"
#define MQ3D_MAX_PEERS 64

struct PeerDesc {                      // host bytes, MQ3D_PEER_DESC_BYTES
    uint32_t magic;
    int32_t rank, world, device, has_color, pad;
    int64_t pid, n_blocks, capacity;
    uint64_t ptr[4];                   // block_keys, tsdf, weight, color: valid inside process `pid`
    cudaIpcMemHandle_t handle[4];
};
static_assert(sizeof(PeerDesc) <= MQ3D_PEER_DESC_BYTES, "descriptor does not fit");
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");

struct PeerView {                      // device-side view of one peer's pool
    const int32_t *keys;
    const float *tsdf, *weight, *color;
    int64_t n_blocks;
};

struct PeerMap {                       // one opened peer allocation
    cudaIpcMemHandle_t handle;
    void *ptr;
};

struct mq3d_peer_state {
    cudaIpcMemHandle_t own_handle[4];  // cache of this grid's exported handles
    void *own_ptr[4];
    PeerMap map[MQ3D_MAX_PEERS][4];
    PeerView *views_host, *views_dev;  // [MQ3D_MAX_PEERS]
    int32_t *pull_keys;                // [pull_cap][3]
    int2 *pull_src;                    // [pull_cap] (peer, block index in the peer's pool)
    int64_t pull_cap;
    int peer_access[MQ3D_MAX_PEERS];   // same-process grids on other devices: peer access enabled
};

static int peer_state(mq3d_grid *g) {
    if (g->peer) return MQ3D_OK;
    mq3d_peer_state *p = (mq3d_peer_state *)calloc(1, sizeof(mq3d_peer_state));
    MQ3D_REQUIRE(p != nullptr, "out of host memory");
    g->peer = p;
    MQ3D_CUDA(cudaMallocHost(&p->views_host, sizeof(PeerView) * MQ3D_MAX_PEERS));
    MQ3D_CUDA(cudaMalloc(&p->views_dev, sizeof(PeerView) * MQ3D_MAX_PEERS));
    return MQ3D_OK;
}

void mq3d_peer_state_free(mq3d_grid *g) {
    mq3d_peer_state *p = g->peer;
    if (!p) return;
    for (int d = 0; d < MQ3D_MAX_PEERS; ++d)
        for (int a = 0; a < 4; ++a)
            if (p->map[d][a].ptr) cudaIpcCloseMemHandle(p->map[d][a].ptr);
    if (p->views_host) cudaFreeHost(p->views_host);
    cudaFree(p->views_dev);
    cudaFree(p->pull_keys);
    cudaFree(p->pull_src);
    free(p);
    g->peer = nullptr;
}

extern "C" int mq3d_grid_peer_descriptor(mq3d_grid *g, void *desc_out, void *stream) {
    MQ3D_REQUIRE(g && desc_out, "null argument");
    MQ3D_REQUIRE(g->part.world <= MQ3D_MAX_PEERS, "peer ghost pull supports at most 64 ranks");
    MQ3D_TRY(mq3d_set_device(g->device));
    MQ3D_TRY(peer_state(g));
    MQ3D_TRY(mq3d_grid_fresh_count(g, as_stream(stream)));     // no synchronisation right after a sequence call
    mq3d_peer_state *p = g->peer;
    PeerDesc d;
    memset(&d, 0, sizeof(d));
    d.magic = MQ3D_PEER_MAGIC;
    d.rank = g->part.rank;
    d.world = g->part.world;
    d.device = g->device;
    d.has_color = g->color != nullptr;
    d.pid = (int64_t)getpid();
    d.n_blocks = g->n_blocks_host;
    d.capacity = g->capacity;
    void *ptrs[4] = {g->block_keys, g->tsdf, g->weight, g->color};
    for (int a = 0; a < 4; ++a) {
        d.ptr[a] = (uint64_t)(uintptr_t)ptrs[a];
        if (!ptrs[a]) continue;
        if (p->own_ptr[a] != ptrs[a]) {           // pool (re)allocated since the last export
            MQ3D_CUDA(cudaIpcGetMemHandle(&p->own_handle[a], ptrs[a]));
            p->own_ptr[a] = ptrs[a];
        }
        d.handle[a] = p->own_handle[a];
    }
    g->ipc_exported = 1;
    memset(desc_out, 0, MQ3D_PEER_DESC_BYTES);
    memcpy(desc_out, &d, sizeof(d));
    return MQ3D_OK;
}

// blocks of peer blockIdx.y that the peer owns and this rank keeps as ghosts
__global__ void k_ghost_scan(const PeerView *__restrict__ peers, Partition part, int *__restrict__ count,
                             int32_t *__restrict__ pull_keys, int2 *__restrict__ pull_src) {
    const int d = blockIdx.y;
    if (d == part.rank) return;
    const PeerView pv = peers[d];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= pv.n_blocks) return;
    const int x = pv.keys[3 * j], y = pv.keys[3 * j + 1], z = pv.keys[3 * j + 2];
    if (mq3d_tile_owner(x, y, z, part) != d) return;   // only the owner's copy is authoritative
    if (!mq3d_block_needed(x, y, z, part)) return;
    const int i = atomicAdd(count, 1);
    pull_keys[3 * i] = x;
    pull_keys[3 * i + 1] = y;
    pull_keys[3 * i + 2] = z;
    pull_src[i] = make_int2(d, (int)j);
}

// one CTA per pulled block: 16 KB tsdf + 16 KB weight (+ 48 KB colour) from the owner's pool into the
// local pool; all loads of a phase are issued before the first store (NVLink latency)
__global__ void __launch_bounds__(256)
k_ghost_copy(const PeerView *__restrict__ peers, const int2 *__restrict__ pull_src, const int32_t *__restrict__ idx,
             const int *__restrict__ count, float *__restrict__ tsdf, float *__restrict__ weight, float *__restrict__ color) {
    const int64_t i = blockIdx.x;
    if (i >= *count) return;           // the grid covers the upper bound; the list length lives on the device
    const int b = idx[i];
    if (b < 0) return;
    const int2 s = pull_src[i];
    const PeerView pv = peers[s.x];
    const int tid = threadIdx.x;
    const float4 *st = reinterpret_cast<const float4 *>(pv.tsdf + (int64_t)s.y * MQ3D_RES3);
    const float4 *sw = reinterpret_cast<const float4 *>(pv.weight + (int64_t)s.y * MQ3D_RES3);
    float4 *dt = reinterpret_cast<float4 *>(tsdf + (int64_t)b * MQ3D_RES3);
    float4 *dw = reinterpret_cast<float4 *>(weight + (int64_t)b * MQ3D_RES3);
    float4 t[4], w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        t[q] = st[tid + 256 * q];
        w[q] = sw[tid + 256 * q];
    }
    if (color && pv.color) {
        const float4 *sc = reinterpret_cast<const float4 *>(pv.color + (int64_t)s.y * 3 * MQ3D_RES3);
        float4 *dc = reinterpret_cast<float4 *>(color + (int64_t)b * 3 * MQ3D_RES3);
        float4 c[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) c[q] = sc[tid + 256 * q];
#pragma unroll
        for (int q = 0; q < 12; ++q) dc[tid + 256 * q] = c[q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        dt[tid + 256 * q] = t[q];
        dw[tid + 256 * q] = w[q];
    }
}

// device pointer of allocation `a` of the peer described by `d`
static int resolve_peer_ptr(mq3d_grid *g, const PeerDesc &d, int a, void **out) {
    mq3d_peer_state *p = g->peer;
    *out = nullptr;
    if (!d.ptr[a]) return MQ3D_OK;
    if (d.pid == (int64_t)getpid()) {      // same process (several grids, rank-by-rank emulation): no IPC
        if (d.device != g->device && !p->peer_access[d.rank]) {
            cudaError_t e = cudaDeviceEnablePeerAccess(d.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
            } else if (e != cudaSuccess) {
                mq3d_set_error("cudaDeviceEnablePeerAccess(%d) -> %s", d.device, cudaGetErrorString(e));
                return MQ3D_ERR_CUDA;
            }
            p->peer_access[d.rank] = 1;
        }
        *out = (void *)(uintptr_t)d.ptr[a];
        return MQ3D_OK;
    }
    PeerMap &m = p->map[d.rank][a];
    if (m.ptr && memcmp(&m.handle, &d.handle[a], sizeof(cudaIpcMemHandle_t)) == 0) {
        *out = m.ptr;
        return MQ3D_OK;
    }
    if (m.ptr) {                            // the peer re-allocated this array: drop the stale mapping
        cudaIpcCloseMemHandle(m.ptr);
        m.ptr = nullptr;
    }
    MQ3D_CUDA(cudaIpcOpenMemHandle(&m.ptr, d.handle[a], cudaIpcMemLazyEnablePeerAccess));
    m.handle = d.handle[a];
    *out = m.ptr;
    return MQ3D_OK;
}

extern "C" int mq3d_grid_ghost_pull(mq3d_grid *g, const void *descs, int64_t *n_pulled, void *stream) {
    MQ3D_REQUIRE(g && descs, "null argument");
    if (n_pulled) *n_pulled = 0;
    const int world = g->part.world, rank = g->part.rank;
    if (world <= 1) return MQ3D_OK;
    MQ3D_REQUIRE(world <= MQ3D_MAX_PEERS, "peer ghost pull supports at most 64 ranks");
    MQ3D_TRY(mq3d_set_device(g->device));
    MQ3D_TRY(peer_state(g));
    mq3d_peer_state *p = g->peer;
    cudaStream_t st = as_stream(stream);
    int64_t total = 0, most = 0;
    for (int d = 0; d < world; ++d) {
        PeerDesc pd;
        memcpy(&pd, (const char *)descs + (size_t)d * MQ3D_PEER_DESC_BYTES, sizeof(pd));
        MQ3D_REQUIRE(pd.magic == MQ3D_PEER_MAGIC && pd.rank == d && pd.world == world, "bad peer descriptor");
        PeerView v;
        memset(&v, 0, sizeof(v));
        if (d != rank && pd.n_blocks > 0) {
            MQ3D_REQUIRE((pd.has_color != 0) == (g->color != nullptr), "peers disagree on the colour attribute");
            void *ptr[4];
            for (int a = 0; a < 4; ++a) MQ3D_TRY(resolve_peer_ptr(g, pd, a, &ptr[a]));
            v.keys = (const int32_t *)ptr[0];
            v.tsdf = (const float *)ptr[1];
            v.weight = (const float *)ptr[2];
            v.color = (const float *)ptr[3];
            v.n_blocks = pd.n_blocks;
            total += pd.n_blocks;
            if (pd.n_blocks > most) most = pd.n_blocks;
        }
        p->views_host[d] = v;
    }
    if (total == 0) return MQ3D_OK;
    if (total > p->pull_cap) {
        cudaFree(p->pull_keys);
        cudaFree(p->pull_src);
        p->pull_keys = nullptr;
        p->pull_src = nullptr;
        p->pull_cap = 0;
        int64_t cap = total + total / 4 + 64;
        MQ3D_CUDA(cudaMalloc(&p->pull_keys, sizeof(int32_t) * 3 * cap));
        MQ3D_CUDA(cudaMalloc(&p->pull_src, sizeof(int2) * cap));
        p->pull_cap = cap;
    }
    MqTrace tr(st);
    MQ3D_CUDA(cudaMemcpyAsync(p->views_dev, p->views_host, sizeof(PeerView) * world, cudaMemcpyHostToDevice, st));
    MQ3D_CUDA(cudaMemsetAsync(g->counter_dev + 5, 0, sizeof(int), st));
    dim3 grid((unsigned)((most + 255) / 256), (unsigned)world);
    k_ghost_scan<<<grid, 256, 0, st>>>(p->views_dev, g->part, g->counter_dev + 5, p->pull_keys, p->pull_src);
    MQ3D_CUDA(cudaGetLastError());
    tr.mark("scan");
    // Everything below is sized by the upper bound `total` (all blocks of all peers) and reads the list length on the
    // device: no host round trip between the scan and the copy.  The pool is reserved for the bound up front.
    MQ3D_TRY(mq3d_grid_activate(g, p->pull_keys, total, /*integrating=*/false, st, g->counter_dev + 5));
    tr.mark("activate");
    k_ghost_copy<<<(unsigned)total, 256, 0, st>>>(p->views_dev, p->pull_src, g->idx_scratch, g->counter_dev + 5, g->tsdf, g->weight,
                                                  g->color);
    MQ3D_CUDA(cudaGetLastError());
    tr.mark("copy");
    if (tr.on) {
        cudaStreamSynchronize(st);
        tr.report("ghost_pull");
    }
    if (n_pulled) {                     // optional: the caller wants the count (synchronises)
        MQ3D_CUDA(cudaMemcpyAsync(g->pinned_host + 5, g->counter_dev + 5, sizeof(int), cudaMemcpyDeviceToHost, st));
        MQ3D_CUDA(cudaStreamSynchronize(st));
        *n_pulled = g->pinned_host[5];
    }
    return MQ3D_OK;
}
