// Halo exchange of the sharded step over NVLink peer memory (SURVEY §8b fbp_halo_*, §8e): the rows of the numerator sums
// that a slab interface shares are STORED straight into the owner's receive buffer by the sending GPU (no NCCL call, no
// packing on the host side), published with a flag, and summed into the owner's rows in a fixed order; the reverse pass
// returns the owners' row cotangents the same way.
//
// Every rank holds one symmetric allocation per constraint (same layout on every rank; the host passes the peers' base
// pointers, obtained from torch.distributed._symmetric_memory or cudaIpc — include/fbpinn_b200.h fbp_halo_peers):
//     int32 flags[4][FBP_HALO_MAX_WORLD]      [2 dir + 0] data flags,  [2 dir + 1] acknowledgement flags   (dir 0 fwd, 1 bwd)
//     float data[dir 0 ...][dir 1 ...]        receive regions, peers in rank order
// Protocol of one exchange in direction d, epoch e (a device counter per direction, the same on every rank):
//   push:  for every peer j it sends to: wait until j has acknowledged epoch e (its previous data was consumed), store the
//          rows into j's receive region, fence, publish data flag e + 1 at j
//   pull:  wait for data flag e + 1 of every peer it receives from, combine (add in peer order / copy), acknowledge e + 1 at
//          those peers; the last block advances the epoch
// Waits are bounded: a lost peer turns into a trap (error 719 at the next synchronisation), not a hung GPU.
// Both kernels are plain stream work: they capture into CUDA graphs with the rest of the step.
#include "fbp_common.cuh"

namespace {

constexpr int HT = 256;

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const int* p, int want) {
    long long t0 = 0;
    for (int spin = 0;; ++spin) {
        if (ld_acquire_sys(p) == want) return;
        if (spin == 256) t0 = clock64();
        if (spin > 256 && clock64() - t0 > 20000000000ll) __trap();      // ~10 s: the peer is gone
        __nanosleep(64);
    }
}

// block b works on peer blk_peer[b], rows [blk_row0[b], blk_row1[b]) of that peer's send list
__global__ void __launch_bounds__(HT) halo_push_kernel(const float* __restrict__ rows, int V, const int32_t* __restrict__ send_idx,
                                                       const int32_t* __restrict__ blk, fbp_halo_peers peers,
                                                       const int64_t* __restrict__ dst_off, int me, int dir,
                                                       const int32_t* __restrict__ epoch, int32_t* __restrict__ ticket) {
    const int j = blk[4 * blockIdx.x + 0], r0 = blk[4 * blockIdx.x + 1], r1 = blk[4 * blockIdx.x + 2], nb = blk[4 * blockIdx.x + 3];
    const int e = epoch[dir];
    __shared__ int last;
    if (threadIdx.x == 0) wait_flag(peers.flags[me] + (2 * dir + 1) * FBP_HALO_MAX_WORLD + j, e);   // j consumed my previous rows
    __syncthreads();
    float* dst = peers.data[j] + dst_off[j];
    // one thread per row: consecutive threads write consecutive rows of the peer's buffer (coalesced NVLink stores)
    for (int r = r0 + threadIdx.x; r < r1; r += HT) {
        const float* src = rows + (int64_t)send_idx[r] * V;
        float* d = dst + (int64_t)r * V;
        for (int v = 0; v < V; ++v) d[v] = src[v];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&ticket[j], 1) == nb - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
        ticket[j] = 0;
        __threadfence_system();
        st_release_sys(peers.flags[j] + (2 * dir + 0) * FBP_HALO_MAX_WORLD + me, e + 1);
    }
}

// mode 0: rows[tgt[i]] += sum over its sources (CSR src_ptr / src_pos into the receive region, peers in rank order);
// mode 1: rows[tgt[i]] = its single source
__global__ void __launch_bounds__(HT) halo_pull_kernel(float* __restrict__ rows, int V, const int32_t* __restrict__ tgt,
                                                       const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src_pos,
                                                       int n_tgt, fbp_halo_peers peers, int64_t my_off, uint32_t from_mask, int me,
                                                       int world, int dir, int mode, int32_t* __restrict__ epoch,
                                                       int32_t* __restrict__ done) {
    const int e = epoch[dir];
    if (threadIdx.x < world && ((from_mask >> threadIdx.x) & 1u))
        wait_flag(peers.flags[me] + (2 * dir + 0) * FBP_HALO_MAX_WORLD + threadIdx.x, e + 1);
    __syncthreads();
    const float* buf = peers.data[me] + my_off;
    // one thread per target row
    for (int t = blockIdx.x * HT + threadIdx.x; t < n_tgt; t += gridDim.x * HT) {
        float* dst = rows + (int64_t)tgt[t] * V;
        if (mode == 1) {                                   // reverse pass: position t of the receive region is row t's value
            const float* src = buf + (int64_t)t * V;
            for (int v = 0; v < V; ++v) dst[v] = src[v];
        } else {
            const int s0 = src_ptr[t], s1 = src_ptr[t + 1];
            for (int v = 0; v < V; ++v) {
                float acc = dst[v];
                for (int s = s0; s < s1; ++s) acc += buf[(int64_t)src_pos[s] * V + v];
                dst[v] = acc;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = (atomicAdd(done, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (last) {
        // every block has read the receive region: acknowledge to the senders, advance the epoch
        if (threadIdx.x < world && ((from_mask >> threadIdx.x) & 1u))
            st_release_sys(peers.flags[threadIdx.x] + (2 * dir + 1) * FBP_HALO_MAX_WORLD + me, e + 1);
        if (threadIdx.x == 0) {
            *done = 0;
            epoch[dir] = e + 1;
        }
    }
}

// dst (n, V) = scale * src[inv[row]] where inv[row] >= 0, else 0   (owned-row scatter of the reverse pass in one launch)
__global__ void __launch_bounds__(HT) scatter_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ inv, int64_t n,
                                                          int V, float scale, float* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * HT + threadIdx.x;
    if (i >= n * V) return;
    const int64_t r = i / V;
    const int v = (int)(i - r * V);
    const int s = inv[r];
    dst[i] = s >= 0 ? scale * src[(int64_t)s * V + v] : 0.0f;
}

}  // namespace

extern "C" {

int fbp_halo_push(const float* d_rows, int32_t row_floats, const int32_t* d_send_idx, const int32_t* d_blocks, int32_t n_blocks,
                  const fbp_halo_peers* peers, const int64_t* d_dst_off, int32_t me, int32_t dir, const int32_t* d_epoch,
                  int32_t* d_ticket, void* stream) {
    FBP_REQUIRE(dir == 0 || dir == 1, "fbp_halo_push: dir must be 0 (forward) or 1 (reverse)");
    if (n_blocks == 0) return 0;                          // nothing to send (the pull still runs: it advances the epoch)
    FBP_REQUIRE(peers && d_blocks && d_dst_off && d_epoch && d_ticket && d_rows && d_send_idx, "fbp_halo_push: null argument");
    halo_push_kernel<<<n_blocks, HT, 0, (cudaStream_t)stream>>>(d_rows, row_floats, d_send_idx, d_blocks, *peers, d_dst_off, me, dir,
                                                              d_epoch, d_ticket);
    FBP_LAUNCH_CHECK();
    return 0;
}

int fbp_halo_pull(float* d_rows, int32_t row_floats, const int32_t* d_tgt, const int32_t* d_src_ptr, const int32_t* d_src_pos,
                  int32_t n_tgt, const fbp_halo_peers* peers, int64_t my_off, uint32_t from_mask, int32_t me, int32_t world,
                  int32_t dir, int32_t mode, int32_t* d_epoch, int32_t* d_done, void* stream) {
    FBP_REQUIRE(peers && d_epoch && d_done, "fbp_halo_pull: null argument");
    FBP_REQUIRE(world >= 1 && world <= FBP_HALO_MAX_WORLD, "fbp_halo_pull: world size out of range");
    FBP_REQUIRE(dir == 0 || dir == 1, "fbp_halo_pull: dir must be 0 (forward) or 1 (reverse)");
    int grid = (n_tgt + HT - 1) / HT;
    grid = grid < 1 ? 1 : (grid > 592 ? 592 : grid);      // every rank runs it each exchange (it advances the epoch)
    halo_pull_kernel<<<grid, HT, 0, (cudaStream_t)stream>>>(d_rows, row_floats, d_tgt, d_src_ptr, d_src_pos, n_tgt, *peers, my_off,
                                                          from_mask, me, world, dir, mode, d_epoch, d_done);
    FBP_LAUNCH_CHECK();
    return 0;
}

int fbp_scatter_rows(const float* d_src, const int32_t* d_inv, int64_t n, int32_t row_floats, float scale, float* d_dst,
                     void* stream) {
    if (n == 0) return 0;
    const int64_t tot = n * row_floats;
    scatter_rows_kernel<<<(unsigned)((tot + HT - 1) / HT), HT, 0, (cudaStream_t)stream>>>(d_src, d_inv, n, row_floats, scale, d_dst);
    FBP_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
