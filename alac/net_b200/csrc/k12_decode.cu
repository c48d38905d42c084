// K1 + K2 kernels: entropy decode (k1_entropy.cuh) and LPC reconstruction
// (k2_lpc.cuh), as two plain kernels and as ONE fused launch in which the LPC
// warps consume a frame's residuals while the entropy lane is still producing
// them.
//
// Why fuse.  Both stages are serial per frame (8192 Rice symbols, then 4096
// predictor steps per channel), and a batch like BASELINE configs[1] (14,063
// frames) has far fewer frames than the GPU has lanes, so the batch takes
// T(entropy of one frame) + T(LPC of one frame) however many SMs idle.  In the
// fused kernel the blocks [0, n_eblocks) run the entropy role and the blocks
// after them the LPC role on the chunk's streams sorted by descending order.
// Channel A's LPC overlaps channel A's and B's entropy decode, channel B's LPC
// overlaps channel B's entropy decode: the batch takes about max(entropy, LPC
// behind it) instead of the sum.
//
// Hand-off: per stream one 32-bit progress word in global memory.  The entropy
// lane stores a row block, and every 32 residuals fences and publishes the
// count (kStreamDone at the end of the channel, also after a decode fault); the
// LPC lane checks the word (ld.acquire.gpu) before it prefetches a residual
// block it has not been granted yet, and reads the plane with ld.global.cg.
// Forward progress: an LPC block only ever waits on entropy blocks, which have
// LOWER block indices and are therefore resident or finished by the time it
// runs (the same assumption decoupled look-back scans rest on); the wait is
// bounded anyway and flags ALACGPU_FRAME_INTERNAL instead of hanging.
#include "k1_entropy.cuh"
#include "k2_lpc.cuh"

namespace alacgpu {

__global__ void __launch_bounds__(kK1Threads)
k1_entropy(const ChunkArgs a, const int lanes_log2)
{
    __shared__ __align__(256) uint8_t smem[kRingBytes * kK1Threads];
    entropy_block<false>(a, lanes_log2, blockIdx.x, smem);
}

__global__ void __launch_bounds__(kK2Threads)
k2_lpc(const ChunkArgs a)
{
    __shared__ int32_t hist_smem[32 * kK2Threads];
    const uint32_t warp = (blockIdx.x * kK2Threads + threadIdx.x) >> 5;
    lpc_role<false>(a, warp, hist_smem + (threadIdx.x >> 5) * 1024);
}

static_assert(kK1Threads == kK2Threads, "the fused kernel uses one block size for both roles");

__global__ void __launch_bounds__(kK1Threads)
k12_entropy_lpc(const ChunkArgs a, const int lanes_log2, const uint32_t n_eblocks)
{
    __shared__ __align__(256) uint8_t smem[kRingBytes * kK1Threads];      // 32 KB: bit rings, or 16 KB of LPC history
    if (blockIdx.x < n_eblocks) {
        entropy_block<true>(a, lanes_log2, blockIdx.x, smem);
    } else {
        const uint32_t warp = ((blockIdx.x - n_eblocks) * kK2Threads + threadIdx.x) >> 5;
        lpc_role<true>(a, warp, reinterpret_cast<int32_t *>(smem) + (threadIdx.x >> 5) * 1024);
    }
}

static int lanes_log2_of(int lanes_per_warp)
{
    if (lanes_per_warp == 16) return 4;
    if (lanes_per_warp == 8) return 3;
    if (lanes_per_warp == 4) return 2;
    return 5;
}

cudaError_t launch_k1(const ChunkArgs &a, int lanes_per_warp, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    const int lg = lanes_log2_of(lanes_per_warp);
    const uint32_t warps = (a.n + (1u << lg) - 1) >> lg;
    k1_entropy<<<(warps + 3) / 4, kK1Threads, 0, st>>>(a, lg);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_sort(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    k0s_order_sort<<<1, kSortThreads, 0, st>>>(a.desc + a.f0, a.n, a.perm, a.perm_count);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_k2(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    const uint32_t warps = (a.n * 2u + ALACGPU_LPC_STREAMS_PER_WARP - 1u) / ALACGPU_LPC_STREAMS_PER_WARP;   // upper bound; warps past n_active exit at once
    k2_lpc<<<(warps + 3) / 4, kK2Threads, 0, st>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_k12(const ChunkArgs &a, int lanes_per_warp, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    const int lg = lanes_log2_of(lanes_per_warp);
    const uint32_t ewarps = (a.n + (1u << lg) - 1) >> lg;
    const uint32_t eblocks = (ewarps + 3) / 4;
    const uint32_t lwarps = (a.n * 2u + ALACGPU_LPC_STREAMS_PER_WARP - 1u) / ALACGPU_LPC_STREAMS_PER_WARP;
    k12_entropy_lpc<<<eblocks + (lwarps + 3) / 4, kK1Threads, 0, st>>>(a, lg, eblocks);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace alacgpu
