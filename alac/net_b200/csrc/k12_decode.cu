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
// lane stores its residuals one by one, and every 64 steps fences and publishes its
// output index (kStreamDone at the end of the channel, also after a decode fault); the
// LPC lane checks the word (ld.acquire.gpu) before it prefetches a residual
// block it has not been granted yet, and reads the plane with ld.global.cg.
// Forward progress: an LPC block only ever waits on entropy blocks, which have
// LOWER block indices and are therefore resident or finished by the time it
// runs (the same assumption decoupled look-back scans rest on); the wait is
// bounded anyway and flags ALACGPU_FRAME_INTERNAL instead of hanging.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "k1_entropy.cuh"
#include "k2_lpc.cuh"
#include "k3_pack.cuh"

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
    lpc_role<false, false>(a, warp, hist_smem + (threadIdx.x >> 5) * 1024);
}

static_assert(kK1Threads == kK2Threads, "the fused kernel uses one block size for both roles");

// Debug aid (ALACGPU_TRACE=<file>): per-warp [start, end] of the fused launch in globaltimer ns, and the SM.
__device__ unsigned long long *g_trace = nullptr;
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_warp(unsigned long long t0)
{
    unsigned long long *tr = g_trace;
    if (tr && (threadIdx.x & 31) == 0) {
        const uint32_t w = blockIdx.x * (kK1Threads / 32) + (threadIdx.x >> 5);
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[3 * w] = t0;
        tr[3 * w + 1] = gtime();
        tr[3 * w + 2] = smid;
    }
}

__global__ void __launch_bounds__(kK1Threads, 3)
k12_entropy_lpc(const ChunkArgs a, const int lanes_log2, const uint32_t n_eblocks)
{
    __shared__ __align__(256) uint8_t smem[kRingBytes * kK1Threads];      // 32 KB: bit rings, or 16 KB of LPC history
    const unsigned long long t0 = gtime();
    if (blockIdx.x < n_eblocks) {
        entropy_block<true>(a, lanes_log2, blockIdx.x, smem);
    } else {
        const uint32_t warp = ((blockIdx.x - n_eblocks) * kK2Threads + threadIdx.x) >> 5;
        lpc_role<true, false>(a, warp, reinterpret_cast<int32_t *>(smem) + (threadIdx.x >> 5) * 1024);
    }
    trace_warp(t0);
}

// ---- pack role of the fully fused launch ---------------------------------------------------------
// A few persistent blocks at the END of the grid (an LPC or entropy warp must not turn into a pack
// worker when it retires: it would then wait, without ever leaving the SM, on producers whose blocks
// may not be resident yet -- measured as a deadlock broken only by the bounded wait).  Their number
// follows the amount of PCM (launch_k123): spinning pack blocks hold SM slots, and too many of them
// starve the producers of the other chunks in flight (148 per chunk: e2e 11 -> 26 ms).
// A task = 256 consecutive sample-frames (one warp, 8 per lane)
constexpr uint32_t kPackGroup = 32 * kK3PerThread;        // sample-frames per task

__device__ __forceinline__ void pack_role(const ChunkArgs &a, uint8_t *stage /* 1536 B of shared memory per warp */)
{
    const int lane = threadIdx.x & 31;
    const uint32_t groups = (a.max_sf + kPackGroup - 1) / kPackGroup;
    const uint32_t total = groups * a.n;
    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(a.pack_next, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= total) break;
        const uint32_t g = task / a.n, slot = task - g * a.n;
        const FrameDesc d = a.desc[a.f0 + slot];
        const uint32_t first = g * kPackGroup;
        // wait for the producers (lane c polls channel c)
        if (d.status0 == FS_OK && !(d.flags & FF_ESCAPE) && first < d.n) {
            const uint32_t need = min((uint32_t)d.n, first + kPackGroup);
            const bool mine = lane < ((d.flags & FF_STEREO) ? 2 : 1);
            const uint32_t sid = slot * 2u + (uint32_t)(lane & 1);
            const uint32_t *word = (mine && a.lpc_flag[sid]) ? a.lpc_done + sid : a.progress + sid;
            uint32_t avail = 0;
            bool stalled = false;
            wait_avail(word, need, avail, mine, stalled);
            if (__any_sync(0xffffffffu, stalled) && lane == 0) { a.desc[a.f0 + slot].status = FS_INTERNAL; atomicAdd(a.faults, 1u); }
        }
        uint32_t w[12], nbytes = 0, cnt = 0;
        uint8_t *dst = nullptr;
        const bool have = pack_group(a, slot, first + (uint32_t)lane * kK3PerThread, w, nbytes, cnt, dst);
        // whole warp, whole groups, 16-byte aligned row: coalesced rows through shared memory
        const uint8_t *dst0 = reinterpret_cast<const uint8_t *>(__shfl_sync(0xffffffffu, (unsigned long long)dst, 0));
        const bool full = have && cnt == kK3PerThread;
        if (__all_sync(0xffffffffu, full) && ((uintptr_t)dst0 & 15u) == 0) {
            const uint32_t words = nbytes >> 2;                       // 4, 6, 8 or 12 per lane
            uint32_t *sw = reinterpret_cast<uint32_t *>(stage);
#pragma unroll
            for (int j = 0; j < 12; j++)
                if ((uint32_t)j < words) sw[(uint32_t)lane * words + j] = w[j];
            __syncwarp();
            const uint32_t row_bytes = nbytes * 32u;
            uint8_t *out = const_cast<uint8_t *>(dst0);
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const uint32_t o = (uint32_t)r * 512u + (uint32_t)lane * 16u;
                if (o < row_bytes) *reinterpret_cast<uint4 *>(out + o) = *reinterpret_cast<const uint4 *>(stage + o);
            }
            __syncwarp();
        } else if (have) {
            store_group(dst, w, nbytes, cnt);
        }
    }
}

__global__ void __launch_bounds__(kK1Threads, 3)
k123_decode(const ChunkArgs a, const int lanes_log2, const uint32_t n_eblocks, const uint32_t n_lblocks)
{
    __shared__ __align__(256) uint8_t smem[kRingBytes * kK1Threads];      // bit rings / LPC history / pack staging
    if (blockIdx.x < n_eblocks) {
        entropy_block<true>(a, lanes_log2, blockIdx.x, smem);
    } else if (blockIdx.x < n_eblocks + n_lblocks) {
        const uint32_t warp = ((blockIdx.x - n_eblocks) * kK2Threads + threadIdx.x) >> 5;
        lpc_role<true, true>(a, warp, reinterpret_cast<int32_t *>(smem) + (threadIdx.x >> 5) * 1024);
    } else {
        pack_role(a, smem + (threadIdx.x >> 5) * 2048);
    }
}

// Frames the entropy stage gave up on AFTER the pack warps had already written part of their PCM
// (status0 OK, status not): the contract is zero PCM of the nominal size.
__global__ void __launch_bounds__(128)
k3_fix_failed(const ChunkArgs a)
{
    const uint32_t slot = blockIdx.x;
    const FrameDesc d = a.desc[a.f0 + slot];
    if (d.status == FS_OK || d.status0 != FS_OK) return;
    uint8_t *dst = a.pcm + (a.frame_off[a.f0 + slot] - a.pcm_base);
    for (uint32_t i = threadIdx.x; i < d.out_len; i += 128) dst[i] = 0;
}

// The entropy lanes skip over zero runs instead of writing them (k1_entropy.cuh): the chunk's rows
// start out cleared, like the reference's output buffer (AlacFile.cs:238-245).
static cudaError_t clear_planes(const ChunkArgs &a, cudaStream_t st)
{
    return cudaMemsetAsync(a.planes, 0, (size_t)a.n * 2u * a.ns * sizeof(int32_t), st);
}

// upper bound of the four-lane (eight-lane) warps, eight (four) streams each; warps past the work lists exit at once
static uint32_t quad_warp_bound(const ChunkArgs &a)
{
    const uint32_t per_frame = ((a.use_quads & 255) ? 1u : 0u) + ((a.use_quads >> 8) & 255 ? 1u : 0u);
    const uint32_t per_warp = ((a.use_quads >> 16) & 1) ? 4u : 8u;
    return (a.n * per_frame + per_warp - 1u) / per_warp;
}

constexpr uint32_t kSmallChunkFrames = 20480;    // up to here a chunk is latency-bound (runtime.cu: kFullFusionMaxFrames)

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
    if (cudaError_t e = clear_planes(a, st)) return e;
    k1_entropy<<<(warps + 3) / 4, kK1Threads, 0, st>>>(a, lg);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_sort(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    k0s_order_sort<<<1, kSortThreads, 0, st>>>(a.desc + a.f0, a.n, a.perm, a.perm_count, a.lpc_flag, a.use_quads);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_k2(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    // upper bound; warps past the work lists exit at once
    const uint32_t warps = ((a.n * 2u + 31u) / 32u + 31u) + quad_warp_bound(a);
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
    const uint32_t lwarps = ((a.n * 2u + 31u) / 32u + 31u) + quad_warp_bound(a);
    if (cudaError_t e = clear_planes(a, st)) return e;
    const uint32_t grid = eblocks + (lwarps + 3) / 4;
    static const char *trace_path = getenv("ALACGPU_TRACE");
    unsigned long long *tr = nullptr;
    if (trace_path && a.n >= 4096) {          // debug: one whole-batch launch, synchronous
        cudaMalloc(&tr, (size_t)grid * 12 * sizeof(unsigned long long));
        cudaMemset(tr, 0, (size_t)grid * 12 * sizeof(unsigned long long));
        cudaMemcpyToSymbol(g_trace, &tr, sizeof(tr));
    }
    // Blocks per SM.  The fused kernel needs 122 registers and 33 KB of shared memory, so four blocks fit
    // an SM.  A small (latency-bound) chunk wants all its warps resident at once (configs[1]: 2.66 ms with
    // four blocks, 2.97 ms with two).  A big chunk fills the machine anyway, and there fewer resident
    // blocks are FASTER (configs[3]-shaped batch: 62.8 Gsamples/s with two blocks per SM, 57.3 with four:
    // with two warps per scheduler the role loops stay in the instruction caches), so big chunks ask for
    // 44 KB of unused dynamic shared memory.
    static const int pad_env = getenv("ALACGPU_SMEM_PAD") ? atoi(getenv("ALACGPU_SMEM_PAD")) : -1;
    const int pad_kb = pad_env >= 0 ? pad_env : (a.n > kSmallChunkFrames ? 44 : 0);
    if (pad_kb > 14) {      // the attribute is per device: set it on the current one, every time (a cheap driver call)
        if (cudaError_t e = cudaFuncSetAttribute(k12_entropy_lpc, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)) return e;
    }
    k12_entropy_lpc<<<grid, kK1Threads, (size_t)pad_kb * 1024, st>>>(a, lg, eblocks);
    if (tr) {
        cudaStreamSynchronize(st);
        std::vector<unsigned long long> h((size_t)grid * 12);
        cudaMemcpy(h.data(), tr, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        std::vector<uint32_t> cnt(2);
        cudaMemcpy(cnt.data(), a.perm_count, 8, cudaMemcpyDeviceToHost);
        if (FILE *fp = fopen(trace_path, "w")) {
            fprintf(fp, "# eblocks %u grid %u one_lane_streams %u quad_streams %u\n", eblocks, grid, cnt[0], cnt[1]);
            for (size_t w = 0; w < (size_t)grid * 4; w++) fprintf(fp, "%zu %llu %llu %llu\n", w, h[3 * w], h[3 * w + 1], h[3 * w + 2]);
            fclose(fp);
        }
        unsigned long long *none = nullptr;
        cudaMemcpyToSymbol(g_trace, &none, sizeof(none));
        cudaFree(tr);
    }
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_k123(const ChunkArgs &a, int lanes_per_warp, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    const int lg = lanes_log2_of(lanes_per_warp);
    const uint32_t ewarps = (a.n + (1u << lg) - 1) >> lg;
    const uint32_t eblocks = (ewarps + 3) / 4;
    const uint32_t lwarps = ((a.n * 2u + 31u) / 32u + 31u) + quad_warp_bound(a);
    const uint32_t lblocks = (lwarps + 3) / 4;
    // pack blocks: one per ~2400 tasks (a task is ~2.5 us of one warp, the decode stages leave ~2.5 ms)
    const uint32_t tasks = ((a.max_sf + kPackGroup - 1) / kPackGroup) * a.n;
    static const uint32_t div = getenv("ALACGPU_PACK_DIV") ? (uint32_t)atoi(getenv("ALACGPU_PACK_DIV")) : 2400u;
    const uint32_t pblocks = tasks / div + 2u < 148u ? tasks / div + 2u : 148u;
    if (cudaError_t e = clear_planes(a, st)) return e;
    k123_decode<<<eblocks + lblocks + pblocks, kK1Threads, 0, st>>>(a, lg, eblocks, lblocks);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_fix(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    k3_fix_failed<<<a.n, 128, 0, st>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace alacgpu
