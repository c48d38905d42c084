// K3 -- stereo un-mix, wasted-byte merge, escape-frame samples and PCM packing.
//
// Replaces Deinterlace16 / Deinterlace24 (ALACDecoder/AlacFile.cs:338-421),
// the mono packers (:527-575), the uncompressed-frame readers (:498-526,
// :663-700), the wasted-byte readers (:476-482, :634-641) and
// AlacContext.FormatSamples (ALACDecoder/AlacContext.cs:214-256).  Output is
// the byte stream AlacContext.Read hands out: interleaved, little-endian,
// left first.
//
// HBM-bound streaming stage.  One thread owns EIGHT consecutive sample-frames
// of one frame: two 16-byte loads per channel from the stream-major planes,
// un-mix with the frame's mixShift/mixWeight, and 16 / 24 / 32 / 48 bytes of
// PCM written with 16-byte stores (consecutive threads write consecutive
// bytes).  Wasted bytes and uncompressed (escape) frames come straight from
// the bitstream: the thread loads the aligned 32-bit words that cover its run
// of bit fields, shifts the run to a word boundary with funnel shifts, and
// extracts the fields at compile-time positions.
#include "k3_pack.cuh"

namespace alacgpu {

__global__ void __launch_bounds__(kK3Threads)
k3_stereo_pack(const ChunkArgs a)
{
    const uint32_t i0 = (blockIdx.x * kK3Threads + threadIdx.x) * kK3PerThread;
    uint32_t w[12], nbytes, cnt;
    uint8_t *dst;
    if (!pack_group(a, blockIdx.y, i0, w, nbytes, cnt, dst)) return;
    store_group(dst, w, nbytes, cnt);
}

cudaError_t launch_k3(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    const uint32_t bx = (a.max_sf + kK3PerBlock - 1) / kK3PerBlock;
    // grid.y is limited to 65535: chunks hold at most 32768 frames
    dim3 grid(bx ? bx : 1, a.n, 1);
    k3_stereo_pack<<<grid, kK3Threads, 0, st>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// ---- position-weighted checksum (alacgpu_pcm_checksum) ------------------------
__global__ void __launch_bounds__(256)
k_checksum(const uint8_t *__restrict__ p, uint64_t len, uint64_t first_word, unsigned long long *__restrict__ sum)
{
    // p is 8-byte aligned and corresponds to global word index first_word
    const uint64_t nwords = (len + 7) >> 3;
    uint64_t acc = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nwords; j += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v;
        if ((j + 1) * 8 <= len) {
            v = reinterpret_cast<const uint64_t *>(p)[j];
        } else {
            v = 0;
            for (uint64_t b = j * 8; b < len; b++) v |= (uint64_t)p[b] << (8 * (b - j * 8));
        }
        acc += v * (2 * (first_word + j) + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, (unsigned long long)acc);
}

cudaError_t launch_checksum(const uint8_t *pcm, uint64_t global_off, uint64_t len, uint64_t *d_sum, cudaStream_t st)
{
    if (len == 0) return cudaSuccess;
    const uint64_t nwords = (len + 7) >> 3;
    uint32_t blocks = (uint32_t)((nwords + 255) / 256);
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    k_checksum<<<blocks, 256, 0, st>>>(pcm, len, global_off >> 3, reinterpret_cast<unsigned long long *>(d_sum));
    return cudaGetLastError();
}

}  // namespace alacgpu
