// K3 -- stereo un-mix, wasted-byte merge, escape-frame samples and PCM packing.
//
// Replaces Deinterlace16 / Deinterlace24 (ALACDecoder/AlacFile.cs:338-421),
// the mono packers (:527-575), the uncompressed-frame readers (:498-526,
// :663-700), the wasted-byte readers (:476-482, :634-641) and
// AlacContext.FormatSamples (ALACDecoder/AlacContext.cs:214-256).  Output is
// the byte stream AlacContext.Read hands out: interleaved, little-endian,
// left first.
//
// HBM-bound stage.  A warp owns a 32-frame x 32-sample block of a tile:
//   phase 1 (lane = frame)  : 128-byte coalesced plane rows -> un-mix with the
//                             lane's own mixShift/mixWeight -> padded smem tile;
//   phase 2 (lane = sample) : per frame, 32 consecutive samples -> one
//                             contiguous run of the frame's PCM (128 B for
//                             16-bit stereo, 192 B for 24-bit stereo).
// Wasted bytes and uncompressed (escape) frames are read straight from the
// bitstream in phase 2, where consecutive lanes read consecutive bit fields.
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

namespace alacgpu {

constexpr int kK3Warps = 4;

struct K3Smem {
    int32_t tl[kK3Warps][32][33];
    int32_t tr[kK3Warps][32][33];
    uint32_t stage[kK3Warps][52];    // 192 B row + slack for the unaligned window
};

__global__ void __launch_bounds__(kK3Warps * 32)
k3_stereo_pack(const ChunkArgs a, const uint32_t blocks_per_tile)
{
    __shared__ K3Smem sm;
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const uint32_t unit = blockIdx.x * kK3Warps + (uint32_t)w;
    const uint32_t tile = unit / blocks_per_tile;
    const uint32_t i0 = (unit % blocks_per_tile) * 32u;
    const uint32_t n_tiles = (a.n + kTile - 1) / kTile;
    if (tile >= n_tiles) return;

    // ---- lane = frame: this lane's frame parameters -----------------------
    const uint32_t slot = tile * kTile + (uint32_t)lane;
    const bool have = slot < a.n;
    FrameDesc d = {};
    uint32_t n_eff = 0;        // sample-frames of PCM this frame emits
    int ss = 16, nch = 2;
    uint64_t out = 0, frame_bit = 0;
    if (have) {
        const uint64_t f = a.f0 + slot;
        d = a.desc[f];
        const FrameRef ref = a.refs[f];
        const TrackCfg cfg = a.cfgs[ref.track];
        ss = cfg.sample_size;
        nch = cfg.num_channels;
        n_eff = d.out_len / (uint32_t)((ss >> 3) * nch);
        out = a.frame_off[f] - a.pcm_base;
        frame_bit = ref.off * 8ull;
    }
    // anything to do for this 32-sample block?
    if (!__any_sync(0xffffffffu, have && i0 < n_eff)) return;

    const bool ok = have && d.status == FS_OK;
    const bool stereo = ok && (d.flags & FF_STEREO);
    const bool from_planes = ok && !(d.flags & FF_ESCAPE);

    // ---- phase 1: plane rows -> un-mixed L/R in the smem tile --------------
    {
        const int32_t *pa = a.planes + ((uint64_t)tile * 2u) * a.ns * kTile + lane;
        const int32_t *pb = pa + (uint64_t)a.ns * kTile;
        const int mw = d.mix_weight, ms = d.mix_shift & 31;
#pragma unroll 8
        for (int r = 0; r < 32; r++) {
            const uint32_t i = i0 + (uint32_t)r;
            int32_t L = 0, R = 0;
            if (from_planes && i < n_eff) {
                const int32_t A = pa[(uint64_t)i * kTile];
                if (stereo) {
                    const int32_t B = pb[(uint64_t)i * kTile];
                    if (mw != 0) {                                   // AlacFile.cs:342-355, :373-380
                        R = (int32_t)((uint32_t)A - (uint32_t)((int32_t)((uint32_t)B * (uint32_t)mw) >> ms));
                        L = (int32_t)((uint32_t)R + (uint32_t)B);
                    } else { L = A; R = B; }                         // :359-366, :401-404
                } else {
                    L = A;                                           // :533-540 (second channel = 0)
                }
            }
            sm.tl[w][r][lane] = L;
            sm.tr[w][r][lane] = R;
        }
    }
    __syncwarp();

    // ---- phase 2: lane = sample; loop over the tile's frames ----------------
    const uint32_t *arena32 = reinterpret_cast<const uint32_t *>(a.arena);
    // pack the per-frame scalars once so each frame costs a handful of shuffles
    const uint32_t meta = (uint32_t)d.flags | ((uint32_t)d.ub << 8) | ((uint32_t)(ok ? 1 : 0) << 16) |
                          ((uint32_t)(ss == 24 ? 1 : 0) << 17) | ((uint32_t)(nch == 2 ? 1 : 0) << 18);
    for (int fr = 0; fr < 32; fr++) {
        const uint32_t f_n = __shfl_sync(0xffffffffu, n_eff, fr);
        if (i0 >= f_n) continue;                                     // warp-uniform
        const uint32_t f_meta = __shfl_sync(0xffffffffu, meta, fr);
        const uint64_t f_out = __shfl_sync(0xffffffffu, out, fr);
        const uint64_t f_bit = __shfl_sync(0xffffffffu, frame_bit, fr);
        const uint32_t f_data = __shfl_sync(0xffffffffu, d.data_bit, fr);
        const uint32_t f_shift = __shfl_sync(0xffffffffu, d.shift_bit, fr);
        const bool f_ok = (f_meta >> 16) & 1u;
        const bool f_24 = (f_meta >> 17) & 1u;
        const bool f_2ch = (f_meta >> 18) & 1u;
        const bool f_stereo = f_meta & FF_STEREO;
        const bool f_escape = f_meta & FF_ESCAPE;
        const int f_ub = (int)((f_meta >> 8) & 0xffu);
        const int f_ss = f_24 ? 24 : 16;
        const int ech = f_stereo ? 2 : 1;

        const uint32_t i = i0 + (uint32_t)lane;
        const uint32_t cnt = min(32u, f_n - i0);                     // samples of this frame in the block
        const bool live = (uint32_t)lane < cnt;
        int32_t L = sm.tl[w][lane][fr];
        int32_t R = sm.tr[w][lane][fr];
        if (live && f_ok && f_escape) {                              // AlacFile.cs:498-524, :663-696
            const uint64_t pos = f_bit + f_data + (uint64_t)i * (uint32_t)(ech * f_ss);
            L = sext((int32_t)arena_bits(arena32, pos, f_ss), f_ss);
            R = f_stereo ? sext((int32_t)arena_bits(arena32, pos + (uint32_t)f_ss, f_ss), f_ss) : 0;
        }
        if (live && f_ok && f_24 && f_ub != 0) {                     // :381-389, :405-413, :549-554
            const int sh = f_ub * 8;
            const uint32_t mask = ~(0xFFFFFFFFu << sh);
            const uint64_t pos = f_bit + f_shift + (uint64_t)i * (uint32_t)(ech * sh);
            L = (int32_t)(((uint32_t)L << sh) | (arena_bits(arena32, pos, sh) & mask));
            if (f_stereo) R = (int32_t)(((uint32_t)R << sh) | (arena_bits(arena32, pos + (uint32_t)sh, sh) & mask));
        }
        uint8_t *dst = a.pcm + f_out;
        if (!f_24) {
            if (f_2ch) {                                             // AlacContext.cs:231-242: low 16 bits, LE
                if (live)
                    reinterpret_cast<uint32_t *>(dst)[i] = ((uint32_t)L & 0xffffu) | ((uint32_t)R << 16);
            } else {
                if (live) reinterpret_cast<uint16_t *>(dst)[i] = (uint16_t)L;
            }
        } else {
            // 24-bit: bytes L0 L1 L2 [R0 R1 R2] (AlacFile.cs:390-395, :555-557) staged per row,
            // then written as aligned 32-bit words with byte-granular edges.
            const uint32_t bpf = f_2ch ? 6u : 3u;
            uint8_t *stg = reinterpret_cast<uint8_t *>(sm.stage[w]);
            __syncwarp();
            if (live) {
                uint8_t *q = stg + (uint32_t)lane * bpf;
                q[0] = (uint8_t)L; q[1] = (uint8_t)((uint32_t)L >> 8); q[2] = (uint8_t)((uint32_t)L >> 16);
                if (f_2ch) { q[3] = (uint8_t)R; q[4] = (uint8_t)((uint32_t)R >> 8); q[5] = (uint8_t)((uint32_t)R >> 16); }
            }
            __syncwarp();
            uint8_t *row = dst + (uint64_t)i0 * bpf;
            const uint32_t len = cnt * bpf;
            const uint32_t head = min(len, (uint32_t)((0u - (uint32_t)(uintptr_t)row) & 3u));
            const uint32_t nwords = (len - head) >> 2;
            const uint32_t tail = len - head - (nwords << 2);
            if ((uint32_t)lane < head) row[lane] = stg[lane];
            uint32_t *row32 = reinterpret_cast<uint32_t *>(row + head);
            for (uint32_t k = (uint32_t)lane; k < nwords; k += 32u) {
                const uint32_t lo = sm.stage[w][k], hi = sm.stage[w][k + 1];
                row32[k] = __funnelshift_r(lo, hi, head * 8u);
            }
            if ((uint32_t)lane < tail) row[head + (nwords << 2) + lane] = stg[head + (nwords << 2) + lane];
        }
    }
}

cudaError_t launch_k3(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    const uint32_t tiles = (a.n + kTile - 1) / kTile;
    const uint32_t bpt = (a.ns + 31) / 32;
    const uint64_t units = (uint64_t)tiles * bpt;
    const uint32_t blocks = (uint32_t)((units + kK3Warps - 1) / kK3Warps);
    k3_stereo_pack<<<blocks, kK3Warps * 32, 0, st>>>(a, bpt);
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
