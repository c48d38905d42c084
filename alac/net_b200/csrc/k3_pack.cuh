// K3 device code -- stereo un-mix, wasted-byte merge, escape-frame samples and PCM packing,
// shared by the standalone kernel (k3_stereo.cu) and the pack role of the fused launch
// (k12_decode.cu).  See k3_stereo.cu for what it replaces in the reference.
#pragma once
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

namespace alacgpu {

constexpr int kK3Threads = 128;
constexpr int kK3PerThread = 8;                              // sample-frames per thread
constexpr int kK3PerBlock = kK3Threads * kK3PerThread;       // 1024 sample-frames per block

// A run of NF big-endian bit fields of W bits each starting at absolute arena bit `pos`.
// out[v] = field v (zero-extended).  NF * W is a multiple of 32.
template <int NF, int W>
__device__ __forceinline__ void read_fields(const uint32_t *__restrict__ arena32, uint64_t pos, uint32_t (&out)[NF])
{
    constexpr int NW = NF * W / 32;
    const uint64_t w0 = pos >> 5;
    const int off = (int)(pos & 31);
    uint32_t raw[NW + 1];
#pragma unroll
    for (int j = 0; j <= NW; j++) raw[j] = bswap32(__ldg(arena32 + w0 + j));
    uint32_t al[NW + 1];
#pragma unroll
    for (int j = 0; j < NW; j++) al[j] = __funnelshift_l(raw[j + 1], raw[j], off);
    al[NW] = 0;
#pragma unroll
    for (int v = 0; v < NF; v++) {
        const int bit = v * W, j = bit >> 5, o = bit & 31;
        const uint32_t win = o ? __funnelshift_l(al[j + 1], al[j], o) : al[j];   // o + W may cross a word
        out[v] = win >> (32 - W);
    }
}

// 8 sample-frames x `ech` channels of W-bit fields, interleaved A,B per sample (AlacFile.cs:634-641,
// :665-696) -> fa[8], fb[8]
template <int W>
__device__ __forceinline__ void read_pairs(const uint32_t *__restrict__ arena32, uint64_t pos, bool two,
                                           uint32_t (&fa)[8], uint32_t (&fb)[8])
{
    if (two) {
        uint32_t f[16];
        read_fields<16, W>(arena32, pos, f);
#pragma unroll
        for (int s = 0; s < 8; s++) { fa[s] = f[2 * s]; fb[s] = f[2 * s + 1]; }
    } else {
        read_fields<8, W>(arena32, pos, fa);
#pragma unroll
        for (int s = 0; s < 8; s++) fb[s] = 0;
    }
}

// The eight sample-frames [i0, i0 + 8) of the chunk's frame `slot`: un-mixed, merged and packed into
// w[0 .. nbytes/4) (little-endian PCM bytes).  Returns false if the group lies past the frame's PCM.
__device__ __forceinline__ bool pack_group(const ChunkArgs &a, const uint32_t slot, const uint32_t i0,
                                           uint32_t (&w)[12], uint32_t &nbytes, uint32_t &cnt, uint8_t *&dst)
{
    if (!ALACGPU_CHECK(a.check, slot < a.n, CK_FRAME)) return false;
    const uint64_t f = a.f0 + slot;
    const FrameDesc d = a.desc[f];
    const FrameRef ref = a.refs[f];
    const TrackCfg cfg = a.cfgs[ref.track];
    const int ss = cfg.sample_size;
    const bool two_ch = cfg.num_channels == 2;
    const bool is24 = ss == 24;
    const uint32_t bpf = (uint32_t)(ss >> 3) * (uint32_t)cfg.num_channels;   // bytes per sample-frame
    const uint32_t n_eff = d.out_len / bpf;                  // sample-frames of PCM this frame emits
    if (i0 >= n_eff) return false;
    cnt = min((uint32_t)kK3PerThread, n_eff - i0);

    const bool ok = d.status == FS_OK;
    const bool stereo = ok && (d.flags & FF_STEREO);
    const bool escape = ok && (d.flags & FF_ESCAPE);
    const uint32_t *arena32 = reinterpret_cast<const uint32_t *>(a.arena);
    const uint64_t frame_bit = ref.off * 8ull;
    const int ech = stereo ? 2 : 1;

    int32_t L[8], R[8];
#pragma unroll
    for (int s = 0; s < 8; s++) { L[s] = 0; R[s] = 0; }

    if (ok && !escape && !ALACGPU_CHECK(a.check, (((uint64_t)slot * 2u + 1u) * a.ns + i0 + 8u) * 4u <= a.plane_bytes, CK_PLANE)) return false;
    if (ok && !escape) {
        // predicted samples: rows are padded to a multiple of 8, so the loads never leave the row
        const int4 *ra = reinterpret_cast<const int4 *>(a.planes + ((uint64_t)slot * 2u) * a.ns + i0);
        const int4 a0 = __ldcg(ra), a1 = __ldcg(ra + 1);
        int32_t A[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (stereo) {
            const int4 *rb = reinterpret_cast<const int4 *>(a.planes + ((uint64_t)slot * 2u + 1u) * a.ns + i0);
            const int4 b0 = __ldcg(rb), b1 = __ldcg(rb + 1);
            int32_t B[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const int mw = d.mix_weight, ms = d.mix_shift & 31;
#pragma unroll
            for (int s = 0; s < 8; s++) {
                if (mw != 0) {                                       // AlacFile.cs:342-355, :373-380
                    R[s] = (int32_t)((uint32_t)A[s] - (uint32_t)((int32_t)((uint32_t)B[s] * (uint32_t)mw) >> ms));
                    L[s] = (int32_t)((uint32_t)R[s] + (uint32_t)B[s]);
                } else { L[s] = A[s]; R[s] = B[s]; }                 // :359-366, :401-404
            }
        } else {
#pragma unroll
            for (int s = 0; s < 8; s++) L[s] = A[s];                 // :533-540 (second channel = 0)
        }
        if (is24 && d.ub != 0) {                                     // :381-389, :405-413, :549-554
            const int sh = d.ub * 8;
            const uint32_t mask = ~(0xFFFFFFFFu << sh);
            const uint64_t pos = frame_bit + d.shift_bit + (uint64_t)i0 * (uint32_t)(ech * sh);
            uint32_t fa[8], fb[8];
            if (d.ub == 1) read_pairs<8>(arena32, pos, stereo, fa, fb);
            else if (d.ub == 2) read_pairs<16>(arena32, pos, stereo, fa, fb);
            else read_pairs<24>(arena32, pos, stereo, fa, fb);
#pragma unroll
            for (int s = 0; s < 8; s++) {
                L[s] = (int32_t)(((uint32_t)L[s] << sh) | (fa[s] & mask));
                if (stereo) R[s] = (int32_t)(((uint32_t)R[s] << sh) | (fb[s] & mask));
            }
        }
    } else if (escape) {                                             // AlacFile.cs:498-524, :663-696
        const uint64_t pos = frame_bit + d.data_bit + (uint64_t)i0 * (uint32_t)(ech * ss);
        uint32_t fa[8], fb[8];
        if (is24) read_pairs<24>(arena32, pos, stereo, fa, fb);
        else read_pairs<16>(arena32, pos, stereo, fa, fb);
#pragma unroll
        for (int s = 0; s < 8; s++) {
            L[s] = sext((int32_t)fa[s], ss);
            R[s] = stereo ? sext((int32_t)fb[s], ss) : 0;
        }
    }

    // ---- pack: little-endian, left first; 16-bit = low 16 bits of each int (AlacContext.cs:231-242),
    // 24-bit = low 24 bits (AlacFile.cs:390-395, :555-557) -----------------------------------------
    if (!is24) {
        if (two_ch) {
#pragma unroll
            for (int s = 0; s < 8; s++) w[s] = ((uint32_t)L[s] & 0xffffu) | ((uint32_t)R[s] << 16);
        } else {
#pragma unroll
            for (int s = 0; s < 4; s++) w[s] = ((uint32_t)L[2 * s] & 0xffffu) | ((uint32_t)L[2 * s + 1] << 16);
        }
    } else {
        if (two_ch) {
            // per pair of sample-frames: L0 R0 L1 R1 (4 x 24 bits) -> 3 words
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const uint32_t l0 = (uint32_t)L[2 * p] & 0xffffffu, r0 = (uint32_t)R[2 * p] & 0xffffffu;
                const uint32_t l1 = (uint32_t)L[2 * p + 1] & 0xffffffu, r1 = (uint32_t)R[2 * p + 1] & 0xffffffu;
                w[3 * p] = l0 | (r0 << 24);
                w[3 * p + 1] = (r0 >> 8) | (l1 << 16);
                w[3 * p + 2] = (l1 >> 16) | (r1 << 8);
            }
        } else {
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const uint32_t s0 = (uint32_t)L[4 * p] & 0xffffffu, s1 = (uint32_t)L[4 * p + 1] & 0xffffffu;
                const uint32_t s2 = (uint32_t)L[4 * p + 2] & 0xffffffu, s3 = (uint32_t)L[4 * p + 3] & 0xffffffu;
                w[3 * p] = s0 | (s1 << 24);
                w[3 * p + 1] = (s1 >> 8) | (s2 << 16);
                w[3 * p + 2] = (s2 >> 16) | (s3 << 8);
            }
        }
    }
    nbytes = cnt * bpf;                                               // 16, 24, 32 or 48 when cnt == 8
    dst = a.pcm + (a.frame_off[f] - a.pcm_base) + (uint64_t)i0 * bpf;
    // (checked build) the group lies inside the PCM buffer
    if (!ALACGPU_CHECK(a.check, a.frame_off[f] >= a.pcm_base && (a.frame_off[f] - a.pcm_base) + (uint64_t)i0 * bpf + nbytes <= a.pcm_bytes, CK_PCM))
        return false;
    return true;
}

// One thread's group straight to memory: 16-byte stores when it is whole and aligned, else bytes.
__device__ __forceinline__ void store_group(uint8_t *dst, const uint32_t (&w)[12], const uint32_t nbytes, const uint32_t cnt)
{
    if (cnt == kK3PerThread && ((uintptr_t)dst & 15u) == 0) {
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        d4[0] = make_uint4(w[0], w[1], w[2], w[3]);
        if (nbytes == 24) {
            reinterpret_cast<uint2 *>(dst)[2] = make_uint2(w[4], w[5]);
        } else if (nbytes >= 32) {
            d4[1] = make_uint4(w[4], w[5], w[6], w[7]);
            if (nbytes == 48) d4[2] = make_uint4(w[8], w[9], w[10], w[11]);
        }
    } else {
        // partial last group of a frame, or a frame that starts off a 16-byte boundary (after an
        // odd-sized partial frame): byte stores from statically indexed registers
#pragma unroll
        for (int j = 0; j < 12; j++) {
#pragma unroll
            for (int bb = 0; bb < 4; bb++)
                if ((uint32_t)(j * 4 + bb) < nbytes) dst[j * 4 + bb] = (uint8_t)(w[j] >> (8 * bb));
        }
    }
}

}  // namespace alacgpu
