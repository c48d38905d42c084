// One tap of the adaptive FIR predictor (ALACDecoder/AlacFile.cs:299-331), shared by the
// stream-lane LPC role (k2_lpc.cuh) and the frame-lane kernels (kf_frame.cu).
#pragma once
#include <cstdint>

namespace alacgpu {

// One tap of one sample: dot-product term + sign-LMS step, branch free.  Written in PTX so
// the update stays two predicated instructions (nvcc otherwise turns `if (E > 0)` into a
// branch per tap and sinks the operand computation into it).
//   dp  = sign * (buf[b] - buf[b+order-p])              (AlacFile.cs:324, :328)
//   acc += coef[p] * dp                                   (:303-304, sign folded out)
//   if (E > 0) { coef[p] -= sgn(dp); E -= ((|dp| + r) >> q) * (order - p); }   (:322-330)
// The weight -(order - p) is an immediate (one-lane warps are homogeneous in order).  Nine
// instructions, split evenly between the two integer pipes of a B200 sub-partition (IMAD on the FMA pipe;
// min/max, shifts and compares on the ALU pipe): |dp| + r is sgn(dp) * dp + r, one IMAD instead of
// IABS + IADD (r1: the LPC role issued 2.5 ALU-pipe instructions for every FMA-pipe one, and the
// ALU pipe is what a machine-filling batch runs out of).
template <int NEGM>
__device__ __forceinline__ void lpc_tap_imm(int32_t &c, int32_t &E, uint32_t &acc, const int32_t h, const int32_t nsg,
                                            const int32_t sgbase, const uint32_t r, const uint32_t q)
{
    asm("{\n\t"
        ".reg .s32 dp, a, u, t;\n\t"
        ".reg .pred act;\n\t"
        "mad.lo.s32 dp, %3, %4, %5;\n\t"
        "setp.gt.s32 act, %1, 0;\n\t"
        "mad.lo.s32 %2, %0, dp, %2;\n\t"
        "max.s32 t, dp, -1;\n\t"
        "min.s32 t, t, 1;\n\t"
        "mad.lo.s32 a, t, dp, %6;\n\t"
        "shr.u32 u, a, %7;\n\t"
        "@act sub.s32 %0, %0, t;\n\t"
        "@act mad.lo.s32 %1, u, %8, %1;\n\t"
        "}"
        : "+r"(c), "+r"(E), "+r"(acc)
        : "r"(h), "r"(nsg), "r"(sgbase), "r"(r), "r"(q), "n"(NEGM));
}
// taps pp = P .. 0 of an order-M stream (AlacFile.cs:322: newest coefficient index first)
template <int M, int P>
__device__ __forceinline__ void lpc_taps(int32_t (&c)[M], const int32_t (&H)[M + 1], int32_t &E, uint32_t &acc, const int32_t nsg,
                                         const int32_t sgbase, const uint32_t r, const uint32_t q)
{
    lpc_tap_imm<P - M>(c[P], E, acc, H[P], nsg, sgbase, r, q);
    if constexpr (P > 0) lpc_taps<M, P - 1>(c, H, E, acc, nsg, sgbase, r, q);
}

}  // namespace alacgpu
