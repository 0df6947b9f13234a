// unpack_iq.cuh -- Unpack8<TYPE> for the codebook formats (see iq_grid.cuh): the 8 values of octet p (0..31) of a super-block
// as fp32, in the reference's / gguf-py's operation order (dequantize.cuh:256-399; gguf-py quants.py classes IQ2_XXS ... IQ1_M):
//   IQ2_*:   db = (d * (0.5 + ls)) * 0.25;  y = +-(db * grid)
//   IQ3_*:   db = (d * (0.5 + ls)) * 0.5;   y = +-(db * grid)     (IQ3_S: the reference's table, top level 62 -- iq_tables.cuh)
//   IQ1_*:   dl = d * (2 s + 1);            y = dl * ((grid + 1) + (-1 +- 0.125))
// every product a separately rounded fp32 operation; the sign is applied to the bits.
#pragma once

#include "unpack.cuh"
#include "iq_grid.cuh"

template <int TYPE> struct UnpackIq8 {
    __device__ __forceinline__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const ggb_iq::Octet o = ggb_iq::octet<TYPE>(b, p >> 2, p & 3);
        const float d = ggb_iq::block_d<TYPE>(b);
        float q[8];
        bytes_to_f32<0>(o.lo, q);
        bytes_to_f32<0>(o.hi, q + 4);
        if constexpr (TYPE == T_IQ1_S || TYPE == T_IQ1_M) {
            const float dl = __fmul_rn(d, (float)o.ls);
            const float shift = o.neg_shift ? -1.125f : -0.875f;
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = __fmul_rn(dl, __fadd_rn(q[i], shift));
        } else {
            constexpr float kStep = (TYPE == T_IQ3_XXS || TYPE == T_IQ3_S) ? 0.5f : 0.25f;
            const float db = __fmul_rn(__fmul_rn(d, __fadd_rn(0.5f, (float)o.ls)), kStep);
#pragma unroll
            for (int i = 0; i < 8; i++)
                y[i] = __uint_as_float(__float_as_uint(__fmul_rn(db, q[i])) ^ (((o.sign8 >> i) & 1u) << 31));
        }
    }
};
template <> struct Unpack8<T_IQ2_XXS> : UnpackIq8<T_IQ2_XXS> {};
template <> struct Unpack8<T_IQ2_XS>  : UnpackIq8<T_IQ2_XS> {};
template <> struct Unpack8<T_IQ2_S>   : UnpackIq8<T_IQ2_S> {};
template <> struct Unpack8<T_IQ3_XXS> : UnpackIq8<T_IQ3_XXS> {};
template <> struct Unpack8<T_IQ3_S>   : UnpackIq8<T_IQ3_S> {};
template <> struct Unpack8<T_IQ1_S>   : UnpackIq8<T_IQ1_S> {};
template <> struct Unpack8<T_IQ1_M>   : UnpackIq8<T_IQ1_M> {};
