// sr_resolve.cu -- the step after the raytrace path: Renderer.PostProcessImage
// (Renderer.cs:819-898, the styles that need no rasteriser depth buffer) followed by
// Renderer.AntiAliasImage (Renderer.cs:937-978), fused into one pass so the frame leaves the GPU final.
// Pure integer / byte work, HBM-bound: every source pixel is read once (coalesced 128-bit loads when
// the super-sampling factor allows), every destination pixel written once.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/softray_cuda.h"

namespace sr {

// Renderer.PostProcessImage's per-pixel colour functions (Surface.ApplyColorFunc, Surface.cs:226-233)
__device__ __forceinline__ uint32_t style_pixel(uint32_t x, int style, uint32_t background)
{
    if (style == SOFTRAY_STYLE_COLOR_SHUFFLE) return ((x & 0xffffu) << 8) + ((x >> 16) & 0xffu);   // ZRGB -> 0GBR (:829)
    if (style == SOFTRAY_STYLE_NEGATIVE) return x == background ? background : 0x00ffffffu - x;    // uint wrap (:833)
    return x;
}

// one thread per destination pixel; AA x AA source pixels each (Renderer.cs:946-975)
template <int AA>
__global__ void __launch_bounds__(256)
resolve_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int dst_w, int dst_h, int aa_rt, int style,
               uint32_t background)
{
    const int aa = AA > 0 ? AA : aa_rt;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dst_w || y >= dst_h) return;
    const size_t src_w = (size_t)dst_w * (size_t)aa;
    if (aa == 1) {
        dst[(size_t)y * dst_w + x] = style_pixel(src[(size_t)y * dst_w + x], style, background);
        return;
    }
    int sum_r = 0, sum_g = 0, sum_b = 0;
    for (int sy = 0; sy < aa; sy++) {
        const uint32_t* row = src + ((size_t)y * aa + sy) * src_w + (size_t)x * aa;
        if (AA == 2) {
            const uint2 p = *reinterpret_cast<const uint2*>(row);
            const uint32_t a = style_pixel(p.x, style, background), b = style_pixel(p.y, style, background);
            sum_r += (int)((a >> 16) & 0xff) + (int)((b >> 16) & 0xff);
            sum_g += (int)((a >> 8) & 0xff) + (int)((b >> 8) & 0xff);
            sum_b += (int)(a & 0xff) + (int)(b & 0xff);
        } else if (AA == 4) {
            const uint4 p = *reinterpret_cast<const uint4*>(row);
            const uint32_t v[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t a = style_pixel(v[k], style, background);
                sum_r += (int)((a >> 16) & 0xff); sum_g += (int)((a >> 8) & 0xff); sum_b += (int)(a & 0xff);
            }
        } else {
            for (int sx = 0; sx < aa; sx++) {
                const uint32_t a = style_pixel(row[sx], style, background);
                sum_r += (int)((a >> 16) & 0xff); sum_g += (int)((a >> 8) & 0xff); sum_b += (int)(a & 0xff);   // Surface.UnpackRgb
            }
        }
    }
    const int nn = aa * aa;
    sum_r /= nn; sum_g /= nn; sum_b /= nn;                                               // :967-969
    dst[(size_t)y * dst_w + x] = (255u << 24) + ((uint32_t)(sum_r & 0xff) << 16) + ((uint32_t)(sum_g & 0xff) << 8) +
                                 (uint32_t)(sum_b & 0xff);                               // Surface.PackRgb
}

// Vector path: one thread produces FOUR adjacent destination pixels (one 128-bit store) from AA rows of
// 4*AA source pixels (AA 128-bit loads per row): 16 B stores and >= 64 B of loads in flight per thread keep
// HBM busy even at AA = 1.  Needs dst_w % 4 == 0 and 16-byte aligned buffers.
template <int AA>
__global__ void __launch_bounds__(256)
resolve_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int dst_w4, int dst_h, int style, uint32_t background)
{
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= dst_w4 || y >= dst_h) return;
    const size_t src_row4 = (size_t)dst_w4 * AA;          // uint4 per source row
    int sum[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    uint32_t styled[4] = {0, 0, 0, 0};
#pragma unroll
    for (int sy = 0; sy < AA; sy++) {
        const uint4* row = src + ((size_t)y * AA + sy) * src_row4 + (size_t)x4 * AA;
#pragma unroll
        for (int j = 0; j < AA; j++) {
            const uint4 p = __ldg(row + j);
            const uint32_t v[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t a = style_pixel(v[k], style, background);
                const int d = (j * 4 + k) / AA;           // destination pixel this source pixel belongs to
                sum[d][0] += (int)((a >> 16) & 0xff); sum[d][1] += (int)((a >> 8) & 0xff); sum[d][2] += (int)(a & 0xff);
                if (AA == 1) styled[k] = a;
            }
        }
    }
    uint32_t o[4];
#pragma unroll
    for (int d = 0; d < 4; d++) {
        if (AA == 1) { o[d] = styled[d]; continue; }      // aa_res == 1: the styled pixel, alpha untouched
        const int nn = AA * AA;
        o[d] = (255u << 24) + ((uint32_t)((sum[d][0] / nn) & 0xff) << 16) + ((uint32_t)((sum[d][1] / nn) & 0xff) << 8) +
               (uint32_t)((sum[d][2] / nn) & 0xff);
    }
    dst[(size_t)y * dst_w4 + x4] = make_uint4(o[0], o[1], o[2], o[3]);
}

cudaError_t launch_resolve(const uint32_t* d_src, uint32_t* d_dst, int dst_w, int dst_h, int aa, int style, uint32_t background,
                           cudaStream_t stream)
{
    const bool vec = (dst_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_src) | reinterpret_cast<uintptr_t>(d_dst)) % 16 == 0) &&
                     (aa == 1 || aa == 2 || aa == 4);
    if (vec) {
        const dim3 block(32, 8);
        const dim3 grid((unsigned)((dst_w / 4 + 31) / 32), (unsigned)((dst_h + 7) / 8));
        const uint4* s4 = reinterpret_cast<const uint4*>(d_src);
        uint4* d4 = reinterpret_cast<uint4*>(d_dst);
        if (aa == 1) resolve_vec_kernel<1><<<grid, block, 0, stream>>>(s4, d4, dst_w / 4, dst_h, style, background);
        else if (aa == 2) resolve_vec_kernel<2><<<grid, block, 0, stream>>>(s4, d4, dst_w / 4, dst_h, style, background);
        else resolve_vec_kernel<4><<<grid, block, 0, stream>>>(s4, d4, dst_w / 4, dst_h, style, background);
        return cudaGetLastError();
    }
    const dim3 block(32, 8);
    const dim3 grid((unsigned)((dst_w + 31) / 32), (unsigned)((dst_h + 7) / 8));
    if (aa == 2) resolve_kernel<2><<<grid, block, 0, stream>>>(d_src, d_dst, dst_w, dst_h, aa, style, background);
    else if (aa == 4) resolve_kernel<4><<<grid, block, 0, stream>>>(d_src, d_dst, dst_w, dst_h, aa, style, background);
    else resolve_kernel<0><<<grid, block, 0, stream>>>(d_src, d_dst, dst_w, dst_h, aa, style, background);
    return cudaGetLastError();
}

}  // namespace sr
