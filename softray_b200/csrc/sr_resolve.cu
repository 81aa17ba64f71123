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

cudaError_t launch_resolve(const uint32_t* d_src, uint32_t* d_dst, int dst_w, int dst_h, int aa, int style, uint32_t background,
                           cudaStream_t stream)
{
    const dim3 block(32, 8);
    const dim3 grid((unsigned)((dst_w + 31) / 32), (unsigned)((dst_h + 7) / 8));
    if (aa == 2) resolve_kernel<2><<<grid, block, 0, stream>>>(d_src, d_dst, dst_w, dst_h, aa, style, background);
    else if (aa == 4) resolve_kernel<4><<<grid, block, 0, stream>>>(d_src, d_dst, dst_w, dst_h, aa, style, background);
    else resolve_kernel<0><<<grid, block, 0, stream>>>(d_src, d_dst, dst_w, dst_h, aa, style, background);
    return cudaGetLastError();
}

}  // namespace sr
