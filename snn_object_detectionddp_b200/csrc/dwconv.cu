// Depthwise 3x3 (stride 1, pad 1) convolution on NHWC bf16 -- the DWConv of the YOLO Detect head's class
// branch (ultralytics nn/modules/head.py `cv3`, instantiated at reference model.py:186).  HBM-bound
// (9 MACs per loaded element): plain coalesced 128-bit kernels, no tensor cores.
//   weights fp32 [9][C] (tap-major so consecutive threads read consecutive channels)
// Also the frame packer of the stand-in feature pyramid (space-to-depth 8x8).
#include "common.cuh"

namespace snn {

// 8 channels / thread
__global__ void __launch_bounds__(256)
dw3x3_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int NB, int H, int W,
                 int C) {
    const int c8 = C >> 3;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)NB * H * W * c8;
    if (idx >= total) return;
    const int cg = (int)(idx % c8);
    const long long pix = idx / c8;
    const int wx = (int)(pix % W), hy = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int h2 = hy + kh - 1;
        if (h2 < 0 || h2 >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int w2 = wx + kw - 1;
            if (w2 < 0 || w2 >= W) continue;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((n * H + h2) * W + w2) * C) + cg);
            const float4 wa = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C) + cg * 2);
            const float4 wb = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C) + cg * 2 + 1);
            acc[0] += bf16_lo(v.x) * wa.x; acc[1] += bf16_hi(v.x) * wa.y; acc[2] += bf16_lo(v.y) * wa.z; acc[3] += bf16_hi(v.y) * wa.w;
            acc[4] += bf16_lo(v.z) * wb.x; acc[5] += bf16_hi(v.z) * wb.y; acc[6] += bf16_lo(v.w) * wb.z; acc[7] += bf16_hi(v.w) * wb.w;
        }
    }
    float4* o = reinterpret_cast<float4*>(y + pix * C) + cg * 2;
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// dx[n,h,w,c] = sum_taps dy[n, h-(kh-1), w-(kw-1), c] * w[kh,kw,c]
__global__ void __launch_bounds__(256)
dw3x3_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, int NB,
                   int H, int W, int C) {
    const int c8 = C >> 3;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)NB * H * W * c8;
    if (idx >= total) return;
    const int cg = (int)(idx % c8);
    const long long pix = idx / c8;
    const int wx = (int)(pix % W), hy = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int h2 = hy - (kh - 1);
        if (h2 < 0 || h2 >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int w2 = wx - (kw - 1);
            if (w2 < 0 || w2 >= W) continue;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(dy + ((n * H + h2) * W + w2) * C) + cg);
            const float4 wa = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C) + cg * 2);
            const float4 wb = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C) + cg * 2 + 1);
            acc[0] += bf16_lo(v.x) * wa.x; acc[1] += bf16_hi(v.x) * wa.y; acc[2] += bf16_lo(v.y) * wa.z; acc[3] += bf16_hi(v.y) * wa.w;
            acc[4] += bf16_lo(v.z) * wb.x; acc[5] += bf16_hi(v.z) * wb.y; acc[6] += bf16_lo(v.w) * wb.z; acc[7] += bf16_hi(v.w) * wb.w;
        }
    }
    uint4 pk;
    pk.x = pack_bf16x2(acc[0], acc[1]); pk.y = pack_bf16x2(acc[2], acc[3]);
    pk.z = pack_bf16x2(acc[4], acc[5]); pk.w = pack_bf16x2(acc[6], acc[7]);
    *(reinterpret_cast<uint4*>(dx + pix * C) + cg) = pk;
}

// dw[tap][c] += sum_pixels dy[pix, c] * x[pix + tap, c];  block = (C/4 threads per pixel) x rows, smem reduce, atomics
__global__ void __launch_bounds__(256)
dw3x3_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw, int NB,
                   int H, int W, int C, int pix_per_block) {
    extern __shared__ float shw[];  // [9][C]
    const int tpp = C >> 2;
    const int rows = 256 / tpp;
    const int cg = threadIdx.x % tpp, row = threadIdx.x / tpp;
    for (int i = threadIdx.x; i < 9 * C; i += 256) shw[i] = 0.f;
    __syncthreads();
    if (row < rows) {
        const long long P = (long long)NB * H * W;
        const long long p0 = (long long)blockIdx.x * pix_per_block, p1 = min(P, p0 + (long long)pix_per_block);
        float acc[9][4];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
        for (long long p = p0 + row; p < p1; p += rows) {
            const int wx = (int)(p % W), hy = (int)((p / W) % H);
            const long long n = p / ((long long)W * H);
            const uint2 g = __ldg(reinterpret_cast<const uint2*>(dy + p * C) + cg);
            const float g4[4] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y), bf16_hi(g.y)};
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int h2 = hy + kh - 1;
                if (h2 < 0 || h2 >= H) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int w2 = wx + kw - 1;
                    if (w2 < 0 || w2 >= W) continue;
                    const uint2 v = __ldg(reinterpret_cast<const uint2*>(x + ((n * H + h2) * W + w2) * C) + cg);
                    acc[kh * 3 + kw][0] += g4[0] * bf16_lo(v.x); acc[kh * 3 + kw][1] += g4[1] * bf16_hi(v.x);
                    acc[kh * 3 + kw][2] += g4[2] * bf16_lo(v.y); acc[kh * 3 + kw][3] += g4[3] * bf16_hi(v.y);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int i = 0; i < 4; ++i) atomicAdd(&shw[t * C + cg * 4 + i], acc[t][i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * C; i += 256) atomicAdd(&dw[i], shw[i]);
}

// frames fp32 [B][T][3][H][W] (or [N][3][H][W] with T=1) -> bf16 NHWC [T*B][H/8][W/8][192], channel = c*64 + dy*8 + dx
__global__ void __launch_bounds__(256)
s2d8_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int T, int H, int W) {
    const int H8 = H >> 3, W8 = W >> 3;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;  // over (n, c, dy, i, j) with j fastest
    const long long total = (long long)B * T * 3 * 8 * H8 * W8;
    if (idx >= total) return;
    const int j = (int)(idx % W8);
    const int i = (int)((idx / W8) % H8);
    const int dy = (int)((idx / ((long long)W8 * H8)) % 8);
    const int c = (int)((idx / ((long long)W8 * H8 * 8)) % 3);
    const long long n = idx / ((long long)W8 * H8 * 24);  // folded index t*B + b
    const int t = (int)(n / B), b = (int)(n % B);
    const float* src = in + ((((long long)b * T + t) * 3 + c) * H + (i * 8 + dy)) * W + j * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), bb = __ldg(reinterpret_cast<const float4*>(src) + 1);
    uint4 pk;
    pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w); pk.z = pack_bf16x2(bb.x, bb.y); pk.w = pack_bf16x2(bb.z, bb.w);
    *reinterpret_cast<uint4*>(out + ((n * H8 + i) * W8 + j) * 192 + c * 64 + dy * 8) = pk;
}

int launch_dw3x3_fwd(const __nv_bfloat16* x, const float* w, float* y, int NB, int H, int W, int C, cudaStream_t st) {
    SNN_REQUIRE(C % 8 == 0, "dw3x3: C must be a multiple of 8");
    const long long total = (long long)NB * H * W * (C / 8);
    dw3x3_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, w, y, NB, H, W, C);
    return check_cuda(cudaGetLastError(), "dw3x3_fwd_kernel");
}
int launch_dw3x3_dgrad(const __nv_bfloat16* dy, const float* w, __nv_bfloat16* dx, int NB, int H, int W, int C, cudaStream_t st) {
    SNN_REQUIRE(C % 8 == 0, "dw3x3: C must be a multiple of 8");
    const long long total = (long long)NB * H * W * (C / 8);
    dw3x3_dgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dy, w, dx, NB, H, W, C);
    return check_cuda(cudaGetLastError(), "dw3x3_dgrad_kernel");
}
int launch_dw3x3_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int NB, int H, int W, int C, cudaStream_t st) {
    SNN_REQUIRE(C % 4 == 0 && C <= 1024, "dw3x3_wgrad: C must be a multiple of 4, <= 1024");
    const int rows = 256 / (C / 4);
    const long long P = (long long)NB * H * W;
    long long want = (long long)num_sms() * 4;
    long long ppb = (P + want - 1) / want;
    if (ppb < rows * 4) ppb = rows * 4;
    ppb = (ppb + rows - 1) / rows * rows;
    dw3x3_wgrad_kernel<<<(unsigned)((P + ppb - 1) / ppb), 256, sizeof(float) * 9 * C, st>>>(x, dy, dw, NB, H, W, C, (int)ppb);
    return check_cuda(cudaGetLastError(), "dw3x3_wgrad_kernel");
}
// uint8 frames (what the dataset decodes, dataset.py:139-152) -> /255 on the device (the reference divides on the host and
// ships fp32: 4x the PCIe bytes); `v / 255.0f` in IEEE fp32 is bit-identical to torch's `.float() / 255.0`
__global__ void __launch_bounds__(256)
s2d8_u8_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int T, int H, int W) {
    const int H8 = H >> 3, W8 = W >> 3;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)B * T * 3 * 8 * H8 * W8;
    if (idx >= total) return;
    const int j = (int)(idx % W8);
    const int i = (int)((idx / W8) % H8);
    const int dy = (int)((idx / ((long long)W8 * H8)) % 8);
    const int c = (int)((idx / ((long long)W8 * H8 * 8)) % 3);
    const long long n = idx / ((long long)W8 * H8 * 24);
    const int t = (int)(n / B), b = (int)(n % B);
    const uint8_t* src = in + ((((long long)b * T + t) * 3 + c) * H + (i * 8 + dy)) * W + j * 8;
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src));
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = __fdiv_rn((float)((raw.x >> (8 * k)) & 0xFFu), 255.0f);
        v[4 + k] = __fdiv_rn((float)((raw.y >> (8 * k)) & 0xFFu), 255.0f);
    }
    uint4 pk;
    pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]); pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + ((n * H8 + i) * W8 + j) * 192 + c * 64 + dy * 8) = pk;
}
int launch_s2d8_u8(const uint8_t* in, __nv_bfloat16* out, int B, int T, int H, int W, cudaStream_t st) {
    SNN_REQUIRE(H % 8 == 0 && W % 8 == 0, "space_to_depth8: H, W must be multiples of 8");
    const long long total = (long long)B * T * 3 * 8 * (H / 8) * (W / 8);
    s2d8_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, B, T, H, W);
    return check_cuda(cudaGetLastError(), "s2d8_u8_kernel");
}
int launch_s2d8(const float* in, __nv_bfloat16* out, int B, int T, int H, int W, cudaStream_t st) {
    SNN_REQUIRE(H % 8 == 0 && W % 8 == 0, "space_to_depth8: H, W must be multiples of 8");
    const long long total = (long long)B * T * 3 * 8 * (H / 8) * (W / 8);
    s2d8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, B, T, H, W);
    return check_cuda(cudaGetLastError(), "s2d8_kernel");
}

}  // namespace snn
