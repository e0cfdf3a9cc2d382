// Depthwise 3x3 (stride 1, pad 1) convolution on NHWC bf16 -- the DWConv of the YOLO Detect head's class
// branch (ultralytics nn/modules/head.py `cv3`, instantiated at reference model.py:186).  HBM-bound
// (9 MACs per loaded element): plain coalesced 128-bit kernels, no tensor cores.
//   weights fp32 [9][C] (tap-major so consecutive threads read consecutive channels)
// Also the frame packer of the stand-in feature pyramid (space-to-depth 8x8).
#include "common.cuh"

namespace snn {

// ------------------------------------------------------------------------------------------
// Column-walking kernels.  A thread owns 8 channels of ONE image column (n, w) and walks h = 0..H-1 with a 3x3 window of
// packed bf16 in registers: every step loads the three 16-byte pieces of the next input row (the two neighbours' pieces
// are L1 hits, the neighbouring columns live in the same block) instead of nine, the nine tap weights sit in shared memory.
// Round 1 used one thread per output pixel: 9 activation + 18 weight loads per 8 outputs -> 1.6 TB/s (24 % of HBM).
//   fprop : y fp32 = conv(x bf16) (+ per-timestep BatchNorm partial sums in the same pass: the separate snn_bn_stats pass over
//           y, 4 B/element, is gone; partials [T][blocks][2][C], one per block, reduced in a fixed order)
//   dgrad : dx bf16 = conv(dy bf16, flipped taps)
//   wgrad : dw[tap][c] += sum_pixels dy * x[pixel + tap] (72 register accumulators per thread, fixed-order block reduce,
//           one atomicAdd per (tap, channel) and block)
// ------------------------------------------------------------------------------------------
SNN_DEVINL void unpack8f(const uint4 v, float* f) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
SNN_DEVINL uint4 ld_px(const __nv_bfloat16* __restrict__ base, long long row_off, int w, int W, int C, int cg) {
    if (w < 0 || w >= W) return make_uint4(0u, 0u, 0u, 0u);
    return __ldg(reinterpret_cast<const uint4*>(base + row_off + (long long)w * C) + cg);
}

// One input row (3 neighbouring pixels x 8 channels) unpacked ONCE to fp32 channel pairs; the 3-row window lives in registers
// and the row loop is unrolled by three so that the rows rotate by renaming, not by register moves.  The first version of this
// kernel kept the window packed and unpacked every tap (72 shift/and + 72 scalar FMA per output row): ncu showed 360
// warp-instructions per row and 42 % issue utilisation at 14 % of DRAM bandwidth -- instruction bound.  Now 24 unpack +
// 36 packed FFMA2 per row.
struct DwRow { float2 v[3][2]; };             // 3 neighbouring pixels x 4 channels (two channel pairs)
// col[k]: this thread's 8-byte piece of pixel (row 0, w + k - 1), or nullptr when that column is outside the map;
// off: row offset in uint2 units.  (Pointers and validity are hoisted out of the row loop: the first cut recomputed 64-bit
// addresses and bounds per load and spent 120 warp-instructions per output row.)
SNN_DEVINL void dw_load_row(DwRow& r, const uint2* const col[3], long long off, bool valid) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint2 p = make_uint2(0u, 0u);
        if (valid && col[k] != nullptr) p = __ldg(col[k] + off);
        r.v[k][0] = make_float2(bf16_lo(p.x), bf16_hi(p.x));
        r.v[k][1] = make_float2(bf16_lo(p.y), bf16_hi(p.y));
    }
}
SNN_DEVINL void dw_mac_row(float2 acc[2], const DwRow& r, const float* __restrict__ shw, int kh, int C, int cg) {
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
        const float4 wa = *reinterpret_cast<const float4*>(shw + (kh * 3 + kw) * C + cg * 4);
        acc[0] = __ffma2_rn(r.v[kw][0], make_float2(wa.x, wa.y), acc[0]);
        acc[1] = __ffma2_rn(r.v[kw][1], make_float2(wa.z, wa.w), acc[1]);
    }
}

// grid (ceil(cols_per_group / slots), groups): group = timestep (fprop with statistics) or 1; a column = (image, w);
// a thread owns 4 channels of one column (8-byte loads: 36 window registers, 3 blocks of 256 threads per SM)
template <bool DGRAD>
__global__ void __launch_bounds__(256)
dw3x3_col_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt, float* __restrict__ y32,
                 __nv_bfloat16* __restrict__ y16, float* __restrict__ part, int imgs_per_group, int H, int W, int C) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sh[];                 // [9][C] weights, then (statistics) [slots][2][C]
    const int c4 = C >> 2, slots = 256 / c4;
    float* shw = sh;
    float* shs = sh + 9 * C;
    for (int i = threadIdx.x; i < 9 * C; i += 256) {
        const int tap = i / C, c = i % C;
        shw[i] = wgt[(DGRAD ? 8 - tap : tap) * C + c];            // dgrad: flipped taps
    }
    __syncthreads();
    const int slot = threadIdx.x / c4, cg = threadIdx.x % c4;
    const int cols = imgs_per_group * W;
    const int col = blockIdx.x * slots + slot;
    float2 ssum[2], ssq[2];
    ssum[0] = ssum[1] = ssq[0] = ssq[1] = make_float2(0.f, 0.f);
    if (slot < slots && col < cols) {
        const int w = col % W;
        const long long n = (long long)blockIdx.y * imgs_per_group + col / W;
        const long long rs4 = ((long long)W * C) >> 2;                 // row stride in 4-channel units
        const long long base4 = (n * H * W + w) * (long long)(C >> 2) + cg;
        const uint2* xin = reinterpret_cast<const uint2*>(x);
        const uint2* colp[3] = {w > 0 ? xin + base4 - (C >> 2) : nullptr, xin + base4, w + 1 < W ? xin + base4 + (C >> 2) : nullptr};
        const bool want_stats = !DGRAD && part != nullptr;
        long long o4 = base4;                                          // output piece of row h (same layout as the input)

        auto emit = [&](const DwRow& a, const DwRow& b, const DwRow& c) {      // rows h-1, h, h+1 -> output row h
            float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            dw_mac_row(acc, a, shw, 0, C, cg);
            dw_mac_row(acc, b, shw, 1, C, cg);
            dw_mac_row(acc, c, shw, 2, C, cg);
            if (DGRAD) {
                reinterpret_cast<uint2*>(y16)[o4] = make_uint2(pack_bf16x2(acc[0].x, acc[0].y), pack_bf16x2(acc[1].x, acc[1].y));
            } else {
                reinterpret_cast<float4*>(y32)[o4] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
                if (want_stats) {
                    ssum[0] = __fadd2_rn(ssum[0], acc[0]); ssq[0] = __ffma2_rn(acc[0], acc[0], ssq[0]);
                    ssum[1] = __fadd2_rn(ssum[1], acc[1]); ssq[1] = __ffma2_rn(acc[1], acc[1], ssq[1]);
                }
            }
            o4 += rs4;
        };

        DwRow r0, r1, r2;
        dw_load_row(r0, colp, 0, false);                               // row -1: zero padding
        dw_load_row(r1, colp, 0, true);
        long long off = rs4;                                           // offset of row h + 1
        int h = 0;
        for (; h + 2 < H; h += 3) {                                    // rows rotate by renaming
            dw_load_row(r2, colp, off, true);
            emit(r0, r1, r2);
            dw_load_row(r0, colp, off + rs4, true);
            emit(r1, r2, r0);
            dw_load_row(r1, colp, off + 2 * rs4, h + 3 < H);
            emit(r2, r0, r1);
            off += 3 * rs4;
        }
        if (h < H) {                                                   // 1 or 2 rows left; window is (r0, r1) = rows (h-1, h)
            dw_load_row(r2, colp, off, h + 1 < H);
            emit(r0, r1, r2);
            if (h + 1 < H) {
                dw_load_row(r0, colp, 0, false);                       // row H: zero padding
                emit(r1, r2, r0);
            }
        }
    }
    if (!DGRAD && part) {
        // per-block partial sums in a FIXED order (deterministic): slot rows in shared memory, column-summed by C threads
        if (slot < slots) {
            float* row = shs + (size_t)slot * 2 * C;
            row[cg * 4 + 0] = ssum[0].x; row[cg * 4 + 1] = ssum[0].y; row[cg * 4 + 2] = ssum[1].x; row[cg * 4 + 3] = ssum[1].y;
            row[C + cg * 4 + 0] = ssq[0].x; row[C + cg * 4 + 1] = ssq[0].y; row[C + cg * 4 + 2] = ssq[1].x; row[C + cg * 4 + 3] = ssq[1].y;
        }
        __syncthreads();
        float* dst = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C;
        for (int i = threadIdx.x; i < 2 * C; i += 256) {
            float a = 0.f;
            for (int sidx = 0; sidx < slots; ++sidx) a += shs[(size_t)sidx * 2 * C + i];
            dst[i] = a;
        }
    }
}

// wgrad: grid (blocks); a block walks columns blockIdx.x, blockIdx.x + gridDim.x, ... (slots columns at a time)
__global__ void __launch_bounds__(256)
dw3x3_wgrad_col_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw, int NB, int H,
                       int W, int C, float* __restrict__ part /* deterministic mode: [gridDim.x][9][C] block partials, else null */) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sh[];                 // [slots][3][C] (one tap row at a time)
    const int c8 = C >> 3, slots = 256 / c8;
    const int slot = threadIdx.x / c8, cg = threadIdx.x % c8;
    const long long cols = (long long)NB * W;
    float acc[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
    if (slot < slots) {
        for (long long col = (long long)blockIdx.x * slots + slot; col < cols; col += (long long)gridDim.x * slots) {
            const int w = (int)(col % W);
            const long long n = col / W;
            const __nv_bfloat16* img = x + n * H * W * C;
            const __nv_bfloat16* gimg = dy + n * H * W * C;
            const long long rs = (long long)W * C;
            uint4 r0[3], r1[3], r2[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { r0[k] = make_uint4(0u, 0u, 0u, 0u); r1[k] = ld_px(img, 0, w + k - 1, W, C, cg); }
            for (int h = 0; h < H; ++h) {
#pragma unroll
                for (int k = 0; k < 3; ++k) r2[k] = (h + 1 < H) ? ld_px(img, (long long)(h + 1) * rs, w + k - 1, W, C, cg) : make_uint4(0u, 0u, 0u, 0u);
                float g[8];
                unpack8f(__ldg(reinterpret_cast<const uint4*>(gimg + (long long)h * rs + (long long)w * C) + cg), g);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        float v[8];
                        unpack8f(kh == 0 ? r0[kw] : (kh == 1 ? r1[kw] : r2[kw]), v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[kh * 3 + kw][i] = fmaf(g[i], v[i], acc[kh * 3 + kw][i]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) { r0[k] = r1[k]; r1[k] = r2[k]; }
            }
        }
    }
    // block reduce, three taps at a time, fixed order; then one atomicAdd per (tap, channel) and block
    for (int kh = 0; kh < 3; ++kh) {
        __syncthreads();
        if (slot < slots) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                for (int i = 0; i < 8; ++i) sh[((size_t)slot * 3 + kw) * C + cg * 8 + i] = acc[kh * 3 + kw][i];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * C; i += 256) {
            float a = 0.f;
            for (int sidx = 0; sidx < slots; ++sidx) a += sh[(size_t)sidx * 3 * C + i];
            if (part) part[((size_t)blockIdx.x * 3 + kh) * 3 * C + i] = a;
            else atomicAdd(&dw[(size_t)kh * 3 * C + i], a);
        }
    }
}

// frames fp32 [B][T][3][H][W] (or [N][3][H][W] with T=1) -> bf16 NHWC [T*B][H/8][W/8][192], channel = c*64 + dy*8 + dx
__global__ void __launch_bounds__(256)
s2d8_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int T, int H, int W) {
    pdl_launch_dependents();
    pdl_wait();
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

// blocks the column kernels launch per statistics group (also the number of partial rows per timestep)
long long dw3x3_stats_blocks(int B, int W, int C) {
    if (C % 8 != 0 || C < 8 || C > 1024 || B < 1 || W < 1) return 0;
    const int slots = 256 / (C / 4);
    return ((long long)B * W + slots - 1) / slots;
}

// groups > 0: per-group (= per-timestep) BatchNorm partial sums -> part [groups][blocks][2][C] (blocks = dw3x3_stats_blocks)
int launch_dw3x3_fwd(const __nv_bfloat16* x, const float* w, float* y, int NB, int H, int W, int C, float* part, int groups, cudaStream_t st) {
    SNN_REQUIRE(C % 8 == 0 && C >= 8 && C <= 1024, "dw3x3: C=%d must be a multiple of 8 in [8,1024]", C);
    if (!part || groups < 1) groups = 1;
    SNN_REQUIRE(NB % groups == 0, "dw3x3: NB=%d is not a multiple of the %d statistics groups", NB, groups);
    const int B = NB / groups, slots = 256 / (C / 4);
    const size_t smem = sizeof(float) * (9 * (size_t)C + (part ? (size_t)slots * 2 * C : 0));
    static PerDeviceOnce once;
    SNN_CUDA_OK(once.run([] { return cudaFuncSetAttribute(dw3x3_col_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); }));
    dim3 grid((unsigned)dw3x3_stats_blocks(B, W, C), groups);
    launch_pdl(dw3x3_col_kernel<false>, grid, dim3(256), smem, st, x, w, y, nullptr, part, B, H, W, C);
    return check_cuda(cudaGetLastError(), "dw3x3_col_kernel<fprop>");
}
int launch_dw3x3_dgrad(const __nv_bfloat16* dy, const float* w, __nv_bfloat16* dx, int NB, int H, int W, int C, cudaStream_t st) {
    SNN_REQUIRE(C % 8 == 0 && C >= 8 && C <= 1024, "dw3x3: C=%d must be a multiple of 8 in [8,1024]", C);
    dim3 grid((unsigned)dw3x3_stats_blocks(NB, W, C), 1);
    launch_pdl(dw3x3_col_kernel<true>, grid, dim3(256), sizeof(float) * 9 * C, st, dy, w, nullptr, dx, nullptr, NB, H, W, C);
    return check_cuda(cudaGetLastError(), "dw3x3_col_kernel<dgrad>");
}
int launch_dw3x3_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int NB, int H, int W, int C, cudaStream_t st) {
    SNN_REQUIRE(C % 8 == 0 && C >= 8 && C <= 1024, "dw3x3_wgrad: C=%d must be a multiple of 8 in [8,1024]", C);
    const int slots = 256 / (C / 8);
    long long blocks = ((long long)NB * W + slots - 1) / slots;
    const long long cap = (long long)num_sms() * 2;
    if (blocks > cap) blocks = cap;
    const size_t smem = sizeof(float) * (size_t)slots * 3 * C;
    static PerDeviceOnce once;
    SNN_CUDA_OK(once.run([] { return cudaFuncSetAttribute(dw3x3_wgrad_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); }));
    if (deterministic()) {
        float* part = static_cast<float*>(det_scratch(sizeof(float) * (size_t)blocks * 9 * C, st));
        if (!part) return 2;
        launch_pdl(dw3x3_wgrad_col_kernel, dim3((unsigned)blocks), dim3(256), smem, st, x, dy, dw, NB, H, W, C, part);
        SNN_CUDA_OK(cudaGetLastError());
        return launch_ordered_combine_f32(part, (int)blocks, 9LL * C, dw, st);
    }
    launch_pdl(dw3x3_wgrad_col_kernel, dim3((unsigned)blocks), dim3(256), smem, st, x, dy, dw, NB, H, W, C, (float*)nullptr);
    return check_cuda(cudaGetLastError(), "dw3x3_wgrad_col_kernel");
}
// uint8 frames (what the dataset decodes, dataset.py:139-152) -> /255 on the device (the reference divides on the host and
// ships fp32: 4x the PCIe bytes); `v / 255.0f` in IEEE fp32 is bit-identical to torch's `.float() / 255.0`
__global__ void __launch_bounds__(256)
s2d8_u8_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int T, int H, int W) {
    pdl_launch_dependents();
    pdl_wait();
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
    launch_pdl(s2d8_u8_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, in, out, B, T, H, W);
    return check_cuda(cudaGetLastError(), "s2d8_u8_kernel");
}
int launch_s2d8(const float* in, __nv_bfloat16* out, int B, int T, int H, int W, cudaStream_t st) {
    SNN_REQUIRE(H % 8 == 0 && W % 8 == 0, "space_to_depth8: H, W must be multiples of 8");
    const long long total = (long long)B * T * 3 * 8 * (H / 8) * (W / 8);
    launch_pdl(s2d8_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, in, out, B, T, H, W);
    return check_cuda(cudaGetLastError(), "s2d8_kernel");
}

}  // namespace snn
