// Task-aligned target assignment for the detection loss, as three small kernels.
//
// Replaces the TaskAlignedAssigner inside ultralytics' v8DetectionLoss (called at reference train.py:74; un-vendored third
// party, published 8.3.x algorithm restated -- PARITY UNPINNED, oracle/detect_oracle.py).  Round 1 expressed it as ~60
// dense ATen launches (topk, scatter_add_, gather, argmax ...) inside the captured training step; this is the same
// arithmetic in 3 launches, no host synchronisation, deterministic (no float atomics: the only atomics are integer max).
//
//   per image b, ground truth m, anchor a (A anchors, M padded ground-truth slots):
//     valid(m,a)  = anchor centre strictly inside box m (min side distance > 1e-9) and slot m holds a box
//     ov(m,a)     = clamp(CIoU(box_m, pred_box_a), 0) * valid          align(m,a) = sqrt(score_{a,label_m}) * ov^6 * valid
//     cand(m)     = the top-k (10) anchors of align(m, :), ties to the lower index; a candidate counts if valid
//     an anchor claimed by several m keeps argmax_m ov(m,a) (first maximum); fg(a) = claimed
//     target score of a foreground anchor = align(m*,a) * max_a' ov(m*,a') / (max_a' align(m*,a') + 1e-9)  over the anchors
//     finally assigned to m*, placed at class label_m*
//
// Layout: predictions are read through a ROW MAP: the Detect head writes its three scales into scale-major buffers
// (scale i occupies rows [B*a_off_i, B*a_off_{i+1}), image-major inside), so no torch.cat ever runs; row(b, a) below.
#include "common.cuh"

namespace snn {

struct GtBox {
    float x1, y1, x2, y2;
    int ok, label;
};
// padded labels (reference train.py:27-37 collate layout, densified): cls int64 [B,M], box fp32 [B,M,4] = normalised
// (cx, cy, w, h), valid uint8 [B,M].  Same arithmetic as the torch formulation: xy*scale, wh*scale/2, * valid, sum > 0.
SNN_DEVINL GtBox load_gt(const long long* __restrict__ cls, const float* __restrict__ box, const uint8_t* __restrict__ valid,
                         int b, int m, int M, float img_w, float img_h) {
    GtBox g;
    const long long i = (long long)b * M + m;
    const float v = valid[i] ? 1.f : 0.f;
    const float cx = box[i * 4] * img_w, cy = box[i * 4 + 1] * img_h;
    const float hw = box[i * 4 + 2] * img_w / 2.f, hh = box[i * 4 + 3] * img_h / 2.f;
    g.x1 = (cx - hw) * v; g.y1 = (cy - hh) * v; g.x2 = (cx + hw) * v; g.y2 = (cy + hh) * v;
    g.ok = (valid[i] != 0) && (((g.x1 + g.y1) + g.x2) + g.y2 > 0.f);
    g.label = (int)cls[i];
    return g;
}

// CIoU(box1 = ground truth, box2 = prediction), ultralytics utils/metrics.py bbox_iou(xywh=False, CIoU=True), eps 1e-7
SNN_DEVINL float ciou_plain(float x1, float y1, float x2, float y2, float X1, float Y1, float X2, float Y2) {
    const float eps = 1e-7f;
    const float w1 = x2 - x1, h1 = y2 - y1 + eps, w2 = X2 - X1, h2 = Y2 - Y1 + eps;
    const float iw = fmaxf(fminf(x2, X2) - fmaxf(x1, X1), 0.f), ih = fmaxf(fminf(y2, Y2) - fmaxf(y1, Y1), 0.f);
    const float inter = iw * ih;
    const float uni = w1 * h1 + w2 * h2 - inter + eps;
    const float iou = inter / uni;
    const float cw = fmaxf(x2, X2) - fminf(x1, X1), ch = fmaxf(y2, Y2) - fminf(y1, Y1);
    const float c2 = cw * cw + ch * ch + eps;
    const float dx = X1 + X2 - x1 - x2, dy = Y1 + Y2 - y1 - y2;
    const float rho2 = (dx * dx + dy * dy) / 4.f;
    const float da = atanf(w2 / h2) - atanf(w1 / h1);
    const float v = 0.40528473456935109f * (da * da);
    const float alpha = v / (v - iou + (1.f + eps));
    return iou - (rho2 / c2 + v * alpha);
}

constexpr int kTopkMax = 16;

// (value, index) arg-max with ties to the LOWER index
SNN_DEVINL void amax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// grid (M, B), 256 threads: metrics of one ground-truth slot against all anchors + its top-k candidates
__global__ void __launch_bounds__(256)
tal_metric_topk_kernel(const float* __restrict__ probs /*[B,A,nc]*/, const float* __restrict__ pboxes /*[B,A,4] xyxy px*/,
                       const float* __restrict__ anchors /*[A,2] grid*/, const float* __restrict__ stride /*[A]*/,
                       const long long* __restrict__ gt_cls, const float* __restrict__ gt_box, const uint8_t* __restrict__ gt_valid,
                       float img_w, float img_h, int A, int M, int nc, int topk,
                       float* __restrict__ align /*[B,M,A]*/, float* __restrict__ ovl /*[B,M,A]*/, int* __restrict__ sel /*[B,M,kTopkMax]*/,
                       unsigned int* __restrict__ pos /*[B,M,2]*/) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_v[8];
    __shared__ int s_i[8];
    __shared__ int s_sel[kTopkMax];
    const int m = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const long long bm = (long long)b * M + m;
    const GtBox g = load_gt(gt_cls, gt_box, gt_valid, b, m, M, img_w, img_h);
    if (tid < 2) pos[bm * 2 + tid] = 0u;
    if (!g.ok) {               // empty slot: its top-k indices all collapse to anchor 0 (count 10 != 1): no candidates
        if (tid < kTopkMax) sel[bm * kTopkMax + tid] = -1;
        return;
    }
    const int label = min(max(g.label, 0), nc - 1);
    float* al_row = align + bm * A;
    float* ov_row = ovl + bm * A;
    for (int a = tid; a < A; a += 256) {
        const float st = stride[a], ax = anchors[2 * a] * st, ay = anchors[2 * a + 1] * st;
        const float dmin = fminf(fminf(ax - g.x1, ay - g.y1), fminf(g.x2 - ax, g.y2 - ay));
        float o = 0.f, al = 0.f;
        if (dmin > 1e-9f) {
            const float4 pb = __ldg(reinterpret_cast<const float4*>(pboxes) + (long long)b * A + a);
            o = fmaxf(ciou_plain(g.x1, g.y1, g.x2, g.y2, pb.x, pb.y, pb.z, pb.w), 0.f);
            const float sc = probs[((long long)b * A + a) * nc + label];
            al = sqrtf(sc) * powf(o, 6.0f);
        }
        al_row[a] = al;
        ov_row[a] = o;
    }
    __syncthreads();
    const int k = min(min(topk, kTopkMax), A);
    for (int r = 0; r < k; ++r) {
        float bv = -1.f;
        int bi = 0x7fffffff;
        for (int a = tid; a < A; a += 256) {
            bool taken = false;
            for (int j = 0; j < r; ++j) taken |= (s_sel[j] == a);
            if (!taken) amax_combine(bv, bi, al_row[a], a);
        }
        for (int o = 16; o; o >>= 1) {
            const float ov2 = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi2 = __shfl_xor_sync(0xffffffffu, bi, o);
            amax_combine(bv, bi, ov2, oi2);
        }
        if ((tid & 31) == 0) { s_v[tid >> 5] = bv; s_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w) amax_combine(bv, bi, s_v[w], s_i[w]);
            s_sel[r] = bi;
        }
        __syncthreads();
    }
    if (tid < kTopkMax) {
        int out = -1;
        if (tid < k) {
            const int a = s_sel[tid];
            const float st = stride[a], ax = anchors[2 * a] * st, ay = anchors[2 * a + 1] * st;
            const float dmin = fminf(fminf(ax - g.x1, ay - g.y1), fminf(g.x2 - ax, g.y2 - ay));
            out = dmin > 1e-9f ? a : -1;          // mask_pos = (count == 1) & valid
        }
        sel[bm * kTopkMax + tid] = out;
    }
}

// grid (ceil(A/128), B): resolve anchors claimed by several ground truths, record per-gt maxima (integer atomicMax on the
// bit patterns of non-negative floats: order-independent -> deterministic)
__global__ void __launch_bounds__(128)
tal_resolve_kernel(const long long* __restrict__ gt_cls, const float* __restrict__ gt_box, const uint8_t* __restrict__ gt_valid,
                   float img_w, float img_h, int A, int M, const float* __restrict__ align, const float* __restrict__ ovl,
                   const int* __restrict__ sel, unsigned int* __restrict__ pos, int* __restrict__ tgt /*[B,A]*/) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ int s_sel[];      // [M][kTopkMax] then ok flags [M]
    int* s_ok = s_sel + M * kTopkMax;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < M * kTopkMax; i += 128) s_sel[i] = sel[(long long)b * M * kTopkMax + i];
    for (int m = threadIdx.x; m < M; m += 128) s_ok[m] = load_gt(gt_cls, gt_box, gt_valid, b, m, M, img_w, img_h).ok;
    __syncthreads();
    const int a = blockIdx.x * 128 + threadIdx.x;
    if (a >= A) return;
    int cnt = 0, first = -1;
    for (int m = 0; m < M; ++m) {
        if (!s_ok[m]) continue;
        bool hit = false;
        for (int r = 0; r < kTopkMax; ++r) hit |= (s_sel[m * kTopkMax + r] == a);
        if (hit) { ++cnt; if (first < 0) first = m; }
    }
    int ms = first;
    if (cnt > 1) {             // overlaps.argmax over ALL slots (empty / non-claiming ones contribute their masked value)
        float best = -1.f;
        for (int m = 0; m < M; ++m) {
            const float o = s_ok[m] ? ovl[((long long)b * M + m) * A + a] : 0.f;
            if (o > best) { best = o; ms = m; }
        }
    }
    tgt[(long long)b * A + a] = cnt > 0 ? ms : -1;
    if (cnt > 0) {
        const long long bm = (long long)b * M + ms;
        // (a slot that did not claim the anchor but won the arg-max may be an empty one: its rows were never written)
        const float al = s_ok[ms] ? align[bm * A + a] : 0.f, o = s_ok[ms] ? ovl[bm * A + a] : 0.f;
        atomicMax(&pos[bm * 2 + 0], __float_as_uint(al));
        atomicMax(&pos[bm * 2 + 1], __float_as_uint(o));
    }
}

// grid (ceil(A/128), B): materialise target boxes (xyxy px), target scores [B,A,nc], fg, and per-block partial sums of the
// target scores (their total normalises the loss; summed later in a fixed order)
__global__ void __launch_bounds__(128)
tal_targets_kernel(const long long* __restrict__ gt_cls, const float* __restrict__ gt_box, const uint8_t* __restrict__ gt_valid,
                   float img_w, float img_h, int A, int M, int nc, const float* __restrict__ align, const unsigned int* __restrict__ pos,
                   const int* __restrict__ tgt, float* __restrict__ tbox /*[B,A,4]*/, float* __restrict__ tscores /*[B,A,nc]*/,
                   uint8_t* __restrict__ fg /*[B,A]*/, float* __restrict__ tss_part /*[B*gridDim.x]*/) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[4];
    const int b = blockIdx.y;
    const int a = blockIdx.x * 128 + threadIdx.x;
    float norm = 0.f;
    if (a < A) {
        const long long n = (long long)b * A + a;
        const int m = tgt[n];
        const GtBox g = load_gt(gt_cls, gt_box, gt_valid, b, m >= 0 ? m : 0, M, img_w, img_h);
        int label = -1;
        if (m >= 0) {
            const long long bm = (long long)b * M + m;
            const float al = g.ok ? align[bm * A + a] : 0.f;
            norm = al * __uint_as_float(pos[bm * 2 + 1]) / (__uint_as_float(pos[bm * 2 + 0]) + 1e-9f);
            label = min(max(g.label, 0), nc - 1);
        }
        reinterpret_cast<float4*>(tbox)[n] = make_float4(g.x1, g.y1, g.x2, g.y2);
        for (int c = 0; c < nc; ++c) tscores[n * nc + c] = (c == label) ? norm : 0.f;
        fg[n] = m >= 0 ? 1 : 0;
    }
    // fixed-order block sum
    float v = norm;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) tss_part[(long long)b * gridDim.x + blockIdx.x] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

// workspace floats / ints the caller allocates: align + ovl [B,M,A] fp32 each, sel [B,M,16] int32, pos [B,M,2] uint32,
// tgt [B,A] int32, tss_part [B*ceil(A/128)] fp32
long long tal_workspace_bytes(int B, int A, int M) {
    const long long blocks = (A + 127) / 128;
    return 4LL * (2LL * B * M * A + (long long)B * M * kTopkMax + 2LL * B * M + (long long)B * A + (long long)B * blocks);
}

int launch_tal_assign(const float* probs, const float* pboxes, const float* anchors, const float* stride, const long long* gt_cls,
                      const float* gt_box, const uint8_t* gt_valid, float img_w, float img_h, int B, int A, int M, int nc, int topk,
                      void* workspace, float* tbox, float* tscores, uint8_t* fg, float** tss_part_out, int* tss_parts, cudaStream_t st) {
    SNN_REQUIRE(B >= 1 && A >= 1 && M >= 1 && nc >= 1, "tal_assign: bad sizes (B=%d A=%d M=%d nc=%d)", B, A, M, nc);
    SNN_REQUIRE(topk >= 1 && topk <= kTopkMax, "tal_assign: topk=%d must be in [1,%d]", topk, kTopkMax);
    SNN_REQUIRE(M <= 1024, "tal_assign: at most 1024 ground-truth slots per image (got %d)", M);
    SNN_REQUIRE(((uintptr_t)pboxes & 15) == 0 && ((uintptr_t)tbox & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
                "tal_assign: pointers must be 16-byte aligned");
    const int blocks = (A + 127) / 128;
    float* align = reinterpret_cast<float*>(workspace);
    float* ovl = align + (long long)B * M * A;
    int* sel = reinterpret_cast<int*>(ovl + (long long)B * M * A);
    unsigned int* pos = reinterpret_cast<unsigned int*>(sel + (long long)B * M * kTopkMax);
    int* tgt = reinterpret_cast<int*>(pos + 2LL * B * M);
    float* tss_part = reinterpret_cast<float*>(tgt + (long long)B * A);
    launch_pdl(tal_metric_topk_kernel, dim3(M, B), dim3(256), 0, st, probs, pboxes, anchors, stride, gt_cls, gt_box, gt_valid, img_w, img_h, A, M, nc,
                                                       topk, align, ovl, sel, pos);
    SNN_CUDA_OK(cudaGetLastError());
    const size_t smem = sizeof(int) * ((size_t)M * kTopkMax + M);
    if (smem > 48 * 1024) {
        static PerDeviceOnce once;
        SNN_CUDA_OK(once.run([] { return cudaFuncSetAttribute(tal_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); }));
    }
    launch_pdl(tal_resolve_kernel, dim3(blocks, B), dim3(128), smem, st, gt_cls, gt_box, gt_valid, img_w, img_h, A, M, align, ovl, sel, pos, tgt);
    SNN_CUDA_OK(cudaGetLastError());
    launch_pdl(tal_targets_kernel, dim3(blocks, B), dim3(128), 0, st, gt_cls, gt_box, gt_valid, img_w, img_h, A, M, nc, align, pos, tgt, tbox, tscores, fg,
                                                        tss_part);
    if (tss_part_out) *tss_part_out = tss_part;
    if (tss_parts) *tss_parts = B * blocks;
    return check_cuda(cudaGetLastError(), "tal_assign kernels");
}

}  // namespace snn
