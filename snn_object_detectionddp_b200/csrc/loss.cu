// Detection-loss tail as two fused kernels (forward sums, backward gradients).
//
// Replaces the per-anchor part of ultralytics' v8DetectionLoss called at reference train.py:74 (un-vendored third
// party, restated from its published algorithm -- PARITY UNPINNED, see oracle/detect_oracle.py):
//   cls : BCE-with-logits(pred_scores, target_scores) summed over [B, A, nc]
//   box : sum over foreground anchors of (1 - CIoU(pred_box, target_box)) * weight
//   dfl : sum over foreground anchors of mean_k[ CE(logits_k, tl)*wl + CE(logits_k, tr)*wr ] * weight
// with pred_box = dist2bbox(softmax-expectation of the 4 x reg_max DFL logits) in grid units,
// weight = sum_c target_scores.  The assigner that produces (fg, target boxes, target scores) runs before it.
//
// One thread per anchor for box/dfl (4*reg_max logits in registers, CIoU differentiated with 4-tangent dual
// numbers), one thread per score element for BCE.  HBM-bound and tiny (B*A ~ 1e5 anchors): the point is launch
// count -- 2 launches instead of the ~60 eager kernels of the torch formulation.
#include "common.cuh"

namespace snn {

constexpr int kRegMax = 16;
constexpr float kEps = 1e-7f;

struct Dual {  // value + tangents w.r.t. pred (x1, y1, x2, y2)
    float v, d[4];
};
SNN_DEVINL Dual dconst(float c) { return Dual{c, {0.f, 0.f, 0.f, 0.f}}; }
SNN_DEVINL Dual dvar(float c, int i) { Dual r = dconst(c); r.d[i] = 1.f; return r; }
SNN_DEVINL Dual operator+(Dual a, Dual b) { Dual r; r.v = a.v + b.v; for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
SNN_DEVINL Dual operator-(Dual a, Dual b) { Dual r; r.v = a.v - b.v; for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
SNN_DEVINL Dual operator*(Dual a, Dual b) { Dual r; r.v = a.v * b.v; for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
SNN_DEVINL Dual operator/(Dual a, Dual b) {
    Dual r; r.v = a.v / b.v;
    const float ib = 1.f / b.v;
    for (int i = 0; i < 4; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
SNN_DEVINL Dual operator+(Dual a, float c) { a.v += c; return a; }
SNN_DEVINL Dual operator*(Dual a, float c) { a.v *= c; for (int i = 0; i < 4; ++i) a.d[i] *= c; return a; }
// min/max against a constant; torch.minimum/maximum backward splits the gradient evenly on ties
SNN_DEVINL Dual dmin(Dual a, float c) {
    if (a.v < c) return a;
    if (a.v > c) return dconst(c);
    Dual r = a * 0.5f; r.v = c; return r;
}
SNN_DEVINL Dual dmax(Dual a, float c) {
    if (a.v > c) return a;
    if (a.v < c) return dconst(c);
    Dual r = a * 0.5f; r.v = c; return r;
}
SNN_DEVINL Dual dclamp0(Dual a) { return a.v >= 0.f ? a : dconst(0.f); }  // clamp_(0): gradient passes at x >= 0
SNN_DEVINL Dual datan(Dual a) {
    Dual r; r.v = atanf(a.v);
    const float g = 1.f / (1.f + a.v * a.v);
    for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * g;
    return r;
}

// CIoU(box1 = pred (dual), box2 = target), ultralytics utils/metrics.py bbox_iou(xywh=False, CIoU=True)
SNN_DEVINL Dual ciou(const Dual b1[4], const float b2[4]) {
    const Dual w1 = b1[2] - b1[0], h1 = (b1[3] - b1[1]) + kEps;
    const float w2 = b2[2] - b2[0], h2 = b2[3] - b2[1] + kEps;
    const Dual iw = dclamp0(dmin(b1[2], b2[2]) - dmax(b1[0], b2[0]));
    const Dual ih = dclamp0(dmin(b1[3], b2[3]) - dmax(b1[1], b2[1]));
    const Dual inter = iw * ih;
    const Dual uni = (w1 * h1 + dconst(w2 * h2)) - inter + kEps;
    const Dual iou = inter / uni;
    const Dual cw = dmax(b1[2], b2[2]) - dmin(b1[0], b2[0]);
    const Dual ch = dmax(b1[3], b2[3]) - dmin(b1[1], b2[1]);
    const Dual c2 = cw * cw + ch * ch + kEps;
    const Dual dx = dconst(b2[0] + b2[2]) - b1[0] - b1[2];
    const Dual dy = dconst(b2[1] + b2[3]) - b1[1] - b1[3];
    const Dual rho2 = (dx * dx + dy * dy) * 0.25f;
    const Dual da = dconst(atanf(w2 / h2)) - datan(w1 / h1);
    const Dual v = (da * da) * 0.40528473456935109f;  // 4 / pi^2
    const float alpha = v.v / (v.v - iou.v + (1.f + kEps));  // torch.no_grad()
    return iou - (rho2 / c2 + v * alpha);
}

struct AnchorTerms {
    float box, dfl;           // (1 - ciou) * weight ; dfl * weight
    float gdist_box[4];       // d box / d dist_k
    float p[4][kRegMax];      // softmax of each side
    float dist[4];
    float wl[4];
    int tl[4];
    float weight;
};

// returns false for background anchors
// n = natural index b*A + a of the targets, r = row of the prediction buffers (RowMap)
SNN_DEVINL bool anchor_terms(const float* __restrict__ distri, const float* __restrict__ tscores,
                             const float* __restrict__ tbox_px, const float* __restrict__ anchors,
                             const float* __restrict__ stride, const uint8_t* __restrict__ fg, long long n, long long r, int a, int nc,
                             AnchorTerms& o) {
    if (!fg[n]) return false;
    float wsum = 0.f;
    for (int c = 0; c < nc; ++c) wsum += tscores[n * nc + c];
    o.weight = wsum;
    const float ax = anchors[2 * a], ay = anchors[2 * a + 1], inv_s = 1.f / stride[a];
    float tb[4];
    for (int k = 0; k < 4; ++k) tb[k] = tbox_px[n * 4 + k] * inv_s;  // target_bboxes /= stride_tensor
    float lse[4];
    for (int k = 0; k < 4; ++k) {
        const float4* lp = reinterpret_cast<const float4*>(distri + r * (4 * kRegMax) + k * kRegMax);
        float l[kRegMax];
#pragma unroll
        for (int j = 0; j < kRegMax / 4; ++j) {
            const float4 t = __ldg(lp + j);
            l[4 * j] = t.x; l[4 * j + 1] = t.y; l[4 * j + 2] = t.z; l[4 * j + 3] = t.w;
        }
        float m = l[0];
#pragma unroll
        for (int j = 1; j < kRegMax; ++j) m = fmaxf(m, l[j]);
        float s = 0.f, e = 0.f;
#pragma unroll
        for (int j = 0; j < kRegMax; ++j) { o.p[k][j] = expf(l[j] - m); s += o.p[k][j]; }
        const float is = 1.f / s;
#pragma unroll
        for (int j = 0; j < kRegMax; ++j) { o.p[k][j] *= is; e += o.p[k][j] * (float)j; }
        o.dist[k] = e;
        lse[k] = m + logf(s);
        // DFL target for this side (bbox2dist + DFLoss clamp to reg_max - 1 - 0.01)
        float t = k == 0 ? ax - tb[0] : (k == 1 ? ay - tb[1] : (k == 2 ? tb[2] - ax : tb[3] - ay));
        t = fminf(fmaxf(t, 0.f), (float)(kRegMax - 1) - 0.01f);
        const int tl = (int)t;
        o.tl[k] = tl;
        o.wl[k] = (float)(tl + 1) - t;
        // CE(logits, j) = lse - l[j]
        float ll = 0.f, lr = 0.f;
#pragma unroll
        for (int j = 0; j < kRegMax; ++j) { if (j == tl) ll = l[j]; if (j == tl + 1) lr = l[j]; }
        lse[k] = (lse[k] - ll) * o.wl[k] + (lse[k] - lr) * (1.f - o.wl[k]);
    }
    o.dfl = 0.25f * (lse[0] + lse[1] + lse[2] + lse[3]) * wsum;
    Dual b1[4] = {dvar(ax - o.dist[0], 0), dvar(ay - o.dist[1], 1), dvar(ax + o.dist[2], 2), dvar(ay + o.dist[3], 3)};
    const Dual c = ciou(b1, tb);
    o.box = (1.f - c.v) * wsum;
    // d box / d dist_k : x1 = ax - d0, y1 = ay - d1, x2 = ax + d2, y2 = ay + d3
    o.gdist_box[0] = c.d[0] * wsum; o.gdist_box[1] = c.d[1] * wsum;
    o.gdist_box[2] = -c.d[2] * wsum; o.gdist_box[3] = -c.d[3] * wsum;
    return true;
}

SNN_DEVINL float block_sum(float v, float* sh) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[w] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}

// Finalisation data of the fused path (NULL tss_part = legacy: raw sums only).  The LAST block to finish (atomic ticket;
// the result does not depend on which block that is) turns the three sums into the loss the reference returns
// (train.py:74: `loss * B` and `loss.detach()`), normalised by max(sum of target scores, 1) and weighted by hyp.box/cls/dfl
// (config.yaml:33-37), and leaves the backward coefficients d(sum_k loss_k * B)/d(sum_k) = gain_k * B / tss.
struct LossFinal {
    const float* tss_part;     // per-block partial sums of the target scores (tal_targets_kernel), fixed-order total
    int n_parts;
    const float* gains;        // device [3]
    unsigned int* counter;     // zeroed; left zeroed
    float* out6;               // loss*B [3], loss [3]
    float* coef3;
    float batch;
    double* part;              // deterministic mode: [gridDim.x][3] block sums, added in block order (else null: fp64 atomics)
};

__global__ void __launch_bounds__(128)
detect_loss_fwd_kernel(const float* __restrict__ distri, const float* __restrict__ scores, const float* __restrict__ anchors,
                       const float* __restrict__ stride, const float* __restrict__ tbox_px, const float* __restrict__ tscores,
                       const uint8_t* __restrict__ fg, long long N, int A, int nc, double* __restrict__ sums, const RowMap rm,
                       const LossFinal fin) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sh[4];
    __shared__ unsigned int s_ticket;
    const long long n = (long long)blockIdx.x * 128 + threadIdx.x;
    float box = 0.f, dfl = 0.f, cls = 0.f;
    if (n < N) {
        const int a = (int)(n % A);
        const long long r = pred_row(rm, (int)(n / A), a, A);
        for (int c = 0; c < nc; ++c) {
            const float x = scores[r * nc + c], t = tscores[n * nc + c];
            cls += fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        }
        AnchorTerms at;
        if (anchor_terms(distri, tscores, tbox_px, anchors, stride, fg, n, r, a, nc, at)) { box = at.box; dfl = at.dfl; }
    }
    box = block_sum(box, sh);
    cls = block_sum(cls, sh);
    dfl = block_sum(dfl, sh);
    if (threadIdx.x == 0) {
        if (fin.part) {
            fin.part[(size_t)blockIdx.x * 3 + 0] = (double)box;
            fin.part[(size_t)blockIdx.x * 3 + 1] = (double)cls;
            fin.part[(size_t)blockIdx.x * 3 + 2] = (double)dfl;
        } else {
            atomicAdd(&sums[0], (double)box);
            atomicAdd(&sums[1], (double)cls);
            atomicAdd(&sums[2], (double)dfl);
        }
    }
    if (fin.tss_part == nullptr) return;
    if (threadIdx.x == 0) {
        __threadfence();
        s_ticket = atomicAdd(fin.counter, 1u);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1 || threadIdx.x != 0) return;
    __threadfence();
    double tss = 0.0;
    for (int i = 0; i < fin.n_parts; ++i) tss += (double)fin.tss_part[i];
    if (tss < 1.0) tss = 1.0;
    for (int k = 0; k < 3; ++k) {
        double s;
        if (fin.part) {
            s = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(&fin.part[(size_t)b * 3 + k]);
            sums[k] = s;
        } else {
            s = __ldcg(&sums[k]);
        }
        const float l = (float)(s / tss) * fin.gains[k];
        fin.out6[k] = l * fin.batch;
        fin.out6[3 + k] = l;
        fin.coef3[k] = (float)((double)fin.gains[k] * (double)fin.batch / tss);
    }
    *fin.counter = 0u;
}

// coef[3] (device) = dL/d(sum_box), dL/d(sum_cls), dL/d(sum_dfl); gout[3] (device, optional) multiplies them (upstream
// gradient of the three returned loss components).  Gradients are written at the prediction ROWS (RowMap), as fp32 or --
// for the fused head path, where they are the bf16 `dy` operands of the closing 1x1 convs' backward -- as bf16.
template <typename TG>
__global__ void __launch_bounds__(128)
detect_loss_bwd_kernel(const float* __restrict__ distri, const float* __restrict__ scores, const float* __restrict__ anchors,
                       const float* __restrict__ stride, const float* __restrict__ tbox_px, const float* __restrict__ tscores,
                       const uint8_t* __restrict__ fg, long long N, int A, int nc, const float* __restrict__ coef,
                       const float* __restrict__ gout, TG* __restrict__ g_distri, TG* __restrict__ g_scores, const RowMap rm) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n = (long long)blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    const int a = (int)(n % A);
    const long long r = pred_row(rm, (int)(n / A), a, A);
    float kb = coef[0], kc = coef[1], kd = coef[2];
    if (gout) { kb *= gout[0]; kc *= gout[1]; kd *= gout[2]; }
    for (int c = 0; c < nc; ++c) {
        const float x = scores[r * nc + c], t = tscores[n * nc + c];
        g_scores[r * nc + c] = (TG)(kc * (1.f / (1.f + expf(-x)) - t));
    }
    TG* gp = g_distri + r * (4 * kRegMax);
    AnchorTerms at;
    if (!anchor_terms(distri, tscores, tbox_px, anchors, stride, fg, n, r, a, nc, at)) {
#pragma unroll
        for (int j = 0; j < 4 * kRegMax; ++j) gp[j] = (TG)0.f;
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float gb = kb * at.gdist_box[k];      // d[(1 - ciou) * w] / d dist_k
        const float kdw = kd * 0.25f * at.weight;
#pragma unroll
        for (int j = 0; j < kRegMax; ++j) {
            const float p = at.p[k][j];
            float v = gb * p * ((float)j - at.dist[k]) + kdw * p;
            if (j == at.tl[k]) v -= kdw * at.wl[k];
            if (j == at.tl[k] + 1) v -= kdw * (1.f - at.wl[k]);
            gp[k * kRegMax + j] = (TG)v;
        }
    }
}

// Detect._inference / bbox_decode: boxes from the DFL logits (softmax expectation -> dist2bbox -> * stride) and sigmoid
// class scores.  xywh != 0: (cx, cy, w, h) as the eval-mode head returns (visualize.py:71-78 feeds it to NMS);
// xywh == 0: (x1, y1, x2, y2) as the assigner wants.  One thread per anchor.
__global__ void __launch_bounds__(128)
detect_decode_kernel(const float* __restrict__ distri, const float* __restrict__ scores, const float* __restrict__ anchors,
                     const float* __restrict__ stride, long long N, int A, int nc, int xywh, float* __restrict__ boxes,
                     float* __restrict__ probs, const RowMap rm) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n = (long long)blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    const int a = (int)(n % A);
    const long long r = pred_row(rm, (int)(n / A), a, A);
    float dist[4];
    for (int k = 0; k < 4; ++k) {
        const float4* lp = reinterpret_cast<const float4*>(distri + r * (4 * kRegMax) + k * kRegMax);
        float l[kRegMax];
#pragma unroll
        for (int j = 0; j < kRegMax / 4; ++j) {
            const float4 t = __ldg(lp + j);
            l[4 * j] = t.x; l[4 * j + 1] = t.y; l[4 * j + 2] = t.z; l[4 * j + 3] = t.w;
        }
        float m = l[0];
#pragma unroll
        for (int j = 1; j < kRegMax; ++j) m = fmaxf(m, l[j]);
        float s = 0.f, e = 0.f;
#pragma unroll
        for (int j = 0; j < kRegMax; ++j) { const float p = expf(l[j] - m); s += p; e += p * (float)j; }
        dist[k] = e / s;
    }
    const float ax = anchors[2 * a], ay = anchors[2 * a + 1], st = stride[a];
    const float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
    float4 o;
    if (xywh) o = make_float4((x1 + x2) * 0.5f * st, (y1 + y2) * 0.5f * st, (x2 - x1) * st, (y2 - y1) * st);
    else o = make_float4(x1 * st, y1 * st, x2 * st, y2 * st);
    reinterpret_cast<float4*>(boxes)[n] = o;
    if (probs)
        for (int c = 0; c < nc; ++c) probs[n * nc + c] = 1.f / (1.f + expf(-scores[r * nc + c]));
}

int make_rowmap(RowMap* rm, int nl, const int* a_off, int B, int A) {
    memset(rm, 0, sizeof(*rm));
    rm->B = B;
    if (nl == 0 || a_off == nullptr) return 0;
    SNN_REQUIRE(nl >= 1 && nl <= 4, "row map: 1..4 scales (got %d)", nl);
    for (int i = 0; i <= nl; ++i) rm->a_off[i] = a_off[i];
    SNN_REQUIRE(a_off[0] == 0 && a_off[nl] == A, "row map: anchor offsets must run from 0 to A=%d", A);
    for (int i = 0; i < nl; ++i) SNN_REQUIRE(a_off[i + 1] > a_off[i], "row map: anchor offsets must increase");
    rm->nl = nl;
    return 0;
}

int launch_detect_decode(const float* distri, const float* scores, const float* anchors, const float* stride, int B, int A,
                         int nc, int reg_max, int xywh, float* boxes, float* probs, int nl, const int* a_off, cudaStream_t st) {
    SNN_REQUIRE(reg_max == kRegMax, "detect_decode: reg_max must be %d (got %d)", kRegMax, reg_max);
    const long long N = (long long)B * A;
    if (N == 0) return 0;
    RowMap rm;
    if (make_rowmap(&rm, nl, a_off, B, A)) return 2;
    launch_pdl(detect_decode_kernel, dim3((unsigned)((N + 127) / 128)), dim3(128), 0, st, distri, scores, anchors, stride, N, A, nc, xywh, boxes, probs, rm);
    return check_cuda(cudaGetLastError(), "detect_decode_kernel");
}

int launch_detect_loss_fwd(const float* distri, const float* scores, const float* anchors, const float* stride,
                           const float* tbox_px, const float* tscores, const uint8_t* fg, int B, int A, int nc, int reg_max,
                           double* sums, int nl, const int* a_off, const float* tss_part, int n_parts, const float* gains,
                           unsigned int* counter, float* out6, float* coef3, cudaStream_t st) {
    SNN_REQUIRE(reg_max == kRegMax, "detect_loss: reg_max must be %d (got %d)", kRegMax, reg_max);
    SNN_CUDA_OK(cudaMemsetAsync(sums, 0, 3 * sizeof(double), st));
    const long long N = (long long)B * A;
    if (N == 0) return 0;
    RowMap rm;
    if (make_rowmap(&rm, nl, a_off, B, A)) return 2;
    LossFinal fin = {tss_part, n_parts, gains, counter, out6, coef3, (float)B, nullptr};
    if (tss_part) SNN_REQUIRE(gains && counter && out6 && coef3 && n_parts >= 1, "detect_loss_fwd: fused finalisation needs gains/counter/out6/coef3");
    const unsigned int blocks = (unsigned)((N + 127) / 128);
    if (deterministic()) {
        fin.part = static_cast<double*>(det_scratch(sizeof(double) * 3 * (size_t)blocks, st));
        if (!fin.part) return 2;
    }
    launch_pdl(detect_loss_fwd_kernel, dim3(blocks), dim3(128), 0, st, distri, scores, anchors, stride, tbox_px, tscores, fg,
                                                                        N, A, nc, sums, rm, fin);
    SNN_CUDA_OK(cudaGetLastError());
    if (fin.part && !tss_part) return launch_ordered_combine_f64(fin.part, (int)blocks, 3, sums, st);    // raw sums only: no finalising block
    return 0;
}

int launch_detect_loss_bwd(const float* distri, const float* scores, const float* anchors, const float* stride,
                           const float* tbox_px, const float* tscores, const uint8_t* fg, int B, int A, int nc, int reg_max,
                           const float* coef, const float* gout, void* g_distri, void* g_scores, int out_bf16, int nl, const int* a_off,
                           cudaStream_t st) {
    SNN_REQUIRE(reg_max == kRegMax, "detect_loss: reg_max must be %d (got %d)", kRegMax, reg_max);
    const long long N = (long long)B * A;
    if (N == 0) return 0;
    RowMap rm;
    if (make_rowmap(&rm, nl, a_off, B, A)) return 2;
    const unsigned grid = (unsigned)((N + 127) / 128);
    if (out_bf16)
        launch_pdl(detect_loss_bwd_kernel<__nv_bfloat16>, grid, dim3(128), 0, st, distri, scores, anchors, stride, tbox_px, tscores, fg, N, A, nc, coef, gout,
                                                                    (__nv_bfloat16*)g_distri, (__nv_bfloat16*)g_scores, rm);
    else
        launch_pdl(detect_loss_bwd_kernel<float>, grid, dim3(128), 0, st, distri, scores, anchors, stride, tbox_px, tscores, fg, N, A, nc, coef, gout,
                                                            (float*)g_distri, (float*)g_scores, rm);
    return check_cuda(cudaGetLastError(), "detect_loss_bwd_kernel");
}

}  // namespace snn
