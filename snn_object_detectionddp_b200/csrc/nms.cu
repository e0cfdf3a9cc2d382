// Batched non-maximum suppression on decoded detections: replaces ultralytics.utils.nms.non_max_suppression
// (+ torchvision.ops.nms) as called by the reference at visualize.py:73-78 (conf 0.3, iou 0.45, multi_label) and
// eval_2.py:108 (conf 0.001, iou 0.6).  ultralytics is un-vendored: the algorithm is restated in
// oracle/detect_oracle.py:non_max_suppression, which is what this kernel is bit-compared against.
//
// One CTA per image, everything on the device, no host round trip:
//   1. candidates: multi_label -> every (anchor, class) with score > conf ; else best class per anchor with score > conf,
//      enumerated anchor-major / class-minor (the row order torch.where / boolean indexing produce);
//      key = (~score_bits) << 32 | enumeration index  (unique; ascending = score descending, ties by enumeration order,
//      i.e. the stable sort torchvision uses)
//   2. bitonic sort of the keys (shared memory up to 16K keys, else in the global workspace)
//   3. greedy suppression in sorted order, 256 candidates per round: each thread tests its candidate against the boxes
//      kept so far and builds its row of the intra-round overlap matrix, one thread resolves the round serially.
//      Stops at max_det kept boxes (later boxes cannot change earlier decisions).
// IoU arithmetic mirrors torchvision's kernel operation by operation (no FMA contraction) so that kept indices agree
// exactly; boxes of different classes are separated by the class offset cls * max_wh unless `agnostic`.
#include "common.cuh"

namespace snn {

struct NmsParams {
    const float* pred;   // [B][4 + nc][A]  (cx, cy, w, h in pixels, then class scores)
    int B, nc, A;
    float conf, iou, max_wh;
    int multi_label, agnostic, max_det, max_nms;
    unsigned long long* keys;   // workspace [B][cap]
    long long cap;
    float* out;          // [B][max_det][6]  x1 y1 x2 y2 conf cls
    int* out_idx;        // [B][max_det]     enumeration index of each kept row (anchor*nc + cls | anchor)
    int* counts;         // [B]
};

constexpr int kNmsThreads = 256;
constexpr int kNmsSmemKeys = 16384;

struct Cand { float x1, y1, x2, y2, area, score; int cls; };

SNN_DEVINL Cand nms_load(const NmsParams& p, const float* img, unsigned e) {
    int a, c;
    if (p.multi_label) { a = (int)(e / (unsigned)p.nc); c = (int)(e % (unsigned)p.nc); } else { a = (int)e; c = -1; }
    const float cx = img[a], cy = img[p.A + a], w = img[2 * p.A + a], h = img[3 * p.A + a];
    float sc;
    if (c >= 0) {
        sc = img[(size_t)(4 + c) * p.A + a];
    } else {
        sc = img[(size_t)4 * p.A + a]; c = 0;
        for (int j = 1; j < p.nc; ++j) {
            const float v = img[(size_t)(4 + j) * p.A + a];
            if (v > sc) { sc = v; c = j; }
        }
    }
    // xywh2xyxy, then the per-class offset (oracle/detect_oracle.py: boxes = x[:, :4] + cls * max_wh)
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    const float off = p.agnostic ? 0.f : __fmul_rn((float)c, p.max_wh);
    Cand k;
    k.x1 = __fsub_rn(cx, hw); k.y1 = __fsub_rn(cy, hh); k.x2 = __fadd_rn(cx, hw); k.y2 = __fadd_rn(cy, hh);
    k.score = sc; k.cls = c;
    // suppression geometry uses the offset boxes
    const float ox1 = __fadd_rn(k.x1, off), oy1 = __fadd_rn(k.y1, off), ox2 = __fadd_rn(k.x2, off), oy2 = __fadd_rn(k.y2, off);
    k.area = __fmul_rn(__fsub_rn(ox2, ox1), __fsub_rn(oy2, oy1));
    return k;
}
struct OBox { float x1, y1, x2, y2, area; };
SNN_DEVINL OBox nms_obox(const NmsParams& p, const Cand& k) {
    const float off = p.agnostic ? 0.f : __fmul_rn((float)k.cls, p.max_wh);
    OBox o;
    o.x1 = __fadd_rn(k.x1, off); o.y1 = __fadd_rn(k.y1, off); o.x2 = __fadd_rn(k.x2, off); o.y2 = __fadd_rn(k.y2, off);
    o.area = k.area;
    return o;
}
SNN_DEVINL bool nms_overlap(const OBox& a, const OBox& b, float thr) {
    const float xx1 = fmaxf(a.x1, b.x1), yy1 = fmaxf(a.y1, b.y1), xx2 = fminf(a.x2, b.x2), yy2 = fminf(a.y2, b.y2);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.area, b.area), inter));
    return ovr > thr;
}

__global__ void __launch_bounds__(kNmsThreads) nms_kernel(const NmsParams p) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ unsigned long long skeys[];     // kNmsSmemKeys keys when the image's candidates fit
    __shared__ int s_scan[kNmsThreads];
    __shared__ int s_total, s_nkept;
    __shared__ OBox s_kept[512];                      // kept offset boxes (max_det <= 512)
    __shared__ OBox s_round[kNmsThreads];
    __shared__ unsigned long long s_row[kNmsThreads][kNmsThreads / 64];
    __shared__ unsigned char s_alive[kNmsThreads];

    const int b = blockIdx.x, tid = threadIdx.x;
    const float* img = p.pred + (size_t)b * (4 + p.nc) * p.A;
    unsigned long long* gkeys = p.keys + (size_t)b * p.cap;

    // ---- 1. ordered candidate compaction ------------------------------------------------------------
    const long long n_enum = p.multi_label ? (long long)p.A * p.nc : p.A;
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (long long base = 0; base < n_enum; base += kNmsThreads) {
        const long long e = base + tid;
        float sc = -1.f;
        if (e < n_enum) {
            if (p.multi_label) {
                const int a = (int)(e / p.nc), c = (int)(e % p.nc);
                sc = img[(size_t)(4 + c) * p.A + a];
            } else {
                sc = img[(size_t)4 * p.A + e];
                for (int j = 1; j < p.nc; ++j) sc = fmaxf(sc, img[(size_t)(4 + j) * p.A + e]);
            }
        }
        const int flag = (e < n_enum && sc > p.conf) ? 1 : 0;
        s_scan[tid] = flag;
        __syncthreads();
        for (int off = 1; off < kNmsThreads; off <<= 1) {     // inclusive Hillis-Steele scan
            const int v = (tid >= off) ? s_scan[tid - off] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int pos = s_total + s_scan[tid] - flag;
        if (flag && pos < p.cap)
            gkeys[pos] = ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(sc)) << 32) | (unsigned long long)(unsigned)e;
        __syncthreads();
        if (tid == kNmsThreads - 1) s_total += s_scan[tid];
        __syncthreads();
    }
    int n = (int)((long long)s_total < p.cap ? (long long)s_total : p.cap);

    // ---- 2. bitonic sort (ascending keys) -----------------------------------------------------------
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    const bool in_smem = np2 <= kNmsSmemKeys;
    unsigned long long* keys = in_smem ? skeys : gkeys;
    if (in_smem) {
        for (int i = tid; i < np2; i += kNmsThreads) skeys[i] = i < n ? gkeys[i] : ~0ull;
    } else {
        for (int i = n + tid; i < np2 && i < p.cap; i += kNmsThreads) gkeys[i] = ~0ull;   // cap is a power of two (host)
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += kNmsThreads) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], c = keys[l];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) { keys[i] = c; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (n > p.max_nms) n = p.max_nms;

    // ---- 3. greedy suppression, 256 candidates per round -----------------------------------------
    if (tid == 0) s_nkept = 0;
    __syncthreads();
    const int max_det = min(p.max_det, 512);
    for (int base = 0; base < n; base += kNmsThreads) {
        const int nkept = s_nkept;
        if (nkept >= max_det) break;
        const int i = base + tid;
        Cand cand;
        OBox ob;
        bool alive = false;
        unsigned e = 0;
        if (i < n) {
            e = (unsigned)(keys[i] & 0xFFFFFFFFull);
            cand = nms_load(p, img, e);
            ob = nms_obox(p, cand);
            alive = true;
            for (int k = 0; k < nkept; ++k)
                if (nms_overlap(s_kept[k], ob, p.iou)) { alive = false; break; }
            s_round[tid] = ob;
        }
        s_alive[tid] = alive ? 1 : 0;
        __syncthreads();
        // row tid: earlier candidates of this round that would suppress candidate tid if kept
        unsigned long long row[kNmsThreads / 64];
#pragma unroll
        for (int w = 0; w < kNmsThreads / 64; ++w) row[w] = 0ull;
        if (alive) {
            for (int j = 0; j < tid; ++j)
                if (s_alive[j] && nms_overlap(s_round[j], ob, p.iou)) row[j >> 6] |= 1ull << (j & 63);
        }
#pragma unroll
        for (int w = 0; w < kNmsThreads / 64; ++w) s_row[tid][w] = row[w];
        __syncthreads();
        if (tid == 0) {
            unsigned long long kept[kNmsThreads / 64];
            for (int w = 0; w < kNmsThreads / 64; ++w) kept[w] = 0ull;
            int nk = nkept;
            const int lim = min(kNmsThreads, n - base);
            for (int j = 0; j < lim && nk < max_det; ++j) {
                if (!s_alive[j]) continue;
                bool sup = false;
                for (int w = 0; w < kNmsThreads / 64; ++w) sup |= (s_row[j][w] & kept[w]) != 0ull;
                if (sup) { s_alive[j] = 0; continue; }
                kept[j >> 6] |= 1ull << (j & 63);
                s_alive[j] = 2;                 // kept: output slot assigned below
                s_scan[j] = nk;
                s_kept[nk] = s_round[j];
                ++nk;
            }
            for (int j = 0; j < lim; ++j) if (s_alive[j] == 1) s_alive[j] = 0;   // beyond max_det
            s_nkept = nk;
        }
        __syncthreads();
        if (i < n && s_alive[tid] == 2) {
            const int slot = s_scan[tid];
            float* o = p.out + ((size_t)b * p.max_det + slot) * 6;
            o[0] = cand.x1; o[1] = cand.y1; o[2] = cand.x2; o[3] = cand.y2; o[4] = cand.score; o[5] = (float)cand.cls;
            p.out_idx[(size_t)b * p.max_det + slot] = (int)e;
        }
        __syncthreads();
    }
    if (tid == 0) p.counts[b] = s_nkept;
}

long long nms_workspace_keys(int nc, int A, int multi_label) {
    long long n = multi_label ? (long long)A * nc : A;
    long long c = 1;
    while (c < n) c <<= 1;
    return c;      // per image
}

int launch_nms(const float* pred, int B, int nc, int A, float conf, float iou, int multi_label, int agnostic, int max_det,
               int max_nms, float max_wh, unsigned long long* keys, long long cap, float* out, int* out_idx, int* counts,
               cudaStream_t st) {
    SNN_REQUIRE(B >= 1 && nc >= 1 && A >= 1, "nms: bad sizes");
    SNN_REQUIRE(max_det >= 1 && max_det <= 512, "nms: max_det=%d must be in [1, 512]", max_det);
    SNN_REQUIRE(cap >= nms_workspace_keys(nc, A, multi_label && nc > 1), "nms: workspace too small");
    NmsParams p;
    p.pred = pred; p.B = B; p.nc = nc; p.A = A; p.conf = conf; p.iou = iou; p.max_wh = max_wh;
    p.multi_label = (multi_label && nc > 1) ? 1 : 0; p.agnostic = agnostic; p.max_det = max_det; p.max_nms = max_nms;
    p.keys = keys; p.cap = cap; p.out = out; p.out_idx = out_idx; p.counts = counts;
    static PerDeviceOnce once;
    SNN_CUDA_OK(once.run([] {
        return cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNmsSmemKeys * (int)sizeof(unsigned long long));
    }));
    launch_pdl(nms_kernel, dim3(B), dim3(kNmsThreads), kNmsSmemKeys * sizeof(unsigned long long), st, p);
    return check_cuda(cudaGetLastError(), "nms_kernel");
}

}  // namespace snn
