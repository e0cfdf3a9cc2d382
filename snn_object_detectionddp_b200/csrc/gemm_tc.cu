// Spike-convolution kernels: implicit-GEMM on the 5th-gen tensor cores (tcgen05.mma, fp32
// accumulators in TMEM), operands staged by TMA into 128B-swizzled shared memory, bf16 x bf16.
//
// Replaces the cuDNN fp32 convolutions behind the reference's ConvBlock / UpBlock / ConvLSTM2d /
// output 1x1 convs (reference model.py:13, 36, 55, 119) and their autograd backward.
//
// Activations are NHWC bf16 with the T*B batch folded into N.  A convolution is a sum over "taps"
// (kh,kw); every tap is a *shifted TMA box* of the activation tensor, so im2col never exists in
// memory: padding is the TMA out-of-bounds zero fill, stride 2 / transposed conv are handled by a
// 5-D "phase view" (N, H/2, 2, W/2, 2*C) of the same memory, channel concatenation
// (torch.cat in model.py:45,66,126,127) by walking two tensor maps in the K loop.
//
//   conv_gemm_kernel  : D[pixel, cout] = sum_taps A_tap[pixel, c] * W[cout, tap, c]   (A, W K-major)
//                       used for fprop, dgrad (with [cin][tap][cout] weights), 1x1, convT (4 phases)
//   wgrad_gemm_kernel : dW[cout, tap, cin] += sum_pixels dY[pixel, cout] * X_tap[pixel, cin]
//                       (both operands MN-major straight from the NHWC tensors, split-K + fp32 red)
//
// Warp roles (192 threads): warp0 = TMA producer + TMEM owner, warp1 = MMA issuer, warps2-5 =
// epilogue (one TMEM lane quarter each); small-K convs launch 320 threads: warps 6-9 = a second epilogue group.
#include <atomic>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace snn {

// ------------------------------------------------------------------------------------------
// parameter blocks (passed __grid_constant__)
// ------------------------------------------------------------------------------------------
struct Seg {  // one K-segment: a tap of one source tensor
    int src, dc, dw, dhp, dh;  // source 0/1 and coordinate offsets (channel, w, row-phase, h)
    int wtap, wc0, nchunk;     // weight tap index, weight channel base, number of 64-wide K chunks
};
struct Phase {
    int nseg, oph, opw, pad;
    Seg seg[18];
};
struct ConvGemmParams {
    int NB, Hd, Wd;                   // tile domain (pixels enumerated by the CTA grid)
    int bn, bh, bw;                   // pixel box (bn*bh*bw == 128)
    int tiles_w, tiles_h, tiles_n;
    int BN;                           // UMMA N
    int n_blocks;                     // ceil(n_store / BN)
    int n_store;                      // valid output channels
    int wn_off;                       // first weight row (N coordinate) of this launch
    int stages, stage_bytes;
    void* out;
    const float* bias;
    int out_f32, accumulate;
    int Ho, Wo, os;                   // output spatial dims and pixel stride (1 or 2)
    long long out_ld;                 // elements between consecutive output pixels
    int out_coff;
    int nphase;
    int b_mn;                         // B operand MN-major: weights [K rows][tap][N contiguous] read in place (dgrad)
    int b_boxes;                      // b_mn: number of 64-column boxes per stage
    float* stats;                     // tma_out + fp32: per-(32-row group, column) sum / sum-of-squares partials [4*m_tiles][2][n_store]
    int tma_out;                      // epilogue through swizzled smem staging + TMA tensor stores (coalesced) instead of per-thread rows
    int cps, chunk_bytes;             // 64-wide K chunks per pipeline stage (1, 2 or 4: narrow-N tiles need more MMA work per
                                      // mbarrier round trip) and bytes of one chunk (A 16 KB + this CTA's B slice)
    int nbuf;                         // epilogue staging tiles per warp (2, or 4 for small-K convs: TMA-store latency bound)
    int epi_groups;                   // 1: warps 2-5 run the epilogue (192 threads); 2: warps 6-9 too (320 threads, small-K convs)
    unsigned int* sched;              // dynamic tile scheduler: {next tile - nworkers, finished workers}, both 0 at launch
    int ksplit;                       // K split over the tap segments (fp32 output, TMA reduce-add): fills the GPU on small-M convs
    int dbg;                          // -DSNN_TIMING_KNOBS builds only: bit 0 = producer skips the TMA loads, bit 1 = no MMA issue
    // Row-strip mode (3x3 stride-1, ntap == 3): a segment is one COLUMN of the 3x3 stencil.  Its A operand is ONE box of
    // bh + 2 rows (a_bytes); the three taps kh = 0..2 are views of it shifted by bw rows (a_tap_off bytes, a multiple of the
    // 1024-byte swizzle atom), each with its own weight slice (b_bytes; weight tap = seg.wtap + tap * tap_wstep).  Operand
    // bytes per flop drop by a third to a half -- these layers sit on the L2->SM fabric cap, not on the tensor pipe.
    int ntap, a_bytes, a_tap_off, b_bytes, tap_wstep;
    int hnw;                          // rows of a pixel box ordered (h, n, w) instead of (n, h, w): tensor maps with the image
                                      // dimension BELOW h, so a one-row shift is bn * bw rows also when a box spans bn > 1 images
    Phase phase[4];
};

struct WTap {
    int g_dc, g_dw, g_dhp, g_dh;
    int x_dc, x_dw, x_dhp, x_dh;
    int wtap, pad0, pad1, pad2;
};
struct WgradParams {
    int NB, Hd, Wd;
    int bn, bh, bw;                   // pixel box (product == 64)
    int tiles_w, tiles_h, tiles_n, total_tiles;
    int ksplit, tiles_per_split;
    int m_items;                      // work items along Cout (128 rows per CTA, 256 per CTA pair)
    int NT;                           // cin per CTA (multiple of 64, <= 256)
    int n_ci_tiles;
    int Cout, Ci;
    float* dw;
    long long dw_ld;                  // elements between consecutive cout rows (= taps*wK)
    int wK, w_coff;
    int stages, stage_bytes;
    int lbo_bytes, sbo_bytes;         // MN-major descriptor strides
    int ntaps;
    // Row-strip mode (3x3 stride-1): an item covers the three taps kh = 0..2 of ONE stencil column (taps[] holds the 3 columns).
    // dY is staged once per 64-pixel step, X as a box of bh + 2 rows (x_box_bytes per 64 channels) whose row-shifted views
    // (x_tap_off bytes apart) feed three accumulators of NT columns each (3 * NT <= 512 TMEM columns, single-buffered).
    int strip, x_box_bytes, x_tap_off;
    int hnw;                          // pixel boxes ordered (h, n, w): see ConvGemmParams::hnw
    unsigned int* sched;              // dynamic work-item scheduler counters (see SchedRing)
    WTap taps[9];
};

constexpr int kMaxStages = 8;
constexpr int kTmemCols = 256;

// ------------------------------------------------------------------------------------------
// fprop-type kernel
// ------------------------------------------------------------------------------------------
// Persistent: one CTA per SM walks tiles (m-tile fastest, so co-resident CTAs share the weight tile in L2); the
// accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the main loop of
// tile i+1.  Pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
//
// PAIR = true: the two CTAs of a cluster work on ONE 256-pixel x BN tile with tcgen05 cta_group::2.  Each CTA stages
// its own 128 pixels of A and only HALF of the weight tile (BN/2 columns), the leader CTA issues the MMAs for both.
// Why: profiles/r1 shows every large conv pinned at 10-13 TB/s of L2->SM traffic (the fabric cap), i.e. bound by
// operand bytes per flop, M*N/(M+N) = 85 flop/B for a 128x256 single-CTA tile; the pair tile is 128 flop/B.
constexpr int kAccCols = 256;

template <bool PAIR>
__global__ void __launch_bounds__(320, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ ConvGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ SchedRing ring;
    __shared__ uint32_t tmem_base_s;

    pdl_launch_dependents();      // the next kernel may start its own set-up; it waits for this grid at its pdl_wait()
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp: provably uniform
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int stages = p.stages;
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes;     // per CTA
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;      // 0 = leader (MMA issuer)
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;    // 128-pixel tiles
    const int m_work = PAIR ? (m_tiles + 1) >> 1 : m_tiles;   // work items along M (pairs of m-tiles)
    const int total_tiles = m_work * p.n_blocks * p.nphase * p.ksplit;   // tile = (m, n-block, phase, k-split), m fastest
    const int bn_cta = PAIR ? p.BN >> 1 : p.BN;               // B columns staged by this CTA

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmA0);
            tma_prefetch_desc(&tmA1);
            tma_prefetch_desc(&tmB);
            if (p.tma_out) tma_prefetch_desc(&tmO);
        }
        __syncwarp();
        if (PAIR) tmem_alloc_pair<2 * kAccCols>(smem_u32(&tmem_base_s));
        else tmem_alloc<2 * kAccCols>(smem_u32(&tmem_base_s));
    } else if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tmem_full_bar[a]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[a]), (PAIR ? 8u : 4u) * (uint32_t)p.epi_groups);   // one arrive per epilogue warp (of both CTAs)
        }
        // readers of a tile id: MMA warp + epilogue warps of the leader; producer warp + epilogue warps of the peer
        sched_init(ring, (PAIR ? 2u : 1u) * (1u + 4u * (uint32_t)p.epi_groups));
        mbar_fence_init();
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();                   // barriers, TMEM and descriptor prefetch above overlapped the previous kernel's tail

    // Both role loops below run WARP-UNIFORM (all 32 lanes walk the loop, one elected lane issues the TMA / MMA
    // instructions): loop state then lives in uniform registers and the per-chunk instruction count of the issuing
    // thread -- which, not the tensor pipe, bounded v1 at ~635 cycles per 64-wide K chunk -- drops several-fold.
    if (warp == 0) {
        const bool lead = elect_one();
        int s = 0;
        uint32_t par = 0;
        uint32_t a_s = sbase;
        const bool dyn = p.sched != nullptr;
        // dynamic: this warp of the leader CTA is the tile scheduler.  Ids are published TWO tiles ahead: `tile` is being
        // loaded, `nxt1` is already in the ring, the atomic for the one after is in flight (`pend`, lane 0).
        int tile = worker, nxt1 = -1, pend = 0, pub = 0;      // pub: sequence number of the next id to publish; -1 once the end marker is out
        if (dyn && rank == 0) {
            if (tile >= total_tiles) tile = -1;
            sched_publish<PAIR>(ring, pub++, tile, lead);
            if (tile < 0) pub = -1;
            if (pub > 0) {
                if (lane == 0) pend = (int)atomicAdd(p.sched, 1u) + nworkers;
                nxt1 = __shfl_sync(0xffffffffu, pend, 0);
                if (nxt1 >= total_tiles) nxt1 = -1;
                sched_publish<PAIR>(ring, pub++, nxt1, lead);
                if (nxt1 < 0) pub = -1;
                else if (lane == 0) pend = (int)atomicAdd(p.sched, 1u) + nworkers;
            }
        }
        for (int it = 0;; ++it) {
            if (!dyn) {
                tile = worker + it * nworkers;
                if (tile >= total_tiles) break;
            } else if (rank == 0) {
                if (tile < 0) break;
            } else {
                tile = sched_next<PAIR>(ring, it, lane);
                if (tile < 0) break;
            }
            const int mt = (tile % m_work) * (PAIR ? 2 : 1) + (int)rank, rest = tile / m_work;
            const int ncol0 = p.wn_off + (rest % p.n_blocks) * p.BN + (int)rank * bn_cta;
            const int pk = rest / p.n_blocks;
            const Phase& ph = p.phase[pk % p.nphase];
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
            const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;   // tn == tiles_n (odd tail of a pair): all OOB -> zeros
            const int spp = ph.nseg / p.ksplit, sg0 = (pk / p.nphase) * spp;
            for (int sg = sg0; sg < sg0 + spp; ++sg) {
                const Seg g = ph.seg[sg];
                const CUtensorMap* tmA = g.src ? &tmA1 : &tmA0;
                const int cw = w0 + g.dw, chh = h0 + g.dh;
                int ca = g.dc, cb = g.wc0;
                for (int kc = 0; kc < g.nchunk; kc += p.cps) {
                    const int nsub = min(p.cps, g.nchunk - kc);       // K chunks carried by this stage
                    mbar_wait(smem_u32(&empty_bar[s]), par ^ 1u);
                    if (lead) {
                        const uint32_t fb = smem_u32(&full_bar[s]);
#ifdef SNN_TIMING_KNOBS
                        if (p.dbg & 1) {
                            if (rank == 0) mbar_arrive(fb);
                        } else
#endif
                        if (!PAIR) {
                            mbar_expect_tx(fb, (uint32_t)(nsub * p.chunk_bytes));
                            for (int j = 0; j < nsub; ++j) {
                                const uint32_t c_s = a_s + (uint32_t)(j * p.chunk_bytes);
                                tma_load_5d(c_s, tmA, fb, ca + 64 * j, cw, g.dhp, p.hnw ? n0 : chh, p.hnw ? chh : n0);
                                for (int tp = 0; tp < p.ntap; ++tp) {
                                    const uint32_t b_s = c_s + (uint32_t)(p.a_bytes + tp * p.b_bytes);
                                    const int wt = g.wtap + tp * p.tap_wstep;
                                    if (!p.b_mn) {
                                        tma_load_3d(b_s, &tmB, fb, cb + 64 * j, wt, ncol0);
                                    } else {  // [64 K rows] x [64 N] boxes, one per 64 output columns
                                        for (int nb = 0; nb < p.b_boxes; ++nb)
                                            tma_load_3d(b_s + (uint32_t)nb * 8192u, &tmB, fb, ncol0 + nb * 64, wt, cb + 64 * j);
                                    }
                                }
                            }
                        } else {
                            if (rank == 0) mbar_expect_tx(fb, 2u * (uint32_t)(nsub * p.chunk_bytes));   // both CTAs' bytes land on the leader's barrier
                            for (int j = 0; j < nsub; ++j) {
                                const uint32_t c_s = a_s + (uint32_t)(j * p.chunk_bytes);
                                tma_load_5d_pair(c_s, tmA, fb, ca + 64 * j, cw, g.dhp, p.hnw ? n0 : chh, p.hnw ? chh : n0);
                                for (int tp = 0; tp < p.ntap; ++tp) {
                                    const uint32_t b_s = c_s + (uint32_t)(p.a_bytes + tp * p.b_bytes);
                                    const int wt = g.wtap + tp * p.tap_wstep;
                                    if (!p.b_mn) {
                                        tma_load_3d_pair(b_s, &tmB, fb, cb + 64 * j, wt, ncol0);
                                    } else {
                                        for (int nb = 0; nb < p.b_boxes; ++nb)
                                            tma_load_3d_pair(b_s + (uint32_t)nb * 8192u, &tmB, fb, ncol0 + nb * 64, wt, cb + 64 * j);
                                    }
                                }
                            }
                        }
                    }
                    ca += 64 * p.cps; cb += 64 * p.cps;
                    a_s += stage_bytes;
                    if (++s == stages) { s = 0; par ^= 1u; a_s = sbase; }
                }
            }
            if (dyn && rank == 0) {          // advance: publish the id two tiles ahead, keep one atomic in flight
                tile = nxt1;
                if (pub > 0) {
                    nxt1 = __shfl_sync(0xffffffffu, pend, 0);                    // lane 0: provably warp-uniform
                    if (nxt1 >= total_tiles) nxt1 = -1;
                    sched_publish<PAIR>(ring, pub++, nxt1, lead);
                    if (nxt1 < 0) pub = -1;
                    else if (lane == 0) pend = (int)atomicAdd(p.sched, 1u) + nworkers;
                } else {
                    nxt1 = -1;
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const bool lead = elect_one();
            const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, p.BN, 0, p.b_mn);
            // smem matrix descriptors: high word constant (SBO = 1024 B, version 1, 128B swizzle); low word = address >> 4
            // plus the leading-dimension offset field (K-major: unused = 16 B; MN-major B: 64-column blocks 8192 B apart)
            const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t a_lo_c = (16u >> 4) << 16;
            const uint32_t b_lo_c = (p.b_mn ? (8192u >> 4) : (16u >> 4)) << 16;
            const uint32_t b_kstep = p.b_mn ? (2048u >> 4) : (32u >> 4);   // 16 K rows of 128 B | 32 B inside the 128 B row
#ifdef SNN_TIMING_KNOBS
            const uint32_t dbg_nomma = (uint32_t)(p.dbg & 2);
#else
            constexpr uint32_t dbg_nomma = 0u;      // the timing-experiment knob is compiled out of the shipped library
#endif
            int s = 0, lt = 0;
            uint32_t par = 0;
            uint32_t a_s = sbase;
            const bool dyn = p.sched != nullptr;
            for (;; ++lt) {
                const int tile = dyn ? sched_next<PAIR>(ring, lt, lane) : (worker + lt * nworkers < total_tiles ? worker + lt * nworkers : -1);
                if (tile < 0) break;
                const int pk = (tile / m_work) / p.n_blocks;
                const Phase& ph = p.phase[pk % p.nphase];
                const int spp = ph.nseg / p.ksplit, sg0 = (pk / p.nphase) * spp;
                const int acc = lt & 1;
                mbar_wait(smem_u32(&tmem_empty_bar[acc]), (uint32_t)(((lt >> 1) & 1) ^ 1));   // epilogue drained this buffer
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * kAccCols);
                uint32_t first = 1u;                                   // first MMA of the tile overwrites the accumulator
                for (int sg = sg0; sg < sg0 + spp; ++sg) {
                    const int nchunk = ph.seg[sg].nchunk;
                    const bool last_seg = sg == sg0 + spp - 1;
                    for (int kc = 0; kc < nchunk; kc += p.cps) {
                        const int nsub = min(p.cps, nchunk - kc);
                        mbar_wait(smem_u32(&full_bar[s]), par);
                        tc_fence_after();
                        if (lead) {
                            if (!dbg_nomma) {
                                for (int j = 0; j < nsub; ++j) {
                                    const uint32_t c_s = a_s + (uint32_t)(j * p.chunk_bytes);
                                    for (int tp = 0; tp < p.ntap; ++tp) {       // row-strip mode: three row-shifted views of one A box
                                        const uint32_t a_lo = a_lo_c | (((c_s + (uint32_t)(tp * p.a_tap_off)) & 0x3FFFFu) >> 4);
                                        const uint32_t b_lo = b_lo_c | (((c_s + (uint32_t)(p.a_bytes + tp * p.b_bytes)) & 0x3FFFFu) >> 4);
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            const uint64_t adesc = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * k);
                                            const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + b_kstep * k);
                                            const uint32_t accf = (k | j | tp) ? 1u : (first ^ 1u);
                                            if (PAIR) umma_bf16_pair(tacc, adesc, bdesc, idesc, accf);
                                            else umma_bf16(tacc, adesc, bdesc, idesc, accf);
                                        }
                                    }
                                }
                            }
                            if (PAIR) umma_commit_pair(smem_u32(&empty_bar[s])); else umma_commit(smem_u32(&empty_bar[s]));
                            if (last_seg && kc + p.cps >= nchunk) {
                                if (PAIR) umma_commit_pair(smem_u32(&tmem_full_bar[acc])); else umma_commit(smem_u32(&tmem_full_bar[acc]));
                            }
                        }
                        first = 0u;
                        a_s += stage_bytes;
                        if (++s == stages) { s = 0; par ^= 1u; a_s = sbase; }
                    }
                }
            }
        }
    } else {
        // epilogue: warp w may touch TMEM lanes [32*(w%4), +32)
        // epi_groups == 2 (small-K convs, 320 threads): warps 6-9 form a second epilogue group on the same TMEM lane quarters
        // and take the odd column chunks.  With K <= 8 chunks per tile the epilogue IS the kernel (ncu, profiles/r2: the MMA
        // warp waits, every epilogue instruction carries equal stall samples = one dependent chain per scheduler); two
        // warps per scheduler overlap those latencies.
        const int q = warp & 3;
        const int grp = warp >= 6 ? 1 : 0, ngrp = p.epi_groups;
        const int row = q * 32 + lane;
        const int wl = row % p.bw;
        const int hl = p.hnw ? row / (p.bw * p.bn) : (row / p.bw) % p.bh, nl = p.hnw ? (row / p.bw) % p.bn : row / (p.bw * p.bh);
        int lt = 0;
        uint32_t gcc = 0;     // staging-tile counter across all tiles of this warp: nbuf buffers in rotation
        const bool dyn = p.sched != nullptr;
        for (;; ++lt) {
            const int tile = dyn ? sched_next<PAIR>(ring, lt, lane) : (worker + lt * nworkers < total_tiles ? worker + lt * nworkers : -1);
            if (tile < 0) break;
            const int mt = (tile % m_work) * (PAIR ? 2 : 1) + (int)rank, rest = tile / m_work;
            const int ncol0 = (rest % p.n_blocks) * p.BN;
            const Phase& ph = p.phase[(rest / p.n_blocks) % p.nphase];
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
            const int n = tn * p.bn + nl, hd = th * p.bh + hl, wd = tw * p.bw + wl;
            const bool valid = (n < p.NB) && (hd < p.Hd) && (wd < p.Wd);
            const long long pix = ((long long)n * p.Ho + (hd * p.os + ph.oph)) * p.Wo + (wd * p.os + ph.opw);
            const int acc = lt & 1;
            mbar_wait(smem_u32(&tmem_full_bar[acc]), (uint32_t)((lt >> 1) & 1));
            tc_fence_after();
            const uint32_t trow = tmem_base + (uint32_t)(acc * kAccCols) + ((uint32_t)(q * 32) << 16);
            if (p.tma_out) {
                // Coalesced epilogue: 32 rows x 128 B of the accumulator go through a 128B-swizzled staging tile (thread =
                // row; the XOR keeps the 16-byte stores of a quarter-warp on distinct banks) and leave with ONE TMA
                // tensor store per warp and chunk.  The v1 epilogue (thread = row writing 64 B pieces 1 KB+ apart) cost
                // 30k cycles per 128x256 tile in LSU transactions -- more than the MMA main loop (profiles/r1 probe).
                const int CH = p.out_f32 ? 32 : 64;                       // columns per 128-byte row
                const int chs = p.out_f32 ? 5 : 6;
                const int row0 = q * 32;
                const int bw_ = p.bw, bh_ = p.bh;
                const int cw = tw * bw_ + row0 % bw_;
                const int chh = th * bh_ + (p.hnw ? row0 / (bw_ * p.bn) : (row0 / bw_) % bh_);
                const int cn = tn * p.bn + (p.hnw ? (row0 / bw_) % p.bn : row0 / (bw_ * bh_));
                const int c3 = p.hnw ? cn : chh, c4 = p.hnw ? chh : cn;       // tensor-map coordinates 3 and 4
                const int cbase = (p.os == 2 ? ph.opw * (int)p.out_ld : 0);
                const int cph = (p.os == 2 ? ph.oph : 0);
                const uint32_t stg0 = sbase + (uint32_t)stages * stage_bytes + (uint32_t)((grp * 4 + q) * p.nbuf) * 4096u;
                const int nch = (min(p.BN, p.n_store - ncol0) + CH - 1) >> chs;
                const bool has_bias = p.bias != nullptr;
                for (int cc = grp; cc < nch; cc += ngrp, ++gcc) {
                    const uint32_t buf = stg0 + (gcc & (uint32_t)(p.nbuf - 1)) * 4096u;       // nbuf is 2 or 4
                    uint32_t r[32];
                    if (p.out_f32) {
                        tmem_ld16(trow + (uint32_t)(cc * 32), r);
                        tmem_ld16(trow + (uint32_t)(cc * 32 + 16), r + 16);
                        tmem_ld_wait();
                        if (has_bias) {         // the same 32 bias values for every lane: 8 broadcast 16-byte loads
                            const int c0 = ncol0 + cc * 32;
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) {
                                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (c0 + 4 * j4 < p.n_store) bv = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + j4);   // n_store % 8 == 0
                                r[4 * j4 + 0] = __float_as_uint(__uint_as_float(r[4 * j4 + 0]) + bv.x);
                                r[4 * j4 + 1] = __float_as_uint(__uint_as_float(r[4 * j4 + 1]) + bv.y);
                                r[4 * j4 + 2] = __float_as_uint(__uint_as_float(r[4 * j4 + 2]) + bv.z);
                                r[4 * j4 + 3] = __float_as_uint(__uint_as_float(r[4 * j4 + 3]) + bv.w);
                            }
                        }
                    } else {
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t t32[32];
                            tmem_ld16(trow + (uint32_t)(cc * 64 + hh * 32), t32);
                            tmem_ld16(trow + (uint32_t)(cc * 64 + hh * 32 + 16), t32 + 16);
                            tmem_ld_wait();
                            if (has_bias) {
                                const int c0 = ncol0 + cc * 64 + hh * 32;
#pragma unroll
                                for (int j4 = 0; j4 < 8; ++j4) {
                                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (c0 + 4 * j4 < p.n_store) bv = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + j4);
                                    r[hh * 16 + 2 * j4] = pack_bf16x2(__uint_as_float(t32[4 * j4]) + bv.x, __uint_as_float(t32[4 * j4 + 1]) + bv.y);
                                    r[hh * 16 + 2 * j4 + 1] = pack_bf16x2(__uint_as_float(t32[4 * j4 + 2]) + bv.z, __uint_as_float(t32[4 * j4 + 3]) + bv.w);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    r[hh * 16 + j] = pack_bf16x2(__uint_as_float(t32[2 * j]), __uint_as_float(t32[2 * j + 1]));
                            }
                        }
                    }
                    if (gcc >= (uint32_t)p.nbuf) {       // the store that read this buffer nbuf chunks ago has drained it
                        if (lane == 0) { if (p.nbuf == 4) bulk_wait_group_read<3>(); else bulk_wait_group_read<1>(); }
                        __syncwarp();
                    }
                    if (p.stats && !valid) {
                        // a pixel row outside the map (box does not divide Hd x Wd): a 3x3 tap of it may still have read
                        // in-range input, so its accumulator is not zero -- it must not reach the statistics (the tensor
                        // store clips it anyway)
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = 0u;
                    }
                    const uint32_t rowaddr = buf + (uint32_t)lane * 128u;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        st_shared_v4(rowaddr + (uint32_t)((c ^ (lane & 7)) << 4), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (p.stats && mt < m_tiles) {
                        // train-mode BatchNorm statistics of the conv output, fused: lane c sums column c of the staged
                        // 32 x 32 tile in a fixed row order (deterministic; rows of out-of-range pixels were zeroed above).
                        // The swizzle makes the 32 lanes of every row read 32 distinct banks.
                        float sm = 0.f, sq = 0.f;
                        const uint32_t cadr = buf + (uint32_t)((lane & 3) << 2);
#pragma unroll 8
                        for (int rr = 0; rr < 32; ++rr) {
                            float v;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(cadr + (uint32_t)rr * 128u + (uint32_t)((((lane >> 2) ^ (rr & 7))) << 4)));
                            sm += v;
                            sq = fmaf(v, v, sq);
                        }
                        const int col = ncol0 + cc * 32 + lane;
                        if (col < p.n_store) {
                            float* dst = p.stats + ((size_t)(mt * 4 + q) * 2) * (size_t)p.n_store + col;
                            dst[0] = sm;
                            dst[p.n_store] = sq;
                        }
                    }
                    if (lane == 0) {
                        if (p.accumulate) tma_reduce_add_5d(&tmO, buf, cbase + ncol0 + cc * CH, cw, cph, c3, c4);
                        else tma_store_5d(&tmO, buf, cbase + ncol0 + cc * CH, cw, cph, c3, c4);
                        bulk_commit_group();
                    }
                }
                // TMEM reads are complete: hand the accumulator back (staging buffers keep rotating across tiles)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster(smem_u32(&tmem_empty_bar[acc]) & kPeerBitMask);
                    else mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
                }
                continue;
            }
            const int nchunks = p.BN >> 4;
            for (int ch = grp; ch < nchunks; ch += ngrp) {
                uint32_t r[16];
                tmem_ld16(trow + (uint32_t)(ch * 16), r);
                tmem_ld_wait();
                const int col = ncol0 + ch * 16;
                if (!valid || col >= p.n_store) continue;
                const int nv = min(16, p.n_store - col);  // 8 or 16 (n_store % 8 == 0)
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < nv) v[j] += __ldg(p.bias + col + j);
                }
                if (p.out_f32) {
                    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.out_ld + p.out_coff + col);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j * 4 < nv) {
                            float4 t = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            if (p.accumulate) {
                                const float4 old = o[j];
                                t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
                            }
                            o[j] = t;
                        }
                    }
                } else {
                    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_ld + p.out_coff + col;
                    uint4* o = reinterpret_cast<uint4*>(ob);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (j * 8 < nv) {
                            if (p.accumulate) {
                                const uint4 old = o[j];
                                v[8 * j + 0] += bf16_lo(old.x); v[8 * j + 1] += bf16_hi(old.x);
                                v[8 * j + 2] += bf16_lo(old.y); v[8 * j + 3] += bf16_hi(old.y);
                                v[8 * j + 4] += bf16_lo(old.z); v[8 * j + 5] += bf16_hi(old.z);
                                v[8 * j + 6] += bf16_lo(old.w); v[8 * j + 7] += bf16_hi(old.w);
                            }
                            uint4 t;
                            t.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); t.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                            t.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); t.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                            o[j] = t;
                        }
                    }
                }
            }
            // all TMEM reads of this warp are complete (tmem_ld_wait above): hand the accumulator back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(smem_u32(&tmem_empty_bar[acc]) & kPeerBitMask);   // the leader's barrier
                else mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
            }
        }
    }
    if (p.tma_out && warp >= 2 && lane == 0) bulk_wait_group<0>();   // all tensor stores of this thread are complete
    __syncwarp();
    tc_fence_before();
    if (PAIR) {
        cluster_sync_all();       // the peer's smem / TMEM / barriers stay alive until the leader's last MMA and commit landed
        if (warp == 0) { __syncwarp(); tmem_dealloc_pair<2 * kAccCols>(tmem_base); }
    } else {
        __syncthreads();
        if (warp == 0) tmem_dealloc<2 * kAccCols>(tmem_base);
    }
    if (p.sched != nullptr && rank == 0 && threadIdx.x == 0) sched_finish(p.sched, nworkers);
}

// ------------------------------------------------------------------------------------------
// wgrad kernel: dW[co, tap, ci] += sum_{pixels} dY[pix, co] * X[pix + tap, ci]
// A = dY tile (M = 128 couts, K = 64 pixels), B = X tile (N = NT cins, K = 64 pixels), both MN-major.
// ------------------------------------------------------------------------------------------
// Persistent: a CTA (pair) walks work items (cout tile, tap x cin tile, pixel range); the accumulator is double-buffered
// in TMEM so the epilogue of item i overlaps the K loop of item i+1 (v1 ran one item per CTA: TMEM allocation, barrier
// set-up and a serialised row-wise red.global.add epilogue cost as much as the 64-stage K loop itself, profiles/r1).
// Epilogue: 32 x 32 fp32 pieces through the 128B-swizzled staging tile, then ONE TMA reduce-add per piece into the flat
// gradient buffer (split-K partial sums meet there; coalesced, asynchronous).
// PAIR: two CTAs share one 256-cout x NT-cin accumulator tile (cta_group::2): each stages its own 128 couts of dY and HALF
// of the X tile.
template <bool PAIR>
__global__ void __launch_bounds__(192, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmW, const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ SchedRing ring;
    __shared__ uint32_t tmem_base_s;

    pdl_launch_dependents();
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int stages = p.stages;
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes;   // per CTA
    constexpr uint32_t kBox = 64 * 128;  // one [64 pixels x 64 channels] bf16 box
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int ny = p.n_ci_tiles * p.ntaps;
    const int total_items = p.m_items * ny * p.ksplit;        // item = (m, tap x cin tile, pixel range), m fastest
    const int nxb = (PAIR ? p.NT >> 1 : p.NT) >> 6;           // X boxes staged by this CTA
    const uint32_t xbb = p.strip ? (uint32_t)p.x_box_bytes : kBox;    // bytes of one X box (64 channels)
    const int ntp = p.strip ? 3 : 1;                          // taps (accumulators) per item
    const int tiles_hw = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmG);
            tma_prefetch_desc(&tmX);
            tma_prefetch_desc(&tmW);
        }
        __syncwarp();
        if (PAIR) tmem_alloc_pair<2 * kTmemCols>(smem_u32(&tmem_base_s)); else tmem_alloc<2 * kTmemCols>(smem_u32(&tmem_base_s));
    } else if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tmem_full_bar[a]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[a]), PAIR ? 8 : 4);
        }
        sched_init(ring, PAIR ? 10u : 5u);        // readers of a work-item id: MMA warp + 4 epilogue warps (leader), producer + 4 (peer)
        mbar_fence_init();
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();                   // barriers, TMEM and descriptor prefetch above overlapped the previous kernel's tail

    // warp-uniform role loops (see conv_gemm_kernel): one elected lane issues, loop state stays in uniform registers
    if (warp == 0) {
        const bool lead = elect_one();
        int s = 0;
        uint32_t par = 0;
        uint32_t g_s = sbase;
        const bool dyn = p.sched != nullptr;
        int item = worker, nxt1 = -1, pend = 0, pub = 0;      // as in conv_gemm_kernel: ids go out two work items ahead
        if (dyn && rank == 0) {
            if (item >= total_items) item = -1;
            sched_publish<PAIR>(ring, pub++, item, lead);
            if (item < 0) pub = -1;
            if (pub > 0) {
                if (lane == 0) pend = (int)atomicAdd(p.sched, 1u) + nworkers;
                nxt1 = __shfl_sync(0xffffffffu, pend, 0);
                if (nxt1 >= total_items) nxt1 = -1;
                sched_publish<PAIR>(ring, pub++, nxt1, lead);
                if (nxt1 < 0) pub = -1;
                else if (lane == 0) pend = (int)atomicAdd(p.sched, 1u) + nworkers;
            }
        }
        for (int seq = 0;; ++seq) {
            if (!dyn) {
                item = worker + seq * nworkers;
                if (item >= total_items) break;
            } else if (rank == 0) {
                if (item < 0) break;
            } else {
                item = sched_next<PAIR>(ring, seq, lane);
                if (item < 0) break;
            }
            const int mi = item % p.m_items, y = (item / p.m_items) % ny, z = item / (p.m_items * ny);
            const int co0 = PAIR ? mi * 256 + (int)rank * 128 : mi * 128;
            const WTap tp = p.taps[y / p.n_ci_tiles];
            const int ci0 = (y % p.n_ci_tiles) * p.NT + (PAIR ? (int)rank * (p.NT >> 1) : 0);
            const int t_begin = z * p.tiles_per_split;
            const int total = min(p.total_tiles, t_begin + p.tiles_per_split) - t_begin;
            int tw = t_begin % p.tiles_w, th = (t_begin / p.tiles_w) % p.tiles_h, tn = t_begin / tiles_hw;
            const int gc0 = tp.g_dc + co0, xc0 = tp.x_dc + ci0;
            for (int it = 0; it < total; ++it) {
                const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
                mbar_wait(smem_u32(&empty_bar[s]), par ^ 1u);
                if (lead) {
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    const uint32_t x_s = g_s + 2 * kBox;
                    if (!PAIR) {
                        mbar_expect_tx(fb, stage_bytes);
                        tma_load_5d(g_s, &tmG, fb, gc0, w0 + tp.g_dw, tp.g_dhp, p.hnw ? n0 : h0 + tp.g_dh, p.hnw ? h0 + tp.g_dh : n0);
                        tma_load_5d(g_s + kBox, &tmG, fb, gc0 + 64, w0 + tp.g_dw, tp.g_dhp, p.hnw ? n0 : h0 + tp.g_dh, p.hnw ? h0 + tp.g_dh : n0);
                        for (int b = 0; b < nxb; ++b)
                            tma_load_5d(x_s + b * xbb, &tmX, fb, xc0 + b * 64, w0 + tp.x_dw, tp.x_dhp, p.hnw ? n0 : h0 + tp.x_dh, p.hnw ? h0 + tp.x_dh : n0);
                    } else {
                        if (rank == 0) mbar_expect_tx(fb, 2u * stage_bytes);
                        tma_load_5d_pair(g_s, &tmG, fb, gc0, w0 + tp.g_dw, tp.g_dhp, p.hnw ? n0 : h0 + tp.g_dh, p.hnw ? h0 + tp.g_dh : n0);
                        tma_load_5d_pair(g_s + kBox, &tmG, fb, gc0 + 64, w0 + tp.g_dw, tp.g_dhp, p.hnw ? n0 : h0 + tp.g_dh, p.hnw ? h0 + tp.g_dh : n0);
                        for (int b = 0; b < nxb; ++b)
                            tma_load_5d_pair(x_s + b * xbb, &tmX, fb, xc0 + b * 64, w0 + tp.x_dw, tp.x_dhp, p.hnw ? n0 : h0 + tp.x_dh, p.hnw ? h0 + tp.x_dh : n0);
                    }
                }
                g_s += stage_bytes;
                if (++s == stages) { s = 0; par ^= 1u; g_s = sbase; }
                if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } }
            }
            if (dyn && rank == 0) {
                item = nxt1;
                if (pub > 0) {
                    nxt1 = __shfl_sync(0xffffffffu, pend, 0);
                    if (nxt1 >= total_items) nxt1 = -1;
                    sched_publish<PAIR>(ring, pub++, nxt1, lead);
                    if (nxt1 < 0) pub = -1;
                    else if (lane == 0) pend = (int)atomicAdd(p.sched, 1u) + nworkers;
                } else {
                    nxt1 = -1;
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const bool lead = elect_one();
            const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, p.NT, 1, 1);
            const uint32_t desc_hi = ((uint32_t)p.sbo_bytes >> 4) | (1u << 14) | (2u << 29);
            const uint32_t lo_c = (((uint32_t)p.lbo_bytes >> 4) & 0x3FFFu) << 16;
            const uint32_t lo_cb = p.strip ? ((((uint32_t)p.x_box_bytes >> 4) & 0x3FFFu) << 16) : lo_c;   // X: 64-channel blocks one box apart
            int s = 0, li = 0;
            uint32_t par = 0;
            uint32_t g_s = sbase;
            const bool dyn = p.sched != nullptr;
            for (;; ++li) {
                const int item = dyn ? sched_next<PAIR>(ring, li, lane) : (worker + li * nworkers < total_items ? worker + li * nworkers : -1);
                if (item < 0) break;
                const int t_begin = (item / (p.m_items * ny)) * p.tiles_per_split;
                const int total = min(p.total_tiles, t_begin + p.tiles_per_split) - t_begin;
                // strip mode: ONE accumulator set of 3 x NT columns (no double buffering: the K loops are long)
                const int acc = p.strip ? 0 : (li & 1);
                mbar_wait(smem_u32(&tmem_empty_bar[acc]), (uint32_t)(((p.strip ? li : (li >> 1)) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * kTmemCols);
                for (int it = 0; it < total; ++it) {
                    mbar_wait(smem_u32(&full_bar[s]), par);
                    tc_fence_after();
                    if (lead) {
                        const uint32_t a_lo = lo_c | ((g_s & 0x3FFFFu) >> 4);
                        for (int t3 = 0; t3 < ntp; ++t3) {     // strip mode: three row-shifted views of the X box
                            const uint32_t b_lo = lo_cb | (((g_s + 2 * kBox + (uint32_t)(t3 * p.x_tap_off)) & 0x3FFFFu) >> 4);
                            const uint32_t tcol = tacc + (uint32_t)(t3 * p.NT);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {  // 16 pixels (rows of 128 B) per MMA
                                const uint64_t adesc = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + (2048u >> 4) * k);
                                const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + (2048u >> 4) * k);
                                if (PAIR) umma_bf16_pair(tcol, adesc, bdesc, idesc, k ? 1u : (uint32_t)(it != 0));
                                else umma_bf16(tcol, adesc, bdesc, idesc, k ? 1u : (uint32_t)(it != 0));
                            }
                        }
                        if (PAIR) umma_commit_pair(smem_u32(&empty_bar[s])); else umma_commit(smem_u32(&empty_bar[s]));
                        if (it == total - 1) {
                            if (PAIR) umma_commit_pair(smem_u32(&tmem_full_bar[acc])); else umma_commit(smem_u32(&tmem_full_bar[acc]));
                        }
                    }
                    g_s += stage_bytes;
                    if (++s == stages) { s = 0; par ^= 1u; g_s = sbase; }
                }
            }
        }
    } else {
        const int q = warp & 3;
        const uint32_t stg0 = sbase + (uint32_t)stages * stage_bytes + (uint32_t)q * 8192u;
        uint32_t gcc = 0;
        int li = 0;
        const bool dyn = p.sched != nullptr;
        for (;; ++li) {
            const int item = dyn ? sched_next<PAIR>(ring, li, lane) : (worker + li * nworkers < total_items ? worker + li * nworkers : -1);
            if (item < 0) break;
            const int mi = item % p.m_items, y = (item / p.m_items) % ny;
            const int co0 = (PAIR ? mi * 256 + (int)rank * 128 : mi * 128) + q * 32;
            const int wtap = p.taps[y / p.n_ci_tiles].wtap;
            const int ci0 = (y % p.n_ci_tiles) * p.NT;
            const int acc = p.strip ? 0 : (li & 1);
            mbar_wait(smem_u32(&tmem_full_bar[acc]), (uint32_t)((p.strip ? li : (li >> 1)) & 1));
            tc_fence_after();
            const uint32_t trow = tmem_base + (uint32_t)(acc * kTmemCols) + ((uint32_t)(q * 32) << 16);
            const int nch = (min(p.NT, p.Ci - ci0) + 31) >> 5;
            for (int c3 = 0; c3 < nch * ntp; ++c3, ++gcc) {
                const int t3 = c3 / nch, cc = c3 - t3 * nch;           // strip mode: tap kh = t3 (weight tap + 3 * t3), accumulator t3
                const uint32_t buf = stg0 + (gcc & 1u) * 4096u;
                uint32_t r[32];
                tmem_ld16(trow + (uint32_t)(t3 * p.NT + cc * 32), r);
                tmem_ld16(trow + (uint32_t)(t3 * p.NT + cc * 32 + 16), r + 16);
                tmem_ld_wait();
                if (gcc >= 2) {
                    if (lane == 0) bulk_wait_group_read<1>();
                    __syncwarp();
                }
                const uint32_t rowaddr = buf + (uint32_t)lane * 128u;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    st_shared_v4(rowaddr + (uint32_t)((c ^ (lane & 7)) << 4), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && co0 < p.Cout) {   // rows >= Cout / columns >= wK are clipped by the tensor map
                    tma_reduce_add_3d(&tmW, buf, p.w_coff + ci0 + cc * 32, wtap + 3 * t3, co0);
                    bulk_commit_group();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(smem_u32(&tmem_empty_bar[acc]) & kPeerBitMask);
                else mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
            }
        }
        if (lane == 0) bulk_wait_group<0>();
    }
    __syncwarp();
    tc_fence_before();
    if (PAIR) {
        cluster_sync_all();
        if (warp == 0) { __syncwarp(); tmem_dealloc_pair<2 * kTmemCols>(tmem_base); }
    } else {
        __syncthreads();
        if (warp == 0) tmem_dealloc<2 * kTmemCols>(tmem_base);
    }
    if (p.sched != nullptr && rank == 0 && threadIdx.x == 0) sched_finish(p.sched, nworkers);
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps
// ------------------------------------------------------------------------------------------
static std::once_flag g_encode_once;

// ------------------------------------------------------------------------------------------
// Tensor-map (TMA descriptor) cache.  A training step launches ~150 tensor-core kernels with 3-4 maps each and the same
// (pointer, shape, box) tuples recur every step (PyTorch's caching allocator hands the same blocks back): the encoded
// 128-byte maps are kept in a process-wide table keyed by every argument of cuTensorMapEncodeTiled (device included),
// guarded by a mutex, cleared when it reaches kTmapCacheMax entries.  The captured-graph path never re-encodes anyway; this
// is for the eager per-frame drop-in path, which is host-bound.
// ------------------------------------------------------------------------------------------
struct TmapKey {
    unsigned long long ptr, dims[5], strides[4];
    unsigned int box[5], dtype, rank, swizzle, l2, dev;
    bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&k);
        unsigned long long h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed as 64-bit words");
constexpr size_t kTmapCacheMax = 16384;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mutex;
static std::atomic<unsigned long long> g_tmap_hits{0}, g_tmap_misses{0};
void tmap_cache_stats(unsigned long long* hits, unsigned long long* misses) {
    *hits = g_tmap_hits.load(); *misses = g_tmap_misses.load();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_raw = nullptr;

// same signature as cuTensorMapEncodeTiled; interleave NONE and OOB fill NONE are the only values on the path
static CUresult g_encode(CUtensorMap* m, CUtensorMapDataType dt, cuuint32_t rank, void* ptr, const cuuint64_t* dims, const cuuint64_t* strides,
                         const cuuint32_t* box, const cuuint32_t* es, CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2,
                         CUtensorMapFloatOOBfill oob) {
    TmapKey k;
    memset(&k, 0, sizeof(k));
    k.ptr = (unsigned long long)(uintptr_t)ptr; k.dtype = (unsigned)dt; k.rank = rank; k.swizzle = (unsigned)sw; k.l2 = (unsigned)l2;
    k.dev = (unsigned)current_device();
    for (cuuint32_t i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; }
    for (cuuint32_t i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
    {
        std::lock_guard<std::mutex> g(g_tmap_mutex);
        auto it = g_tmap_cache.find(k);
        if (it != g_tmap_cache.end()) { *m = it->second; g_tmap_hits.fetch_add(1, std::memory_order_relaxed); return CUDA_SUCCESS; }
    }
    const CUresult r = g_encode_raw(m, dt, rank, ptr, dims, strides, box, es, il, sw, l2, oob);
    if (r == CUDA_SUCCESS) {
        std::lock_guard<std::mutex> g(g_tmap_mutex);
        if (g_tmap_cache.size() >= kTmapCacheMax) g_tmap_cache.clear();
        g_tmap_cache.emplace(k, *m);
        g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
    }
    return r;
}

// Plan-only mode (snn_conv_plan): the host-side launch planning below runs exactly as for a launch, but tensor maps are not
// encoded, no CUDA call is made and nothing is launched; the chosen configuration lands in t_plan.v.  Usable without a GPU
// (the SM count then defaults to 148), which is how the CPU test suite covers this logic.
struct PlanSink { bool active = false; int v[24]; };
static thread_local PlanSink t_plan;

static int get_encode() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode_raw = (EncodeTiledFn)fn;
    });
    SNN_REQUIRE(g_encode_raw != nullptr, "cuTensorMapEncodeTiled not available from the CUDA driver");
    return 0;
}

// NHWC bf16 activation view -> 5-D map (c, w, row-phase, h, n).
//   plain : tensor (NB, H, W, C) with pixel stride ld           -> dims (C, W, 1, H, NB)
//   phase : same memory seen as (NB, H/2, 2, W/2, [2 pixels])   -> dims (ld + C, W/2, 2, H/2, NB)
//   hnw   : (plain only) dimensions 3 and 4 swapped -> (C, W, 1, NB, H): a box lands in shared memory ordered (h, n, w)
static int make_act_map(CUtensorMap* m, const void* ptr, int NB, int H, int W, int C, long long ld, int phase_view,
                        int box_n, int box_h, int box_w, int hnw = 0) {
    if (t_plan.active) return 0;
    if (get_encode()) return 2;
    SNN_REQUIRE(((uintptr_t)ptr & 15) == 0, "activation pointer must be 16-byte aligned");
    SNN_REQUIRE(ld % 8 == 0 && C % 8 == 0, "activation channels/stride must be multiples of 8 (C=%d ld=%lld)", C, ld);
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5] = {64, (cuuint32_t)box_w, 1, (cuuint32_t)box_h, (cuuint32_t)box_n}, es[5] = {1, 1, 1, 1, 1};
    if (!phase_view && hnw) {
        // strides[i] = byte stride of dimension i + 1: row-phase (size 1), then n, then h
        dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = NB; dims[4] = H;
        strides[0] = ld * 2; strides[1] = (cuuint64_t)W * ld * 2; strides[2] = (cuuint64_t)H * W * ld * 2;
        strides[3] = (cuuint64_t)W * ld * 2;
        box[3] = (cuuint32_t)box_n; box[4] = (cuuint32_t)box_h;
    } else if (!phase_view) {
        dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = NB;
        strides[0] = ld * 2; strides[1] = (cuuint64_t)W * ld * 2; strides[2] = (cuuint64_t)W * ld * 2;
        strides[3] = (cuuint64_t)H * W * ld * 2;
    } else {
        SNN_REQUIRE(!hnw, "make_act_map: (h, n, w) order is for plain views only");
        SNN_REQUIRE(H % 2 == 0 && W % 2 == 0, "stride-2 / transposed conv needs even H, W (got %dx%d)", H, W);
        SNN_REQUIRE(C % 64 == 0, "stride-2 / transposed conv needs C %% 64 == 0 (got %d)", C);
        dims[0] = ld + C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = NB;
        strides[0] = 2 * ld * 2; strides[1] = (cuuint64_t)W * ld * 2; strides[2] = (cuuint64_t)2 * W * ld * 2;
        strides[3] = (cuuint64_t)H * W * ld * 2;
    }
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SNN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(act) failed: %d (NB=%d H=%d W=%d C=%d ld=%lld phase=%d box=%d,%d,%d)",
                (int)r, NB, H, W, C, ld, phase_view, box_n, box_h, box_w);
    return 0;
}

// weights bf16 [wN][wT][wK] -> 3-D map (k, tap, n), box (64, 1, box_n)
static int make_w_map(CUtensorMap* m, const void* ptr, int wN, int wT, int wK, int box_n) {
    if (t_plan.active) return 0;
    if (get_encode()) return 2;
    SNN_REQUIRE(((uintptr_t)ptr & 15) == 0, "weight pointer must be 16-byte aligned");
    SNN_REQUIRE(wK % 8 == 0, "weight K extent must be a multiple of 8 (got %d)", wK);
    cuuint64_t dims[3] = {(cuuint64_t)wK, (cuuint64_t)wT, (cuuint64_t)wN};
    cuuint64_t strides[2] = {(cuuint64_t)wK * 2, (cuuint64_t)wT * wK * 2};
    cuuint32_t box[3] = {64, 1, (cuuint32_t)box_n}, es[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SNN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed: %d (N=%d T=%d K=%d box_n=%d)", (int)r, wN, wT, wK, box_n);
    return 0;
}

// output tensor (NB, Ho, Wo, n_store) with pixel stride ld, fp32 or bf16 -> 5-D STORE map; box = one epilogue warp's
// 32 rows x 128 bytes.  os == 2 (transposed conv / stride-2 dgrad): the phase view of make_act_map, the store's channel
// coordinate selects the column phase.
static int make_out_map(CUtensorMap* m, void* ptr, int f32, int NB, int Ho, int Wo, int n_store, long long ld, int os,
                        int sbw, int sbh, int sbn, int hnw = 0) {
    if (t_plan.active) return 0;
    if (get_encode()) return 2;
    const cuuint64_t es = f32 ? 4 : 2;
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)sbw, 1, (cuuint32_t)sbh, (cuuint32_t)sbn}, est[5] = {1, 1, 1, 1, 1};
    if (os == 1 && hnw) {
        dims[0] = n_store; dims[1] = Wo; dims[2] = 1; dims[3] = NB; dims[4] = Ho;
        strides[0] = ld * es; strides[1] = (cuuint64_t)Wo * ld * es; strides[2] = (cuuint64_t)Ho * Wo * ld * es;
        strides[3] = (cuuint64_t)Wo * ld * es;
        box[3] = (cuuint32_t)sbn; box[4] = (cuuint32_t)sbh;
    } else if (os == 1) {
        dims[0] = n_store; dims[1] = Wo; dims[2] = 1; dims[3] = Ho; dims[4] = NB;
        strides[0] = ld * es; strides[1] = (cuuint64_t)Wo * ld * es; strides[2] = (cuuint64_t)Wo * ld * es;
        strides[3] = (cuuint64_t)Ho * Wo * ld * es;
    } else {
        dims[0] = ld + n_store; dims[1] = Wo / 2; dims[2] = 2; dims[3] = Ho / 2; dims[4] = NB;
        strides[0] = 2 * ld * es; strides[1] = (cuuint64_t)Wo * ld * es; strides[2] = (cuuint64_t)2 * Wo * ld * es;
        strides[3] = (cuuint64_t)Ho * Wo * ld * es;
    }
    CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ptr, dims, strides, box, est,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SNN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(out) failed: %d (NB=%d Ho=%d Wo=%d C=%d ld=%lld os=%d box=%d,%d,%d)",
                (int)r, NB, Ho, Wo, n_store, ld, os, sbw, sbh, sbn);
    return 0;
}

static int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// choose a power-of-two pixel box (bn, bh, bw) with bn*bh*bw == npix that wastes the least
static void pick_box(int NB, int H, int W, int npix, int* bn, int* bh, int* bw) {
    double best = -1;
    for (int w = npix; w >= 1; w >>= 1) {
        if (w > 256) continue;
        for (int h = npix / w; h >= 1; h >>= 1) {
            const int n = npix / (w * h);
            if (w > pow2_ceil(W) || h > pow2_ceil(H)) continue;
            if (n > 256) continue;
            const double ew = (double)W / (((W + w - 1) / w) * w), eh = (double)H / (((H + h - 1) / h) * h),
                         en = (double)NB / (((NB + n - 1) / n) * n);
            const double e = ew * eh * en + 1e-6 * w;  // tie -> wider rows
            if (e > best) { best = e; *bn = n; *bh = h; *bw = w; }
        }
    }
}

enum { GEOM_3x3_S1 = 0, GEOM_3x3_S2 = 1, GEOM_1x1 = 2, GEOM_T2x2_S2 = 3 };

// Test knobs (snn_debug_set): plain relaxed atomics, read once per launch on the calling host thread.  Knob 7 (timing
// experiments with invalid results) only exists in -DSNN_TIMING_KNOBS builds.
struct DebugFlag {
    std::atomic<int> v{0};
    operator int() const { return v.load(std::memory_order_relaxed); }
};
static DebugFlag g_debug_flags[16];
void neuron_debug_set(int k, int v);
void debug_set(int k, int v) {
#ifndef SNN_TIMING_KNOBS
    if (k == 7) return;
#endif
    if ((k >= 0 && k < 8) || (k >= 12 && k < 16)) g_debug_flags[k].v.store(v, std::memory_order_relaxed);
    else if (k >= 8 && k < 12) neuron_debug_set(k - 8, v);
}

static int smem_budget() { return 220 * 1024; }

// Counter pairs of the dynamic tile scheduler: a per-device pool, allocated and zeroed once (library-owned scratch, 32 KB;
// the only device memory the library holds).  Every launch takes the next pair; a kernel leaves its pair zeroed, launches that
// share a pair are 4096 launches apart (stream-ordered in every use on the path; a captured graph re-uses its nodes' pairs on
// every replay).
constexpr int kSchedPairs = 4096;
static unsigned int* g_sched_pool[64];
static std::atomic<unsigned int> g_sched_next{0};
static std::atomic<int> g_dynamic_tiles{0};
void set_dynamic_tiles(int on) { g_dynamic_tiles.store(on ? 1 : 0, std::memory_order_relaxed); }
static int sched_counters(unsigned int** out) {
    *out = nullptr;
    if (!g_dynamic_tiles.load(std::memory_order_relaxed)) return 0;      // static tile walk
    const int dev = current_device();
    SNN_REQUIRE(dev >= 0 && dev < 64, "bad CUDA device %d", dev);
    static PerDeviceOnce once;
    SNN_CUDA_OK(once.run([dev] {
        cudaError_t e = cudaMalloc(&g_sched_pool[dev], sizeof(unsigned int) * 2 * kSchedPairs);
        if (e == cudaSuccess) e = cudaMemset(g_sched_pool[dev], 0, sizeof(unsigned int) * 2 * kSchedPairs);
        return e;
    }));
    *out = g_sched_pool[dev] + 2 * (g_sched_next.fetch_add(1u, std::memory_order_relaxed) % kSchedPairs);
    return 0;
}

// w / w_dims: the weight tensor behind the B operand (the map is built here because its box depends on PAIR)
static int make_w_map(CUtensorMap* m, const void* ptr, int wN, int wT, int wK, int box_n);
struct WDesc { const void* ptr; int wN, wT, wK; };

// CTA pairs need an even split of the B tile: K-major rows in multiples of 8 and N % 16 == 0 (cta_group::2),
// MN-major 64-column boxes; g_debug_flags[6] == 1 forces the single-CTA kernel (tests / A-B timing)
static bool pair_possible(const ConvGemmParams& p) {
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    if (m_tiles < 2 || g_debug_flags[6] == 1) return false;
    return p.b_mn ? (p.BN % 128 == 0) : (p.BN % 16 == 0);
}

// Row-strip mode for a 3x3 stride-1 conv (see ConvGemmParams::ntap)?  The box is staged with its rows ordered (h, n, w)
// (ConvGemmParams::hnw), so the three views of a [bh + 2][bn][bw] box shifted by one image row are contiguous 128-row
// ranges bn * bw rows apart also when a box spans several images (8x8 and 4x4 maps); needs bn * bw to be a multiple of 8
// (view shift = whole 1024-byte swizzle atoms) and room for >= 2 of the deeper stages (knob 12 == 2: >= 3, which leaves the 256-column tiles
// tap-by-tap; measured 0.8 % slower per step, profiles/README.md); knob 12 == 1 turns the mode off.  Set domain, BN and b_mn
// before calling.
static bool strip_mode_ok(const ConvGemmParams& p) {
    if (g_debug_flags[12] == 1) return false;
    if ((p.bn * p.bw) % 8 != 0) return false;
    // 4x4 maps: the box covers the whole height, 2 of its 6 rows are always padding, and the 2-stage pipeline costs more than
    // the saved bytes (measured: 1024->4096@4x4 1488 -> 1360 TF/s); 8x8 maps gain (512<-512 dgrad 1198 -> 1362)
    if (p.bh < 8 && p.tiles_h == 1) return false;
    if (g_debug_flags[12] == 3 && p.bn != 1) return false;          // A/B timing: only boxes inside one image
    const int bn_cta = pair_possible(p) ? p.BN / 2 : p.BN;
    const int b_bytes = p.b_mn ? ((bn_cta + 63) / 64) * 8192 : bn_cta * 128;
    const int chunk = (p.bh + 2) * p.bn * p.bw * 128 + 3 * b_bytes;
    const int stages = (220 * 1024 - 4 * 2 * 4096) / chunk;          // epilogue staging of the TMA-store path taken out
    return stages >= (g_debug_flags[12] == 2 ? 3 : 2);
}

static int launch_conv_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const WDesc& wd, ConvGemmParams& p, cudaStream_t st) {
    static PerDeviceOnce once;
    if (!t_plan.active) {
        SNN_CUDA_OK(once.run([] {
            cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
            return e;
        }));
    }
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    bool pair = pair_possible(p);
    // K chunks per work item: small-K convs (1x1, transposed) are bound by per-tile latencies, not by operand traffic:
    // single-CTA tiles (no cluster handshakes) and four epilogue staging tiles instead of two
    if (p.ksplit < 1) p.ksplit = 1;
    const bool strip = p.ntap > 1;            // set by the caller after strip_mode_ok()
    if (!strip) { p.ntap = 1; p.tap_wstep = 0; }
    int chunks_per_item = 0;
    for (int i = 0; i < p.phase[0].nseg; ++i) chunks_per_item += p.phase[0].seg[i].nchunk * p.ntap;
    chunks_per_item /= p.ksplit;
    const bool small_k = chunks_per_item <= 8;
    if (small_k && g_debug_flags[6] != 2) pair = false;
    SNN_REQUIRE(!(strip && small_k), "conv_gemm: row-strip mode on a small-K conv");
    const int bn_cta = pair ? p.BN / 2 : p.BN;
    p.b_boxes = (bn_cta + 63) / 64;
    p.b_bytes = p.b_mn ? p.b_boxes * 8192 : bn_cta * 128;
    p.a_bytes = strip ? (p.bh + 2) * p.bn * p.bw * 128 : 16384;
    p.a_tap_off = strip ? p.bn * p.bw * 128 : 0;
    p.hnw = strip ? 1 : 0;
    p.chunk_bytes = p.a_bytes + p.ntap * p.b_bytes;
    // chunks per stage: keep >= ~512 MMA cycles behind every mbarrier round trip (N = 256: 1 chunk, 128: 2, <= 64: 4)
    p.cps = (small_k || strip) ? 1 : (p.BN > 128 ? 1 : (p.BN > 64 ? 2 : 4));
    for (int ph = 0; ph < p.nphase && p.cps > 1; ++ph)          // ragged segments (e.g. 144 channels = 3 chunks) would leave
        for (int i = 0; i < p.phase[ph].nseg; ++i)               // half-empty stages behind: fall back to one chunk per stage
            while (p.cps > 1 && p.phase[ph].seg[i].nchunk % p.cps != 0) p.cps >>= 1;
    if (g_debug_flags[2] > 0) p.cps = g_debug_flags[2];
    // coalesced TMA-store epilogue whenever the output view is expressible as a tensor map
    const int es = p.out_f32 ? 4 : 2, CH = p.out_f32 ? 32 : 64;
    char* obase = reinterpret_cast<char*>(p.out) + (long long)p.out_coff * es;
    // (accumulating into a bf16 output = cp.reduce.async.bulk.tensor .add on a bf16 tensor map: one rounding of the sum)
    p.tma_out = g_debug_flags[0] != 1 && ((uintptr_t)obase % 16 == 0) && ((p.out_ld * es) % 16 == 0) &&
                (p.os == 1 || (p.n_store % CH == 0 && p.Ho % 2 == 0 && p.Wo % 2 == 0));
    SNN_REQUIRE(!p.stats || p.tma_out, "conv_fprop: fused statistics need the TMA-store epilogue (output alignment)");
    p.nbuf = 2;                       // (four staging tiles per warp were measured to change nothing, profiles/README.md)
    p.epi_groups = (small_k && g_debug_flags[3] != 1) ? 2 : 1;          // knob 3 = 1: one epilogue group everywhere (A/B timing, tests)
    const int stage_extra = p.tma_out ? 4 * p.epi_groups * p.nbuf * 4096 : 0;     // epilogue warps x nbuf staging tiles of 32 rows x 128 B
    while (p.cps > 1 && (smem_budget() - stage_extra) / (p.cps * p.chunk_bytes) < 3) p.cps >>= 1;   // keep >= 3 stages
    p.stage_bytes = p.cps * p.chunk_bytes;
    int stages = (smem_budget() - stage_extra) / p.stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (g_debug_flags[1] > 0 && stages > g_debug_flags[1]) stages = g_debug_flags[1];
    SNN_REQUIRE(stages >= 2, "conv_gemm: not enough shared memory for 2 stages");
    p.stages = stages;
    const size_t smem = (size_t)stages * p.stage_bytes + stage_extra + 1024;
    p.n_blocks = (p.n_store + p.BN - 1) / p.BN;
    if (sched_counters(&p.sched)) return 2;
#ifdef SNN_TIMING_KNOBS
    p.dbg = g_debug_flags[7];
#else
    p.dbg = 0;
#endif
    if (t_plan.active) {
        const int tiles = (pair ? (m_tiles + 1) / 2 : m_tiles) * p.n_blocks * p.nphase * p.ksplit;
        const int workers = pair ? num_sms() / 2 : num_sms();
        const int v[20] = {p.bn, p.bh, p.bw, p.BN, pair ? 1 : 0, p.ntap > 1 ? 1 : 0, p.hnw, p.stages, p.stage_bytes, p.cps, p.ksplit,
                           tiles, p.n_blocks, workers < tiles ? workers : tiles, p.epi_groups, small_k ? 1 : 0, p.tma_out, (int)smem,
                           p.a_bytes, p.b_bytes};
        for (int i = 0; i < 20; ++i) t_plan.v[i] = v[i];
        return 0;
    }
    CUtensorMap b, o;
    if (make_w_map(&b, wd.ptr, wd.wN, wd.wT, wd.wK, p.b_mn ? 64 : bn_cta)) return 2;
    if (p.tma_out) {
        // one epilogue warp's 32 rows as a box: (n, h, w) order fills w, then h, then n; (h, n, w) order w, then n, then h
        const int sbw = p.bw >= 32 ? 32 : p.bw;
        int sbh, sbn;
        if (p.hnw) { sbn = (32 / sbw) < p.bn ? (32 / sbw) : p.bn; sbh = 32 / (sbw * sbn); }
        else { sbh = (32 / sbw) < p.bh ? (32 / sbw) : p.bh; sbn = 32 / (sbw * sbh); }
        if (make_out_map(&o, obase, p.out_f32, p.NB, p.Ho, p.Wo, p.n_store, p.out_ld, p.os, sbw, sbh, sbn, p.hnw)) return 2;
    } else {
        o = b;
    }
    if (!pair) {
        const int total_tiles = m_tiles * p.n_blocks * p.nphase * p.ksplit;
        int grid = num_sms();
        if (g_debug_flags[5] > 0) grid = g_debug_flags[5];
        if (grid > total_tiles) grid = total_tiles;
        return check_cuda(launch_pdl(conv_gemm_kernel<false>, dim3(grid), dim3(64 + 128 * p.epi_groups), smem, st, a0, a1, b, o, p),
                          "conv_gemm_kernel launch");
    }
    const int total_tiles = ((m_tiles + 1) / 2) * p.n_blocks * p.nphase * p.ksplit;
    int pairs = num_sms() / 2;
    if (g_debug_flags[5] > 0) pairs = g_debug_flags[5];
    if (pairs > total_tiles) pairs = total_tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    cfg.blockDim = dim3(64 + 128 * p.epi_groups, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true>, a0, a1, b, o, p), "conv_gemm_kernel<pair> launch");
}

static int pick_bn(int n_store) {
    // largest UMMA N (multiple of 16, <= 256) that tiles n_store with little padding
    if (n_store <= 256) return ((n_store + 15) / 16) * 16;
    int best = 256; double beste = -1;
    for (int bn = 256; bn >= 64; bn -= 16) {
        const int tiles = (n_store + bn - 1) / bn;
        const double e = (double)n_store / (tiles * bn) + 1e-4 * bn / 256.0;
        if (e > beste) { beste = e; best = bn; }
    }
    return best;
}

// nb_for_box: the image count the pixel box is chosen for.  Convs with fused BatchNorm statistics pass B (frames per
// timestep) instead of NB = T*B: the 32-pixel partial-sum groups then cover the same pixels however many timesteps are
// folded into the launch, so the fused T-step sequence and T per-frame calls give bit-identical statistics.
static void set_domain(ConvGemmParams& p, int NB, int Hd, int Wd, int nb_for_box = 0) {
    p.NB = NB; p.Hd = Hd; p.Wd = Wd;
    pick_box(nb_for_box > 0 ? nb_for_box : NB, Hd, Wd, 128, &p.bn, &p.bh, &p.bw);
    p.tiles_w = (Wd + p.bw - 1) / p.bw; p.tiles_h = (Hd + p.bh - 1) / p.bh; p.tiles_n = (NB + p.bn - 1) / p.bn;
}

static Seg mkseg(int src, int dc, int dw, int dhp, int dh, int wtap, int wc0, int C) {
    Seg s; s.src = src; s.dc = dc; s.dw = dw; s.dhp = dhp; s.dh = dh; s.wtap = wtap; s.wc0 = wc0; s.nchunk = (C + 63) / 64;
    return s;
}

// stride-2 3x3 (pad 1) tap -> phase-view offsets: input row 2*o + k - 1
static void s2_tap(int k, int* phase, int* d) {
    if (k == 0) { *phase = 1; *d = -1; } else if (k == 1) { *phase = 0; *d = 0; } else { *phase = 1; *d = 0; }
}

// Fused BatchNorm statistics (conv_fprop(..., stats, B)): the epilogue writes one (sum, sumsq) partial per 32-row group and
// column; 4 groups per 128-pixel tile, tiles enumerated (n, h, w).  Available when every tile lies inside ONE timestep:
// the pixel box covers bn images and B (frames per timestep) is a multiple of bn.  Returns the number of groups
// (0 = not available: use snn_bn_stats) and the number of groups per timestep.
long long conv_stats_groups(int geom, int NB, int H, int W, int B, int* groups_per_t) {
    if (geom != GEOM_3x3_S1 && geom != GEOM_3x3_S2 && geom != GEOM_1x1) return 0;
    if (B <= 0 || NB % B != 0) return 0;
    int Hd = H, Wd = W;
    if (geom == GEOM_3x3_S2) { Hd = H / 2; Wd = W / 2; }
    ConvGemmParams p;
    set_domain(p, NB, Hd, Wd, B);
    if (p.bn > 1 && B % p.bn != 0) return 0;
    const int tiles_hw = p.tiles_w * p.tiles_h;
    if (groups_per_t) *groups_per_t = 4 * (B / p.bn) * tiles_hw;
    return 4LL * tiles_hw * p.tiles_n;
}

// ------------------------------------------------------------------------------------------
// fprop: out[NB,Ho,Wo,Cout] = conv(cat(x0,x1)) (+bias) ; geometry decides Ho, Wo
// ------------------------------------------------------------------------------------------
int conv_fprop(int geom, int NB, int H, int W, const void* x0, int C0, long long ld0, const void* x1, int C1, long long ld1,
               const void* w, int w_rows, int w_K, int w_coff, int Cout, int w_row_off, const float* bias, void* out,
               int out_f32, long long out_ld, int out_coff, int accumulate, cudaStream_t st, float* stats, int frames_per_step) {
    SNN_REQUIRE(Cout % 8 == 0, "conv_fprop: Cout=%d must be a multiple of 8", Cout);
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.stats = stats;
    const int taps = geom == GEOM_3x3_S1 || geom == GEOM_3x3_S2 ? 9 : (geom == GEOM_1x1 ? 1 : 4);
    const int phase_view = geom == GEOM_3x3_S2;
    int Hd = H, Wd = W;
    p.os = 1; p.Ho = H; p.Wo = W;
    if (geom == GEOM_3x3_S2) { Hd = H / 2; Wd = W / 2; p.Ho = Hd; p.Wo = Wd; }
    if (geom == GEOM_T2x2_S2) { p.os = 2; p.Ho = 2 * H; p.Wo = 2 * W; }
    set_domain(p, NB, Hd, Wd, stats ? frames_per_step : 0);
    p.BN = pick_bn(Cout);
    p.n_store = Cout; p.wn_off = w_row_off;
    p.out = out; p.bias = bias; p.out_f32 = out_f32; p.accumulate = accumulate; p.out_ld = out_ld; p.out_coff = out_coff;
    const bool strip = geom == GEOM_3x3_S1 && strip_mode_ok(p);
    const int box_h = strip ? p.bh + 2 : p.bh;
    CUtensorMap a0, a1;
    if (make_act_map(&a0, x0, NB, H, W, C0, ld0, phase_view, p.bn, box_h, p.bw, strip)) return 2;
    if (x1) { if (make_act_map(&a1, x1, NB, H, W, C1, ld1, phase_view, p.bn, box_h, p.bw, strip)) return 2; } else a1 = a0;
    const WDesc wd = {w, w_rows, taps, w_K};
    if (stats) {
        int gpt = 0;
        SNN_REQUIRE(conv_stats_groups(geom, NB, H, W, frames_per_step, &gpt) > 0 && out_f32 && !accumulate && !bias && out_coff == 0,
                    "conv_fprop: fused BatchNorm statistics are not available for this call (geom %d NB %d %dx%d B %d)", geom,
                    NB, H, W, frames_per_step);
    }
    if (geom == GEOM_T2x2_S2) {
        p.nphase = 4;
        for (int a = 0; a < 2; ++a)
            for (int bb = 0; bb < 2; ++bb) {
                Phase& ph = p.phase[a * 2 + bb];
                ph.oph = a; ph.opw = bb; ph.nseg = 0;
                ph.seg[ph.nseg++] = mkseg(0, 0, 0, 0, 0, a * 2 + bb, w_coff, C0);
                if (x1) ph.seg[ph.nseg++] = mkseg(1, 0, 0, 0, 0, a * 2 + bb, w_coff + C0, C1);
            }
    } else if (strip) {
        // one segment per stencil column kw: box rows h0 - 1 .. h0 + bh, taps kh = 0..2 = weight taps kw, kw + 3, kw + 6
        p.nphase = 1; p.ntap = 3; p.tap_wstep = 3;
        Phase& ph = p.phase[0];
        ph.nseg = 0;
        for (int kw = 0; kw < 3; ++kw) {
            ph.seg[ph.nseg++] = mkseg(0, 0, kw - 1, 0, -1, kw, w_coff, C0);
            if (x1) ph.seg[ph.nseg++] = mkseg(1, 0, kw - 1, 0, -1, kw, w_coff + C0, C1);
        }
    } else {
        p.nphase = 1;
        Phase& ph = p.phase[0];
        ph.nseg = 0;
        const int k = geom == GEOM_1x1 ? 1 : 3;
        for (int kh = 0; kh < k; ++kh)
            for (int kw = 0; kw < k; ++kw) {
                int dcm = 0, dw = 0, dhp = 0, dh = 0;
                if (geom == GEOM_3x3_S1) { dw = kw - 1; dh = kh - 1; }
                if (geom == GEOM_3x3_S2) { int pc; s2_tap(kh, &dhp, &dh); s2_tap(kw, &pc, &dw); dcm = pc; }
                ph.seg[ph.nseg++] = mkseg(0, (int)(dcm * ld0), dw, dhp, dh, kh * k + kw, w_coff, C0);
                if (x1) ph.seg[ph.nseg++] = mkseg(1, (int)(dcm * ld1), dw, dhp, dh, kh * k + kw, w_coff + C0, C1);
            }
    }
    return launch_conv_gemm(a0, a1, wd, p, st);
}

// ------------------------------------------------------------------------------------------
// dgrad: dx[NB,H,W,Ci] = conv^T(dy).  The weights are read IN PLACE from the fprop layout [Cout][tap][Cin_tot]
// (B operand MN-major: N = input channel contiguous, K = output channel = row) -- no transposed copy exists.
// (H, W) are the conv's INPUT spatial dims; dy has the geometry's output dims.
// ------------------------------------------------------------------------------------------
int conv_dgrad(int geom, int NB, int H, int W, const void* dy, int Cout, long long ld_dy, const void* wt, int w_K,
               int ci_off, int Ci, void* dx, int dx_f32, long long dx_ld, int dx_coff, int accumulate, cudaStream_t st) {
    SNN_REQUIRE(Ci % 8 == 0, "conv_dgrad: Ci=%d must be a multiple of 8", Ci);
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    const int taps = geom == GEOM_3x3_S1 || geom == GEOM_3x3_S2 ? 9 : (geom == GEOM_1x1 ? 1 : 4);
    int Hd = H, Wd = W, Hy = H, Wy = W, phase_view = 0;
    p.os = 1; p.Ho = H; p.Wo = W;
    if (geom == GEOM_3x3_S2) { Hd = H / 2; Wd = W / 2; Hy = Hd; Wy = Wd; p.os = 2; }
    if (geom == GEOM_T2x2_S2) { Hy = 2 * H; Wy = 2 * W; phase_view = 1; }
    set_domain(p, NB, Hd, Wd);
    p.BN = pick_bn(Ci);
    p.n_store = Ci; p.wn_off = ci_off; p.b_mn = 1;
    p.out = dx; p.bias = nullptr; p.out_f32 = dx_f32; p.accumulate = accumulate; p.out_ld = dx_ld; p.out_coff = dx_coff;
    // Small-M convs (ConvLSTM recurrent dgrad: 8 pixel tiles x 4 column blocks = 16 work items for 74 CTA pairs) split K
    // over the taps; the partial products meet in the fp32 output through TMA reduce-add (zero-filled first).
    int ks = 1;
    if (dx_f32 && !accumulate && geom != GEOM_3x3_S2 && dx_coff == 0 && dx_ld == Ci && g_debug_flags[0] != 1) {
        const int m_work = (p.tiles_w * p.tiles_h * p.tiles_n + 1) / 2;
        const int items = m_work * ((Ci + p.BN - 1) / p.BN);
        const int nseg = taps, pairs = num_sms() / 2;           // one K segment per tap in these geometries
        for (int cand = 1; cand <= nseg; ++cand)
            if (nseg % cand == 0 && items * cand <= 2 * pairs) ks = cand;
        if (!(ks > 1 && items <= pairs / 2)) ks = 1;
    }
    if (deterministic()) ks = 1;
    const bool strip = ks == 1 && geom == GEOM_3x3_S1 && strip_mode_ok(p);
    CUtensorMap a0;
    if (make_act_map(&a0, dy, NB, Hy, Wy, Cout, ld_dy, phase_view, p.bn, strip ? p.bh + 2 : p.bh, p.bw, strip)) return 2;
    const WDesc wd = {wt, Cout, taps, w_K};                  // box = 64 input channels x 1 tap x 64 output-channel rows
    if (strip) {
        // dx[h, w] = sum dy[h + 1 - kh, w + 1 - kw] * W[kh, kw]: column offset d = 1 - kw; box row r = 0..2 reads dy row
        // h - 1 + r, i.e. kh = 2 - r -> weight taps (2 - r) * 3 + kw = 6 + kw, 3 + kw, kw
        p.nphase = 1; p.ntap = 3; p.tap_wstep = -3;
        Phase& ph = p.phase[0];
        for (int kw = 0; kw < 3; ++kw) ph.seg[ph.nseg++] = mkseg(0, 0, 1 - kw, 0, -1, 6 + kw, 0, Cout);
    } else if (geom == GEOM_3x3_S1 || geom == GEOM_1x1) {
        p.nphase = 1;
        Phase& ph = p.phase[0];
        const int k = geom == GEOM_1x1 ? 1 : 3;
        for (int kh = 0; kh < k; ++kh)
            for (int kw = 0; kw < k; ++kw)
                ph.seg[ph.nseg++] = mkseg(0, 0, k == 3 ? 1 - kw : 0, 0, k == 3 ? 1 - kh : 0, kh * k + kw, 0, Cout);
    } else if (geom == GEOM_3x3_S2) {
        // input pixel (2a+pr, 2b+pc): pr==0 -> kh=1 (dh 0); pr==1 -> kh=0 (dh +1), kh=2 (dh 0)
        p.nphase = 4;
        for (int pr = 0; pr < 2; ++pr)
            for (int pc = 0; pc < 2; ++pc) {
                Phase& ph = p.phase[pr * 2 + pc];
                ph.oph = pr; ph.opw = pc; ph.nseg = 0;
                for (int kh = 0; kh < 3; ++kh) {
                    if ((kh == 1) != (pr == 0)) continue;
                    const int dh = kh == 0 ? 1 : 0;
                    for (int kw = 0; kw < 3; ++kw) {
                        if ((kw == 1) != (pc == 0)) continue;
                        const int dw = kw == 0 ? 1 : 0;
                        ph.seg[ph.nseg++] = mkseg(0, 0, dw, 0, dh, kh * 3 + kw, 0, Cout);
                    }
                }
            }
    } else {  // transposed conv: dx[h,w] = sum_{a,b} dy[2h+a, 2w+b] * W[:, :, a, b]
        p.nphase = 1;
        Phase& ph = p.phase[0];
        for (int a = 0; a < 2; ++a)
            for (int bb = 0; bb < 2; ++bb) ph.seg[ph.nseg++] = mkseg(0, (int)(bb * ld_dy), 0, a, 0, a * 2 + bb, 0, Cout);
    }
    if (ks > 1) {
        if (!t_plan.active) SNN_CUDA_OK(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)NB * H * W * Ci, st));
        p.ksplit = ks;
        p.accumulate = 1;
    }
    return launch_conv_gemm(a0, a0, wd, p, st);
}

// ------------------------------------------------------------------------------------------
// wgrad: dw[Cout][taps][w_K] (fp32, +=) at channel offset w_coff, from x (conv input) and dy
// ------------------------------------------------------------------------------------------
int conv_wgrad(int geom, int NB, int H, int W, const void* x, int Ci, long long ld_x, const void* dy, int Cout,
               long long ld_dy, float* dw, int w_K, int w_coff, cudaStream_t st) {
    static PerDeviceOnce once;
    if (!t_plan.active) {
        SNN_CUDA_OK(once.run([] {
            cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
            return e;
        }));
    }
    SNN_REQUIRE(Ci % 8 == 0 && Cout % 8 == 0 && w_K % 4 == 0 && w_coff % 4 == 0, "conv_wgrad: channel counts must be multiples of 8");
    WgradParams p;
    memset(&p, 0, sizeof(p));
    const int taps = geom == GEOM_3x3_S1 || geom == GEOM_3x3_S2 ? 9 : (geom == GEOM_1x1 ? 1 : 4);
    int Hd = H, Wd = W, Hy = H, Wy = W, x_phase = 0, g_phase = 0;
    if (geom == GEOM_3x3_S2) { Hd = H / 2; Wd = W / 2; Hy = Hd; Wy = Wd; x_phase = 1; }
    if (geom == GEOM_T2x2_S2) { Hy = 2 * H; Wy = 2 * W; g_phase = 1; }
    p.NB = NB; p.Hd = Hd; p.Wd = Wd;
    pick_box(NB, Hd, Wd, 64, &p.bn, &p.bh, &p.bw);
    p.tiles_w = (Wd + p.bw - 1) / p.bw; p.tiles_h = (Hd + p.bh - 1) / p.bh; p.tiles_n = (NB + p.bn - 1) / p.bn;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    // cin per CTA: multiple of 64 up to 256 with little padding
    {
        const int c64 = (Ci + 63) / 64;
        int nt = c64 >= 4 ? 4 : c64;
        if (c64 > 4) { // prefer an exact tiling
            for (int cand = 4; cand >= 2; --cand) if (c64 % cand == 0) { nt = cand; break; }
        }
        p.NT = nt * 64;
        p.n_ci_tiles = (Ci + p.NT - 1) / p.NT;
    }
    p.Cout = Cout; p.Ci = Ci; p.dw = dw; p.dw_ld = (long long)taps * w_K; p.wK = w_K; p.w_coff = w_coff;
    // Row-strip mode (WgradParams::strip): 3x3 stride-1, bn * bw a multiple of 8 (boxes ordered (h, n, w)), cin in tiles of
    // 128 without padding (64 for a 64-channel input): dY is read once per stencil column instead of once per tap and X once
    // per three taps -- 32 KB instead of 72 KB of operands per CTA and three taps of a 128-wide cin tile.  Knob 13 = 1: off.
    p.strip = geom == GEOM_3x3_S1 && (p.bn * p.bw) % 8 == 0 && (Ci % 128 == 0 || Ci == 64) && g_debug_flags[13] != 1 &&
              (p.bh >= 8 || p.tiles_h > 1) &&                         // not on 4x4 maps (a third of the strip box would be padding)
              !(g_debug_flags[13] == 2 && p.bn != 1);                 // knob 13 == 2 (A/B timing): only boxes inside one image
    if (p.strip) {
        p.NT = Ci == 64 ? 64 : 128;
        p.n_ci_tiles = (Ci + p.NT - 1) / p.NT;
        p.x_box_bytes = (p.bh + 2) * p.bn * p.bw * 128;
        p.x_tap_off = p.bn * p.bw * 128;
        p.hnw = 1;            // both pixel boxes ordered (h, n, w): the K (pixel) index of dY and X must agree
    }
    const bool pair = Cout > 128 && p.NT % 128 == 0 && g_debug_flags[6] != 1;
    p.stage_bytes = p.strip ? 2 * 8192 + ((pair ? p.NT / 2 : p.NT) / 64) * p.x_box_bytes : (2 + (pair ? p.NT / 2 : p.NT) / 64) * 8192;
    const int stage_extra = 4 * 8192;      // epilogue staging: 4 warps x 2 tiles of 32 x 32 fp32
    int stages = (smem_budget() - stage_extra) / p.stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    p.stages = stages;
    p.lbo_bytes = 8192;     // MN-major operands: 64-channel blocks 8192 B apart, 8-row (pixel) groups 1024 B apart
    p.sbo_bytes = 1024;
    p.ntaps = p.strip ? 3 : taps;
    const int k = geom == GEOM_1x1 ? 1 : (geom == GEOM_T2x2_S2 ? 2 : 3);
    if (p.strip) {
        for (int kw = 0; kw < 3; ++kw) {          // one entry per stencil column: X box rows h0 - 1 .. h0 + bh, weight taps kw, kw + 3, kw + 6
            WTap& t = p.taps[kw];
            memset(&t, 0, sizeof(t));
            t.wtap = kw; t.x_dw = kw - 1; t.x_dh = -1;
        }
    } else
    for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) {
            WTap& t = p.taps[kh * k + kw];
            memset(&t, 0, sizeof(t));
            t.wtap = kh * k + kw;
            if (geom == GEOM_3x3_S1) { t.x_dw = kw - 1; t.x_dh = kh - 1; }
            if (geom == GEOM_3x3_S2) { int pc; s2_tap(kh, &t.x_dhp, &t.x_dh); s2_tap(kw, &pc, &t.x_dw); t.x_dc = (int)(pc * ld_x); }
            if (geom == GEOM_T2x2_S2) { t.g_dhp = kh; t.g_dc = (int)(kw * ld_dy); }
        }
    p.m_items = pair ? (Cout + 255) / 256 : (Cout + 127) / 128;
    const int base = p.m_items * p.n_ci_tiles * p.ntaps;          // work items before splitting K
    const int workers = pair ? num_sms() / 2 : num_sms();
    // split K (pixel tiles) so that the item count is just below one or two full rounds of the persistent workers; every
    // item keeps >= 16 pipeline stages when two rounds are used (more items = more reduce-add traffic into dW)
    int ksplit = 1;
    if (base < workers) {
        const int k2 = (2 * workers) / base, k1 = workers / base;
        ksplit = (k2 >= 1 && p.total_tiles / k2 >= 16) ? k2 : (k1 >= 1 ? k1 : 1);
        // strip mode: the epilogue (3 accumulators, not overlapped with the next item's K loop) favours ONE round of long items
        if (p.strip && k1 >= 1 && g_debug_flags[14] != 1) ksplit = k1;
    }
    const int max_split = (p.total_tiles + 3) / 4;  // >= 4 pixel tiles per item
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
    if (g_debug_flags[4] > 0) ksplit = g_debug_flags[4];
    if (deterministic()) ksplit = 1;       // one reduce-add per dW element and launch: the order of additions is the stream order
    p.tiles_per_split = (p.total_tiles + ksplit - 1) / ksplit;
    p.ksplit = (p.total_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    CUtensorMap mg, mx, mw;
    if (make_act_map(&mg, dy, NB, Hy, Wy, Cout, ld_dy, g_phase, p.bn, p.bh, p.bw, p.hnw)) return 2;
    if (make_act_map(&mx, x, NB, H, W, Ci, ld_x, x_phase, p.bn, p.strip ? p.bh + 2 : p.bh, p.bw, p.hnw)) return 2;
    if (!t_plan.active) {   // dW fp32 [Cout][taps][w_K] -> 3-D reduce-add map (k, tap, n), box 32 x 1 x 32
        if (get_encode()) return 2;
        SNN_REQUIRE(((uintptr_t)dw & 15) == 0, "conv_wgrad: dw must be 16-byte aligned");
        cuuint64_t dims[3] = {(cuuint64_t)w_K, (cuuint64_t)taps, (cuuint64_t)Cout};
        cuuint64_t strides[2] = {(cuuint64_t)w_K * 4, (cuuint64_t)taps * w_K * 4};
        cuuint32_t box[3] = {32, 1, 32}, es[3] = {1, 1, 1};
        CUresult r = g_encode(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dw, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SNN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(dw) failed: %d (Cout=%d taps=%d K=%d)", (int)r, Cout, taps, w_K);
    }
    const size_t smem = (size_t)p.stages * p.stage_bytes + stage_extra + 1024;
    if (sched_counters(&p.sched)) return 2;
    const int items = base * p.ksplit;
    int nw = workers < items ? workers : items;
    if (g_debug_flags[5] > 0 && nw > g_debug_flags[5]) nw = g_debug_flags[5];
    if (t_plan.active) {
        const int v[20] = {p.bn, p.bh, p.bw, p.NT, pair ? 1 : 0, p.strip, p.hnw, p.stages, p.stage_bytes, 1, p.ksplit, items, p.n_ci_tiles, nw, 1, 0, 1,
                           (int)smem, p.strip ? p.x_box_bytes : 8192, 2 * 8192};
        for (int i = 0; i < 20; ++i) t_plan.v[i] = v[i];
        return 0;
    }
    if (!pair) {
        return check_cuda(launch_pdl(wgrad_gemm_kernel<false>, dim3(nw), dim3(192), smem, st, mg, mx, mw, p), "wgrad_gemm_kernel launch");
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * nw, 1, 1);
    cfg.blockDim = dim3(192, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, wgrad_gemm_kernel<true>, mg, mx, mw, p), "wgrad_gemm_kernel<pair> launch");
}

// kind: 0 fprop, 1 dgrad, 2 wgrad.  out[20] = {bn, bh, bw, N tile (BN | NT), CTA pair, row-strip mode, (h,n,w) row order, stages,
// stage bytes, K chunks per stage, K split, work items, N blocks (| cin tiles), CTAs (pairs) launched, epilogue warp groups, small-K,
// TMA-store epilogue, dynamic shared memory bytes, A (| X box) bytes per stage and tap group, B (| dY) bytes per tap}
int conv_plan(int kind, int geom, int NB, int H, int W, int Cin, int Cout, int out_f32, int frames_per_step, int accumulate, int* out) {
    void* const dummy = reinterpret_cast<void*>((uintptr_t)1 << 20);      // aligned, never dereferenced
    t_plan.active = true;
    for (int i = 0; i < 24; ++i) t_plan.v[i] = 0;
    int rc;
    if (kind == 0)
        rc = conv_fprop(geom, NB, H, W, dummy, Cin, Cin, nullptr, 0, 0, dummy, Cout, Cin, 0, Cout, 0, nullptr, dummy, out_f32, Cout, 0, accumulate,
                        nullptr, frames_per_step > 0 ? static_cast<float*>(dummy) : nullptr, frames_per_step);
    else if (kind == 1)
        rc = conv_dgrad(geom, NB, H, W, dummy, Cout, Cout, dummy, Cin, 0, Cin, dummy, out_f32, Cin, 0, accumulate, nullptr);
    else
        rc = conv_wgrad(geom, NB, H, W, dummy, Cin, Cin, dummy, Cout, Cout, static_cast<float*>(dummy), Cin, 0, nullptr);
    t_plan.active = false;
    if (rc == 0) for (int i = 0; i < 20; ++i) out[i] = t_plan.v[i];
    return rc;
}

}  // namespace snn
