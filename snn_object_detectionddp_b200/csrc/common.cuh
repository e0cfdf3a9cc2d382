// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers.
// Hand-written inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <utility>

#define SNN_DEVINL __device__ __forceinline__

namespace snn {

// ------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define SNN_CUDA_OK(expr)                                   \
    do {                                                    \
        int _rc = ::snn::check_cuda((expr), #expr);         \
        if (_rc) return _rc;                                \
    } while (0)
#define SNN_REQUIRE(cond, ...)                              \
    do {                                                    \
        if (!(cond)) {                                      \
            ::snn::set_error(__VA_ARGS__);                  \
            return 2;                                       \
        }                                                   \
    } while (0)

int num_sms();          // SM count of the CURRENT device (cached per device ordinal)
int current_device();   // cudaGetDevice, -1 on error

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol): a kernel launched through launch_pdl() may become resident while its
// predecessor on the stream is still draining -- launch latency, CTA scheduling and its own barrier / TMEM / table set-up
// overlap the predecessor's tail.  CONTRACT: every kernel launched this way executes pdl_wait() before its first access
// to global memory (reads AND writes: the predecessor may still be reading what this kernel overwrites); the wait returns
// once the preceding grid has completed and its writes are visible.  pdl_launch_dependents() at the top lets the NEXT
// kernel do the same with respect to this one (it still waits for this grid's completion at its own pdl_wait()).
// Without the launch attribute both instructions are no-ops.  Process-wide switch: set_pdl() / snn_set_pdl().
// ------------------------------------------------------------------------------------------
SNN_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
SNN_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
void set_pdl(int on);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------------------------------
// Deterministic mode (set_deterministic / snn_set_deterministic): every sum whose order is otherwise decided by atomics or by
// the arrival order of TMA reduce-adds is taken in a FIXED order, so two runs from the same state give bit-identical gradients:
//   * wgrad and the small-M dgrad run without split-K (one contribution per output element and launch);
//   * block-level reductions (BatchNorm backward sums, bias column sums, depthwise wgrad, gradient norm, loss sums) write one
//     partial per block into a library-owned scratch buffer (det_scratch) and a second tiny kernel adds them in block order;
//     inside a block, per-row partials go through shared-memory slots and are summed in row order.
// Slower (the 128-channel wgrads lose their K split) -- a reproducibility / debugging mode, off by default.  The scratch is
// allocated per device at the first deterministic launch on that device: make it outside CUDA-graph capture.
// ------------------------------------------------------------------------------------------
bool deterministic();
void set_deterministic(int on);
// >= bytes of device scratch on the current device (stream-ordered reuse: all users run on the caller's stream); nullptr + error set on failure
void* det_scratch(size_t bytes, cudaStream_t st);
// out[i] += sum over b = 0 .. nblocks-1 (in this order) of part[b * n + i]
int launch_ordered_combine_f32(const float* part, int nblocks, long long n, float* out, cudaStream_t st);
int launch_ordered_combine_f64(const double* part, int nblocks, long long n, double* out, cudaStream_t st);

// One-time initialisation PER DEVICE (function attributes such as the dynamic shared-memory limit live in the device's
// context: a process that touches cuda:3 after cuda:0 must set them again).  Thread-safe; remembers the first error.
struct PerDeviceOnce {
    std::mutex m;
    unsigned long long done = 0;
    cudaError_t err[64];
    template <typename F>
    cudaError_t run(F f) {
        const int d = current_device();
        if (d < 0 || d >= 64) return f();
        std::lock_guard<std::mutex> g(m);
        if (!((done >> d) & 1ull)) { err[d] = f(); done |= 1ull << d; }
        return err[d];
    }
};

// ------------------------------------------------------------------------------------------
// Row map of the Detect head's prediction buffers.  The head writes its scales into SCALE-MAJOR buffers (scale i occupies
// rows [B*a_off_i, B*a_off_{i+1}), image-major inside), so no torch.cat ever runs; nl == 0 = natural [B, A] layout.
// ------------------------------------------------------------------------------------------
struct RowMap {
    int nl;
    int a_off[5];      // first anchor of scale i (a_off[nl] = A)
    int B;
};
__device__ __forceinline__ long long pred_row(const RowMap& rm, int b, int a, int A) {
    if (rm.nl == 0) return (long long)b * A + a;
    int i = 0;
    while (i + 1 < rm.nl && a >= rm.a_off[i + 1]) ++i;
    const int hw = rm.a_off[i + 1] - rm.a_off[i];
    return (long long)rm.B * rm.a_off[i] + (long long)b * hw + (a - rm.a_off[i]);
}
int make_rowmap(RowMap* rm, int nl, const int* a_off, int B, int A);

// ------------------------------------------------------------------------------------------
// shared-memory address / mbarrier
// ------------------------------------------------------------------------------------------
SNN_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

SNN_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
SNN_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SNN_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SNN_DEVINL void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
SNN_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
SNN_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) tiled loads, 3-D and 5-D, completing on an mbarrier
// ------------------------------------------------------------------------------------------
SNN_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
SNN_DEVINL void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
SNN_DEVINL void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------
template <int kCols>
SNN_DEVINL void tmem_alloc(uint32_t dst_smem) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
SNN_DEVINL void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
SNN_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SNN_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
SNN_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
SNN_DEVINL void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns
SNN_DEVINL void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
SNN_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, 128-byte swizzle (TMA SWIZZLE_128B tiles, 1024 B aligned).
//   K-major  : rows of 64 bf16 (128 B), 8-row groups 1024 B apart  -> SBO = 1024, LBO unused (1)
//   MN-major : 64 MN-contiguous bf16 per 128 B row, K advances by rows; 8-row (K) groups SBO apart,
//              64-wide MN blocks LBO apart.
SNN_DEVINL uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
// instruction descriptor: bf16 A/B, fp32 D, M x N, operand majors (0 = K-major, 1 = MN-major)
SNN_DEVINL uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2): the two CTAs of a pair hold 128 accumulator rows each and
// HALF of the B operand each; one thread of the leader CTA (cluster rank 0) issues the MMAs for both.
// Shared-window addresses carry the CTA's cluster rank in bit 24: clearing it names the same offset in the leader.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
SNN_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
SNN_DEVINL void cluster_sync_all() {   // every thread of every CTA in the cluster
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// arrive on an mbarrier that may live in the peer CTA (shared::cluster address)
SNN_DEVINL void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in the executing CTA, completion bytes go to the LEADER's mbarrier
SNN_DEVINL void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)m), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
SNN_DEVINL void tma_load_5d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"((uint64_t)m), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
template <int kCols>
SNN_DEVINL void tmem_alloc_pair(uint32_t dst_smem) {  // one whole warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
SNN_DEVINL void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the pair: M = 256 (128 rows per CTA), B = the two CTAs' halves
SNN_DEVINL void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all prior MMAs of this thread completed) on the mbarrier at this offset in BOTH CTAs of the pair
SNN_DEVINL void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// ------------------------------------------------------------------------------------------
// TMA tensor STORES (shared -> global) of a swizzled staging tile, optionally as a reduction (+=)
// ------------------------------------------------------------------------------------------
SNN_DEVINL void tma_store_5d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"((uint64_t)m), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
SNN_DEVINL void tma_reduce_add_5d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"((uint64_t)m), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
SNN_DEVINL void tma_reduce_add_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"((uint64_t)m), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
SNN_DEVINL void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
SNN_DEVINL void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
SNN_DEVINL void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
SNN_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SNN_DEVINL void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------------------------------
// Dynamic tile scheduler of the persistent GEMM kernels.
// Round 1 walked tiles statically (tile = worker, worker + nworkers, ...): with gradient all-reduces running beside backward,
// the SMs that also host NCCL CTAs run their tiles slower and every launch waits for them (conv dgrad 896 -> 810 TF/s from 1
// to 8 GPUs in SCALE_r01).  Now a worker's first tile is static and every further one comes from a global atomic counter;
// the leader CTA's producer warp fetches it and hands it to the other warps (and to the peer CTA of a pair, through
// distributed shared memory) over an 8-slot ring guarded by mbarriers, TWO tiles ahead of its own loads (the peer's
// producer must never wait for an id at a tile boundary: published one tile ahead only, the 128-channel layers lost 25 %).
// The last worker to finish zeroes the counters, so a launch always finds them at 0 (kernels that share a counter pair are
// stream-ordered).  sched == nullptr selects the static walk (single-GPU default: nothing to steal from, 3-8 % faster).
// ------------------------------------------------------------------------------------------
constexpr int kSchedSlots = 8;
struct __align__(8) SchedRing {
    uint64_t full[kSchedSlots];
    uint64_t empty[kSchedSlots];
    int tile[kSchedSlots];
};
SNN_DEVINL uint32_t mapa_u32(uint32_t addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
SNN_DEVINL void st_shared_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
SNN_DEVINL void mbar_wait_cluster(uint32_t bar, uint32_t parity) {      // acquire at cluster scope (data written by the peer CTA)
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1u << 26)) __trap();
    }
}
SNN_DEVINL void sched_init(SchedRing& r, uint32_t consumers) {           // one thread, before the block/cluster barrier
    for (int s = 0; s < kSchedSlots; ++s) {
        mbar_init(smem_u32(&r.full[s]), 1);
        mbar_init(smem_u32(&r.empty[s]), consumers);
    }
}
// scheduler side: the leader CTA's producer warp (all 32 lanes call; `lead` = the elected lane)
template <bool PAIR>
SNN_DEVINL void sched_publish(SchedRing& r, int it, int tile, bool lead) {
    const int s = it & (kSchedSlots - 1);
    mbar_wait(smem_u32(&r.empty[s]), (uint32_t)(((it / kSchedSlots) & 1) ^ 1));      // every reader of the slot's previous tile has it
    if (lead) {
        *reinterpret_cast<volatile int*>(&r.tile[s]) = tile;
        if (PAIR) st_shared_cluster_u32(mapa_u32(smem_u32(&r.tile[s]), 1u), (uint32_t)tile);
        mbar_arrive(smem_u32(&r.full[s]));
        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&r.full[s]), 1u));
    }
    __syncwarp();
}
// reader side: any other role warp of either CTA (all 32 lanes call); returns the tile with sequence number `it`, -1 = done
template <bool PAIR>
SNN_DEVINL int sched_next(SchedRing& r, int it, int lane) {
    const int s = it & (kSchedSlots - 1);
    if (PAIR) mbar_wait_cluster(smem_u32(&r.full[s]), (uint32_t)((it / kSchedSlots) & 1));
    else mbar_wait(smem_u32(&r.full[s]), (uint32_t)((it / kSchedSlots) & 1));
    int t = *reinterpret_cast<volatile int*>(&r.tile[s]);
    t = __shfl_sync(0xffffffffu, t, 0);        // provably warp-uniform: the role loops keep their state in uniform registers
    if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&r.empty[s]), 0u));           // the LEADER's ring collects all readers
        else mbar_arrive(smem_u32(&r.empty[s]));
    }
    return t;
}
// end of kernel, one thread per worker: the last worker to finish leaves both counters at zero for the next launch
SNN_DEVINL void sched_finish(unsigned int* sched, int nworkers) {
    __threadfence();
    const unsigned int old = atomicAdd(sched + 1, 1u);
    if (old == (unsigned)(nworkers - 1)) {
        sched[0] = 0u;
        sched[1] = 0u;
        __threadfence();
    }
}

SNN_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------
SNN_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
SNN_DEVINL float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
SNN_DEVINL float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

}  // namespace snn
