// ConvBR as a slab-streaming implicit GEMM on the Blackwell tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA) - models/operations_3d.py:31-47 for k in {1,3}, stride 1.
//
// GEMM view:  Out[voxel, co] = sum_{tap, ci} In[voxel + tap, ci] * W[tap, ci, co]
//   M = 128 output voxels (8 along w x 16 along h of one depth slice) = the 128 TMEM lanes,
//   N = c_out (padded to 16/32/64) x number of weight planes, K = 16 input channels per tcgen05.mma.
//
// Operand layout.  Activations live in HBM as bf16 "planes" (hi, lo[, lo2]) blocked by 8 channels
// (include/leastereo_b200.h), so a (voxel, 8 channels, plane) group is exactly one 16-byte row of a UMMA
// no-swizzle K-major core matrix.  One TMA box load per pipeline stage brings a HALO slab
// (18 h x 10 w voxels x 16 channels x P planes of ONE input depth) into shared memory as [block][h][w][8];
// every one of the 9 (kh,kw) taps - and all three kd taps, which feed three different output depths - is then
// just a different START ADDRESS of the same staged slab (rows are 16 B apart, 8-row groups 160 B apart), so a
// loaded byte is reused 27x from shared memory and each input slab is fetched once per CTA sweep.
//
// Split precision ("bf16xN").  With A = a0+a1(+a2), W = w0+w1(+w2) (bf16 planes) the triangular product
// sum_{i+j<P} a_i*w_j is issued term by term, every term accumulating into the same fp32 TMEM accumulator.
// P=2 gives the 3-term bf16x3 product, P=3 the 6-term one (~fp32).
//
// kd batching.  The three kd taps of one (kh,kw) read the SAME staged window of input slab s and feed the three
// output depths s+1, s, s-1.  Accumulators of a region are laid out in TMEM in DESCENDING depth order, so those
// three targets are adjacent columns and ONE tcgen05.mma with N = 3*N_out against the weight tile
// [W(kd=0); W(kd=1); W(kd=2)] does all three: the 4 KB activation window - the operand that bounds this kernel,
// shared-memory bandwidth being the limit at small N - is fetched once instead of three times.
//
// Schedule.  Persistent CTAs (one per SM), 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2-5 = epilogue.  A work item is (batch, depth chunk of Dc slices, h tile, w tile).  For every group of 16
// input channels the producer streams that group's 27-tap weight part (cp.async.bulk) and then the Dc+2 input
// slabs; the issuer accumulates into Dc resident TMEM accumulators.  Two accumulator sets ping-pong so the
// epilogue of item i (TMEM -> registers -> BN/ReLU/+res -> plane split -> 16-byte stores) overlaps item i+1.
#include "lea_common.h"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define LEA_TC_TW 8      // output tile: 8 voxels along w ...
#define LEA_TC_TH 16     // ... x 16 along h = 128 GEMM rows = 128 TMEM lanes

namespace {

// Epilogue warps: one per TMEM lane quarter.  Measured: two per quarter (alternate depth batches) speed up only the
// 8-channel layers (155 -> 139 us) and lose overall (174 -> 170 pairs/s): 10 warps put three warps on one SM
// sub-partition (16 K registers), capping the kernel at 168 registers per thread.
// The 8-output-channel instances (E8 == 1, 2 planes: cell 10's single ops with residual, epilogue-bound) run two warps per
// quarter; their epilogue fits the 168-register cap of 10 warps.
__host__ __device__ constexpr int epi_warps(int planes, int e8) { return (planes == 2 && e8 == 1) ? 8 : 4; }
__host__ __device__ constexpr int tc_threads(int planes, int e8) { return 64 + 32 * epi_warps(planes, e8); }
constexpr int kMaxTerms = 6;
constexpr int kMaxSets = 8;      // TMEM accumulator sets (work items in flight between the issuer and the epilogue)
constexpr int kMaxStages = 24;   // deep enough that 8 KB 1x1x1 stages keep ~1.5 us of HBM latency covered
constexpr int kSmemBudget = 227 * 1024;
constexpr int kHeaderBytes = 2048;
static_assert(8 * (2 * kMaxStages + 4 + 2 * kMaxSets + 6) <= 1024, "barriers must fit below the BN vectors at byte 1024");
#ifdef LEA_TC_SOFT_TIMEOUT
constexpr unsigned long long kWaitTimeoutCycles = 100000000ull;
#else
constexpr unsigned long long kWaitTimeoutCycles = 4000000000ull;
#endif    // ~2 s: a stuck pipeline traps instead of hanging

// Timing-ablation switches of the chunked kernel (lea_tc_opts.debug) exist only in builds with -DLEA_TC_ABLATION: even as
// never-taken branches they cost the epilogue-bound 8-channel convs 8 % (629 -> 684 us, measured).
#ifdef LEA_TC_ABLATION
#define TC_DBG(p, bit) (((p).dbg & (bit)) != 0)
#else
#define TC_DBG(p, bit) false
#endif

struct TcParams {
    int B, D, H, W;
    int g0_stride_b;          // tensor-map dim-4 blocks per batch element = src channel blocks * P
    int g0_first;             // first dim-4 block of the channel slice = src_cb0 * P
    int P;
    int ncg, blocks_per_cg;
    int ks, taps;
    int tiles_w, tiles_h, dchunks, Dc, total_items;
    int flat, PD;             // flat = depth-1 volume (2-D conv): the 8-wide w tiles of a row of tiles take the place of the
                              // depth slices of an item (PD of them), so that one accumulator hand-over serves Dc tiles
    int step_tw, step_th, step_dc, step_b;   // mixed-radix digits of the grid stride (ItemCursor)
    int NP, c_out, ngroups;   // padded N, real c_out, TMEM regions (weight planes) summed by the epilogue
    int nterm;
    int term_aoff[kMaxTerms], term_lbo_blocks[kMaxTerms], term_btile[kMaxTerms], term_region[kMaxTerms];
    int term_first[kMaxTerms];   // 1 = first term written into its region (carries the zero-init)
    int term_skip[kMaxTerms];    // the MMA writes from this accumulator column of the slab's range on (N shrinks by as much)
    int nbt, nb_rows, btile_bytes, wpart_bytes;
    int slab_vox, pitch_vox, blk_bytes, stage_bytes, stage_stride;   // stride = bytes rounded up to 128 (TMA alignment)
    int nstages, nwbuf;
    int wsplit;               // single streamed weight buffer handed over in three parts (one per kh), see the producer
    int tw_log2;              // tile = (1 << tw_log2) voxels along w x (128 >> tw_log2) along h  (3 for k = 3)
    int cv_skip;              // collapsed stem0: skip the voxels lea_stem0_assemble writes (lea_cv_interior)
    int dbg;                  // development switches (bit 0: epilogue skips its stores, bit 1: skips the TMEM loads, bit 2: skips the
                              // residual reads, bit 3: the issuer issues no MMAs) - timing ablations only, results are wrong
    int wres;                 // 1 = the weight parts of ALL channel groups stay resident in shared memory (loaded once per CTA)
    int fused_cv, ncg_half;   // fused cost volume: channel groups [0,ncg_half) come from x, the rest from y(w-d)
    const CUtensorMap* cvmaps; // [2*D]: x maps for d = 0..D-1, then y maps
    int fold;                 // folded terms: NP = 2 x padded c_out "virtual" channels [main | correction] (see tc_shape)
    int nsets;                // TMEM accumulator sets: 2 = epilogue of item i overlaps MMAs of item i+1, 1 = larger Dc
    int swap_lbo_sbo;         // debug switch for the descriptor convention
    const uint8_t* wimg;
    const float* bn_scale; const float* bn_shift; int relu;
    lea_vol dst; int dst_c0; lea_vol res; int res_c0; int has_res; float* dst_f32;
};

__device__ int g_lea_tc_status;
#ifdef LEA_TC_ABLATION
// development profile of CTA 0 (ablation builds only): per role [loop cycles, cycles inside barrier waits, items]
__device__ long long g_lea_tc_prof[12];
#define TC_PROF_DECL long long prof_t0 = clock64(), prof_wait = 0, prof_items = 0
#define TC_PROF_WAIT(stmt) { const long long w0_ = clock64(); stmt; prof_wait += clock64() - w0_; }
#define TC_PROF_ITEM ++prof_items
#define TC_PROF_END(role) if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { g_lea_tc_prof[(role) * 3] = clock64() - prof_t0; \
        g_lea_tc_prof[(role) * 3 + 1] = prof_wait; g_lea_tc_prof[(role) * 3 + 2] = prof_items; }
#else
#define TC_PROF_DECL
#define TC_PROF_WAIT(stmt) stmt
#define TC_PROF_ITEM
#define TC_PROF_END(role)
#endif

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking probe (mbarrier.test_wait never suspends the thread)
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code) {
#ifdef LEA_TC_SPIN
    if (mbar_test(bar, parity)) return;
    const unsigned long long t0 = clock64();
    while (!mbar_test(bar, parity)) {
#else
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
#endif
        if (clock64() - t0 > kWaitTimeoutCycles) {
#ifdef LEA_TC_SOFT_TIMEOUT
            atomicCAS(&g_lea_tc_status, 0, code * 1000 + (int)(threadIdx.x));   // debug build: record the first stuck wait, go on
            return;
#else
            atomicExch(&g_lea_tc_status, code);
            __trap();
#endif
        }
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// UMMA shared-memory descriptor, no swizzle, K-major: rows 16 B apart inside an 8-row core matrix,
// LBO = byte distance between the two 8-element K halves, SBO = byte distance between 8-row groups.
// (the kernels assemble the two 32-bit halves of this descriptor directly: address >> 4 | LBO >> 4 << 16, SBO >> 4 | 1 << 14)
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n
__device__ __host__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct ItemGeom { int b, d0, d_hi, h0, w0, dlo, dhi; };
// Work items are numbered (b, depth chunk, h tile, w tile) and a CTA takes every gridDim.x-th one.  The mixed-radix
// digits of the current item are carried along and advanced by the digits of the stride (computed by the host) instead
// of being re-derived with three integer divisions per item and role: with depth-1 volumes (the 2-D feature net) an item
// is a single slab and those dependent divisions were a large part of the ~1400 cycles every item cost each role.
struct ItemCursor {
    int tw, th, dc, b;
    __device__ __forceinline__ void init(const TcParams& p, int item) {
        int r = item;
        tw = r % p.tiles_w; r /= p.tiles_w;
        th = r % p.tiles_h; r /= p.tiles_h;
        dc = r % p.dchunks; r /= p.dchunks;
        b = r;
    }
    __device__ __forceinline__ void next(const TcParams& p) {
        tw += p.step_tw;
        int carry = tw >= p.tiles_w ? 1 : 0;
        tw -= carry ? p.tiles_w : 0;
        th += p.step_th + carry;
        carry = th >= p.tiles_h ? 1 : 0;
        th -= carry ? p.tiles_h : 0;
        dc += p.step_dc + carry;
        carry = dc >= p.dchunks ? 1 : 0;
        dc -= carry ? p.dchunks : 0;
        b += p.step_b + carry;
    }
};
__device__ __forceinline__ ItemGeom decode_item(const TcParams& p, const ItemCursor& c) {
    ItemGeom g;
    const int tw = c.tw, th = c.th, dc = c.dc;
    g.b = c.b;
    g.d0 = dc * p.Dc;
    g.d_hi = min(g.d0 + p.Dc, p.flat ? p.PD : p.D);
    g.h0 = th * (128 >> p.tw_log2); g.w0 = tw << p.tw_log2;
    if (p.cv_skip) {
        // collapsed stem0: keep only the depths of the chunk that lea_stem0_assemble does not write.  For one w tile
        // they are contiguous (depth 0 | the band 8 tw - 1 .. 8 tw + 9 | depth D-1), so trimming both ends is exact;
        // an edge chunk then streams 2 slabs instead of Dc + 2.  Empty range = the item is skipped by every role.
        while (g.d0 < g.d_hi && lea_cv_collapsed(g.d0, tw, p.D, p.W)) ++g.d0;
        while (g.d_hi > g.d0 && lea_cv_collapsed(g.d_hi - 1, tw, p.D, p.W)) --g.d_hi;
    }
    if (p.ks == 3 && !p.flat) { g.dlo = max(g.d0 - 1, 0); g.dhi = min(g.d_hi, p.D - 1); }
    else           { g.dlo = g.d0;             g.dhi = g.d_hi - 1; }
    return g;
}

// collapsed stem0: true when every voxel of the item is written by lea_stem0_assemble (the item is skipped by all roles)
__device__ __forceinline__ bool item_skipped(const TcParams& p, const ItemGeom& g) {
    return p.cv_skip && g.d0 >= g.d_hi;
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred) : "r"(0xffffffffu));
    return pred;
}
// issue one MMA from the elected lane; descriptors are (lo, hi) 32-bit halves
__device__ __forceinline__ void tc_mma_issue(uint32_t elected, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                             uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.ne.b32 q, %0, 0;\n\t"
        "mov.b64 da, {%2, %3};\n\t"
        "mov.b64 db, {%4, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %6, p;\n\t}"
        ::"r"(elected), "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_if(uint32_t elected, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t}"
        ::"r"(elected), "r"(bar) : "memory");
}
// ---- epilogue helpers with the plane count known at compile time -------------------------------------------
template <int PL>
__device__ __forceinline__ void ep_load_raw8(const lea_vol& v, int b, int cb, int d, int h, int w, uint4* q) {
    const uint4* base = reinterpret_cast<const uint4*>(v.data);
    const int64_t g = lea_vol_group(v, b, cb, 0, d, h, w);
    const int64_t ps = lea_vol_plane_stride(v);
#pragma unroll
    for (int pl = 0; pl < PL; ++pl) q[pl] = __ldg(base + g + pl * ps);
}
template <int PL>
__device__ __forceinline__ void ep_add_raw8(const uint4* q, float* f) {
#pragma unroll
    for (int pl = 0; pl < PL; ++pl) {
        const uint32_t w4[4] = {q[pl].x, q[pl].y, q[pl].z, q[pl].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] += __uint_as_float(w4[i] << 16);
            f[2 * i + 1] += __uint_as_float(w4[i] & 0xffff0000u);
        }
    }
}
// two fp32 -> packed bf16x2 (round-to-nearest-even), low half = first argument
__device__ __forceinline__ uint32_t ep_pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// split 8 fp32 values into PL bf16 planes (same bits as lea_split_planes for finite inputs) and store them
template <int PL>
__device__ __forceinline__ void ep_store8(const lea_vol& v, int b, int cb, int d, int h, int w, const float* f) {
    uint4* base = reinterpret_cast<uint4*>(v.data);
    const int64_t g = lea_vol_group(v, b, cb, 0, d, h, w);
    const int64_t ps = lea_vol_plane_stride(v);
    uint32_t q[PL][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float x0 = f[2 * i], x1 = f[2 * i + 1];
#pragma unroll
        for (int pl = 0; pl < PL; ++pl) {
            const uint32_t hq = ep_pack_bf16x2(x0, x1);
            q[pl][i] = hq;
            if (pl + 1 < PL) {
                x0 -= __uint_as_float(hq << 16);
                x1 -= __uint_as_float(hq & 0xffff0000u);
            }
        }
    }
#pragma unroll
    for (int pl = 0; pl < PL; ++pl) base[g + pl * ps] = make_uint4(q[pl][0], q[pl][1], q[pl][2], q[pl][3]);
}

// raw TMEM loads: 16 consecutive fp32 columns of this thread's lane, no wait
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_touch8(uint32_t* r) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// orders later uses of r[] after the preceding tcgen05.wait::ld (the registers are "rewritten" by an empty asm)
__device__ __forceinline__ void tc_touch16(uint32_t* r) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}

// Epilogue of one depth batch [j0, j0 + JB) of an item for the channel group at c16: NV = 8 or 16 accumulator columns
// per region, SB depths per tcgen05.wait::ld.  rq = residual groups loaded by the caller (valid when p.has_res).
// FLAT: the "depth" index of the item is a w tile of a depth-1 volume: voxel (0, h, tile_w * d + lw).
template <int PL, int NV, int SB, int JB, bool FLAT>
__device__ __forceinline__ void ep_depth_batch(const TcParams& p, const ItemGeom& g, uint32_t tcol, uint32_t r1off,
                                               bool two_regions, int j0, int nd, int c16, bool valid, int h, int w_in,
                                               int lw, int64_t sp, const float* s_scale, const float* s_shift,
                                               const uint4 (*rq)[2][PL]) {
#pragma unroll
    for (int jb = 0; jb < JB; jb += SB) {
        if (j0 + jb >= nd) break;
        uint32_t ra[SB][NV], rb[SB][NV];
#pragma unroll
        for (int s = 0; s < SB; ++s) {
            const int j = j0 + jb + s;
            if (j < nd) {
                const uint32_t col = tcol + (uint32_t)((nd - 1 - j) * p.NP);
                if (NV == 8) { tc_ld8_nowait(col, ra[s]); if (two_regions) tc_ld8_nowait(col + r1off, rb[s]); }
                else         { tc_ld16_nowait(col, ra[s]); if (two_regions) tc_ld16_nowait(col + r1off, rb[s]); }
            }
        }
        tc_wait_ld();
#pragma unroll
        for (int s = 0; s < SB; ++s) {
            const int jj = jb + s;
            if (j0 + jj >= nd) break;
            const int dj = g.d0 + j0 + jj;
            const int d = FLAT ? 0 : dj;
            const int w = FLAT ? (dj << p.tw_log2) + lw : w_in;
            float acc[NV];
            if (NV == 8) tc_touch8(ra[s]); else tc_touch16(ra[s]);
            if (two_regions) {
                if (NV == 8) tc_touch8(rb[s]); else tc_touch16(rb[s]);
#pragma unroll
                for (int i = 0; i < NV; ++i) acc[i] = __uint_as_float(ra[s][i]) + __uint_as_float(rb[s][i]);
            } else {
#pragma unroll
                for (int i = 0; i < NV; ++i) acc[i] = __uint_as_float(ra[s][i]);
            }
            if (!valid || (FLAT && w >= p.W) || (p.cv_skip && lea_cv_collapsed(d, g.w0 >> 3, p.D, p.W))) continue;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                float v = acc[i] * s_scale[c16 + i] + s_shift[c16 + i];
                if (p.relu) v = fmaxf(v, 0.0f);
                acc[i] = v;
            }
            if (p.dst_f32) {
                float* o = p.dst_f32 + (int64_t)g.b * p.c_out * sp + ((int64_t)d * p.H + h) * p.W + w;
                for (int n = 0; n < NV && c16 + n < p.c_out; ++n) o[(c16 + n) * sp] = acc[n];
            } else {
                if (p.has_res) {
                    ep_add_raw8<PL>(rq[jj][0], acc);
                    if (NV == 16) ep_add_raw8<PL>(rq[jj][1], acc + 8);
                }
                ep_store8<PL>(p.dst, g.b, (p.dst_c0 + c16) >> 3, d, h, w, acc);
                if (NV == 16) ep_store8<PL>(p.dst, g.b, ((p.dst_c0 + c16) >> 3) + 1, d, h, w, acc + 8);
            }
        }
    }
}

// what the MMA issuer needs for one activation slab (computed one slab ahead)
template <int NTERM>
struct SlabPar {
    uint32_t idesc_t[NTERM], a_base[NTERM], b_base[NTERM], d_base[NTERM];
    uint32_t idesc_fresh, idesc_rest, fresh_cols;
    int nfresh, nrest;
};

template <int KS, int NTERM, int PL, int E8>
__global__ void __launch_bounds__(tc_threads(PL, E8), 1)
lea_conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // header: barriers, TMEM base, BN scale/shift
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* full = bars;                          // [kMaxStages]
    uint64_t* empty = bars + kMaxStages;            // [kMaxStages]
    uint64_t* wfull = bars + 2 * kMaxStages;        // [2]
    uint64_t* wempty = wfull + 2;                   // [2]
    uint64_t* accfull = wempty + 2;                 // [kMaxSets]
    uint64_t* accempty = accfull + kMaxSets;        // [kMaxSets]
    uint64_t* w3full = accempty + kMaxSets;         // [3]  weight parts (kh = 0, 1, 2) of the single streamed buffer (p.wsplit)
    uint64_t* w3empty = w3full + 3;                 // [3]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 1536);
    float* s_scale = reinterpret_cast<float*>(smem + 1024);     // [64]  (barriers occupy the first KB)
    float* s_shift = s_scale + 64;                              // [64]
    uint8_t* wbuf = smem + kHeaderBytes;
    const int wbuf_stride = (p.wpart_bytes + 127) & ~127;
    uint8_t* stages = wbuf + (size_t)p.nwbuf * wbuf_stride;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // Programmatic dependent launch: the next launch of the stream may be scheduled as soon as every CTA of this grid
    // has started (its CTAs take an SM when one of ours exits - one CTA fits per SM - and run their prologue there);
    // no global memory is touched before griddepcontrol.wait, which returns once the PREVIOUS grid has completed and
    // flushed, so the launch chain keeps stream order for every buffer, including the re-used arena.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nstages; ++i) { mbar_init(smem_u32(full + i), 1); mbar_init(smem_u32(empty + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(wfull + i), 1); mbar_init(smem_u32(wempty + i), 1); }
        for (int i = 0; i < 3; ++i) { mbar_init(smem_u32(w3full + i), 1); mbar_init(smem_u32(w3empty + i), 1); }
        for (int i = 0; i < kMaxSets; ++i) {
            mbar_init(smem_u32(accfull + i), 1); mbar_init(smem_u32(accempty + i), 32 * epi_warps(PL, E8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
        const int n = threadIdx.x - 64;
        s_scale[n] = (p.bn_scale && n < p.c_out) ? __ldg(p.bn_scale + n) : 1.0f;
        s_shift[n] = (p.bn_shift && n < p.c_out) ? __ldg(p.bn_shift + n) : 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // KS: 1 = 1x1x1, 3 = 3x3x3, 5 = 3x3x3 with the single streamed weight buffer handed over in three parts (p.wsplit; a
    // variant of its own because even never-taken hooks in the tap loops cost every 3x3x3 kernel 2-4 % of its issuer cycles);
    // on depth-1 volumes ("flat": w tiles as the depth slices of an item) 2 = 3x3, 4 = 1x1
    constexpr int KT = (KS == 1 || KS == 4) ? 1 : 3;           // taps per in-plane axis
    constexpr bool FLAT = (KS == 2 || KS == 4);
    constexpr bool K3 = (KS == 3 || KS == 5);                  // three tap planes along depth, halo slabs
    constexpr bool WSPLIT = (KS == 5);
    constexpr int kHalo = (KT == 3) ? 1 : 0;
    constexpr int kPitch = LEA_TC_TW + 2 * kHalo;              // voxels per staged row

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0, sphase = 0, wb = 0, wphase = 0, wuse = 0;
            TC_PROF_DECL;
            if (p.wres) {                         // all weight parts fit: load them once, they stay for every item
                mbar_arrive_expect_tx(smem_u32(wfull), (uint32_t)(p.ncg * p.wpart_bytes));
                for (int cg = 0; cg < p.ncg; ++cg)
                    bulk_load(smem_u32(wbuf + (size_t)cg * wbuf_stride), p.wimg + (size_t)cg * p.wpart_bytes,
                              (uint32_t)p.wpart_bytes, smem_u32(wfull));
            }
            ItemCursor cur;
            cur.init(p, blockIdx.x);
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, cur.next(p)) {
                const ItemGeom g = decode_item(p, cur);
                if (item_skipped(p, g)) continue;
                const int gbase = g.b * p.g0_stride_b + p.g0_first;
                for (int cg = 0; cg < p.ncg; ++cg) {
                    if (!p.wres && !p.wsplit) {
                        TC_PROF_WAIT(mbar_wait(smem_u32(wempty + wb), wphase ^ 1, 101));
                        mbar_arrive_expect_tx(smem_u32(wfull + wb), (uint32_t)p.wpart_bytes);
                        bulk_load(smem_u32(wbuf + (size_t)wb * wbuf_stride), p.wimg + (size_t)cg * p.wpart_bytes,
                                  (uint32_t)p.wpart_bytes, smem_u32(wfull + wb));
                        if (++wb == p.nwbuf) { wb = 0; wphase ^= 1; }
                    }
                    for (int d_in = g.dlo; d_in <= g.dhi; ++d_in) {
                        TC_PROF_WAIT(mbar_wait(smem_u32(empty + stage), sphase ^ 1, 102));
                        mbar_arrive_expect_tx(smem_u32(full + stage), (uint32_t)p.stage_bytes);
                        if (p.fused_cv) {
                            // cost volume built by the loader: disparity d_in selects the tensor map; the map's
                            // origin/width make TMA's zero fill reproduce [w >= d] (see lea_build_fused_cv_maps)
                            const bool left = cg < p.ncg_half;
                            const CUtensorMap* m = p.cvmaps + (left ? 0 : p.D) + d_in;
                            tma_load_3d(smem_u32(stages + (size_t)stage * p.stage_stride), m, smem_u32(full + stage),
                                        (g.w0 - kHalo - d_in) * 8, g.h0 - kHalo,
                                        gbase + (left ? cg : cg - p.ncg_half) * p.blocks_per_cg);
                        } else {
                            tma_load_4d(smem_u32(stages + (size_t)stage * p.stage_stride), &tmap, smem_u32(full + stage),
                                        ((FLAT ? (d_in << p.tw_log2) : g.w0) - kHalo) * 8, g.h0 - kHalo, FLAT ? 0 : d_in,
                                        gbase + cg * p.blocks_per_cg);
                        }
                        if (++stage == p.nstages) { stage = 0; sphase ^= 1; }
                    }
                    if (p.wsplit) {
                        // conv1 / conv2: the 111 KB weight part of a channel group leaves room for ONE weight buffer, and
                        // waiting for the issuer to finish a group before its successor's weights are even requested cost
                        // ~1900 idle tensor-pipe cycles per group (measured: issuer 22 % in waits).  The buffer is handed
                        // over in three parts, one per kh: a part is reloaded as soon as the issuer has run the last slab's
                        // taps of that kh, while the taps of the later kh still execute.  The group's activation slabs are
                        // requested BEFORE its weights so that they do not queue behind the hand-over.
                        const uint32_t part = (uint32_t)p.wpart_bytes / 3u;
#pragma unroll 1
                        for (int kh = 0; kh < 3; ++kh) {
                            TC_PROF_WAIT(mbar_wait(smem_u32(w3empty + kh), (uint32_t)((wuse & 1) ^ 1), 103));
                            mbar_arrive_expect_tx(smem_u32(w3full + kh), part);
                            bulk_load(smem_u32(wbuf) + (uint32_t)kh * part, p.wimg + (size_t)cg * p.wpart_bytes + (size_t)kh * part,
                                      part, smem_u32(w3full + kh));
                        }
                        ++wuse;
                    }
                }
                TC_PROF_ITEM;
            }
            TC_PROF_END(0);
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the whole warp walks the loop (warp-uniform descriptor arithmetic),
        //                   one elected lane issues tcgen05.mma / tcgen05.commit =================
        const uint32_t elected = elect_one();
        uint32_t a_term16[NTERM], a_lbo_field[NTERM], b_term16[NTERM], reg_col[NTERM];
        bool t_first[NTERM];
        uint32_t t_skip_idesc[NTERM];
#pragma unroll
        for (int t = 0; t < NTERM; ++t) {
            t_skip_idesc[t] = ((uint32_t)p.term_skip[t] >> 3) << 17;         // make_idesc is linear in N
            a_term16[t] = (uint32_t)(p.term_aoff[t] * p.blk_bytes) >> 4;
            a_lbo_field[t] = ((uint32_t)(p.term_lbo_blocks[t] * p.blk_bytes) >> 4) << 16;
            b_term16[t] = (uint32_t)(p.term_btile[t] * p.btile_bytes) >> 4;
            reg_col[t] = (uint32_t)(p.term_region[t] * p.Dc * p.NP) + (uint32_t)p.term_skip[t];    // (a term's skipped columns)
            t_first[t] = p.term_first[t] != 0;
        }
        const int NPv = p.NP;
        const uint32_t idesc0 = make_idesc(0), idesc_step = make_idesc(NPv) - idesc0;
        const uint32_t a_hi = (uint32_t)kPitch | (1u << 14);                      // SBO = kPitch*16 B, version 1
        const uint32_t b_hi = 8u | (1u << 14);                                    // SBO = 128 B
        const uint32_t b_lbo_field = (uint32_t)p.nb_rows << 16;                   // LBO = nb_rows*16 B
        const uint32_t tap16 = (uint32_t)(p.nbt * p.btile_bytes) >> 4;            // weight bytes per (kh,kw) / 16
        const uint32_t set_cols = (uint32_t)(p.ngroups * p.Dc * p.NP);
        const uint32_t stages16 = smem_u32(stages) >> 4, stride16 = (uint32_t)p.stage_stride >> 4;   // (128 B multiples)
        int stage = 0, sphase = 0, wb = 0, wphase = 0, set_run = 0, aphase_run = 0, wuse = 0;
        uint32_t probed = 0;     // the NEXT stage's full barrier, tested while this stage's MMAs issue (tcgen05.mma issue
                                 // is synchronous with the pipe: a barrier round trip between slabs is a tensor-pipe bubble)
        TC_PROF_DECL;
        auto slab_par = [&](int d_in, int cg, int stage_v, uint32_t w16, uint32_t set_base, const ItemGeom& g) {
            SlabPar<NTERM> r;
            const uint32_t s16 = stages16 + (uint32_t)stage_v * stride16;
            // valid kd range of this slab: output depth d = d_in + 1 - kd must lie in [d0, d_hi):
            //   kd_a = 0 / 1 / 2 for d_in <= d_hi-2 / = d_hi-1 / = d_hi;   kd_b = 2 / 1 / 0 for d_in >= d0+1 / = d0 / = d0-1.
            // Branch-free on purpose (as ternaries it compiled to a chain of ~10 uniform branches).
            int kd_a = (KS == 2) ? 1 : 0, kd_b = kd_a;        // flat 3x3: the middle tap plane only, one "depth" per slab
            if (K3) {
                // (measured: sums of compares instead of these clamps cost the issuer 4-7 % more cycles per item)
                kd_a = min(max(d_in - (g.d_hi - 2), 0), 2);
                kd_b = min(max(d_in - g.d0 + 1, 0), 2);
            }
            const int nkd = kd_b - kd_a + 1;
            // accumulators are stored in descending depth order: depth d sits at column (d_hi-1-d)*NP
            const int d_top = K3 ? d_in + 1 - kd_a : d_in;
            const uint32_t col0 = (uint32_t)((g.d_hi - 1 - d_top) * NPv);
            // depths touched for the first time by this slab (only while the first channel group runs):
            // kd = 0 always opens depth d_in+1; at d_in == 0 depth 0 (kd = 1) opens too
            int nfresh = 1;
            if (K3) {
                const int z = (d_in == 0) ? 1 : 0;
                nfresh = ((kd_a == 0) ? 1 : 0) + (z & ((((kd_a == 0) ? 1 : 0) & ((kd_b >= 1) ? 1 : 0)) | ((kd_a == 1) ? 1 : 0)));
            }
            nfresh = (cg == 0) ? nfresh : 0;
            const uint32_t brow16 = (uint32_t)(kd_a * NPv);               // first weight row used, x16 B
            // the instruction descriptor is linear in N: idesc(k * NP) = idesc0 + k * idesc_step
            const uint32_t idesc_all = idesc0 + (uint32_t)nkd * idesc_step;
#pragma unroll
            for (int t = 0; t < NTERM; ++t) {
                r.idesc_t[t] = idesc_all - t_skip_idesc[t];            // per term: N less the columns the term skips
                r.a_base[t] = (s16 + a_term16[t]) | a_lbo_field[t];   // (the tap offset added later cannot carry into the LBO field)
                r.b_base[t] = w16 + b_term16[t] + brow16;
                r.d_base[t] = set_base + reg_col[t] + col0;
            }
            r.nfresh = nfresh;
            r.nrest = nkd - nfresh;
            r.fresh_cols = (uint32_t)(nfresh * NPv);
            r.idesc_fresh = idesc0 + (uint32_t)nfresh * idesc_step;
            r.idesc_rest = idesc0 + (uint32_t)r.nrest * idesc_step;
            return r;
        };
        if (p.wres) mbar_wait(smem_u32(wfull), 0, 202);
        ItemCursor cur;
        cur.init(p, blockIdx.x);
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, cur.next(p)) {
            const ItemGeom g = decode_item(p, cur);
            if (item_skipped(p, g)) continue;
            const int set = set_run, aphase = aphase_run;      // (running counters: it % nsets was an integer division per item)
            if (++set_run == p.nsets) { set_run = 0; aphase_run ^= 1; }
            TC_PROF_WAIT(mbar_wait(smem_u32(accempty + set), aphase ^ 1, 201));
            tc_fence_after();
            const uint32_t set_base = tmem_base + (uint32_t)set * set_cols;
            for (int cg = 0; cg < p.ncg; ++cg) {
                if (!p.wres && !p.wsplit) TC_PROF_WAIT(mbar_wait(smem_u32(wfull + wb), wphase, 202));
                const uint32_t w16 = (smem_u32(wbuf + (size_t)(p.wres ? cg : wb) * wbuf_stride) >> 4) | b_lbo_field;
                // Slab parameters are computed one slab ahead (before the last tap of the previous slab is issued): the
                // bookkeeping then runs while MMAs execute instead of between the last MMA of a slab and the first of the next.
                SlabPar<NTERM> sp = slab_par(g.dlo, cg, stage, w16, set_base, g);
                for (int d_in = g.dlo; d_in <= g.dhi; ++d_in) {
                    if (!probed) TC_PROF_WAIT(mbar_wait(smem_u32(full + stage), sphase, 203));
                    tc_fence_after();
                    SlabPar<NTERM> nx = sp;
                    const bool wrap = (stage + 1 == p.nstages);
                    const int stage_n = wrap ? 0 : stage + 1;
                    if (TC_DBG(p, 8)) {                // development: no MMAs - what the TMA ring and the epilogue cost alone
                        probed = mbar_test(smem_u32(full + stage_n), (uint32_t)(wrap ? sphase ^ 1 : sphase));
                        if (p.wsplit) {
                            for (int kh = 0; kh < 3; ++kh) {
                                if (d_in == g.dlo) mbar_wait(smem_u32(w3full + kh), (uint32_t)(wuse & 1), 204);
                                if (d_in == g.dhi) tc_commit_if(elected, smem_u32(w3empty + kh));
                            }
                        }
                        nx = slab_par(d_in + 1, cg, stage_n, w16, set_base, g);
                    } else
#pragma unroll
                    for (int kh = 0; kh < KT; ++kh) {
                        if (WSPLIT && d_in == g.dlo)                        // this group's weight part for kh has landed
                            TC_PROF_WAIT(mbar_wait(smem_u32(w3full + kh), (uint32_t)(wuse & 1), 204));
#pragma unroll
                        for (int kw = 0; kw < KT; ++kw) {
                            if (KT > 1 && kh == KT - 1 && kw == KT - 1) nx = slab_par(d_in + 1, cg, stage_n, w16, set_base, g);
#pragma unroll
                            for (int t = 0; t < NTERM; ++t) {
                                // per-slab, per-term bases + one add each per MMA: the operands of an MMA must not sit
                                // at the end of a chain of dependent uniform-datapath instructions (each ~10 cycles)
                                const uint32_t a_lo = sp.a_base[t] + (uint32_t)(kh * kPitch + kw);
                                const uint32_t b_lo = sp.b_base[t] + (uint32_t)(kh * KT + kw) * tap16;
                                const uint32_t dcol = sp.d_base[t];
                                if (kh == 0 && kw == 0 && t_first[t] && sp.nfresh > 0) {
                                    tc_mma_issue(elected, dcol, a_lo, a_hi, b_lo, b_hi, sp.idesc_fresh, 0u);
                                    if (sp.nrest > 0)
                                        tc_mma_issue(elected, dcol + sp.fresh_cols, a_lo, a_hi,
                                                     b_lo + sp.fresh_cols, b_hi, sp.idesc_rest, 1u);
                                } else {
                                    tc_mma_issue(elected, dcol, a_lo, a_hi, b_lo, b_hi, sp.idesc_t[t], 1u);
                                }
                            }
                            if (kh == 0 && kw == 0)
                                probed = mbar_test(smem_u32(full + stage_n), (uint32_t)(wrap ? sphase ^ 1 : sphase));
                            if (KT == 1) nx = slab_par(d_in + 1, cg, stage_n, w16, set_base, g);
                        }
                        if (WSPLIT && d_in == g.dhi)                        // last slab of the group: part kh may be reloaded
                            tc_commit_if(elected, smem_u32(w3empty + kh));
                    }
                    tc_commit_if(elected, smem_u32(empty + stage));
                    if (++stage == p.nstages) { stage = 0; sphase ^= 1; }
                    sp = nx;
                }
                if (p.wsplit) ++wuse;
                else if (!p.wres) {
                    tc_commit_if(elected, smem_u32(wempty + wb));
                    if (++wb == p.nwbuf) { wb = 0; wphase ^= 1; }
                }
            }
            tc_commit_if(elected, smem_u32(accfull + set));
            TC_PROF_ITEM;
        }
        TC_PROF_END(1);
    } else {
        // ================= epilogue warps 2..5 =================
        // Thread m owns output voxel m of the tile (TMEM lane m).  Depth slices are processed kJB at a time so that
        // the residual reads of kJB slices are in flight together (the epilogue is otherwise latency-bound).
        // depth slices per batch: 4 where the 8-column path exists, 2 elsewhere (the 16-column batch below holds
        // 2 depths x 2 regions x 16 columns in registers; with 4-deep residual prefetch on top it spills)
        // (E8: 0 = every group is a full 16-channel group, 1 = c_out == 8, 2 = 16-channel groups and an 8-channel tail;
        //  with kJB = 2 everywhere the 8-channel convs lose 21 %, with kJB = 4 and both batch paths the kernel spills)
        constexpr int kJB = (E8 == 1) ? 4 : 2;
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;                 // tile row = TMEM lane
        const int lh = m >> p.tw_log2, lw = m & ((1 << p.tw_log2) - 1);
        const int64_t sp = (int64_t)p.D * p.H * p.W;
        int set_run = 0, aphase_run = 0;
        TC_PROF_DECL;
        ItemCursor cur;
        cur.init(p, blockIdx.x);
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, cur.next(p)) {
            const ItemGeom g = decode_item(p, cur);
            if (item_skipped(p, g)) continue;
            const int set = set_run, aphase = aphase_run;      // (running counters: it % nsets was an integer division per item)
            if (++set_run == p.nsets) { set_run = 0; aphase_run ^= 1; }
            TC_PROF_ITEM;
            const int h = g.h0 + lh, w = g.w0 + lw, w_item = w;
            const bool valid = (h < p.H) && (FLAT || w < p.W);
            const int nd = g.d_hi - g.d0;
            bool waited = false;
            for (int j0 = ((warp - 2) >> 2) * kJB; j0 < nd; j0 += (epi_warps(PL, E8) / 4) * kJB) {
                for (int c16 = 0; c16 < p.c_out; c16 += 16) {
                    const bool two = (c16 + 8 < p.c_out);
                    uint4 rq[kJB][2][PL];
                    if (p.has_res && valid && !TC_DBG(p, 4)) {
#pragma unroll
                        for (int jj = 0; jj < kJB; ++jj) {
                            const int dj = g.d0 + j0 + jj;
                            const int rd = FLAT ? 0 : dj, rw = FLAT ? (dj << p.tw_log2) + lw : w;
                            if (j0 + jj < nd && (!FLAT || rw < p.W)) {
                                ep_load_raw8<PL>(p.res, g.b, (p.res_c0 + c16) >> 3, rd, h, rw, rq[jj][0]);
                                if (two)
                                    ep_load_raw8<PL>(p.res, g.b, ((p.res_c0 + c16) >> 3) + 1, rd, h, rw, rq[jj][1]);
                            }
                        }
                    }
                    if (!waited) {                    // residual loads of the first batch are issued before the wait
                        TC_PROF_WAIT(mbar_wait(smem_u32(accfull + set), aphase, 301));
                        tc_fence_after();
                        waited = true;
                    }
                    if (E8 != 0 && !two) {
                        // The group holds ONE 8-channel block (8-channel convs, the tail of 24): read only its 8 columns,
                        // and the whole depth batch with one tcgen05.wait::ld so that the TMEM latencies overlap.
                        // Measured: the epilogue-bound 8-channel convs 640 -> 456 us.  Compiled only into the instances
                        // that serve c_out % 16 == 8: merely present in the other kernels it costs them 4-5 %.
                        const bool two_regions = (p.ngroups == 2) || p.fold;
                        const uint32_t r1off = (uint32_t)(p.fold ? (p.NP >> 1) : p.Dc * p.NP);
                        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) +
                                              (uint32_t)(set * p.ngroups * p.Dc * p.NP + c16);
                        ep_depth_batch<PL, 8, kJB, kJB, FLAT>(p, g, tcol, r1off, two_regions, j0, nd, c16, valid, h, w, lw, sp,
                                                              s_scale, s_shift, rq);
                        continue;
                    }
                    if (E8 != 1 && two) {
                        // full 16-channel group: 2 depths per tcgen05.wait::ld.  Measured (KITTI, 4 pairs): 16-channel
                        // cell ops 125 -> 106 us, batched ones 240 -> 211, conv1/conv2 -3 %; 199.8 -> 209.2 pairs/s.
                        const bool two_regions = (p.ngroups == 2) || p.fold;
                        const uint32_t r1off = (uint32_t)(p.fold ? (p.NP >> 1) : p.Dc * p.NP);
                        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) +
                                              (uint32_t)(set * p.ngroups * p.Dc * p.NP + c16);
                        ep_depth_batch<PL, 16, 2, kJB, FLAT>(p, g, tcol, r1off, two_regions, j0, nd, c16, valid, h, w, lw, sp,
                                                             s_scale, s_shift, rq);
                        continue;
                    }
#pragma unroll
                    for (int jj = 0; jj < kJB; ++jj) {
                        if (j0 + jj >= nd) break;
                        const int dj = g.d0 + j0 + jj;
                        const int d = FLAT ? 0 : dj;
                        const int w = FLAT ? (dj << p.tw_log2) + lw : w_item;
                        // depth d of region r: column set*ngroups*Dc*NP + r*Dc*NP + (d_hi-1-d)*NP
                        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) +
                                              (uint32_t)(set * p.ngroups * p.Dc * p.NP + (nd - 1 - (j0 + jj)) * p.NP);
                        float acc[16];
                        if (TC_DBG(p, 2)) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
                        } else {
                            uint32_t ra[16], rb[16];
                            tc_ld16_nowait(trow + (uint32_t)c16, ra);
                            const bool two_regions = (p.ngroups == 2) || p.fold;
                            if (two_regions)
                                tc_ld16_nowait(trow + (uint32_t)((p.fold ? (p.NP >> 1) : p.Dc * p.NP) + c16), rb);
                            tc_wait_ld();
                            tc_touch16(ra);
                            if (two_regions) {
                                tc_touch16(rb);
#pragma unroll
                                for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(ra[i]) + __uint_as_float(rb[i]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(ra[i]);
                            }
                        }
                        if (!valid || (FLAT && w >= p.W) || TC_DBG(p, 1) ||
                            (p.cv_skip && lea_cv_collapsed(d, g.w0 >> 3, p.D, p.W))) continue;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float v = acc[i] * s_scale[c16 + i] + s_shift[c16 + i];
                            if (p.relu) v = fmaxf(v, 0.0f);
                            acc[i] = v;
                        }
                        if (p.dst_f32) {
                            float* o = p.dst_f32 + (int64_t)g.b * p.c_out * sp + ((int64_t)d * p.H + h) * p.W + w;
                            for (int n = 0; n < 16 && c16 + n < p.c_out; ++n) o[(c16 + n) * sp] = acc[n];
                        } else {
                            if (p.has_res && !TC_DBG(p, 4)) {
                                ep_add_raw8<PL>(rq[jj][0], acc);
                                if (two) ep_add_raw8<PL>(rq[jj][1], acc + 8);
                            }
                            ep_store8<PL>(p.dst, g.b, (p.dst_c0 + c16) >> 3, d, h, w, acc);
                            if (two) ep_store8<PL>(p.dst, g.b, ((p.dst_c0 + c16) >> 3) + 1, d, h, w, acc + 8);
                        }
                    }
                }
            }
            if (!waited) {        // a warp group without a depth batch in this item still paces itself on the item
                TC_PROF_WAIT(mbar_wait(smem_u32(accfull + set), aphase, 302));
                tc_fence_after();
            }
            tc_fence_before();
            mbar_arrive(smem_u32(accempty + set));
        }
        if (warp == 2) { TC_PROF_END(2); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

typedef void (*TcKernelFn)(const CUtensorMap, const TcParams);
template <int KS, int E8>
static TcKernelFn tc_kernel_for_ks(int nterm, int planes) {
    if (planes == 1) return lea_conv_tc_kernel<KS, 1, 1, E8>;
    if (planes == 2) {
        if (nterm == 1) return lea_conv_tc_kernel<KS, 1, 2, E8>;
        if (nterm == 2) return lea_conv_tc_kernel<KS, 2, 2, E8>;  // 8-channel layout / folded terms
        return lea_conv_tc_kernel<KS, 3, 2, E8>;                   // bf16x3
    }
    if (nterm == 1) return lea_conv_tc_kernel<KS, 1, 3, E8>;
    if (nterm == 2) return lea_conv_tc_kernel<KS, 2, 3, E8>;      // 8-channel layout, 3 planes, folded
    if (nterm == 3) return lea_conv_tc_kernel<KS, 3, 3, E8>;      // 8-channel layout, 3 planes
    return lea_conv_tc_kernel<KS, 6, 3, E8>;                       // bf16x6
}
template <int E8>
static TcKernelFn tc_kernel_for_e8(int ks, int nterm, int planes) {      // ks: 1, 3, or flat (depth-1 volumes): 2 = 3x3, 4 = 1x1
    switch (ks) {
        case 3: return tc_kernel_for_ks<3, E8>(nterm, planes);
        case 2: return tc_kernel_for_ks<2, E8>(nterm, planes);
        case 4: return tc_kernel_for_ks<4, E8>(nterm, planes);
        case 5: return tc_kernel_for_ks<5, E8>(nterm, planes);
        default: return tc_kernel_for_ks<1, E8>(nterm, planes);
    }
}
static TcKernelFn tc_kernel_for(int ks, int nterm, int planes, int e8) {
    if (e8 == 1) return tc_kernel_for_e8<1>(ks, nterm, planes);
    if (e8 == 2) return tc_kernel_for_e8<2>(ks, nterm, planes);
    return tc_kernel_for_e8<0>(ks, nterm, planes);
}

// ---------------------------------------------------------------------------------------------------------
// weight image:  [cg][kh*KS+kw][btile][khalf(2)][row = kd*NP + co][8] bf16        (kd-major rows: one MMA spans 3 kd)
//   standard (c_in % 16 == 0): tile j = weight plane j; row holds plane j of W[co][cg*16 + khalf*8 + 0..7][kd,kh,kw]
//   c_in == 8, P == 2 ("C8"):  the A operand is [hi(8) | lo(8)] of the same 8 channels, so
//       tile 0: khalf0 = Whi, khalf1 = Whi   (hi*Whi + lo*Whi)        tile 1: khalf0 = Wlo, khalf1 = 0   (hi*Wlo)
//   FOLDED (k = 3, P == 2, c_out <= 16):  one tcgen05.mma costs max(64, N/2) cycles, so for N = 3*16 two thirds of
//       the instruction are idle.  Product terms that share their A operand are therefore issued as ONE MMA against a
//       weight tile with 2*16 "virtual" output channels per kd, row = kd*32 + region*16 + co: region 0 accumulates the
//       main term, region 1 the 2^-8-scaled corrections (the two accumulator regions of accum_split, now adjacent
//       columns of the same depth), and the epilogue adds column c and c+16:
//         standard: tile 0 (A = a_hi) = [Whi | Wlo],  tile 1 (A = a_lo) = [0 | Whi]      3 MMAs -> 2 per tap
//         C8:       tile 0 (A = [hi|lo]) = [[Whi;Whi] | [Wlo;0]]                          2 MMAs -> 1 per tap
// ---------------------------------------------------------------------------------------------------------
struct TcShape {
    bool ok; bool c8; bool fold;
    int NP, nb_rows, nbt, btile_bytes, taps2d, ncg, wpart_bytes, ngroups;
};
__host__ __device__ inline TcShape tc_shape(int c_in, int c_out, int ks, int P, int allow_fold) {
    TcShape s{};
    s.ok = false;
    if (!(ks == 1 || ks == 3) || P < 1 || P > 3) return s;
    if (!(c_out == 1 || (c_out % 8 == 0 && c_out >= 8 && c_out <= 64))) return s;
    s.c8 = (c_in == 8);
    if (s.c8 ? (P < 2) : (c_in % 16 != 0 || c_in < 16 || c_in > 1024)) return s;
    s.NP = (c_out + 15) & ~15;                                 // UMMA N granularity at M = 128
    // (k = 1 too: the 1x1x1 convs were issue-bound at 3 tiny MMAs per slab; 8 input channels on 3 planes - the convs of
    //  the 2-D feature net, issue-bound at 27 MMAs per depth-1 item - fold their three tiles into two)
    // 32 output channels (stem1, the level-2 cell ops, stem0's band conv) fold to 64 virtual channels with ONE weight tile
    // [W_hi | W_lo] (a second tile [0 | W_hi] would double the image and the weights would no longer stay resident): the
    // a_lo MMA multiplies the same tile from its first row but writes one block further into the accumulators (term_skip),
    // i.e. a_lo x [W_hi(kd0) W_lo(kd0) W_hi(kd1) W_lo(kd1) W_hi(kd2)] into [corr(kd0) main(kd1) corr(kd1) main(kd2) corr(kd2)]:
    // a_lo*W_hi lands in the correction columns of the right depth; the blocks in between add a_lo*W_lo of the NEIGHBOURING
    // kd tap (<= 2^-18 relative, the size of the lo*lo term every split-precision mode drops) to the next depth's main
    // columns - an approximation (LEA_TC_FOLD=1 switches it off), gated by the parity tests at BASELINE sizes: KITTI
    // calibrated 99.987 % / 0.00571 px -> 99.980 % / 0.00589 px within 0.1 px / mean, stem1 -10 %, level-2 ops -13 %.
    const bool fold32 = (s.NP == 32 && P == 2 && !s.c8 && allow_fold >= 2);
    s.fold = allow_fold && ((s.NP == 16 && (P == 2 || (P == 3 && s.c8))) || fold32);
    if (s.fold) s.NP *= 2;                                     // [main | correction] virtual channels
    s.taps2d = ks * ks;
    s.nb_rows = ks * s.NP;
    if (s.c8) { s.nbt = s.fold ? (P == 3 ? 2 : 1) : P; s.ncg = 1; s.ngroups = 1; }
    else      { s.nbt = fold32 ? 1 : P; s.ncg = c_in / 16; s.ngroups = P; }
    s.btile_bytes = 2 * s.nb_rows * 16;
    s.wpart_bytes = s.taps2d * s.nbt * s.btile_bytes;
    s.ok = true;
    return s;
}

__global__ void lea_pack_weights_tc_kernel(const float* __restrict__ w, lea_u4* __restrict__ img,
                                           int c_in, int c_out, int ks, int P, int total_groups, int allow_fold,
                                           int dgrad) {
    const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= total_groups) return;
    const TcShape s = tc_shape(c_in, c_out, ks, P, allow_fold);
    int r = gidx;
    const int row = r % s.nb_rows; r /= s.nb_rows;
    const int khalf = r % 2; r /= 2;
    const int bt = r % s.nbt; r /= s.nbt;
    const int tap2d = r % s.taps2d; r /= s.taps2d;
    const int cg = r;
    const int kd = row / s.NP;
    int co = row % s.NP, region = 0;
    if (s.fold) { const int half = s.NP >> 1; region = co / half; co -= region * half; }
    const int tap = kd * s.taps2d + tap2d;                      // PyTorch order (kd, kh, kw)
    const int ntaps = s.taps2d * ks;
    uint32_t q[4] = {0, 0, 0, 0};
    int plane, ci0;
    bool zero = false;
    if (s.fold) {
        if (s.c8 && P == 3) {
            // tile 0 (A = [a0|a1]) = [[w0;w0] | [w1;w1]],  tile 1 (A = [a0|a2]) = [[w2;w0] | 0]       3 MMAs -> 2 per tap
            ci0 = 0;
            if (bt == 0) plane = region;
            else { plane = (khalf == 0) ? 2 : 0; zero = (region == 1); }
        } else if (s.c8) { ci0 = 0; plane = region; zero = (region == 1 && khalf == 1); }
        else {
            ci0 = cg * 16 + khalf * 8;
            if (bt == 0) plane = region;                       // A = a_hi:  [Whi | Wlo]
            else { plane = 0; zero = (region == 0); }          // A = a_lo:  [ 0  | Whi]
        }
    } else if (s.c8) {
        // K16 = two plane blocks of the same 8 channels.  P=2: tile0 [w0;w0], tile1 [w1;0].
        //                                                  P=3: tile0 [w0;w0], tile1 [w1;w1], tile2 [w2;w0].
        ci0 = 0;
        if (bt == 0) plane = 0;
        else if (P == 2) { plane = 1; zero = (khalf == 1); }
        else if (bt == 1) plane = 1;
        else plane = (khalf == 0) ? 2 : 0;
    } else {
        plane = bt; ci0 = cg * 16 + khalf * 8;
    }
    if (!zero && co < c_out) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint16_t a[3], b[3];
            // dgrad: the image of the data-gradient conv W'[co][ci][tap] = w[ci][co][flipped tap], read in place from
            // the forward weight (c_in, c_out here are the data-gradient conv's: c_in = the forward c_out)
            const int ca = ci0 + 2 * i, cb = ci0 + 2 * i + 1;
            // (dgrad = row length of the forward weight in channels, >= c_out: a launch may take a slice of them)
            const int64_t ia = dgrad ? ((int64_t)ca * dgrad + co) * ntaps + (ntaps - 1 - tap) : ((int64_t)co * c_in + ca) * ntaps + tap;
            const int64_t ib = dgrad ? ((int64_t)cb * dgrad + co) * ntaps + (ntaps - 1 - tap) : ((int64_t)co * c_in + cb) * ntaps + tap;
            lea_split_planes(w[ia], P, a);
            lea_split_planes(w[ib], P, b);
            q[i] = (uint32_t)a[plane] | ((uint32_t)b[plane] << 16);
        }
    }
    lea_u4 o; o.x = q[0]; o.y = q[1]; o.z = q[2]; o.w = q[3];
    img[gidx] = o;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

int device_sm_count() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

// Folded weight images (see the layout comment above tc_shape) are the default; LEA_TC_FOLD=0 in the environment of the
// process restores the term-by-term images for A/B measurements.  Read once: pack and launch must agree.
int tc_fold_enabled() {      // 0 = term by term, 1 = folds for <= 16 output channels, 2 (default) = also the single-tile fold for 32
    static const int v = [] { const char* e = getenv("LEA_TC_FOLD"); return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2; }();
    return v;
}

int tc_launch(const lea_conv* c, const void* wimg, const lea_tc_opts* opts, void* stream, int swap_lbo_sbo) {
    const bool fused = opts && opts->fused_cv;
    const int P = fused ? opts->fx.P : c->src.P;
    const TcShape s = tc_shape(c->c_in, c->c_out, c->ksize, P, tc_fold_enabled());
    if (fused) {
        LEA_CHECK(opts->cv_maps != nullptr, "conv3d_tc: fused_cv needs cv_maps (lea_build_fused_cv_maps)");
        LEA_CHECK(c->ksize == 3 && c->c_in == 2 * opts->fx.C && (opts->fx.C % 16) == 0,
                  "conv3d_tc: fused_cv needs ksize 3 and c_in == 2*C with C %% 16 == 0");
        LEA_CHECK(opts->fx.D == 1 && opts->fy.D == 1 && opts->fx.P == opts->fy.P && opts->fx.C == opts->fy.C &&
                  opts->fx.B == opts->fy.B && opts->fx.H == opts->fy.H && opts->fx.W == opts->fy.W,
                  "conv3d_tc: fused_cv feature volumes must be matching 2-D planes volumes");
        LEA_CHECK(opts->d3 >= 1 && opts->d3 <= opts->fx.W, "conv3d_tc: fused_cv needs 1 <= D3 <= W3");
    }
    LEA_CHECK(s.ok, "conv3d_tc: shape c_in=%d c_out=%d k=%d planes=%d is not taken by the tensor-core kernel",
              c->c_in, c->c_out, c->ksize, P);
    LEA_CHECK(c->dst_f32 != nullptr || (c->dst.P == P && (!c->has_res || c->res.P == P)),
              "conv3d_tc: src/dst/res plane counts differ");
    LEA_CHECK(c->dst_f32 == nullptr || c->c_out <= 8, "conv3d_tc: fp32 output supports c_out <= 8");
    const int single = opts && opts->mma_terms == 1;

    TcParams p{};
    if (fused) {
        p.B = opts->fx.B; p.D = opts->d3; p.H = opts->fx.H; p.W = opts->fx.W;
        p.g0_stride_b = (opts->fx.C >> 3) * P;
        p.g0_first = 0;
        p.fused_cv = 1;
        p.cvmaps = reinterpret_cast<const CUtensorMap*>(opts->cv_maps);
    } else {
        p.B = c->src.B; p.D = c->src.D; p.H = c->src.H; p.W = c->src.W;
        p.g0_stride_b = (c->src.C >> 3) * P;
        p.g0_first = (c->src_c0 >> 3) * P;
    }
    p.P = P;
    LEA_CHECK(c->dst_f32 != nullptr || (c->dst.B == p.B && c->dst.D == p.D && c->dst.H == p.H && c->dst.W == p.W),
              "conv3d_tc: output volume does not match the input geometry");
    p.ks = c->ksize; p.taps = s.taps2d * c->ksize;
    p.NP = s.NP; p.c_out = c->c_out;
    p.ncg = s.ncg;
    p.ncg_half = s.ncg / 2;
    p.nbt = s.nbt; p.nb_rows = s.nb_rows; p.btile_bytes = s.btile_bytes; p.wpart_bytes = s.wpart_bytes;
    p.nterm = 0;
    auto add_term = [&](int aoff, int lbo, int btile, int region, int first) {
        p.term_aoff[p.nterm] = aoff; p.term_lbo_blocks[p.nterm] = lbo; p.term_btile[p.nterm] = btile;
        p.term_region[p.nterm] = region; p.term_first[p.nterm] = first; p.term_skip[p.nterm] = 0; ++p.nterm;
    };
    p.fold = s.fold ? 1 : 0;
    if (s.fold) {
        // terms that share their A operand are one MMA over [main | correction] virtual channels (layout comment above)
        p.ngroups = 1;
        if (s.c8) {
            p.blocks_per_cg = P;
            add_term(0, 1, 0, 0, 1);                           // [a0|a1] x [[w0;w0] | [w1;0]]     (P = 3: [[w0;w0] | [w1;w1]])
            if (P == 3 && !single) add_term(0, 2, 1, 0, 0);    // [a0|a2] x [[w2;w0] | 0]
        } else {
            p.blocks_per_cg = 2 * P;
            add_term(0, P, 0, 0, 1);                           // a0 x [w0 | w1]
            if (!single && s.nbt == 1) {                       // 64 virtual channels: a1 x the SAME tile, one block later
                add_term(1, P, 0, 0, 0);
                p.term_skip[p.nterm - 1] = s.NP >> 1;
            } else if (!single) add_term(1, P, 1, 0, 0);       // a1 x [ 0 | w0]
        }
    } else if (s.c8) {
        p.blocks_per_cg = P;                                   // 1 channel block x P planes
        p.ngroups = 1;
        add_term(0, 1, 0, 0, 1);                               // [a0|a1] x [w0;w0]
        if (!single && P == 2) add_term(0, 1, 1, 0, 0);        // [a0|a1] x [w1;0]
        if (!single && P == 3) {
            add_term(0, 1, 1, 0, 0);                           // [a0|a1] x [w1;w1]
            add_term(0, 2, 2, 0, 0);                           // [a0|a2] x [w2;w0]
        }
    } else {
        p.blocks_per_cg = 2 * P;                               // 2 channel blocks x P planes, order [cb][plane]
        if (single) { add_term(0, P, 0, 0, 1); p.ngroups = 1; }
        else {
            // Product terms a_t * w_pw with t + pw < P.  The tensor core's fp32 accumulate loses ~1 ulp per
            // accumulation step (measured: error grows linearly with the number of steps), so the dominant term
            // hi*hi gets an accumulator of its own (region 0) and all correction terms, 2^-8 smaller, share
            // region 1; the epilogue adds the two regions in fp32.
            // accum_split 3: two regions only for long reductions (more than one 16-channel group at k = 3); a single
            // 16-channel group accumulates 9 taps x 3 terms = 27 steps, where one region costs little accuracy and
            // halves the TMEM columns per depth (twice the depth per item: fewer halo slabs)
            bool split = !(opts && opts->accum_split == 2);
            if (opts && opts->accum_split == 3) split = (s.ncg * s.taps2d > 9);
            if (opts && opts->accum_split == 4) split = (s.ncg * s.taps2d > 18);      // ... more than two groups
            p.ngroups = (P > 1 && split) ? 2 : 1;
            // mma_terms = 2 on 3-plane volumes: only the terms of order < 2 (a0 w0, a1 w0, a0 w1) - bf16x3 products on
            // exactly stored operands, half the MMAs of the full 6-term product
            const int order = (opts && opts->mma_terms >= 2 && opts->mma_terms < P) ? opts->mma_terms : P;
            for (int pw = 0; pw < order; ++pw)                 // weight plane pw = weight tile pw
                for (int t = 0; t + pw < order; ++t) {         // activation plane t (plane t of cb 0; cb 1 is P blocks on)
                    const bool main_term = (pw == 0 && t == 0);
                    const bool first_small = (pw == 0 && t == 1);
                    add_term(t, P, pw, (main_term || p.ngroups == 1) ? 0 : 1,
                             main_term || (first_small && p.ngroups == 2));
                }
        }
    }
    const int accw = p.ngroups * p.NP;                         // TMEM columns per output depth
    // long reductions: the epilogue is a few % of an item, so spend all of TMEM on depth (fewer halo slabs);
    // short reductions: keep two accumulator sets so that the epilogue overlaps the next item's MMAs
    // two accumulator sets let the epilogue of item i overlap the MMAs of item i+1; when that would leave fewer than
    // 4 depth slices per item (wide N), one set with twice the depth wastes fewer halo slabs
    // One or two TMEM accumulator sets.  Two sets let the epilogue of item i overlap the MMAs of item i+1 at half the
    // depth per item (more halo slabs); when that leaves fewer than 4 depths a measured cycle model decides: a slab
    // costs M = groups * taps * terms * max(64, N/2) cycles of tensor pipe, the epilogue E ~ 650 cycles per 16 output
    // channels and depth.
    // (Round 1 also carried two schedules that drain depths while their item is still accumulating - a rolling TMEM
    // ring and an "early drain" single set.  Both saved halo slabs and measured slower; removed in round 2.
    // Round 2 measured a shared-memory-staged TMA-store epilogue - cp.async.bulk.tensor stores of each warp's 32-voxel box
    // per depth and 8-channel block, then one proxy fence per 4 boxes: the epilogue-bound convs lose 12-20 % against the
    // per-thread 16-byte stores below (8 -> 8 at 64x128x416, 4 pairs: 450 us -> 565 / 517 us); not kept.)
    // Flat mode: a 3x3(x3) conv on a depth-1 volume (the 2-D feature net, the 2-D maps of the collapsed stem0) takes the
    // 8-wide w tiles of a row of tiles as the "depth slices" of an item - no halo slabs, only the middle tap plane - so
    // that one accumulator hand-over and one item set-up serve up to 16 tiles instead of one (they cost each role
    // ~600 cycles per item, a third of a depth-1 item's time).
    p.flat = (!fused && p.D == 1 && c->dst_f32 == nullptr && !(opts && opts->debug & 16)) ? 1 : 0;
    const bool dhalo = (p.ks == 3 && !p.flat);                 // halo slabs along the chunked axis
    p.nsets = (512 / (2 * accw) >= 4) ? 2 : 1;
    if (p.nsets == 1 && 512 / (2 * accw) >= 2) {
        const int halo = dhalo ? 2 : 0;
        const int nmax = (dhalo ? 3 : 1) * s.NP;
        const double M = (double)s.ncg * s.taps2d * p.nterm * (nmax / 2 > 64 ? nmax / 2 : 64);
        const double E = 650.0 * ((c->c_out + 15) / 16);
        const int d1 = 512 / accw, d2 = 512 / (2 * accw);
        const double t1 = ((d1 + halo) * M + d1 * E) / d1;
        const double me = (d2 + halo) * M, ee = d2 * E;
        const double t2 = (me > ee ? me : ee) / d2;
        if (t2 < t1) p.nsets = 2;
    }
    if (opts && (opts->acc_sets == 1 || opts->acc_sets == 2)) p.nsets = opts->acc_sets;
    // tile shape: 8 x 16 for k = 3 (the tap windows need the 8-row core-matrix groups to be rows of the slab); a
    // 1x1x1 conv has no halo, its slab is 128 consecutive rows whatever the shape, so it takes the widest tile the
    // volume fills without padding (longer contiguous runs for TMA and for the epilogue's stores)
    p.tw_log2 = 3;
    if (p.ks == 1) {
        int64_t best_area = -1;
        for (int l2 = 3; l2 <= 5; ++l2) {
            const int tw = 1 << l2, th = 128 >> l2;
            const int64_t area = (int64_t)((p.W + tw - 1) / tw) * tw * (((p.H + th - 1) / th) * th);
            if (best_area < 0 || area <= best_area) { best_area = area; p.tw_log2 = l2; }
        }
        if (opts && opts->tile_w_log2 >= 3 && opts->tile_w_log2 <= 7) p.tw_log2 = opts->tile_w_log2;
    }
    const int tile_w = 1 << p.tw_log2, tile_h = 128 >> p.tw_log2;
    p.pitch_vox = (p.ks == 3) ? tile_w + 2 : tile_w;
    p.slab_vox = p.pitch_vox * ((p.ks == 3) ? tile_h + 2 : tile_h);
    p.blk_bytes = p.slab_vox * 16;
    p.stage_bytes = p.blocks_per_cg * p.blk_bytes;
    p.stage_stride = (p.stage_bytes + 127) & ~127;
    p.PD = (p.W + tile_w - 1) / tile_w;
    const int depth_n = p.flat ? p.PD : p.D;                   // slices an item chunks over
    p.tiles_w = p.flat ? 1 : (p.W + tile_w - 1) / tile_w;
    p.tiles_h = (p.H + tile_h - 1) / tile_h;
    p.dbg = opts ? opts->debug : 0;
    p.cv_skip = (fused && opts->cv_skip == 1) ? 1 : 0;
    int dc_max = 512 / (p.nsets * accw);
    if (dc_max > 16) dc_max = 16;
    if (dc_max > depth_n) dc_max = depth_n;
    const int num_sms = (opts && opts->num_sms > 0) ? opts->num_sms : device_sm_count();
    // Depth slices per work item: an item of Dc slices streams Dc + 2*halo slabs at a fixed MMA cost per slab, and the
    // persistent grid runs ceil(items / SMs) items per CTA - pick the Dc that minimises slabs on the busiest SM.
    int Dc = 1;
    int64_t best = -1;
    for (int cand = 1; cand <= dc_max; ++cand) {
        const int64_t items = (int64_t)p.B * ((depth_n + cand - 1) / cand) * p.tiles_h * p.tiles_w;
        // (flat mode: an item's set-up costs each role about half a slab of a small 2-D conv - without that term one tile
        //  per item always wins the count of slabs and nothing is amortised)
        const int64_t cost = ((items + num_sms - 1) / num_sms) * (p.flat ? 2 * cand + 1 : 2 * (cand + (dhalo ? 2 : 0)));
        if (best < 0 || cost < best || (cost == best && cand > Dc)) { best = cost; Dc = cand; }
    }
    if (opts && opts->depth_chunk > 0) Dc = opts->depth_chunk < dc_max ? opts->depth_chunk : dc_max;
    p.Dc = Dc;
    // TMEM columns a thin depth chunk leaves unused become further accumulator sets (more items in flight between issuer
    // and epilogue).  Measured: no effect on the depth-1 convs it was tried for - their per-item cost was the item set-up
    // of each role, not the hand-over round trip - and none anywhere else; kept because it is free.
    if (!(opts && (opts->acc_sets == 1 || opts->acc_sets == 2))) {
        int more = 512 / (accw * Dc);
        if (more > kMaxSets) more = kMaxSets;
        if (more > p.nsets) p.nsets = more;
    }
    p.dchunks = (depth_n + Dc - 1) / Dc;
    const int64_t total = (int64_t)p.B * p.dchunks * p.tiles_h * p.tiles_w;
    LEA_CHECK(total < (1ll << 31), "conv3d_tc: too many work items");
    p.total_items = (int)total;
    const int wstride = (p.wpart_bytes + 127) & ~127;
    p.nwbuf = (2 * wstride + 3 * p.stage_stride + kHeaderBytes <= kSmemBudget) ? 2 : 1;
    // every channel group's weight part resident (one load per CTA instead of one per item and group) when that
    // still leaves a useful activation ring
    p.wres = (kHeaderBytes + (int64_t)p.ncg * wstride + 6 * (int64_t)p.stage_stride <= kSmemBudget) ? 1 : 0;
    if (opts && opts->resident_weights == 2) p.wres = 0;
    if (p.wres) p.nwbuf = p.ncg;

    int nst = (kSmemBudget - kHeaderBytes - p.nwbuf * wstride) / p.stage_stride;
    if (nst > kMaxStages) nst = kMaxStages;
    LEA_CHECK(nst >= 2, "conv3d_tc: shared memory too small for this shape (weights part %d B)", p.wpart_bytes);
    p.nstages = nst;
    // (the producer requests a group's slabs before its weight parts: the ring must hold all of them, or it would wait for
    //  slots that the issuer - waiting for those weights - never frees)
    p.wsplit = (!p.wres && p.nwbuf == 1 && p.ks == 3 && !p.flat && !fused && (p.wpart_bytes % 48) == 0 &&
                p.nstages >= p.Dc + 3 && !(opts && opts->debug & 32)) ? 1 : 0;
    p.swap_lbo_sbo = swap_lbo_sbo;
    p.wimg = reinterpret_cast<const uint8_t*>(wimg);
    p.bn_scale = c->bn_scale; p.bn_shift = c->bn_shift; p.relu = c->relu;
    p.dst = c->dst; p.dst_c0 = c->dst_c0; p.res = c->res; p.res_c0 = c->res_c0; p.has_res = c->has_res;
    p.dst_f32 = c->dst_f32;
    const size_t smem = (size_t)kHeaderBytes + (size_t)p.nwbuf * wstride + (size_t)p.nstages * p.stage_stride;

    PFN_encodeTiled encode = get_encode_fn();
    LEA_CHECK(encode != nullptr, "conv3d_tc: cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap tmap;
    // (with fused_cv the 5-D map is unused by the kernel; it is encoded over fx only to keep one launch signature)
    const int mapC = fused ? opts->fx.C : c->src.C;
    const int mapD = fused ? 1 : p.D;
    void* map_base = fused ? opts->fx.data : c->src.data;
    const cuuint64_t G = (cuuint64_t)p.B * (mapC >> 3) * P;
    // the 8-channel group and w are contiguous in memory: one merged inner dimension of W*8 elements, so that a box
    // row is one 128..512-byte run for the TMA engine instead of 8..32 separate 16-byte rows
    cuuint64_t gdim[4] = {(cuuint64_t)p.W * 8, (cuuint64_t)p.H, (cuuint64_t)mapD, G};
    cuuint64_t gstr[3] = {(cuuint64_t)p.W * 16, (cuuint64_t)p.H * p.W * 16, (cuuint64_t)mapD * p.H * p.W * 16};
    cuuint32_t box[4] = {(cuuint32_t)p.pitch_vox * 8, (cuuint32_t)(p.slab_vox / p.pitch_vox), 1,
                         (cuuint32_t)p.blocks_per_cg};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, map_base, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LEA_CHECK(cr == CUDA_SUCCESS, "conv3d_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    const int e8 = (p.c_out & 15) == 8 ? (p.c_out == 8 ? 1 : 2) : 0;
    TcKernelFn kernel = tc_kernel_for(p.flat ? (p.ks == 3 ? 2 : 4) : (p.wsplit ? 5 : p.ks), p.nterm, P, e8);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    LEA_CHECK(e == cudaSuccess, "conv3d_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    const int grid = p.total_items < num_sms ? p.total_items : num_sms;
    {
        int r = grid;
        p.step_tw = r % p.tiles_w; r /= p.tiles_w;
        p.step_th = r % p.tiles_h; r /= p.tiles_h;
        p.step_dc = r % p.dchunks; r /= p.dchunks;
        p.step_b = r;
    }
    // always request the full budget so that exactly one CTA (which owns all 512 TMEM columns) fits per SM
    // programmatic stream serialization (see the kernel's prologue): opt-in with LEA_TC_PDL=1.  Measured inside the CUDA
    // graph of the KITTI step: 196.0 pairs/s with, 196.1 without - launch gaps are not what the step loses.
    static const int use_pdl = [] { const char* v = getenv("LEA_TC_PDL"); return (v && v[0] == '1') ? 1 : 0; }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)tc_threads(P, e8));
    cfg.dynamicSmemBytes = kSmemBudget; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, kernel, tmap, p);
    (void)smem;
    if (e == cudaSuccess) e = cudaGetLastError();
    LEA_CHECK(e == cudaSuccess, "conv3d_tc: launch failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace

extern "C" int64_t lea_tc_weight_image_bytes(int32_t c_in, int32_t c_out, int32_t ksize, int32_t planes) {
    const TcShape s = tc_shape(c_in, c_out, ksize, planes, tc_fold_enabled());
    if (!s.ok) return 0;
    return (int64_t)s.ncg * s.wpart_bytes;
}

static int pack_weights_tc(const float* weight, void* wimg, int32_t c_in, int32_t c_out, int32_t ksize,
                           int32_t planes, void* stream, int dgrad) {
    const TcShape s = tc_shape(c_in, c_out, ksize, planes, tc_fold_enabled());
    LEA_CHECK(s.ok, "pack_weights_tc: unsupported shape c_in=%d c_out=%d k=%d planes=%d", c_in, c_out, ksize, planes);
    LEA_CHECK(weight && wimg && ((((uintptr_t)wimg) & 15) == 0), "pack_weights_tc: bad pointer");
    const int total = (int)((int64_t)s.ncg * s.wpart_bytes / 16);
    lea_pack_weights_tc_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        weight, reinterpret_cast<lea_u4*>(wimg), c_in, c_out, ksize, planes, total, tc_fold_enabled(), dgrad);
    cudaError_t e = cudaGetLastError();
    LEA_CHECK(e == cudaSuccess, "pack_weights_tc: launch failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int lea_pack_weights_tc(const float* weight, void* wimg, int32_t c_in, int32_t c_out, int32_t ksize,
                                   int32_t planes, void* stream) {
    return pack_weights_tc(weight, wimg, c_in, c_out, ksize, planes, stream, 0);
}

extern "C" int lea_pack_weights_tc_dgrad(const float* weight, void* wimg, int32_t c_in, int32_t c_out, int32_t ksize,
                                         int32_t planes, int32_t fwd_c_in, int32_t fwd_ci0, void* stream) {
    LEA_CHECK(fwd_ci0 >= 0 && fwd_ci0 + c_out <= fwd_c_in, "pack_weights_tc_dgrad: channel slice outside the forward weight");
    const int64_t ntaps = (int64_t)ksize * ksize * ksize;
    return pack_weights_tc(weight ? weight + (int64_t)fwd_ci0 * ntaps : weight, wimg, c_in, c_out, ksize, planes, stream,
                           fwd_c_in);
}

extern "C" int lea_conv3d_tc(const lea_conv* c, const void* wimg, const lea_tc_opts* opts, void* stream) {
    LEA_CHECK(c != nullptr && wimg != nullptr, "conv3d_tc: null argument");
    return tc_launch(c, wimg, opts, stream, 0);
}

extern "C" int lea_conv3d_tc_debug(const lea_conv* c, const void* wimg, const lea_tc_opts* opts, void* stream,
                                   int swap_lbo_sbo) {
    return tc_launch(c, wimg, opts, stream, swap_lbo_sbo);
}

#ifdef LEA_TC_ABLATION
extern "C" int lea_tc_prof(long long* out12) {
    return cudaMemcpyFromSymbol(out12, g_lea_tc_prof, sizeof(long long) * 12) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int lea_tc_status(void) {
    int v = 0;
    cudaMemcpyFromSymbol(&v, g_lea_tc_status, sizeof(int));
    return v;
}

extern "C" int64_t lea_fused_cv_maps_bytes(int32_t d3) { return d3 > 0 ? (int64_t)2 * d3 * sizeof(CUtensorMap) : 0; }

extern "C" int lea_build_fused_cv_maps(const lea_vol* fx, const lea_vol* fy, int32_t d3, void* maps_dev, void* stream) {
    LEA_CHECK(fx && fy && maps_dev && fx->data && fy->data, "build_fused_cv_maps: null argument");
    LEA_CHECK(fx->D == 1 && fy->D == 1 && fx->P == fy->P && fx->C == fy->C && fx->B == fy->B && fx->H == fy->H &&
              fx->W == fy->W && (fx->C % 16) == 0, "build_fused_cv_maps: feature volumes must match (D == 1, C %% 16 == 0)");
    LEA_CHECK(d3 >= 1 && d3 <= fx->W, "build_fused_cv_maps: need 1 <= D3 <= W3");
    LEA_CHECK((((uintptr_t)maps_dev) & 63) == 0, "build_fused_cv_maps: maps must be 64-byte aligned");
    PFN_encodeTiled encode = get_encode_fn();
    LEA_CHECK(encode != nullptr, "build_fused_cv_maps: cuTensorMapEncodeTiled is not available from the driver");
    const int P = fx->P, H = fx->H, W = fx->W;
    const cuuint64_t G = (cuuint64_t)fx->B * (fx->C >> 3) * P;
    std::vector<CUtensorMap> maps(2 * (size_t)d3);
    for (int side = 0; side < 2; ++side) {
        for (int d = 0; d < d3; ++d) {
            // x: origin shifted right by d voxels  -> coordinate w' = w - d addresses x[h, w];   w' < 0  <=> w < d -> 0
            // y: origin unchanged                  -> coordinate w' = w - d addresses y[h, w-d]; w' < 0  <=> w < d -> 0
            // both: width W - d                    -> w' >= W - d <=> w >= W (right halo) -> 0
            uint8_t* base = reinterpret_cast<uint8_t*>(side == 0 ? fx->data : fy->data) + (side == 0 ? (size_t)d * 16 : 0);
            cuuint64_t gdim[3] = {(cuuint64_t)(W - d) * 8, (cuuint64_t)H, G};
            cuuint64_t gstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
            cuuint32_t box[3] = {(cuuint32_t)(LEA_TC_TW + 2) * 8, (cuuint32_t)(LEA_TC_TH + 2), (cuuint32_t)(2 * P)};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult cr = encode(&maps[(size_t)side * d3 + d], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gdim, gstr, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            LEA_CHECK(cr == CUDA_SUCCESS, "build_fused_cv_maps: cuTensorMapEncodeTiled failed (%d) at d=%d", (int)cr, d);
        }
    }
    cudaError_t e = cudaMemcpyAsync(maps_dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice,
                                    (cudaStream_t)stream);
    LEA_CHECK(e == cudaSuccess, "build_fused_cv_maps: copy failed: %s", cudaGetErrorString(e));
    e = cudaStreamSynchronize((cudaStream_t)stream);       // the host vector goes out of scope on return
    LEA_CHECK(e == cudaSuccess, "build_fused_cv_maps: sync failed: %s", cudaGetErrorString(e));
    return 0;
}
