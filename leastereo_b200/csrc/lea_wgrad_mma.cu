// Weight gradient of the 3x3x3 and 1x1x1 ConvBR convolutions on the tensor cores (train.py:158, models/operations_3d.py:37):
//
//     dW[co, ci, kd, kh, kw] = sum_{b,d,h,w} dY[b, co, d, h, w] * X[b, ci, d+kd-1, h+kh-1, w+kw-1]
//
// GEMM view per tap: D_tap[ci, co] = sum_v X_tap[v, ci] * dY[v, co] - the contraction runs over VOXELS (millions) and the
// output is tiny (8..32 x 8..32 per tap and channel tile).  That shape is the opposite of what tcgen05.mma wants
// (128 x N x 16 per instruction: with ci, co <= 32 at most 6 % of the array would do useful work), so this kernel uses
// the warp-level tensor path, mma.sync.m16n8k16 (bf16 x bf16 -> fp32), whose 16 x 8 tiles fit the problem exactly
// (measured on B200: 556 TFLOP/s dense for this instruction, tools/micro/hmma_rate.cu).
//
// Both operands live in HBM as bf16 planes blocked by 8 channels (include/leastereo_b200.h): one (voxel, 8 channels)
// group is a 16-byte row, exactly a row of an 8x8 ldmatrix tile with the VOXEL as row index.  The MMA needs the voxel
// index as K, i.e. the transposed tile - ldmatrix.trans delivers it, so no data is re-laid-out anywhere:
//   A fragment (16 ci x 16 voxels) = four 8x8 tiles (2 channel blocks x 2 tile rows) of the X halo slab, its start
//                                    shifted by the tap (kh, kw) - every tap is just another start address;
//   B fragment (16 voxels x 8 co)  = two 8x8 tiles of the dY tile.
// Split precision (2 planes: value = hi + lo): the three product terms hi*hi, hi*lo, lo*hi accumulate in fp32.
//
// Schedule.  Persistent CTAs; a work item is a column (b, 16 x 8 voxel tile) of Dc depths.  One producer warp streams
// the X halo slabs (18 x 10 voxels, TMA box load, zero fill outside the volume = the conv's padding) through a 4-deep
// ring - slab d serves output depths d-1, d, d+1, so each slab is loaded once per column - and the dY tiles through a
// 2-deep ring.  Nine compute warps own the nine (kd, kh) pairs and loop over kw; each keeps its 3 taps x CI_T x CO_T
// accumulators in registers for the whole kernel and adds them to dW with atomics once at the end.
// Channel tiles (CI_T = CO_T = 8, 16 or 32) are mapped to blockIdx.y, so 64 -> 32 (stem0) and 192 -> 64 (conv1/conv2) run
// as 2 and 12 independent channel-tile problems over the same voxels.
// 1x1x1 convs (KS = 1; the cells' pre-processing layers) have one tap and no halo: the slab is the tile itself, and
// the compute warps split the 8 row pairs of a tile between them instead of the taps (HBM-bound: every operand byte
// feeds only CI_T x CO_T / 8 MACs).
#include "lea_common.h"
#include <cuda.h>
#include <cstdio>

namespace {

constexpr int kWgWarps = 9;                    // compute warps: (kd, kh) = (warp / 3, warp % 3)
constexpr int kWgThreads = 32 * (kWgWarps + 1);
constexpr int kXSlots = 4, kYSlots = 2;
constexpr int kYBlk = 16 * 8 * 16;             // bytes of one (channel block, plane) of a dY tile
__host__ __device__ constexpr int wg_xblk(int ks) { return ks == 3 ? 18 * 10 * 16 : kYBlk; }   // ... of an X (halo) slab

struct WgParams {
    int B, D, H, W;
    int c_in, c_out;                           // full channel counts of the weight tensor
    int gx_stride_b, gx_first;                 // X tensor-map block index: b * gx_stride_b + gx_first + ci_tile * (CI_T/8) * 2
    int gy_stride_b, gy_first;
    int tiles_h, tiles_w, dchunks, Dc, total_items;
    int ci_tiles;                              // blockIdx.y = co_tile * ci_tiles + ci_tile
    float* dw;
};

__device__ __forceinline__ uint32_t wg_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wg_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void wg_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wg_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wg_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void wg_tma_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void wg_ldm_x4_t(uint32_t addr, uint32_t* r) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void wg_ldm_x2_t(uint32_t addr, uint32_t* r) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void wg_mma(float* c, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

struct WgItem { int b, h0, w0, d0, d1; };
__device__ __forceinline__ WgItem wg_item(const WgParams& p, int item) {
    WgItem g;
    int r = item;
    const int dc = r % p.dchunks; r /= p.dchunks;          // depth chunks of one column are consecutive items
    const int tw = r % p.tiles_w; r /= p.tiles_w;
    const int th = r % p.tiles_h; r /= p.tiles_h;
    g.b = r; g.h0 = th * 16; g.w0 = tw * 8;
    g.d0 = dc * p.Dc; g.d1 = min(g.d0 + p.Dc, p.D);
    return g;
}

// CI_T, CO_T: channel tile (8, 16 or 32 each).  MB = number of 16-row MMA tiles along ci (CI_T == 8: one tile holding the
// taps kw = 0 | 1 in its two row halves and a second one for kw = 2), NB = 8-column tiles along co.
template <int KS, int CI_T, int CO_T>
__global__ void __launch_bounds__(kWgThreads, (CI_T == 32 ? 1 : 2))
lea_wgrad_mma_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap,
                     const __grid_constant__ WgParams p) {
    constexpr int kXBlk = wg_xblk(KS);
    constexpr int XB = (CI_T / 8) * 2, YB = (CO_T / 8) * 2;           // blocks (channel block x plane) per slab / tile
    constexpr int XSLAB = XB * kXBlk, YTILE = YB * kYBlk;
    constexpr int NB = CO_T / 8;
    constexpr bool C8 = (CI_T == 8);
    constexpr int MB = C8 ? 1 : CI_T / 16;
    constexpr int NACC = (KS == 1) ? 1 : (C8 ? 2 : 3);               // accumulator groups: C8: (kw0|kw1), (kw2|-); else kw = 0,1,2
    constexpr int kHalo = (KS == 3) ? 1 : 0;
    constexpr int kRowB = (KS == 3) ? 160 : 128;                      // bytes of one slab row (10 or 8 voxels)
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);               // xfull[4] xempty[4] yfull[2] yempty[2]
    uint8_t* xring = smem + 1024;
    uint8_t* yring = xring + kXSlots * XSLAB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci_tile = blockIdx.y % p.ci_tiles, co_tile = blockIdx.y / p.ci_tiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kXSlots; ++i) { wg_mbar_init(wg_smem(bars + i), 1); wg_mbar_init(wg_smem(bars + kXSlots + i), kWgWarps); }
        for (int i = 0; i < kYSlots; ++i) {
            wg_mbar_init(wg_smem(bars + 2 * kXSlots + i), 1); wg_mbar_init(wg_smem(bars + 2 * kXSlots + kYSlots + i), kWgWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t xfull = wg_smem(bars), xempty = wg_smem(bars + kXSlots);
    const uint32_t yfull = wg_smem(bars + 2 * kXSlots), yempty = wg_smem(bars + 2 * kXSlots + kYSlots);

    if (warp == kWgWarps) {
        // ================= producer =================
        if (lane == 0) {
            uint32_t xs = 0, ys = 0;                                  // slabs / tiles issued so far
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const WgItem g = wg_item(p, item);
                const int gx = g.b * p.gx_stride_b + p.gx_first + ci_tile * XB;
                const int gy = g.b * p.gy_stride_b + p.gy_first + co_tile * YB;
                // interleave: X slab d0-1, d0, then (X slab d+1, dY tile d) for every depth of the chunk
                // (KS = 1: X slab d, dY tile d)
                for (int j = -kHalo; j < g.d1 - g.d0 + kHalo; ++j) {
                    const int d = g.d0 + j;                           // X slab depth (outside [0, D): zero fill)
                    const uint32_t slot = xs % kXSlots, ph = (xs / kXSlots) & 1;
                    wg_mbar_wait(xempty + 8 * slot, ph ^ 1);
                    wg_mbar_expect(xfull + 8 * slot, (uint32_t)XSLAB);
                    wg_tma_4d(wg_smem(xring + (size_t)slot * XSLAB), &xmap, xfull + 8 * slot, (g.w0 - kHalo) * 8, g.h0 - kHalo,
                              d, gx);
                    ++xs;
                    if (j >= kHalo) {                                 // X slabs up to d0 + j are issued: dY tile of depth d0 + j - halo
                        const uint32_t ysl = ys % kYSlots, yph = (ys / kYSlots) & 1;
                        wg_mbar_wait(yempty + 8 * ysl, yph ^ 1);
                        wg_mbar_expect(yfull + 8 * ysl, (uint32_t)YTILE);
                        wg_tma_4d(wg_smem(yring + (size_t)ysl * YTILE), &ymap, yfull + 8 * ysl, g.w0 * 8, g.h0, d - kHalo, gy);
                        ++ys;
                    }
                }
            }
        }
        return;
    }

    // ================= compute warps =================
    const int kd = (KS == 3) ? warp / 3 : 0, kh = (KS == 3) ? warp % 3 : 0;
    float acc[NACC][MB][NB][4];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int m = 0; m < MB; ++m)
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[a][m][n][i] = 0.0f;

    // per-lane byte offsets of the ldmatrix row addresses
    //   B (x2): lanes 0-7 -> tile row 2r, voxel lane; lanes 8-15 -> tile row 2r+1  (lanes 16-31: any valid address)
    const uint32_t b_lane = (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * 16);
    //   A (x4): matrices (block 0, row), (block 1, row), (block 0, row + 1), (block 1, row + 1); block = channel block of the
    //           16-channel MMA tile - or, for 8-channel volumes, the tap kw = 0 / 1 (shift of one voxel = 16 bytes)
    //           (KS = 1 with 8 channels: both row halves read the same block; rows 8-15 of the result are ignored)
    const uint32_t a_lane = (uint32_t)((lane & 7) * 16 + ((lane >> 4) & 1) * kRowB +
                                       ((lane >> 3) & 1) * (C8 ? (KS == 3 ? 16 : 0) : 2 * kXBlk));

    uint32_t xs = 0, ys = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const WgItem g = wg_item(p, item);
        const int nd = g.d1 - g.d0;
        for (int i = 0; i < nd; ++i) {
            // X slabs of depths d-1, d, d+1 are sequence numbers xs + i, xs + i + 1, xs + i + 2; this warp reads slab i + kd
            // (a slot is re-used only after ALL warps released it, so warps stay within kXSlots depths of each other and a
            // barrier can never be two phases ahead of a waiter)
            const uint32_t xq = xs + (uint32_t)i + (uint32_t)kd;
            wg_mbar_wait(xfull + 8 * (xq % kXSlots), (xq / kXSlots) & 1);
            const uint32_t yq = ys + (uint32_t)i;
            wg_mbar_wait(yfull + 8 * (yq % kYSlots), (yq / kYSlots) & 1);
            const uint32_t xbase = wg_smem(xring + (size_t)(xq % kXSlots) * XSLAB) + (uint32_t)(kh * kRowB) + a_lane;
            const uint32_t ybase = wg_smem(yring + (size_t)(yq % kYSlots) * YTILE) + b_lane;
            // KS = 3: every warp walks the 8 row pairs for its 3 taps; KS = 1: one tap, warp w takes row pair w (warp 8 idles)
            const int r_lo = (KS == 3) ? 0 : warp, r_hi = (KS == 3) ? 8 : (warp < 8 ? warp + 1 : warp);
#pragma unroll 2
            for (int r = r_lo; r < r_hi; ++r) {                       // 16 voxels per step: tile rows 2r, 2r+1
                uint32_t bh[NB][2], bl[NB][2];
#pragma unroll
                for (int n = 0; n < NB; ++n) {
                    wg_ldm_x2_t(ybase + (uint32_t)((n * 2) * kYBlk + r * 256), bh[n]);
                    wg_ldm_x2_t(ybase + (uint32_t)((n * 2 + 1) * kYBlk + r * 256), bl[n]);
                }
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    // tap shift along w: C8 packs kw = (0 | 1) into one tile and kw = 2 into the next, else a = kw
                    const uint32_t shift = (uint32_t)((C8 ? 2 * a : a) * 16);
#pragma unroll
                    for (int m = 0; m < MB; ++m) {
                        uint32_t ah[4], al[4];
                        const uint32_t addr = xbase + shift + (uint32_t)(r * 2 * kRowB) + (uint32_t)(C8 ? 0 : m * 4 * kXBlk);
                        wg_ldm_x4_t(addr, ah);
                        wg_ldm_x4_t(addr + kXBlk, al);
                        // ldmatrix delivers (block 0, k lo), (block 1, k lo), (block 0, k hi), (block 1, k hi) = a0, a1, a2, a3
#pragma unroll
                        for (int n = 0; n < NB; ++n) {
                            wg_mma(acc[a][m][n], ah, bh[n]);
                            wg_mma(acc[a][m][n], ah, bl[n]);
                            wg_mma(acc[a][m][n], al, bh[n]);
                        }
                    }
                }
            }
            // depth d is done for this warp: slab d-1 (sequence xs + i; KS = 1: slab d) and the dY tile are free
            __syncwarp();
            if (lane == 0) {
                wg_mbar_arrive(xempty + 8 * ((xs + (uint32_t)i) % kXSlots));
                wg_mbar_arrive(yempty + 8 * (yq % kYSlots));
            }
        }
        // the last two slabs of the column are not re-used by the next item
        __syncwarp();
        if (KS == 3 && lane == 0) {
            wg_mbar_arrive(xempty + 8 * ((xs + (uint32_t)nd) % kXSlots));
            wg_mbar_arrive(xempty + 8 * ((xs + (uint32_t)nd + 1u) % kXSlots));
        }
        xs += (uint32_t)nd + 2u * kHalo;
        ys += (uint32_t)nd;
    }

    // ---- write-out: acc[.][m][n] holds D[ci = 16 m + g (+8), co = 8 n + 2 t (+1)] of its tap; dW is (co, ci, kd, kh, kw)
    const int gq = lane >> 2, tq = lane & 3;
    const int ci0 = ci_tile * CI_T, co0 = co_tile * CO_T;
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int m = 0; m < MB; ++m)
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = gq + ((i >> 1) ? 8 : 0), col = 8 * n + 2 * tq + (i & 1);
                    int ci, kw;
                    if (C8) { kw = 2 * a + (row >> 3); ci = row & 7; if (kw > KS - 1) continue; }
                    else    { kw = a; ci = 16 * m + row; }
                    const int tap = (kd * 3 + kh) * 3 + kw;
                    atomicAdd(p.dw + ((int64_t)(co0 + col) * p.c_in + (ci0 + ci)) * (KS * KS * KS) + tap, acc[a][m][n][i]);
                }
}

typedef CUresult (*PFN_encodeTiledW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiledW wg_encode_fn() {
    static PFN_encodeTiledW fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiledW>(ptr);
    }
    return fn;
}

int wg_encode(CUtensorMap* map, const lea_vol* v, int box_w_vox, int box_h, int nblk) {
    PFN_encodeTiledW encode = wg_encode_fn();
    if (!encode) return 1;
    const cuuint64_t G = (cuuint64_t)v->B * (v->C >> 3) * v->P;
    cuuint64_t gdim[4] = {(cuuint64_t)v->W * 8, (cuuint64_t)v->H, (cuuint64_t)v->D, G};
    cuuint64_t gstr[3] = {(cuuint64_t)v->W * 16, (cuuint64_t)v->H * v->W * 16, (cuuint64_t)v->D * v->H * v->W * 16};
    cuuint32_t box[4] = {(cuuint32_t)box_w_vox * 8, (cuuint32_t)box_h, 1, (cuuint32_t)nblk};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, v->data, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

template <int KS, int T>
int wg_launch(const lea_vol* in, int in_c0, int c_in, const lea_vol* dout, int dout_c0, int c_out, float* dw, void* stream) {
    constexpr int XB = (T / 8) * 2, YB = (T / 8) * 2;
    constexpr int kXBlk = wg_xblk(KS);
    CUtensorMap xmap, ymap;
    LEA_CHECK(wg_encode(&xmap, in, KS == 3 ? 10 : 8, KS == 3 ? 18 : 16, XB) == 0 && wg_encode(&ymap, dout, 8, 16, YB) == 0,
              "conv3d_wgrad_tc: cuTensorMapEncodeTiled failed");
    WgParams p{};
    p.B = in->B; p.D = in->D; p.H = in->H; p.W = in->W;
    p.c_in = c_in; p.c_out = c_out;
    p.gx_stride_b = (in->C >> 3) * 2;   p.gx_first = (in_c0 >> 3) * 2;
    p.gy_stride_b = (dout->C >> 3) * 2; p.gy_first = (dout_c0 >> 3) * 2;
    p.tiles_h = (p.H + 15) / 16; p.tiles_w = (p.W + 7) / 8;
    p.ci_tiles = c_in / T;
    const int ytiles = p.ci_tiles * (c_out / T);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // depth chunk: as long as possible (two halo slabs per chunk) while the persistent grid still gets >= 2 items per CTA
    const int64_t cols = (int64_t)p.B * p.tiles_h * p.tiles_w;
    // small channel tiles leave room for two CTAs per SM (63 KB / 32 KB of shared memory, few registers)
    int grid_x = sms * (T == 32 ? 1 : 2) / ytiles;
    if (grid_x < 1) grid_x = 1;
    int Dc = p.D;
    while (Dc > 4 && cols * ((p.D + Dc - 1) / Dc) < 2 * (int64_t)grid_x) Dc = (Dc + 1) / 2;
    p.Dc = Dc; p.dchunks = (p.D + Dc - 1) / Dc;
    const int64_t total = cols * p.dchunks;
    LEA_CHECK(total < (1ll << 31), "conv3d_wgrad_tc: too many work items");
    p.total_items = (int)total;
    if (grid_x > p.total_items) grid_x = p.total_items;
    p.dw = dw;
    const size_t smem = 1024 + (size_t)kXSlots * XB * kXBlk + (size_t)kYSlots * YB * kYBlk;
    auto kernel = lea_wgrad_mma_kernel<KS, T, T>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LEA_CHECK(e == cudaSuccess, "conv3d_wgrad_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    kernel<<<dim3((unsigned)grid_x, (unsigned)ytiles), kWgThreads, smem, (cudaStream_t)stream>>>(xmap, ymap, p);
    e = cudaGetLastError();
    LEA_CHECK(e == cudaSuccess, "conv3d_wgrad_tc: launch failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace

// 1 when lea_conv3d_wgrad_tc takes this shape (k = 3, two planes, channel counts that tile by 8, 16 or 32).
extern "C" int lea_conv3d_wgrad_tc_supported(int32_t c_in, int32_t c_out, int32_t ksize, int32_t planes) {
    if ((ksize != 3 && ksize != 1) || planes != 2 || c_in < 8 || c_out < 8 || (c_in & 7) || (c_out & 7)) return 0;
    if (c_in % 32 == 0 && c_out % 32 == 0) return 1;
    if (c_in % 16 == 0 && c_out % 16 == 0) return (c_in / 16) * (c_out / 16) <= 64;
    return (c_in / 8) * (c_out / 8) <= 64;
}

extern "C" int lea_conv3d_wgrad_tc(const lea_vol* in, int32_t in_c0, int32_t c_in, const lea_vol* dout, int32_t dout_c0,
                                   int32_t c_out, int32_t ksize, float* dw, void* stream) {
    LEA_CHECK(in && dout && dw && in->data && dout->data, "conv3d_wgrad_tc: null argument");
    LEA_CHECK(lea_conv3d_wgrad_tc_supported(c_in, c_out, ksize, in->P) && dout->P == 2,
              "conv3d_wgrad_tc: shape c_in=%d c_out=%d k=%d planes=%d is not taken (use lea_conv3d_wgrad)", c_in, c_out, ksize, in->P);
    LEA_CHECK(in->B == dout->B && in->D == dout->D && in->H == dout->H && in->W == dout->W, "conv3d_wgrad_tc: shapes differ");
    LEA_CHECK((in_c0 & 7) == 0 && (dout_c0 & 7) == 0 && in_c0 + c_in <= in->C && dout_c0 + c_out <= dout->C,
              "conv3d_wgrad_tc: bad channel slice");
    if (ksize == 3) {
        if (c_in % 32 == 0 && c_out % 32 == 0) return wg_launch<3, 32>(in, in_c0, c_in, dout, dout_c0, c_out, dw, stream);
        if (c_in % 16 == 0 && c_out % 16 == 0) return wg_launch<3, 16>(in, in_c0, c_in, dout, dout_c0, c_out, dw, stream);
        return wg_launch<3, 8>(in, in_c0, c_in, dout, dout_c0, c_out, dw, stream);
    }
    if (c_in % 32 == 0 && c_out % 32 == 0) return wg_launch<1, 32>(in, in_c0, c_in, dout, dout_c0, c_out, dw, stream);
    if (c_in % 16 == 0 && c_out % 16 == 0) return wg_launch<1, 16>(in, in_c0, c_in, dout, dout_c0, c_out, dw, stream);
    return wg_launch<1, 8>(in, in_c0, c_in, dout, dout_c0, c_out, dw, stream);
}
