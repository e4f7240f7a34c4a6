// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d (all bias=False) of
//   networks/generator.py:8-20 (ResidualBlock 3x3), :80-95 (7x7 stem, 3x3 stride-2 encoders),
//   :107-124 (ConvTranspose 3x3 s2 p1 op1 decoders, 3x3 "skippers" on cat[skip, d]).
//
// GEMM view (no im2col buffer is ever materialised):
//   D[m, n] = sum_{tap, c} A_tap[m, c] * W[tap][n][c]
//   m = one output pixel of a 16 x 8 spatial tile (M = 128 = one UMMA),  n = output channel,
//   A_tap = the same NHWC fp16 activation tensor fetched by TMA at the tap's spatial offset
//           (out-of-bounds rows/cols are zero-filled by the TMA unit = the conv padding),
//   W     = weights repacked once to [tap][Cout][Cin] fp16 (K-major rows of 128 bytes).
// Both operands land in shared memory in the canonical K-major SWIZZLE_128B layout, so each
// K = 64 stage is 4 x tcgen05.mma (M128 x N_TILE x K16) issued by ONE thread; accumulators
// live in TMEM (double buffered: the epilogue of tile i overlaps the main loop of tile i+1).
//
// Precision: x ~= x_hi + x_lo, w ~= w_hi + w_lo in fp16; with SPLIT the accumulator receives
// x_hi*w_hi + x_hi*w_lo + x_lo*w_hi (fp32 accumulate), which reproduces the fp32 reference
// convolution to ~1e-5 (single pass fp16 cannot meet the 1e-3 parity bar).
// In f8 mode (ConvParams::f8, the default of the host side) the two small products are issued as ONE
// kind::f8f6f4 MMA per K step on e4m3 operand pairs stored in the lo buffers (elementwise.cu, kF8*).
//
// Kernels: k_conv_tc2 (cta_group::2, a CTA pair shares every weight tile; the default), k_conv_tc (one CTA
// per tile, optional weight multicast), k_conv_halo (activation halo + shifted descriptors; off by default).
//
// Variants, all through the same kernel:
//   stride 2      : four parity views of the input (plain strided tensor maps), tap -> view
//   transposed    : four sub-pixel output phases, each a 1/2/2/4-tap stride-1 conv
//   concat input  : K chunks 0..chunks0-1 from tensor 0, the rest from tensor 1 (torch.cat free)
//   7x7 stem      : "row-K" trick -- with 8 channels per pixel, 8 consecutive pixels of a padded
//                   NHWC8 row are 64 contiguous fp16, so one K = 64 stage covers a whole filter
//                   row (overlapping-stride tensor map); 7 stages instead of 49.
// Epilogue: TMEM -> registers -> fp32 NHWC global (128 B per thread) + per-(n, c) sum / sum of
// squares for the InstanceNorm that follows every conv (warp butterfly -> smem -> one f64 atomic
// per column per tile).
#include <cuda.h>

#include <stdlib.h>

#include <new>

#include <cuda_fp8.h>

#include "common.cuh"
#include "sample.cuh"

namespace {

constexpr int TILE_H = 16, TILE_W = 8;          // output pixels per tile (M = 128)
constexpr int KCHUNK = 64;                       // fp16 elements per K stage (128 B swizzle span)
constexpr int A_BYTES = 128 * 128;
constexpr int MAX_TAPS = 49;
constexpr int NUM_THREADS = 192;                 // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

// InstanceNorm (+ReLU, +residual, +LWB warp-add) fused into the conv epilogue (k_conv_tc2<.., FUSED = true>): the operands
// of the NEXT layer leave the kernel directly, no fp32 raw tensor and no second pass over HBM.  See epilogue_fused_loop.
struct FusedNorm {
    const float* gamma; const float* beta; float eps; int relu;
    const float* residual;                                // [n,h,w,c] fp32, nullable
    const float* warp_src; int src_batch; const float* T; int th, tw, align_corners;      // nullable LWB source (NHWC fp32)
    float* y_f32; __half* y_hi; uint8_t* y_lo; int lo_format;
    int* range_flag;
    int* counters;                                        // [n_img * n_tiles_n], zero before every run: tiles of a unit done
    int tiles_per_image;
    double inv_hw;
};

struct ConvParams {
    FusedNorm fn;
    CUtensorMap a_hi[4];
    CUtensorMap a_lo[4];
    CUtensorMap w_hi;
    CUtensorMap w_lo;
    int n_img, tiles_y, tiles_x, n_tiles_n;
    int dom_h, dom_w;                 // extent of the tile domain (output grid, or phase grid)
    int ntaps, chunks0, chunks1;
    signed char dy[MAX_TAPS], dx[MAX_TAPS], tmap[MAX_TAPS];
    short wtap[MAX_TAPS];
    float* out; int out_h, out_w, cout;
    int oy_mul, oy_add, ox_mul, ox_add;
    double* stats;
    int stages;                       // pipeline depth actually used (<= Cfg::STAGES; LWB_STAGES, diagnostic)
    int f8;                           // SPLIT stages hold [A_hi | A_lo8 | B_hi | B_lo8]: 1 f16 + 1 f8f6f4 MMA per K step
    float out_scale;                  // accumulator -> output (2^-w_exp in f8 mode, else 1)
    // y-halo schedule (k_conv_tc2y): taps re-ordered into groups that share the input view, dx and the channel chunk and whose
    // dy are consecutive; the group's activation tile (TILE_H + cnt - 1 rows) is fetched ONCE and tap i reads it i rows down
    int ngroups; unsigned char g_first[MAX_TAPS], g_cnt[MAX_TAPS];
    int a_rows;                       // rows of the activation box = TILE_H + max group size - 1
    int na_stages, nb_stages;         // activation / weight ring depths
    int phase_cols;                   // > 0: merged transposed conv -- column block col / phase_cols = sub-pixel phase (a, b) =
                                      // (ph >> 1, ph & 1) of output pixel (2y + a, 2x + b), channel = col % phase_cols
};

template <int N_TILE, bool SPLIT, int KC = KCHUNK>
struct Cfg {
    static constexpr int ROW_BYTES = KC * 2;                 // one operand row of a stage (128 B or 64 B)
    static constexpr int A_TILE = 128 * ROW_BYTES;
    static constexpr int B_BYTES = N_TILE * ROW_BYTES;
    static constexpr int STAGE_BYTES = (A_TILE + B_BYTES) * (SPLIT ? 2 : 1);
    static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    // NACC = 3 gives each of the three split products its own TMEM accumulator (summed in the epilogue).  Measured
    // on B200: no gain (0.528 vs 0.518 ms on the 64-channel skipper) -- the ~50-cycle per-instruction overhead of a
    // tcgen05.mma is not an accumulator dependency -- so one accumulator is used.  (Kept: 6*N_TILE <= 512 allows 3.)
    static constexpr int NACC = 1;
    static constexpr int ACC_COLS = 2 * NACC * N_TILE;
    static constexpr int TMEM_COLS = (ACC_COLS <= 32) ? 32 : (ACC_COLS <= 64) ? 64 : (ACC_COLS <= 128) ? 128
                                   : (ACC_COLS <= 256) ? 256 : 512;
    static constexpr int BAR_BYTES = 256;
    static constexpr int STATS_BYTES = 4 * N_TILE * 2 * 4;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES + STATS_BYTES;
    static_assert(STAGES >= 2, "need at least two pipeline stages");
    static_assert(ACC_COLS <= 512, "accumulator double buffer exceeds TMEM");
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            printf("lwb conv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
// One lane of a CONVERGED warp.  Issuing TMA / tcgen05 under `if (elect_one())` (instead of `if (lane == 0)`)
// keeps the surrounding control flow warp-uniform, so the compiler holds descriptors, coordinates and TMEM
// addresses in uniform registers; under a lane-id branch every UTCHMMA/UTMALDG operand goes through an
// ELECT + R2UR.BROADCAST waterfall loop (~12 extra dependent instructions per MMA, measured issue-bound).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// K-major SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(=1)<<16 |
// SBO(=1024B>>4)<<32 | version(1)<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Same for a K stage of KC fp16: KC = 64 -> 128 B rows, SWIZZLE_128B (2), 8-row group = 1024 B;
//                               KC = 32 ->  64 B rows, SWIZZLE_64B  (4), 8-row group =  512 B.
template <int KC>
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)((8 * KC * 2) >> 4) << 32) | (1ull << 46)
         | ((KC == 64 ? 2ull : 4ull) << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Same instruction on 8-bit operands (E4M3 x E4M3: the format fields of the instruction descriptor are both 0, like
// F16 x F16): K = 32 per instruction, i.e. the same 32 B descriptor step, at twice the fp16 rate.
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// cluster variants: one CTA's TMA load lands in every CTA of the cluster (same smem offset) and completes
// tx bytes on each destination CTA's own mbarrier; one commit arrives on every CTA's mbarrier.
__device__ __forceinline__ void tma_load_3d_mc(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

// Persistent tile schedule shared by the three warp roles.  A cluster of CL CTAs takes CL consecutive
// M tiles of ONE N tile per step ("super tile"), so that the weight tile is common to the cluster.
struct Sched {
    int first, step, total, msup, cl, rank;
    __device__ __forceinline__ void decode(int sup, int& n_idx, int& m_idx) const { n_idx = sup / msup; m_idx = (sup % msup) * cl + rank; }
};
__device__ __forceinline__ Sched make_sched(int m_tiles, int n_tiles_n, int cl) {
    Sched s;
    s.cl = cl; s.rank = cl > 1 ? (int)cluster_ctarank() : 0;
    s.msup = m_tiles / cl; s.total = s.msup * n_tiles_n;
    s.first = (int)blockIdx.x / cl; s.step = (int)gridDim.x / cl;
    return s;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Column sums over the 32 lanes of a warp for NV per-lane values: after the butterfly, lane L
// (L < NV) holds sum over lanes of v[L].  NV = 32: 31 shuffles; NV = 16: 16 + 15 shuffles.
template <int NV>
__device__ __forceinline__ float warp_col_sums(float* v, unsigned lane) {
    if (NV == 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
    }
#pragma unroll
    for (int off = (NV == 32 ? 16 : 8), cnt = (NV == 32 ? 16 : 8); off >= 1; off >>= 1, cnt >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < cnt; j++) {
            const float send = up ? v[j] : v[j + cnt];
            const float keep = up ? v[j + cnt] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}


__device__ __forceinline__ uint32_t mapa_cta0(uint32_t local_addr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local_addr));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}

// Epilogue warps (4): TMEM accumulator -> fp32 NHWC global + InstanceNorm partial statistics.
template <int N_TILE, int NACC, bool TWO_SM = false, class PT>
__device__ __forceinline__ void epilogue_loop(const PT& P, int warp, unsigned lane, const Sched& sch,
                                              uint32_t tmem_base, uint64_t* bar_tfull, uint64_t* bar_tempty, float2* s_stats)
{
    const int q = warp & 3;                                     // TMEM lane quarter this warp may access
    const int row = q * 32 + (int)lane;
    const int ty = row >> 3, tx = row & 7;
    const int et = threadIdx.x - 64;                            // 0..127
    int abuf = 0; uint32_t aphase = 0;
    for (int sup = sch.first; sup < sch.total; sup += sch.step) {
        int n_idx, m_idx;
        sch.decode(sup, n_idx, m_idx);
        const int img = m_idx / (P.tiles_y * P.tiles_x);
        const int rem = m_idx % (P.tiles_y * P.tiles_x);
        const int y = (rem / P.tiles_x) * TILE_H + ty, x = (rem % P.tiles_x) * TILE_W + tx;
        const bool valid = y < P.dom_h && x < P.dom_w;
        float* optr = P.out + (((size_t)img * P.out_h + (P.oy_mul * y + P.oy_add)) * P.out_w + (P.ox_mul * x + P.ox_add)) * P.cout
                    + (size_t)n_idx * N_TILE;
        const int out_w_c = P.out_w * P.cout;                   // one output row, in floats (merged transposed conv)
        mbar_wait(bar_tfull + abuf, aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(abuf * NACC * N_TILE);
        constexpr int CW = N_TILE >= 32 ? 32 : 16;
#pragma unroll 1
        for (int c = 0; c < (P.oy_mul < 0 ? 0 : N_TILE / CW); c++) {      // oy_mul < 0: debug, skip the epilogue body
            uint32_t r[CW];
            if (CW == 32) tmem_ld32(taddr + c * CW, r); else tmem_ld16(taddr + c * CW, r);
            if (NACC > 1) {
                uint32_t r1[CW], r2[CW];
                if (CW == 32) { tmem_ld32(taddr + N_TILE + c * CW, r1); tmem_ld32(taddr + 2 * N_TILE + c * CW, r2); }
                else          { tmem_ld16(taddr + N_TILE + c * CW, r1); tmem_ld16(taddr + 2 * N_TILE + c * CW, r2); }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CW; j++)
                    r[j] = __float_as_uint(__uint_as_float(r[j]) + (__uint_as_float(r1[j]) + __uint_as_float(r2[j])));
            }
            tmem_ld_wait();
            if (P.out_scale != 1.f) {
#pragma unroll
                for (int j = 0; j < CW; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) * P.out_scale);
            }
            if (valid && P.out) {
                float* o = optr + c * CW;
                if (P.phase_cols > 0) {
                    const int col = n_idx * N_TILE + c * CW, ph = col / P.phase_cols;
                    o = P.out + (((size_t)img * P.out_h + 2 * y) * P.out_w + 2 * x) * P.cout
                      + (size_t)(ph >> 1) * out_w_c + (ph & 1) * P.cout + (col - ph * P.phase_cols);
                }
                // 32 bytes (one full sector) per store: rows of a warp are different pixels, so nothing coalesces across lanes
#pragma unroll
                for (int j = 0; j < CW / 8; j++) lwb::stg_f32x8(o + 8 * j, reinterpret_cast<const float*>(r + 8 * j));
            }
            if (P.stats) {
                float v[CW], v2[CW];
#pragma unroll
                for (int j = 0; j < CW; j++) { const float t = valid ? __uint_as_float(r[j]) : 0.f; v[j] = t; v2[j] = t * t; }
                const float s1 = warp_col_sums<CW>(v, lane);
                const float s2 = warp_col_sums<CW>(v2, lane);
                if ((int)lane < CW) s_stats[q * N_TILE + c * CW + lane] = make_float2(s1, s2);
            }
        }
        // all TMEM reads of this buffer are complete: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (TWO_SM) mbar_arrive_cluster(mapa_cta0(smem_u32(bar_tempty + abuf)));     // the leader's barrier collects both CTAs
            else        mbar_arrive(bar_tempty + abuf);
        }
        if (P.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int col = et; col < N_TILE; col += 128) {
                const float2 a = s_stats[col], b = s_stats[N_TILE + col], cc = s_stats[2 * N_TILE + col], d = s_stats[3 * N_TILE + col];
                int ch = n_idx * N_TILE + col;
                if (P.phase_cols > 0) ch %= P.phase_cols;      // the four phases of a channel share its statistics
                double* dst = P.stats + 2 * ((size_t)img * P.cout + ch);
                atomicAdd(dst, (double)((a.x + b.x) + (cc.x + d.x)));
                atomicAdd(dst + 1, (double)((a.y + b.y) + (cc.y + d.y)));
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (++abuf == 2) { abuf = 0; aphase ^= 1; }
    }
}


__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint8_t cvt_e4m3(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }

// Fused epilogue (cta_group::2 kernel only).  A "unit" = (image, N tile): its InstanceNorm statistics need every M tile of
// the image.  Each CTA (a) reduces its tile's per-channel sum / sum of squares out of TMEM and adds them to the global f64
// statistics, (b) publishes "tile done" on the unit's counter and waits until all tiles_per_image tiles are in, (c) reads
// the finished statistics, and (d) normalises its tile out of TMEM a second time, applying ReLU / residual / LWB warp-add and
// writing the next layer's operands (fp16 hi + fp16 or e4m3-pair lo, optional fp32).
// Why the wait cannot deadlock: the grid is persistent with one CTA per SM (all CTAs co-resident), tiles are taken in
// rounds of gridDim.x, a unit is a run of consecutive tiles no longer than one round, so a CTA waiting on its round-r tile
// depends only on statistics of rounds <= r of other CTAs plus round r+1 tiles of LOWER-numbered pairs -- which those pairs
// reach through their own (earlier-ordered) waits; the MMA warp meanwhile fills the second TMEM accumulator.  The host
// enables this mode only when tiles_per_image / 2 <= gridDim.x / 2 and never together with multi-stream sub-batches.
template <int N_TILE>
__device__ __forceinline__ void epilogue_fused_loop(const ConvParams& P, int warp, unsigned lane, const Sched& sch,
                                                    uint32_t tmem_base, uint64_t* bar_tfull, uint64_t* bar_tempty,
                                                    float2* s_stats, float2* s_ss)
{
    const FusedNorm& F = P.fn;
    const int q = warp & 3;
    const int row = q * 32 + (int)lane;
    const int ty = row >> 3, tx = row & 7;
    const int et = threadIdx.x - 64;                            // 0..127
    constexpr int CW = 32;
    int abuf = 0; uint32_t aphase = 0;
    for (int sup = sch.first; sup < sch.total; sup += sch.step) {
        int n_idx, m_idx;
        sch.decode(sup, n_idx, m_idx);
        const int img = m_idx / (P.tiles_y * P.tiles_x);
        const int rem = m_idx % (P.tiles_y * P.tiles_x);
        const int y = (rem / P.tiles_x) * TILE_H + ty, x = (rem % P.tiles_x) * TILE_W + tx;
        const bool valid = y < P.dom_h && x < P.dom_w;
        mbar_wait(bar_tfull + abuf, aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(abuf * N_TILE);
        // ---- pass 1: per-channel sums of this tile -> global statistics
#pragma unroll 1
        for (int c = 0; c < N_TILE / CW; c++) {
            uint32_t r[CW];
            tmem_ld32(taddr + c * CW, r);
            tmem_ld_wait();
            float v[CW], v2[CW];
#pragma unroll
            for (int j = 0; j < CW; j++) { const float t = valid ? __uint_as_float(r[j]) * P.out_scale : 0.f; v[j] = t; v2[j] = t * t; }
            const float s1 = warp_col_sums<CW>(v, lane);
            const float s2 = warp_col_sums<CW>(v2, lane);
            s_stats[q * N_TILE + c * CW + lane] = make_float2(s1, s2);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int col = et; col < N_TILE; col += 128) {
            const float2 a = s_stats[col], b = s_stats[N_TILE + col], cc = s_stats[2 * N_TILE + col], d = s_stats[3 * N_TILE + col];
            double* dst = P.stats + 2 * ((size_t)img * P.cout + (size_t)n_idx * N_TILE + col);
            atomicAdd(dst, (double)((a.x + b.x) + (cc.x + d.x)));
            atomicAdd(dst + 1, (double)((a.y + b.y) + (cc.y + d.y)));
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // ---- publish + wait for the unit
        if (et == 0) {
            int* ctr = F.counters + img * P.n_tiles_n + n_idx;
            atomicAdd(ctr, 1);
            if (ld_acquire_gpu(ctr) < F.tiles_per_image) {
                const long long t0 = clock64();
                while (ld_acquire_gpu(ctr) < F.tiles_per_image) {
                    __nanosleep(64);
                    if (clock64() - t0 > 4000000000ll) {
                        printf("lwb conv_tc: fused InstanceNorm wait timed out (block %d, unit %d/%d, count %d of %d)\n",
                               blockIdx.x, img, n_idx, ld_acquire_gpu(ctr), F.tiles_per_image);
                        __trap();
                    }
                }
            }
        }
        __syncwarp();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int col = et; col < N_TILE; col += 128) {
            const int ch = n_idx * N_TILE + col;
            const double* sp = P.stats + 2 * ((size_t)img * P.cout + ch);
            const double mean = __ldcg(sp) * F.inv_hw;
            double var = __ldcg(sp + 1) * F.inv_hw - mean * mean;
            if (var < 0) var = 0;
            const float rstd = (float)(1.0 / sqrt(var + (double)F.eps));
            const float g = F.gamma ? __ldg(F.gamma + ch) : 1.f, bt = F.beta ? __ldg(F.beta + ch) : 0.f;
            s_ss[col] = make_float2(g * rstd, bt - (float)mean * g * rstd);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // ---- pass 2: normalise this thread's pixel, emit operands
        const size_t pix = ((size_t)img * P.out_h + y) * P.out_w + x;          // fused plans have oy_mul = ox_mul = 1
        lwb::Taps tp;
        tp.m = 0;
        if (F.warp_src && valid) {
            float gx, gy;
            lwb::flow_at(F.T + (size_t)img * F.th * F.tw * 2, F.th, F.tw, P.out_h, P.out_w, y, x, gx, gy);
            lwb::make_taps(gx, gy, P.out_h, P.out_w, F.align_corners, tp);
        }
        const float* wsrc = F.warp_src ? F.warp_src + (size_t)(F.src_batch == 1 ? 0 : img) * P.out_h * P.out_w * P.cout : nullptr;
        unsigned hmax = 0;
#pragma unroll 1
        for (int c = 0; c < N_TILE / CW; c++) {
            uint32_t r[CW];
            __syncwarp();                                       // tcgen05.ld is warp-collective: reconverge after the per-pixel branches
            tmem_ld32(taddr + c * CW, r);
            tmem_ld_wait();
            if (!valid) continue;
            const int ch0 = n_idx * N_TILE + c * CW;
            float v[CW];
#pragma unroll
            for (int j = 0; j < CW; j++) {
                const float2 ss = s_ss[c * CW + j];
                float t = fmaf(__uint_as_float(r[j]) * P.out_scale, ss.x, ss.y);
                v[j] = F.relu ? fmaxf(t, 0.f) : t;
            }
            if (F.residual) {
                const float4* rp = reinterpret_cast<const float4*>(F.residual + pix * P.cout + ch0);
#pragma unroll
                for (int j = 0; j < CW / 4; j++) {
                    const float4 t = __ldg(rp + j);
                    v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                }
            }
            if (tp.m) {
                const int offs[4] = {tp.o00, tp.o00 + 1, tp.o00 + P.out_w, tp.o00 + P.out_w + 1};
                const float wt[4] = {tp.w00, tp.w01, tp.w10, tp.w11};
                float acc[CW];
#pragma unroll
                for (int j = 0; j < CW; j++) acc[j] = 0.f;
#pragma unroll
                for (int t4 = 0; t4 < 4; t4++) {
                    if (tp.m & (1 << t4)) {
                        const float4* sp4 = reinterpret_cast<const float4*>(wsrc + (size_t)offs[t4] * P.cout + ch0);
#pragma unroll
                        for (int j = 0; j < CW / 4; j++) {
                            const float4 t = __ldg(sp4 + j);
                            acc[4 * j] += t.x * wt[t4]; acc[4 * j + 1] += t.y * wt[t4];
                            acc[4 * j + 2] += t.z * wt[t4]; acc[4 * j + 3] += t.w * wt[t4];
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < CW; j++) v[j] += acc[j];
            }
            if (F.y_f32) {
                float4* o = reinterpret_cast<float4*>(F.y_f32 + pix * P.cout + ch0);
#pragma unroll
                for (int j = 0; j < CW / 4; j++) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (F.y_hi) {
                __align__(16) __half hh[CW];
#pragma unroll
                for (int j = 0; j < CW; j++) hh[j] = __float2half_rn(v[j]);
                const uint4* hv = reinterpret_cast<const uint4*>(hh);
                uint4* oh = reinterpret_cast<uint4*>(F.y_hi + pix * P.cout + ch0);
#pragma unroll
                for (int j = 0; j < CW / 8; j++) {
                    oh[j] = hv[j];
                    hmax = __vmaxu2(hmax, __vmaxu2(__vmaxu2(hv[j].x & 0x7fff7fffu, hv[j].y & 0x7fff7fffu),
                                                   __vmaxu2(hv[j].z & 0x7fff7fffu, hv[j].w & 0x7fff7fffu)));
                }
                if (F.y_lo && F.lo_format == 0) {
                    __align__(16) __half ll[CW];
#pragma unroll
                    for (int j = 0; j < CW; j++) ll[j] = __float2half_rn(v[j] - __half2float(hh[j]));
                    uint4* ol = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(F.y_lo) + pix * P.cout + ch0);
#pragma unroll
                    for (int j = 0; j < CW / 8; j++) ol[j] = reinterpret_cast<const uint4*>(ll)[j];
                } else if (F.y_lo) {
                    // e4m3 pair block of this pixel's 64-channel group: bytes [ch % 64] = e4m3(x / 16), [64 + ch % 64] = e4m3(x_lo * 2^10)
                    __align__(16) uint8_t x8[CW];
                    __align__(16) uint8_t l8[CW];
#pragma unroll
                    for (int j = 0; j < CW; j++) {
                        x8[j] = cvt_e4m3(v[j] * (1.f / 16.f));
                        l8[j] = cvt_e4m3((v[j] - __half2float(hh[j])) * 1024.f);
                    }
                    uint8_t* blk = F.y_lo + (pix * P.cout + (size_t)(ch0 & ~63)) * 2 + (ch0 & 63);
                    reinterpret_cast<uint4*>(blk)[0] = reinterpret_cast<const uint4*>(x8)[0];
                    reinterpret_cast<uint4*>(blk)[1] = reinterpret_cast<const uint4*>(x8)[1];
                    reinterpret_cast<uint4*>(blk + 64)[0] = reinterpret_cast<const uint4*>(l8)[0];
                    reinterpret_cast<uint4*>(blk + 64)[1] = reinterpret_cast<const uint4*>(l8)[1];
                }
            }
        }
        if (F.range_flag) {
            const unsigned m = max(hmax & 0xffffu, hmax >> 16);
            if (m >= 0x6400u) atomicOr(F.range_flag, m >= 0x7b53u ? 3 : 1);
        }
        // all TMEM reads of this buffer are complete: hand it back to the MMA warp (the leader's barrier collects both CTAs)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_cta0(smem_u32(bar_tempty + abuf)));
        if (++abuf == 2) { abuf = 0; aphase ^= 1; }
    }
}

// ----------------------------------------------------------------------------------- kernel
template <int N_TILE, bool SPLIT, int CL, int KC>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_tc(const __grid_constant__ ConvParams P)
{
    using C = Cfg<N_TILE, SPLIT, KC>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* bar_empty = bar_full + C::STAGES;
    uint64_t* bar_tfull = bar_empty + C::STAGES;
    uint64_t* bar_tempty = bar_tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
    float2* s_stats = reinterpret_cast<float2*>(smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES);   // [4][N_TILE]

    const int warp = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;

    if (warp == 1) {
        if (lane == 0) {
            // empty[s] collects one commit from EVERY CTA of the cluster: a stage may be refilled (by
            // any CTA's multicast) only when all of them have finished reading it.
            for (int s = 0; s < C::STAGES; s++) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, CL); }
            for (int b = 0; b < 2; b++) { mbar_init(bar_tfull + b, 1); mbar_init(bar_tempty + b, 4); }
            fence_barrier_init();
            fence_proxy_async();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();              // peers' barriers are initialised before any multicast can land
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    lwb::pdl_wait();                             // see k_conv_tc2: global memory only after the preceding kernel completed
    lwb::pdl_trigger();

    const int nchunks = P.chunks0 + P.chunks1;
    const int ksteps = P.ntaps * nchunks;
    const int nstages = P.stages;
    const int m_tiles = P.n_img * P.tiles_y * P.tiles_x;
    const Sched sch = make_sched(m_tiles, P.n_tiles_n, CL);
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    constexpr int B_SLICE = N_TILE / CL;         // weight rows this CTA fetches (and multicasts) per stage

    if (warp == 0) {
        // ================================ TMA producer (whole warp, one elected lane issues) ==========
        {
            if (elect_one()) {
                asm volatile("prefetch.tensormap [%0];" :: "l"(&P.a_hi[0]) : "memory");
                asm volatile("prefetch.tensormap [%0];" :: "l"(&P.w_hi) : "memory");
                if (SPLIT) {
                    asm volatile("prefetch.tensormap [%0];" :: "l"(&P.a_lo[0]) : "memory");
                    asm volatile("prefetch.tensormap [%0];" :: "l"(&P.w_lo) : "memory");
                }
            }
            int stage = 0; uint32_t phase = 0;
            for (int sup = sch.first; sup < sch.total; sup += sch.step) {
                int n_idx, m_idx;
                sch.decode(sup, n_idx, m_idx);
                const int img = m_idx / (P.tiles_y * P.tiles_x);
                const int rem = m_idx % (P.tiles_y * P.tiles_x);
                const int y0 = (rem / P.tiles_x) * TILE_H, x0 = (rem % P.tiles_x) * TILE_W;
                for (int s = 0; s < ksteps; s++) {
                    mbar_wait(bar_empty + stage, phase ^ 1);
                    uint8_t* st = smem + stage * C::STAGE_BYTES;
                    const int tap = s / nchunks, chunk = s % nchunks;
                    const bool second = chunk >= P.chunks0;
                    const int mi = second ? 1 : P.tmap[tap];
                    const int c0 = (second ? chunk - P.chunks0 : chunk) * KC;
                    const int xx = x0 + P.dx[tap], yy = y0 + P.dy[tap];
                    const int wt = P.wtap[tap];
                    uint8_t* sb = st + C::A_TILE * (SPLIT ? 2 : 1);
                    if (elect_one()) {
                        mbar_expect_tx(bar_full + stage, (uint32_t)C::STAGE_BYTES);
                        tma_load_4d(&P.a_hi[mi], st, bar_full + stage, c0, xx, yy, img);
                        if (SPLIT) tma_load_4d(&P.a_lo[mi], st + C::A_TILE, bar_full + stage, c0, xx, yy, img);
                        if (CL == 1) {
                            tma_load_3d(&P.w_hi, sb, bar_full + stage, chunk * KC, n_idx * N_TILE, wt);
                            if (SPLIT) tma_load_3d(&P.w_lo, sb + C::B_BYTES, bar_full + stage, chunk * KC, n_idx * N_TILE, wt);
                        } else {
                            const int r0 = sch.rank * B_SLICE;
                            tma_load_3d_mc(&P.w_hi, sb + r0 * C::ROW_BYTES, bar_full + stage, chunk * KC, n_idx * N_TILE + r0, wt, kMask);
                            if (SPLIT) tma_load_3d_mc(&P.w_lo, sb + C::B_BYTES + r0 * C::ROW_BYTES, bar_full + stage, chunk * KC, n_idx * N_TILE + r0, wt, kMask);
                        }
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (whole warp, one elected lane issues) ============
        {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, K-major both, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t smem_base = smem_u32(smem);
            int stage = 0; uint32_t phase = 0;
            int abuf = 0; uint32_t aphase = 0;
            for (int sup = sch.first; sup < sch.total; sup += sch.step) {
                mbar_wait(bar_tempty + abuf, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(abuf * C::NACC * N_TILE);
                for (int s = 0; s < ksteps; s++) {
                    mbar_wait(bar_full + stage, phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t a_lo = a_hi + C::A_TILE;
                    const uint32_t b_hi = a_hi + C::A_TILE * (SPLIT ? 2 : 1);
                    const uint32_t b_lo = b_hi + C::B_BYTES;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < KC / 16; k++) {
                            const uint64_t da = make_desc_k<KC>(a_hi + k * 32), db = make_desc_k<KC>(b_hi + k * 32);
                            const uint32_t first = (s > 0 || k > 0) ? 1u : 0u;
                            umma_f16(d_tmem, da, db, idesc, first);
                            if (SPLIT && P.f8) {
                                umma_f8(d_tmem, make_desc_k<KC>(a_lo + k * 32), make_desc_k<KC>(b_lo + k * 32), idesc, 1u);
                            } else if (SPLIT) {
                                constexpr uint32_t A1 = C::NACC > 1 ? N_TILE : 0, A2 = C::NACC > 1 ? 2 * N_TILE : 0;
                                umma_f16(d_tmem + A1, da, make_desc_k<KC>(b_lo + k * 32), idesc, C::NACC > 1 ? first : 1u);
                                umma_f16(d_tmem + A2, make_desc_k<KC>(a_lo + k * 32), db, idesc, C::NACC > 1 ? first : 1u);
                            }
                        }
                        // smem slot free (in every CTA of the cluster) once these MMAs retire
                        if (CL == 1) umma_commit(bar_empty + stage); else umma_commit_mc(bar_empty + stage, kMask);
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit(bar_tfull + abuf);     // accumulator complete -> epilogue
                __syncwarp();
                if (++abuf == 2) { abuf = 0; aphase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue (4 warps) ===========================
        epilogue_loop<N_TILE, C::NACC>(P, warp, lane, sch, tmem_base, bar_tfull, bar_tempty, s_stats);
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();              // nobody exits while a peer may still multicast into it
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// =====================================================================================================
// 2-CTA variant (tcgen05 cta_group::2): a CTA pair (cluster 2x1x1, one TPC) computes TWO M tiles of the
// same N tile with ONE M256 x N x K16 instruction issued by the leader CTA.  Each CTA stages its own
// activation tile and only HALF of the weight tile (N/2 rows); the tensor cores of the pair read the
// other half from the peer's shared memory.  Per CTA and stage this halves the weight bytes written by
// TMA and read by the MMAs -- the kernel is shared-memory-bandwidth bound -- and
// frees shared memory for a third stage at N = 256 (64 KB instead of 96 KB per stage).
//   full[s]   lives in the leader; both CTAs' TMA loads (cta_group::2) complete their bytes on it
//   empty[s]  one per CTA; the leader's commit is multicast to both
//   tfull[b]  one per CTA (multicast commit); tempty[b] lives in the leader, 4 warps x 2 CTAs arrive on it
// =====================================================================================================
template <int N_TILE, bool SPLIT>
struct Cfg2 {
    static constexpr int B_HALF = (N_TILE / 2) * 128;
    static constexpr int STAGE_BYTES = (A_BYTES + B_HALF) * (SPLIT ? 2 : 1);       // per CTA
    static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int TMEM_COLS = Cfg<N_TILE, SPLIT>::TMEM_COLS;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES + 4 * N_TILE * 2 * 4 + N_TILE * 8;
};

__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, void* dst, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* map, void* dst, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f8_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
template <int N_TILE, bool SPLIT, bool FUSED>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_tc2(const __grid_constant__ ConvParams P)
{
    using C = Cfg2<N_TILE, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* bar_empty = bar_full + C::STAGES;
    uint64_t* bar_tfull = bar_empty + C::STAGES;
    uint64_t* bar_tempty = bar_tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
    float2* s_stats = reinterpret_cast<float2*>(smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES);
    float2* s_ss = s_stats + 4 * N_TILE;                       // [N_TILE] (scale, shift), fused mode

    const int warp = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < C::STAGES; s++) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
            for (int b = 0; b < 2; b++) { mbar_init(bar_tfull + b, 1); mbar_init(bar_tempty + b, 8); }
            fence_barrier_init();
            fence_proxy_async();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above touched only this CTA's shared memory / TMEM; global memory (operands, output, statistics) is
    // produced / still read by the preceding kernel of the stream: wait for it (no-op without a programmatic launch)
    lwb::pdl_wait();
    lwb::pdl_trigger();                                  // the statistics-finalize kernel may be scheduled early (it waits too)

    const int nchunks = P.chunks0 + P.chunks1;
    const int ksteps = P.ntaps * nchunks;
    const int m_tiles = P.n_img * P.tiles_y * P.tiles_x;
    const Sched sch = make_sched(m_tiles, P.n_tiles_n, 2);

    if (warp == 0) {
        // ================================ TMA producer (both CTAs) ======================
        int stage = 0; uint32_t phase = 0;
        for (int sup = sch.first; sup < sch.total; sup += sch.step) {
            int n_idx, m_idx;
            sch.decode(sup, n_idx, m_idx);
            const int img = m_idx / (P.tiles_y * P.tiles_x);
            const int rem = m_idx % (P.tiles_y * P.tiles_x);
            const int y0 = (rem / P.tiles_x) * TILE_H, x0 = (rem % P.tiles_x) * TILE_W;
            for (int s = 0; s < ksteps; s++) {
                mbar_wait(bar_empty + stage, phase ^ 1);
                uint8_t* st = smem + stage * C::STAGE_BYTES;
                const int tap = s / nchunks, chunk = s % nchunks;
                const bool second = chunk >= P.chunks0;
                const int mi = second ? 1 : P.tmap[tap];
                const int c0 = (second ? chunk - P.chunks0 : chunk) * KCHUNK;
                const int xx = x0 + P.dx[tap], yy = y0 + P.dy[tap];
                const int wt = P.wtap[tap];
                const int nrow = n_idx * N_TILE + (int)rank * (N_TILE / 2);
                uint8_t* sb = st + A_BYTES * (SPLIT ? 2 : 1);
                const uint32_t fullc = mapa_cta0(smem_u32(bar_full + stage));       // the leader's barrier
                if (elect_one()) {
                    if (leader) mbar_expect_tx(bar_full + stage, 2u * (uint32_t)C::STAGE_BYTES);   // both CTAs' bytes
                    tma_load_4d_2sm(&P.a_hi[mi], st, fullc, c0, xx, yy, img);
                    if (SPLIT) tma_load_4d_2sm(&P.a_lo[mi], st + A_BYTES, fullc, c0, xx, yy, img);
                    tma_load_3d_2sm(&P.w_hi, sb, fullc, chunk * KCHUNK, nrow, wt);
                    if (SPLIT) tma_load_3d_2sm(&P.w_lo, sb + C::B_HALF, fullc, chunk * KCHUNK, nrow, wt);
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only) ==================
        if (leader) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t smem_base = smem_u32(smem);
            int stage = 0; uint32_t phase = 0;
            int abuf = 0; uint32_t aphase = 0;
            for (int sup = sch.first; sup < sch.total; sup += sch.step) {
                mbar_wait(bar_tempty + abuf, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(abuf * N_TILE);
                for (int s = 0; s < ksteps; s++) {
                    mbar_wait(bar_full + stage, phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t a_lo = a_hi + A_BYTES;
                    const uint32_t b_hi = a_hi + A_BYTES * (SPLIT ? 2 : 1);
                    const uint32_t b_lo = b_hi + C::B_HALF;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < KCHUNK / 16; k++) {
                            const uint64_t da = make_desc(a_hi + k * 32), db = make_desc(b_hi + k * 32);
                            umma_f16_2sm(d_tmem, da, db, idesc, (s > 0 || k > 0) ? 1u : 0u);
                            if (SPLIT && P.f8) {
                                umma_f8_2sm(d_tmem, make_desc(a_lo + k * 32), make_desc(b_lo + k * 32), idesc, 1u);
                            } else if (SPLIT) {
                                umma_f16_2sm(d_tmem, da, make_desc(b_lo + k * 32), idesc, 1u);
                                umma_f16_2sm(d_tmem, make_desc(a_lo + k * 32), db, idesc, 1u);
                            }
                        }
                        umma_commit_2sm(bar_empty + stage, 3);          // stage free in BOTH CTAs
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit_2sm(bar_tfull + abuf, 3);  // accumulators complete in both CTAs
                __syncwarp();
                if (++abuf == 2) { abuf = 0; aphase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue (both CTAs, own TMEM rows) ===========
        if constexpr (FUSED) epilogue_fused_loop<N_TILE>(P, warp, lane, sch, tmem_base, bar_tfull, bar_tempty, s_stats, s_ss);
        else                 epilogue_loop<N_TILE, 1, true>(P, warp, lane, sch, tmem_base, bar_tfull, bar_tempty, s_stats);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// =====================================================================================================
// y-halo variant of the 2-CTA kernel (k_conv_tc2y).  The kernel above re-fetches the 16 x 8-pixel activation tile for every
// filter tap: per K step a CTA writes 32 KB (A hi + lo) + its weight half by TMA and the MMAs read ~48-96 KB -- the SM's
// 128 B/clk shared-memory port is the bound, the tensor pipe idles.  Taps that differ only in dy read
// the SAME pixels shifted by whole image rows, and one image row of the tile (8 pixels x 128 B) is exactly one 1024-byte
// SWIZZLE_128B atom: a tile of TILE_H + cnt - 1 rows fetched once serves all cnt taps of the group through the UMMA
// descriptor's start address (+ i * 1024 B; the swizzle phase lives in address bits 7-9 and does not change).  For a 3x3
// filter the activation fill drops to 18/48 of its bytes, for the 7-tap heads / stem to 22/112.
//   A ring: na_stages (2-4) entries of a_rows KB (x2 with lo);  full[s] in the leader (both CTAs' bytes), empty[s] per CTA
//   B ring: nb_stages (3-8) entries of N/2 weight rows (x2 with lo), one entry per tap
// (an activation entry now lasts cnt taps of MMA time, so the ring is as deep as shared memory allows: 2 entries at N = 256,
// 3-4 for the narrow layers, where two entries could not cover the TMA latency)
// =====================================================================================================
constexpr int Y_MAX_NA = 4;
constexpr int Y_MAX_NB = 8;

template <int N_TILE, bool SPLIT>
struct CfgY {
    static constexpr int B_HALF = (N_TILE / 2) * 128;
    static constexpr int B_STAGE = B_HALF * (SPLIT ? 2 : 1);
    static constexpr int RING_BYTES = 196 * 1024;
    static constexpr int TMEM_COLS = Cfg<N_TILE, SPLIT>::TMEM_COLS;
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM_BYTES = 1024 + RING_BYTES + BAR_BYTES + 4 * N_TILE * 2 * 4;
};

template <int N_TILE, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_tc2y(const __grid_constant__ ConvParams P)
{
    using C = CfgY<N_TILE, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_plane = P.a_rows * 1024;                        // one of hi / lo
    const int a_stage = a_plane * (SPLIT ? 2 : 1);
    const int na = P.na_stages;
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + na * a_stage;
    uint64_t* bar_afull = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES);
    uint64_t* bar_aempty = bar_afull + Y_MAX_NA;
    uint64_t* bar_bfull = bar_aempty + Y_MAX_NA;
    uint64_t* bar_bempty = bar_bfull + Y_MAX_NB;
    uint64_t* bar_tfull = bar_bempty + Y_MAX_NB;
    uint64_t* bar_tempty = bar_tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
    float2* s_stats = reinterpret_cast<float2*>(smem + C::RING_BYTES + C::BAR_BYTES);

    const int warp = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int nb = P.nb_stages;

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Y_MAX_NA; s++) { mbar_init(bar_afull + s, 1); mbar_init(bar_aempty + s, 1); }
            for (int s = 0; s < Y_MAX_NB; s++) { mbar_init(bar_bfull + s, 1); mbar_init(bar_bempty + s, 1); }
            for (int b = 0; b < 2; b++) { mbar_init(bar_tfull + b, 1); mbar_init(bar_tempty + b, 8); }
            fence_barrier_init();
            fence_proxy_async();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    lwb::pdl_wait();
    lwb::pdl_trigger();

    const int nchunks = P.chunks0 + P.chunks1;
    const int m_tiles = P.n_img * P.tiles_y * P.tiles_x;
    const Sched sch = make_sched(m_tiles, P.n_tiles_n, 2);

    if (warp == 0) {
        // ================================ TMA producer (both CTAs) ======================
        int sa = 0; uint32_t pa = 0;
        int sb = 0; uint32_t pb = 0;
        for (int sup = sch.first; sup < sch.total; sup += sch.step) {
            int n_idx, m_idx;
            sch.decode(sup, n_idx, m_idx);
            const int img = m_idx / (P.tiles_y * P.tiles_x);
            const int rem = m_idx % (P.tiles_y * P.tiles_x);
            const int y0 = (rem / P.tiles_x) * TILE_H, x0 = (rem % P.tiles_x) * TILE_W;
            const int nrow = n_idx * N_TILE + (int)rank * (N_TILE / 2);
            for (int g = 0; g < P.ngroups; g++) {
                const int t0 = P.g_first[g], cnt = P.g_cnt[g];
                for (int chunk = 0; chunk < nchunks; chunk++) {
                    const bool second = chunk >= P.chunks0;
                    const int mi = second ? 1 : P.tmap[t0];
                    const int c0 = (second ? chunk - P.chunks0 : chunk) * KCHUNK;
                    mbar_wait(bar_aempty + sa, pa ^ 1);
                    uint8_t* sta = a_ring + sa * a_stage;
                    const uint32_t afullc = mapa_cta0(smem_u32(bar_afull + sa));
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(bar_afull + sa, 2u * (uint32_t)a_stage);
                        tma_load_4d_2sm(&P.a_hi[mi], sta, afullc, c0, x0 + P.dx[t0], y0 + P.dy[t0], img);
                        if (SPLIT) tma_load_4d_2sm(&P.a_lo[mi], sta + a_plane, afullc, c0, x0 + P.dx[t0], y0 + P.dy[t0], img);
                    }
                    __syncwarp();
                    if (++sa == na) { sa = 0; pa ^= 1; }
                    for (int i = 0; i < cnt; i++) {
                        mbar_wait(bar_bempty + sb, pb ^ 1);
                        uint8_t* stb = b_ring + sb * C::B_STAGE;
                        const uint32_t bfullc = mapa_cta0(smem_u32(bar_bfull + sb));
                        const int wt = P.wtap[t0 + i];
                        if (elect_one()) {
                            if (leader) mbar_expect_tx(bar_bfull + sb, 2u * (uint32_t)C::B_STAGE);
                            tma_load_3d_2sm(&P.w_hi, stb, bfullc, chunk * KCHUNK, nrow, wt);
                            if (SPLIT) tma_load_3d_2sm(&P.w_lo, stb + C::B_HALF, bfullc, chunk * KCHUNK, nrow, wt);
                        }
                        __syncwarp();
                        if (++sb == nb) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only) ==================
        if (leader) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t a_ring_u = smem_u32(a_ring), b_ring_u = smem_u32(b_ring);
            int sa = 0; uint32_t pa = 0;
            int sb = 0; uint32_t pb = 0;
            int abuf = 0; uint32_t aphase = 0;
            for (int sup = sch.first; sup < sch.total; sup += sch.step) {
                mbar_wait(bar_tempty + abuf, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(abuf * N_TILE);
                uint32_t started = 0;
                for (int g = 0; g < P.ngroups; g++) {
                    const int cnt = P.g_cnt[g];
                    for (int chunk = 0; chunk < nchunks; chunk++) {
                        mbar_wait(bar_afull + sa, pa);
                        tc_fence_after();
                        const uint32_t a_base = a_ring_u + sa * a_stage;
                        for (int i = 0; i < cnt; i++) {
                            mbar_wait(bar_bfull + sb, pb);
                            tc_fence_after();
                            const uint32_t a_hi = a_base + (uint32_t)i * 1024u;        // tap i: the tile i image rows further down
                            const uint32_t a_lo = a_hi + a_plane;
                            const uint32_t b_hi = b_ring_u + sb * C::B_STAGE;
                            const uint32_t b_lo = b_hi + C::B_HALF;
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < KCHUNK / 16; k++) {
                                    const uint64_t da = make_desc(a_hi + k * 32), db = make_desc(b_hi + k * 32);
                                    umma_f16_2sm(d_tmem, da, db, idesc, (started | (uint32_t)k) ? 1u : 0u);
                                    if (SPLIT && P.f8) {
                                        umma_f8_2sm(d_tmem, make_desc(a_lo + k * 32), make_desc(b_lo + k * 32), idesc, 1u);
                                    } else if (SPLIT) {
                                        umma_f16_2sm(d_tmem, da, make_desc(b_lo + k * 32), idesc, 1u);
                                        umma_f16_2sm(d_tmem, make_desc(a_lo + k * 32), db, idesc, 1u);
                                    }
                                }
                                umma_commit_2sm(bar_bempty + sb, 3);            // weight slot free in BOTH CTAs
                            }
                            __syncwarp();
                            started = 1;
                            if (++sb == nb) { sb = 0; pb ^= 1; }
                        }
                        if (elect_one()) umma_commit_2sm(bar_aempty + sa, 3);   // activation slot free in BOTH CTAs
                        __syncwarp();
                        if (++sa == na) { sa = 0; pa ^= 1; }
                    }
                }
                if (elect_one()) umma_commit_2sm(bar_tfull + abuf, 3);
                __syncwarp();
                if (++abuf == 2) { abuf = 0; aphase ^= 1; }
            }
        }
    } else {
        epilogue_loop<N_TILE, 1, true>(P, warp, lane, sch, tmem_base, bar_tfull, bar_tempty, s_stats);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// =====================================================================================================
// Halo variant (stride-1 k x k convs, and the row-K stem): the activation tile is fetched ONCE per
// 64-channel chunk together with its halo -- (16 + kh - 1) rows of 16 pixels (8 + kw - 1 <= 16) -- and
// every filter tap reads a shifted window of that one shared-memory tile through the UMMA descriptor:
//   start address = tile + ky * pitch + kx * 128 B,  8-row-group stride (SBO) = pitch = 16 px * 128 B.
//   The window then no longer starts on a 1024 B swizzle-atom boundary; measured on B200
//   (tools/halo_diag.py): the tensor core derives the 128B-swizzle phase from the ADDRESS bits
//   [7:9] of every row it fetches -- exactly what TMA used when it wrote the tile -- so the
//   descriptor's base_offset field must stay 0 (setting it to kx or -kx gives garbage).
// Activation traffic L2->SMEM drops from taps x 16 KB to (16+kh-1) x 2 KB per chunk (3x3: 4x less,
// 7x7: 17x less); weights stream per (chunk, tap) through their own ring.  This is what makes the
// narrow layers (Cout = 64 at 256^2, the 7x7 heads) tensor-bound instead of L2-bound.
// =====================================================================================================
struct HaloParams {
    CUtensorMap a_hi[2];
    CUtensorMap a_lo[2];
    CUtensorMap w_hi;
    CUtensorMap w_lo;
    int n_img, tiles_y, tiles_x, n_tiles_n;
    int dom_h, dom_w;
    int kh, kw;                       // taps = kh * kw (row-K: kw = 1, the filter row lives in K)
    int chunks0, chunks1;
    int x_off, y_off;                 // halo box origin relative to the tile origin (-pad, or 0 for row-K)
    int a_rows;                       // TILE_H + kh - 1
    int pitch_bytes;                  // smem bytes per halo row: 2048 (16 px), or 1024 for row-K (8 positions)
    int a_plane_bytes;                // a_rows * pitch_bytes (one of hi / lo)
    int a_stage_bytes;                // a_plane_bytes * (SPLIT ? 2 : 1)
    int nb_stages;
    int resident_b;                   // 1: the whole weight tensor (ntaps x nchunks tiles) stays in smem for the kernel's lifetime
    int bo_mode;                      // descriptor base_offset for a kx-shifted window: 0 (correct), 1 +kx, 2 -kx (LWB_HALO_BO, diagnostic only)
    float* out; int out_h, out_w, cout;
    int oy_mul, oy_add, ox_mul, ox_add;
    double* stats;
    float out_scale;                  // always 1 (shared epilogue)
    int phase_cols;                   // always 0 (shared epilogue)
};

constexpr int HALO_NA = 2;            // activation ring depth
constexpr int HALO_MAX_NB = 8;        // weight ring depth (max)

__device__ __forceinline__ uint64_t make_desc_ex(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46)
         | ((uint64_t)(base_offset & 7) << 49) | (2ull << 61);
}

template <int N_TILE, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_halo(const __grid_constant__ HaloParams P)
{
    constexpr int B_BYTES = N_TILE * 128;
    constexpr int B_STAGE = B_BYTES * (SPLIT ? 2 : 1);
    constexpr int TMEM_COLS = Cfg<N_TILE, SPLIT>::TMEM_COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + HALO_NA * P.a_stage_bytes;
    uint8_t* tail = b_ring + P.nb_stages * B_STAGE;
    uint64_t* bar_afull = reinterpret_cast<uint64_t*>(tail);
    uint64_t* bar_aempty = bar_afull + HALO_NA;
    uint64_t* bar_bfull = bar_aempty + HALO_NA;
    uint64_t* bar_bempty = bar_bfull + HALO_MAX_NB;
    uint64_t* bar_tfull = bar_bempty + HALO_MAX_NB;
    uint64_t* bar_tempty = bar_tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
    float2* s_stats = reinterpret_cast<float2*>(tail + 256);

    const int warp = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;
    const int nb = P.nb_stages;

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < HALO_NA; s++) { mbar_init(bar_afull + s, 1); mbar_init(bar_aempty + s, 1); }
            for (int s = 0; s < nb; s++) { mbar_init(bar_bfull + s, 1); mbar_init(bar_bempty + s, 1); }
            for (int b = 0; b < 2; b++) { mbar_init(bar_tfull + b, 1); mbar_init(bar_tempty + b, 4); }
            fence_barrier_init();
            fence_proxy_async();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nchunks = P.chunks0 + P.chunks1;
    const int ntaps = P.kh * P.kw;
    const int m_tiles = P.n_img * P.tiles_y * P.tiles_x;
    const int total_tiles = m_tiles * P.n_tiles_n;

    if (warp == 0) {
        // ================================ TMA producer =================================
        // Work items = (tile, chunk) in execution order.  The activation halo of item i+1 is requested
        // BEFORE the weight taps of item i are streamed, so the big A transfer has a whole chunk of MMA
        // time to land (the weight ring alone would only let it start 2-4 taps before it is needed).
        {
            int sa = 0; uint32_t pa = 0;
            int sb = 0; uint32_t pb = 0;
            const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const int items = my_tiles * nchunks;
            auto issue_a = [&](int item) {
                const int tile = (int)blockIdx.x + (item / nchunks) * (int)gridDim.x;
                const int chunk = item % nchunks;
                const int m_idx = tile % m_tiles;
                const int img = m_idx / (P.tiles_y * P.tiles_x);
                const int rem = m_idx % (P.tiles_y * P.tiles_x);
                const int y0 = (rem / P.tiles_x) * TILE_H + P.y_off, x0 = (rem % P.tiles_x) * TILE_W + P.x_off;
                const bool second = chunk >= P.chunks0;
                const int mi = second ? 1 : 0;
                const int c0 = (second ? chunk - P.chunks0 : chunk) * KCHUNK;
                mbar_wait(bar_aempty + sa, pa ^ 1);
                uint8_t* sta = a_ring + sa * P.a_stage_bytes;
                if (elect_one()) {
                    mbar_expect_tx(bar_afull + sa, (uint32_t)P.a_stage_bytes);
                    tma_load_4d(&P.a_hi[mi], sta, bar_afull + sa, c0, x0, y0, img);
                    if (SPLIT) tma_load_4d(&P.a_lo[mi], sta + P.a_plane_bytes, bar_afull + sa, c0, x0, y0, img);
                }
                __syncwarp();
                if (++sa == HALO_NA) { sa = 0; pa ^= 1; }
            };
            if (P.resident_b && items > 0) {
                // weights of every (chunk, tap) are fetched once; slot = chunk * ntaps + tap, barrier phase 0 forever
                for (int ct = 0; ct < nchunks * ntaps; ct++) {
                    uint8_t* stb = b_ring + ct * B_STAGE;
                    if (elect_one()) {
                        mbar_expect_tx(bar_bfull + ct, (uint32_t)B_STAGE);
                        tma_load_3d(&P.w_hi, stb, bar_bfull + ct, (ct / ntaps) * KCHUNK, 0, ct % ntaps);
                        if (SPLIT) tma_load_3d(&P.w_lo, stb + B_BYTES, bar_bfull + ct, (ct / ntaps) * KCHUNK, 0, ct % ntaps);
                    }
                    __syncwarp();
                }
            }
            if (items > 0) issue_a(0);
            for (int item = 0; item < items; item++) {
                if (item + 1 < items) issue_a(item + 1);
                if (P.resident_b) continue;
                const int tile = (int)blockIdx.x + (item / nchunks) * (int)gridDim.x;
                const int chunk = item % nchunks;
                const int n_idx = tile / m_tiles;
                for (int tap = 0; tap < ntaps; tap++) {
                    mbar_wait(bar_bempty + sb, pb ^ 1);
                    uint8_t* stb = b_ring + sb * B_STAGE;
                    if (elect_one()) {
                        mbar_expect_tx(bar_bfull + sb, (uint32_t)B_STAGE);
                        tma_load_3d(&P.w_hi, stb, bar_bfull + sb, chunk * KCHUNK, n_idx * N_TILE, tap);
                        if (SPLIT) tma_load_3d(&P.w_lo, stb + B_BYTES, bar_bfull + sb, chunk * KCHUNK, n_idx * N_TILE, tap);
                    }
                    __syncwarp();
                    if (++sb == nb) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (whole warp, elected lane issues) ==
        {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int sa = 0; uint32_t pa = 0;
            int sb = 0; uint32_t pb = 0;
            int abuf = 0; uint32_t aphase = 0;
            const uint32_t a_ring_u = smem_u32(a_ring), b_ring_u = smem_u32(b_ring);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(bar_tempty + abuf, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(abuf * N_TILE);
                for (int chunk = 0; chunk < nchunks; chunk++) {
                    mbar_wait(bar_afull + sa, pa);
                    tc_fence_after();
                    const uint32_t a_hi = a_ring_u + sa * P.a_stage_bytes;
                    const uint32_t a_lo = a_hi + P.a_plane_bytes;
                    for (int tap = 0; tap < ntaps; tap++) {
                        if (P.resident_b) { sb = chunk * ntaps + tap; pb = 0; }
                        mbar_wait(bar_bfull + sb, pb);
                        tc_fence_after();
                        const int ky = tap / P.kw, kx = tap - ky * P.kw;
                        const uint32_t a_off = (uint32_t)(ky * P.pitch_bytes + kx * 128);
                        const uint32_t bo = P.bo_mode == 1 ? (uint32_t)kx : (P.bo_mode == 2 ? (uint32_t)((8 - kx) & 7) : 0u);
                        const uint32_t b_hi = b_ring_u + sb * B_STAGE;
                        const uint32_t b_lo = b_hi + B_BYTES;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < KCHUNK / 16; k++) {
                                const uint64_t da = make_desc_ex(a_hi + a_off + k * 32, (uint32_t)P.pitch_bytes, bo);
                                const uint64_t db = make_desc(b_hi + k * 32);
                                umma_f16(d_tmem, da, db, idesc, (chunk > 0 || tap > 0 || k > 0) ? 1u : 0u);
                                if (SPLIT) {
                                    umma_f16(d_tmem, da, make_desc(b_lo + k * 32), idesc, 1u);
                                    umma_f16(d_tmem, make_desc_ex(a_lo + a_off + k * 32, (uint32_t)P.pitch_bytes, bo), db, idesc, 1u);
                                }
                            }
                            if (!P.resident_b) umma_commit(bar_bempty + sb);
                        }
                        __syncwarp();
                        if (!P.resident_b) { if (++sb == nb) { sb = 0; pb ^= 1; } }
                    }
                    if (elect_one()) umma_commit(bar_aempty + sa);
                    __syncwarp();
                    if (++sa == HALO_NA) { sa = 0; pa ^= 1; }
                }
                if (elect_one()) umma_commit(bar_tfull + abuf);
                __syncwarp();
                if (++abuf == 2) { abuf = 0; aphase ^= 1; }
            }
        }
    } else {
        const Sched sch = make_sched(m_tiles, P.n_tiles_n, 1);
        epilogue_loop<N_TILE, 1>(P, warp, lane, sch, tmem_base, bar_tfull, bar_tempty, s_stats);
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp16 tensor map, 128B swizzle, zero OOB fill. dims/strides innermost first; strides in bytes for dims 1..rank-1.
int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) { lwb::set_error("cuTensorMapEncodeTiled entry point not available"); return LWB_E_CUDA; }
    cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; i++) gstr[i] = strides_bytes[i];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, box[0] == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        lwb::set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] strides [%llu,%llu,%llu] box [%u,%u,%u,%u]",
                       (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                       (unsigned long long)(rank > 3 ? dims[3] : 0), (unsigned long long)strides_bytes[0],
                       (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), (unsigned long long)(rank > 3 ? strides_bytes[2] : 0),
                       box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return LWB_E_CUDA;
    }
    return LWB_OK;
}

struct Launch {
    ConvParams p;
    HaloParams h;
    bool halo;
    int halo_smem;
    int n_tile;
    bool split;
    int cl;
    int kc;
    bool two_sm;
    bool fused;
    bool yhalo;
    int grid;
};

template <int N_TILE, bool SPLIT>
int launch_halo_one(const Launch& L, cudaStream_t st)
{
    static int attr_smem_dev[lwb::kMaxDevices] = {};          // function attributes are per device
    int& attr_smem = attr_smem_dev[lwb::device_slot()];
    if (L.halo_smem > attr_smem) {
        LWB_CUDA_OK(cudaFuncSetAttribute(k_conv_halo<N_TILE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.halo_smem));
        attr_smem = L.halo_smem;
    }
    k_conv_halo<N_TILE, SPLIT><<<L.grid, NUM_THREADS, L.halo_smem, st>>>(L.h);
    LWB_LAUNCH_OK();
    return LWB_OK;
}

template <int N_TILE, bool SPLIT, int CL, int KC>
int launch_cl(const Launch& L, cudaStream_t st)
{
    using C = Cfg<N_TILE, SPLIT, KC>;
    static bool attr_set_dev[lwb::kMaxDevices] = {};          // function attributes are per device
    bool& attr_set = attr_set_dev[lwb::device_slot()];
    if (!attr_set) {
        LWB_CUDA_OK(cudaFuncSetAttribute(k_conv_tc<N_TILE, SPLIT, CL, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    ConvParams& pp = const_cast<ConvParams&>(L.p);
    { static int forced = -1; if (forced < 0) { const char* e = getenv("LWB_STAGES"); forced = e ? atoi(e) : 0; }
      pp.stages = (forced >= 2 && forced < C::STAGES) ? forced : C::STAGES; }
    if (CL == 1) {
        LWB_CUDA_OK(lwb::launch_pdl(k_conv_tc<N_TILE, SPLIT, CL, KC>, dim3(L.grid), dim3(NUM_THREADS), C::SMEM_BYTES, st, L.p));
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        LWB_CUDA_OK(cudaLaunchKernelEx(&cfg, k_conv_tc<N_TILE, SPLIT, CL, KC>, L.p));
    }
    LWB_LAUNCH_OK();
    return LWB_OK;
}

template <int N_TILE, bool SPLIT, bool FUSED = false>
int launch_2sm(const Launch& L, cudaStream_t st)
{
    using C = Cfg2<N_TILE, SPLIT>;
    static bool attr_set_dev[lwb::kMaxDevices] = {};          // function attributes are per device
    bool& attr_set = attr_set_dev[lwb::device_slot()];
    if (!attr_set) {
        LWB_CUDA_OK(cudaFuncSetAttribute(k_conv_tc2<N_TILE, SPLIT, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    // Programmatic dependent launch (LWB_PDL=1): the CTAs may be scheduled while the previous kernel of the stream
    // drains; they set up barriers / TMEM and then block in griddepcontrol.wait until that kernel has completed.
    const int pdl = lwb::pdl_enabled() ? 1 : 0;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
    LWB_CUDA_OK(cudaLaunchKernelEx(&cfg, k_conv_tc2<N_TILE, SPLIT, FUSED>, L.p));
    LWB_LAUNCH_OK();
    return LWB_OK;
}

template <int N_TILE, bool SPLIT>
int launch_2sm_y(const Launch& L, cudaStream_t st)
{
    using C = CfgY<N_TILE, SPLIT>;
    static bool attr_set_dev[lwb::kMaxDevices] = {};
    bool& attr_set = attr_set_dev[lwb::device_slot()];
    if (!attr_set) {
        LWB_CUDA_OK(cudaFuncSetAttribute(k_conv_tc2y<N_TILE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = lwb::pdl_enabled() ? 2 : 1;
    LWB_CUDA_OK(cudaLaunchKernelEx(&cfg, k_conv_tc2y<N_TILE, SPLIT>, L.p));
    LWB_LAUNCH_OK();
    return LWB_OK;
}

template <int N_TILE, bool SPLIT>
int launch_one(const Launch& L, cudaStream_t st)
{
    if (L.two_sm && L.yhalo) {
        if constexpr (N_TILE >= 32) return launch_2sm_y<N_TILE, SPLIT>(L, st);
    }
    if (L.two_sm) {
        if constexpr (SPLIT && N_TILE >= 128) { if (L.fused) return launch_2sm<N_TILE, SPLIT, true>(L, st); }
        if constexpr (N_TILE >= 64) return launch_2sm<N_TILE, SPLIT>(L, st);
    }
    if (L.kc == 32) return launch_cl<N_TILE, SPLIT, 1, 32>(L, st);
    if (L.cl == 2) return launch_cl<N_TILE, SPLIT, 2, 64>(L, st);
    if (L.cl == 4 && N_TILE >= 32) return launch_cl<N_TILE, SPLIT, (N_TILE >= 32 ? 4 : 1), 64>(L, st);
    return launch_cl<N_TILE, SPLIT, 1, 64>(L, st);
}

int launch(const Launch& L, cudaStream_t st)
{
    if (L.halo) {
        switch (L.n_tile) {
            case 16:  return L.split ? launch_halo_one<16, true>(L, st) : launch_halo_one<16, false>(L, st);
            case 64:  return L.split ? launch_halo_one<64, true>(L, st) : launch_halo_one<64, false>(L, st);
            case 128: return L.split ? launch_halo_one<128, true>(L, st) : launch_halo_one<128, false>(L, st);
            case 256: return L.split ? launch_halo_one<256, true>(L, st) : launch_halo_one<256, false>(L, st);
        }
        lwb::set_error("conv_halo: unsupported N tile %d", L.n_tile);
        return LWB_E_UNSUPPORTED;
    }
    switch (L.n_tile) {
        case 16:  return L.split ? launch_one<16, true>(L, st) : launch_one<16, false>(L, st);
        case 32:  return L.split ? launch_one<32, true>(L, st) : launch_one<32, false>(L, st);
        case 64:  return L.split ? launch_one<64, true>(L, st) : launch_one<64, false>(L, st);
        case 128: return L.split ? launch_one<128, true>(L, st) : launch_one<128, false>(L, st);
        case 256: return L.split ? launch_one<256, true>(L, st) : launch_one<256, false>(L, st);
    }
    lwb::set_error("conv_tc: unsupported N tile %d", L.n_tile);
    return LWB_E_UNSUPPORTED;
}

// y-halo schedule: sort the taps by (input view, dx, dy) and cut them into runs of consecutive dy.  Returns the activation box
// rows (TILE_H + longest run - 1) and fills the group tables + the weight-ring depth, or 0 when the variant does not apply
// (disabled, longest run 1, rings do not fit).
int plan_yhalo(ConvParams& p, int n_tile, bool split)
{
    const char* e = getenv("LWB_YHALO");                     // read per plan: tests switch it between plans
    const int want = e ? atoi(e) : 1;
    if (!want || p.ntaps < 2) return 0;
    int order[MAX_TAPS];
    for (int t = 0; t < p.ntaps; t++) order[t] = t;
    auto key = [&](int t) { return ((int)p.tmap[t] << 16) + (((int)p.dx[t] + 128) << 8) + ((int)p.dy[t] + 128); };
    for (int i = 1; i < p.ntaps; i++) {                      // insertion sort (<= 49 taps)
        const int v = order[i];
        int j = i - 1;
        while (j >= 0 && key(order[j]) > key(v)) { order[j + 1] = order[j]; j--; }
        order[j + 1] = v;
    }
    signed char dy[MAX_TAPS], dx[MAX_TAPS], tm[MAX_TAPS]; short wt[MAX_TAPS];
    for (int i = 0; i < p.ntaps; i++) { dy[i] = p.dy[order[i]]; dx[i] = p.dx[order[i]]; tm[i] = p.tmap[order[i]]; wt[i] = p.wtap[order[i]]; }
    int ng = 0, maxcnt = 1;
    unsigned char first[MAX_TAPS], cnt[MAX_TAPS];
    for (int i = 0; i < p.ntaps; i++) {
        if (i > 0 && tm[i] == tm[i - 1] && dx[i] == dx[i - 1] && dy[i] == dy[i - 1] + 1 && cnt[ng - 1] < 7) { cnt[ng - 1]++; }
        else { first[ng] = (unsigned char)i; cnt[ng] = 1; ng++; }
        if (cnt[ng - 1] > maxcnt) maxcnt = cnt[ng - 1];
    }
    if (maxcnt < 2) return 0;
    const int a_rows = TILE_H + maxcnt - 1;
    const int a_stage = a_rows * 1024 * (split ? 2 : 1);
    const int b_stage = (n_tile / 2) * 128 * (split ? 2 : 1);
    int na = 0, nb = 0;
    for (int cand = Y_MAX_NA; cand >= 2 && !na; cand--) {    // deepest activation ring that leaves a useful weight ring
        const int left = (196 * 1024 - cand * a_stage) / b_stage;
        if (left >= (cand == 2 ? 3 : 4)) { na = cand; nb = left; }
    }
    if (!na) return 0;
    if (nb > Y_MAX_NB) nb = Y_MAX_NB;
    { const char* f = getenv("LWB_YHALO_NA"); if (f && atoi(f) >= 2 && atoi(f) <= na) { na = atoi(f); nb = (196 * 1024 - na * a_stage) / b_stage; if (nb > Y_MAX_NB) nb = Y_MAX_NB; } }
    for (int i = 0; i < p.ntaps; i++) { p.dy[i] = dy[i]; p.dx[i] = dx[i]; p.tmap[i] = tm[i]; p.wtap[i] = wt[i]; }
    for (int g = 0; g < ng; g++) { p.g_first[g] = first[g]; p.g_cnt[g] = cnt[g]; }
    p.ngroups = ng; p.a_rows = a_rows; p.na_stages = na; p.nb_stages = nb;
    return a_rows;
}

int pick_n_tile(int cout, bool split, int forced)
{
    if (forced > 0) return forced;
    if (cout % 256 == 0) return 256;      // measured on B200: split N=256 (2 stages) 473 TF/s vs N=128 (3 stages) 420
    if (cout % 128 == 0) return 128;
    if (cout % 64 == 0) return 64;
    if (cout % 32 == 0) return 32;
    if (cout % 16 == 0) return 16;
    return -1;
}

}  // namespace

struct lwb_conv_plan {
    int num;
    Launch launches[4];
};

// Plain NHWC activation map: dims [C, W, H, N].
static int map_nhwc(CUtensorMap* m, const uint16_t* base, int n, int h, int w, int c, int kc = KCHUNK, int rows = TILE_H)
{
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)c * 2, (uint64_t)w * c * 2, (uint64_t)h * w * c * 2};
    const uint32_t box[4] = {(uint32_t)kc, TILE_W, (uint32_t)rows, 1};
    return encode_map(m, base, 4, dims, str, box);
}
// Parity view (py, px) of an NHWC tensor for stride-2 convs: element (y', x') = input (2y'+py, 2x'+px).
static int map_nhwc_parity(CUtensorMap* m, const uint16_t* base, int n, int h, int w, int c, int py, int px, int kc = KCHUNK,
                           int rows = TILE_H)
{
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)((w - px + 1) / 2), (uint64_t)((h - py + 1) / 2), (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)2 * c * 2, (uint64_t)2 * w * c * 2, (uint64_t)h * w * c * 2};
    const uint32_t box[4] = {(uint32_t)kc, TILE_W, (uint32_t)rows, 1};
    return encode_map(m, base + ((size_t)py * w + px) * c, 4, dims, str, box);
}

extern "C" int lwb_conv_plan_create(const lwb_conv_desc* d,
                                    const uint16_t* x0_hi, const uint16_t* x0_lo,
                                    const uint16_t* x1_hi, const uint16_t* x1_lo,
                                    const uint16_t* w_hi, const uint16_t* w_lo,
                                    float* out_raw, double* stats, lwb_conv_plan** plan_out)
{
    LWB_CHECK_ARG(d && x0_hi && w_hi && out_raw && plan_out, "null pointer");
    const bool split = d->split != 0;
    const bool f8 = d->split == 2;     // lo operands are fp8 pairs (lwb_pack_conv_weight_f8 / lo_format 1 of lwb_norm_act_nhwc)
    LWB_CHECK_ARG(d->split >= 0 && d->split <= 2, "split must be 0, 1 or 2");
    LWB_CHECK_ARG(!split || (x0_lo && w_lo), "split mode needs the lo operands");
    LWB_CHECK_ARG(!f8 || (!d->rowk && !d->halo), "the fp8 lo mode is not available for row-K / halo plans");
    LWB_CHECK_ARG(d->n > 0 && d->h_in > 0 && d->w_in > 0 && d->cout > 0, "non-positive size");
    LWB_CHECK_ARG(d->cout % 16 == 0, "cout must be a multiple of 16");
    LWB_CHECK_ARG(!f8 || (d->w_exp >= -40 && d->w_exp <= 60), "w_exp out of range");
    int n_tile = pick_n_tile(d->cout, split, d->n_tile);
    if (d->halo && split && n_tile == 256 && d->n_tile == 0) n_tile = 128;      // halo + split: the 256-wide weight ring does not fit
    LWB_CHECK_ARG(n_tile > 0 && d->cout % n_tile == 0, "no N tile divides cout");

    const int sms = lwb::sm_count();
    // Cluster size for weight-tile multicast (LWB_CLUSTER = 1 | 2 | 4, default 1): the CTAs of a cluster work on
    // consecutive M tiles of the same N tile, each fetches 1/CL of the weight tile and multicasts it.
    // Measured on B200 (tools/conv_microbench.py): CL=2 is within 2% of CL=1 on every layer, CL=4 is slower --
    // weight-tile L2 traffic is not what bounds this kernel -- so the plain launch stays the default.
    const int dom_h0 = d->transposed ? d->h_in : d->h_out, dom_w0 = d->transposed ? d->w_in : d->w_out;
    const long m_tiles0 = (long)d->n * lwb::ceil_div(dom_h0, TILE_H) * lwb::ceil_div(dom_w0, TILE_W);
    int cl = 1;
    { const char* e = getenv("LWB_CLUSTER"); if (e) cl = atoi(e); }
    if (cl != 1 && cl != 2 && cl != 4) cl = 1;
    while (cl > 1 && (m_tiles0 % cl != 0 || (n_tile / cl) % 8 != 0 || n_tile / cl < 8)) cl >>= 1;
    if (d->halo) cl = 1;
    // 2-CTA MMA (cta_group::2, LWB_2SM=0 disables): pairs of M tiles share one weight tile, half of it per CTA
    bool two_sm = false;
    bool want_2sm = true;
    { const char* e = getenv("LWB_2SM"); const int want = e ? atoi(e) : 1;
      want_2sm = want != 0;
      two_sm = want && !d->halo && n_tile >= 64 && (m_tiles0 % 2 == 0) && sms >= 2;
      if (d->transposed && n_tile <= 64 && want != 2) two_sm = false;      // measured: 0.276 vs 0.240 ms on the 128->64 phases
    }
    if (two_sm) cl = 2;
    // K elements per pipeline stage: 64 (128 B rows, SWIZZLE_128B) or 32 (64 B rows, SWIZZLE_64B: twice the stages)
    int kc = KCHUNK;
    { const char* e = getenv("LWB_KC"); if (e && atoi(e) == 32 && !d->rowk && !d->halo && !f8) { kc = 32; cl = 1; two_sm = false; } }

    lwb_conv_plan* plan = new (std::nothrow) lwb_conv_plan();
    LWB_CHECK_ARG(plan, "out of host memory");
    plan->num = 0;
    int rc = LWB_OK;
    auto fail = [&](int code) { delete plan; return code; };

    auto finish = [&](Launch& L, int dom_h, int dom_w) {
        ConvParams& p = L.p;
        p.n_img = d->n;
        p.dom_h = dom_h; p.dom_w = dom_w;
        p.tiles_y = lwb::ceil_div(dom_h, TILE_H); p.tiles_x = lwb::ceil_div(dom_w, TILE_W);
        p.n_tiles_n = d->cout / n_tile;
        if (getenv("LWB_DEBUG_NOEPI")) { p.oy_mul = -1; }
        p.out = getenv("LWB_DEBUG_NOSTORE") ? nullptr : out_raw; p.out_h = d->h_out; p.out_w = d->w_out; p.cout = d->cout;
        p.stats = stats;
        p.f8 = f8 ? 1 : 0;
        p.out_scale = f8 ? ldexpf(1.f, -d->w_exp) : 1.f;      // weights are packed x 2^w_exp in f8 mode (lwb_pack_conv_weight_f8)
        L.n_tile = n_tile; L.split = split; L.halo = false; L.halo_smem = 0; L.cl = cl; L.kc = d->rowk ? KCHUNK : kc;
        L.two_sm = two_sm; L.fused = false; L.yhalo = false;
        p.stages = 64;      // clamped to Cfg::STAGES at launch
        const long total_super = (long)p.n_img * p.tiles_y * p.tiles_x * p.n_tiles_n / cl;
        const long max_clusters = sms / cl;
        L.grid = (int)((total_super < max_clusters ? total_super : max_clusters) * cl);
    };

    if (d->halo) {
        // Halo variant: stride-1 k x k conv (pad = k/2, dil 1, optional concat input) or the row-K stem.
        LWB_CHECK_ARG(d->stride == 1 && !d->transposed && d->dil == 1, "halo mode needs stride 1, dilation 1, not transposed");
        Launch& L = plan->launches[plan->num++];
        memset(&L.h, 0, sizeof(L.h));
        HaloParams& h = L.h;
        L.halo = true;
        const int b_stage = n_tile * 128 * (split ? 2 : 1);
        if (d->rowk) {
            LWB_CHECK_ARG(d->kw <= 8 && d->cin0 == 8 && d->cin1 == 0 && d->row_pitch >= d->w_in + 8, "row-K shape");
            const int hp = d->h_in + d->kh - 1;
            h.kh = d->kh; h.kw = 1; h.x_off = 0; h.y_off = 0;
            h.a_rows = TILE_H + d->kh - 1; h.pitch_bytes = TILE_W * 128;
            const uint64_t dims[4] = {64, (uint64_t)d->w_in, (uint64_t)hp, (uint64_t)d->n};
            const uint64_t str[3] = {16, (uint64_t)d->row_pitch * 16, (uint64_t)hp * d->row_pitch * 16};
            const uint32_t box[4] = {KCHUNK, TILE_W, (uint32_t)h.a_rows, 1};
            if ((rc = encode_map(&h.a_hi[0], x0_hi, 4, dims, str, box)) != LWB_OK) return fail(rc);
            if (split && (rc = encode_map(&h.a_lo[0], x0_lo, 4, dims, str, box)) != LWB_OK) return fail(rc);
            const uint64_t wd[3] = {64, (uint64_t)d->cout, (uint64_t)d->kh};
            const uint64_t ws[2] = {128, (uint64_t)d->cout * 128};
            const uint32_t wb[3] = {KCHUNK, (uint32_t)n_tile, 1};
            if ((rc = encode_map(&h.w_hi, w_hi, 3, wd, ws, wb)) != LWB_OK) return fail(rc);
            if (split && (rc = encode_map(&h.w_lo, w_lo, 3, wd, ws, wb)) != LWB_OK) return fail(rc);
            h.chunks0 = 1; h.chunks1 = 0;
        } else {
            LWB_CHECK_ARG(d->kw <= 9 && (d->kw & 1) && (d->kh & 1), "halo mode needs an odd kernel, kw <= 9, with 'same' padding (kh/2, kw/2)");
            LWB_CHECK_ARG(d->cin0 % KCHUNK == 0 && d->cin1 % KCHUNK == 0 && d->cin0 > 0, "input channels must be multiples of 64");
            LWB_CHECK_ARG(d->cin1 == 0 || (x1_hi && (!split || x1_lo)), "second input missing");
            h.kh = d->kh; h.kw = d->kw; h.x_off = -(d->kw / 2); h.y_off = -(d->kh / 2);
            h.a_rows = TILE_H + d->kh - 1; h.pitch_bytes = 16 * 128;
            auto amap = [&](CUtensorMap* m, const uint16_t* base, int c) {
                const uint64_t dims[4] = {(uint64_t)c, (uint64_t)d->w_in, (uint64_t)d->h_in, (uint64_t)d->n};
                const uint64_t str[3] = {(uint64_t)c * 2, (uint64_t)d->w_in * c * 2, (uint64_t)d->h_in * d->w_in * c * 2};
                const uint32_t box[4] = {KCHUNK, 16, (uint32_t)h.a_rows, 1};
                return encode_map(m, base, 4, dims, str, box);
            };
            if ((rc = amap(&h.a_hi[0], x0_hi, d->cin0)) != LWB_OK) return fail(rc);
            if (split && (rc = amap(&h.a_lo[0], x0_lo, d->cin0)) != LWB_OK) return fail(rc);
            if (d->cin1) {
                if ((rc = amap(&h.a_hi[1], x1_hi, d->cin1)) != LWB_OK) return fail(rc);
                if (split && (rc = amap(&h.a_lo[1], x1_lo, d->cin1)) != LWB_OK) return fail(rc);
            }
            const int cin_total = d->cin0 + d->cin1;
            const uint64_t wd[3] = {(uint64_t)cin_total, (uint64_t)d->cout, (uint64_t)(d->kh * d->kw)};
            const uint64_t ws[2] = {(uint64_t)cin_total * 2, (uint64_t)d->cout * cin_total * 2};
            const uint32_t wb[3] = {KCHUNK, (uint32_t)n_tile, 1};
            if ((rc = encode_map(&h.w_hi, w_hi, 3, wd, ws, wb)) != LWB_OK) return fail(rc);
            if (split && (rc = encode_map(&h.w_lo, w_lo, 3, wd, ws, wb)) != LWB_OK) return fail(rc);
            h.chunks0 = d->cin0 / KCHUNK; h.chunks1 = d->cin1 / KCHUNK;
        }
        { const char* e = getenv("LWB_HALO_BO"); h.bo_mode = e ? atoi(e) : 0; }
        h.a_plane_bytes = h.a_rows * h.pitch_bytes;
        h.a_stage_bytes = h.a_plane_bytes * (split ? 2 : 1);
        const int fixed = 1024 + HALO_NA * h.a_stage_bytes + 256 + 4 * n_tile * 8;
        int nbs = (227 * 1024 - fixed) / b_stage;
        {   // whole weight tensor resident when it fits (the 7x7 stem: 7 x 16 KB): then only activations stream
            const int slots = (h.chunks0 + h.chunks1) * h.kh * h.kw;
            h.resident_b = (d->cout == n_tile && slots <= HALO_MAX_NB && slots <= nbs) ? 1 : 0;
            if (h.resident_b) nbs = slots;
        }
        if (nbs > HALO_MAX_NB) nbs = HALO_MAX_NB;
        if (nbs < 2) { lwb::set_error("conv_halo: tile does not fit shared memory (n_tile %d, kh %d)", n_tile, d->kh); return fail(LWB_E_UNSUPPORTED); }
        h.nb_stages = nbs;
        L.halo_smem = fixed + nbs * b_stage;
        h.n_img = d->n; h.dom_h = d->h_out; h.dom_w = d->w_out;
        h.tiles_y = lwb::ceil_div(d->h_out, TILE_H); h.tiles_x = lwb::ceil_div(d->w_out, TILE_W);
        h.n_tiles_n = d->cout / n_tile;
        h.out = out_raw; h.out_h = d->h_out; h.out_w = d->w_out; h.cout = d->cout;
        h.oy_mul = 1; h.ox_mul = 1; h.oy_add = 0; h.ox_add = 0;
        h.stats = stats;
        h.out_scale = 1.f;
        L.n_tile = n_tile; L.split = split; L.cl = 1; L.two_sm = false; L.fused = false; L.yhalo = false;
        const long total = (long)h.n_img * h.tiles_y * h.tiles_x * h.n_tiles_n;
        L.grid = (int)(total < sms ? total : sms);
        *plan_out = plan;
        return LWB_OK;
    }

    if (d->rowk) {
        // 7x7 stem through the row-K trick.  Input: padded NHWC8 buffer [n, h_in + kh - 1, wp, 8] whose
        // pixel (y + pad, x + pad) holds input pixel (y, x); wp >= w_in + 8.  One K = 64 stage = 8
        // consecutive pixels x 8 channels of a padded row = one whole filter row (8th tap weight = 0).
        LWB_CHECK_ARG(d->stride == 1 && !d->transposed && d->kw <= 8 && d->cin0 == 8 && d->cin1 == 0, "row-K needs stride 1, kw <= 8, 8 channels");
        LWB_CHECK_ARG(d->h_out == d->h_in && d->w_out == d->w_in && d->row_pitch >= d->w_in + 8, "row-K shape");
        Launch& L = plan->launches[plan->num++];
        memset(&L.p, 0, sizeof(L.p));
        const int hp = d->h_in + d->kh - 1;
        const uint64_t dims[4] = {64, (uint64_t)d->w_in, (uint64_t)hp, (uint64_t)d->n};
        const uint64_t str[3] = {16, (uint64_t)d->row_pitch * 16, (uint64_t)hp * d->row_pitch * 16};
        const uint32_t box[4] = {KCHUNK, TILE_W, TILE_H, 1};
        if ((rc = encode_map(&L.p.a_hi[0], x0_hi, 4, dims, str, box)) != LWB_OK) return fail(rc);
        if (split && (rc = encode_map(&L.p.a_lo[0], x0_lo, 4, dims, str, box)) != LWB_OK) return fail(rc);
        const uint64_t wd[3] = {64, (uint64_t)d->cout, (uint64_t)d->kh};
        const uint64_t ws[2] = {128, (uint64_t)d->cout * 128};
        const uint32_t wb[3] = {KCHUNK, (uint32_t)(n_tile / cl), 1};
        if ((rc = encode_map(&L.p.w_hi, w_hi, 3, wd, ws, wb)) != LWB_OK) return fail(rc);
        if (split && (rc = encode_map(&L.p.w_lo, w_lo, 3, wd, ws, wb)) != LWB_OK) return fail(rc);
        L.p.ntaps = d->kh; L.p.chunks0 = 1; L.p.chunks1 = 0;
        for (int ky = 0; ky < d->kh; ky++) { L.p.dy[ky] = (signed char)ky; L.p.dx[ky] = 0; L.p.tmap[ky] = 0; L.p.wtap[ky] = (short)ky; }
        L.p.oy_mul = 1; L.p.ox_mul = 1; L.p.oy_add = 0; L.p.ox_add = 0;
        bool yh = false;
        if (two_sm) {
            const int r = plan_yhalo(L.p, n_tile, split);
            if (r) {
                yh = true;
                const uint32_t boxy[4] = {KCHUNK, TILE_W, (uint32_t)r, 1};
                if ((rc = encode_map(&L.p.a_hi[0], x0_hi, 4, dims, str, boxy)) != LWB_OK) return fail(rc);
                if (split && (rc = encode_map(&L.p.a_lo[0], x0_lo, 4, dims, str, boxy)) != LWB_OK) return fail(rc);
            }
        }
        finish(L, d->h_out, d->w_out);
        L.yhalo = yh;
        *plan_out = plan;
        return LWB_OK;
    }

    LWB_CHECK_ARG(d->cin0 % KCHUNK == 0 && d->cin1 % KCHUNK == 0 && d->cin0 > 0, "input channels must be multiples of 64");
    LWB_CHECK_ARG(d->cin1 == 0 || (x1_hi && (!split || x1_lo)), "second input missing");
    const int cin_total = d->cin0 + d->cin1;
    const int ntaps_w = d->kh * d->kw;
    LWB_CHECK_ARG(ntaps_w <= MAX_TAPS, "too many filter taps");
    const uint64_t wd[3] = {(uint64_t)cin_total, (uint64_t)d->cout, (uint64_t)ntaps_w};
    const uint64_t ws[2] = {(uint64_t)cin_total * 2, (uint64_t)d->cout * cin_total * 2};
    const uint32_t wb[3] = {(uint32_t)kc, (uint32_t)(n_tile / cl), 1};

    if (d->transposed == 2) {
        // Merged transposed conv: ONE stride-1 pass over the input grid with the four taps (dy, dx) in {0,1}^2 and
        // N = 4 x cout columns = the four sub-pixel phases (weights from the host in [tap][phase*cout + co][cin] layout,
        // zero where a phase does not use a tap: 9 of 16 blocks are non-zero).  One launch, one read of every input tile.
        LWB_CHECK_ARG(d->kh == 3 && d->kw == 3 && d->stride == 2 && d->pad == 1 && d->cin1 == 0, "transposed conv: only k3 s2 p1 op1");
        LWB_CHECK_ARG(d->h_out == 2 * d->h_in && d->w_out == 2 * d->w_in, "transposed conv output must be 2x input");
        const int ncols = 4 * d->cout;
        n_tile = ncols % 256 == 0 ? 256 : (ncols % 128 == 0 ? 128 : 64);
        LWB_CHECK_ARG(d->cout % 32 == 0 && ncols % n_tile == 0, "merged transposed conv needs cout in multiples of 32");
        two_sm = sms >= 2 && (m_tiles0 % 2 == 0);
        { const char* e = getenv("LWB_2SM"); if (e && atoi(e) == 0) two_sm = false; }
        cl = two_sm ? 2 : 1;
        Launch& L = plan->launches[plan->num++];
        memset(&L.p, 0, sizeof(L.p));
        if ((rc = map_nhwc(&L.p.a_hi[0], x0_hi, d->n, d->h_in, d->w_in, d->cin0, kc)) != LWB_OK) return fail(rc);
        if (split && (rc = map_nhwc(&L.p.a_lo[0], x0_lo, d->n, d->h_in, d->w_in, d->cin0, kc)) != LWB_OK) return fail(rc);
        const uint64_t mwd[3] = {(uint64_t)cin_total, (uint64_t)ncols, 4};
        const uint64_t mws[2] = {(uint64_t)cin_total * 2, (uint64_t)ncols * cin_total * 2};
        const uint32_t mwb[3] = {(uint32_t)kc, (uint32_t)(n_tile / cl), 1};
        if ((rc = encode_map(&L.p.w_hi, w_hi, 3, mwd, mws, mwb)) != LWB_OK) return fail(rc);
        if (split && (rc = encode_map(&L.p.w_lo, w_lo, 3, mwd, mws, mwb)) != LWB_OK) return fail(rc);
        for (int t = 0; t < 4; t++) { L.p.dy[t] = (signed char)(t >> 1); L.p.dx[t] = (signed char)(t & 1); L.p.tmap[t] = 0; L.p.wtap[t] = (short)t; }
        L.p.ntaps = 4; L.p.chunks0 = d->cin0 / kc; L.p.chunks1 = 0;
        L.p.oy_mul = 2; L.p.ox_mul = 2; L.p.oy_add = 0; L.p.ox_add = 0;
        bool yh = false;
        if (two_sm && kc == KCHUNK) {
            const int r = plan_yhalo(L.p, n_tile, split);
            if (r) {
                yh = true;
                if ((rc = map_nhwc(&L.p.a_hi[0], x0_hi, d->n, d->h_in, d->w_in, d->cin0, kc, r)) != LWB_OK) return fail(rc);
                if (split && (rc = map_nhwc(&L.p.a_lo[0], x0_lo, d->n, d->h_in, d->w_in, d->cin0, kc, r)) != LWB_OK) return fail(rc);
            }
        }
        finish(L, d->h_in, d->w_in);
        L.yhalo = yh;
        L.p.phase_cols = d->cout;
        L.p.n_tiles_n = ncols / n_tile;
        const long total_super = (long)L.p.n_img * L.p.tiles_y * L.p.tiles_x * L.p.n_tiles_n / cl;
        const long max_clusters = sms / cl;
        L.grid = (int)((total_super < max_clusters ? total_super : max_clusters) * cl);
        *plan_out = plan;
        return LWB_OK;
    }

    if (d->transposed) {
        // ConvTranspose2d(k=3, s=2, p=1, output_padding=1): out[2i+a, 2j+b] gathers, per axis,
        //   a = 0: (k=1, d=0)            a = 1: (k=2, d=0), (k=0, d=+1)        (oy = 2*iy - 1 + ky)
        LWB_CHECK_ARG(d->kh == 3 && d->kw == 3 && d->stride == 2 && d->pad == 1 && d->cin1 == 0, "transposed conv: only k3 s2 p1 op1");
        LWB_CHECK_ARG(d->h_out == 2 * d->h_in && d->w_out == 2 * d->w_in, "transposed conv output must be 2x input");
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
            Launch& L = plan->launches[plan->num++];
            memset(&L.p, 0, sizeof(L.p));
            if ((rc = map_nhwc(&L.p.a_hi[0], x0_hi, d->n, d->h_in, d->w_in, d->cin0, kc)) != LWB_OK) return fail(rc);
            if (split && (rc = map_nhwc(&L.p.a_lo[0], x0_lo, d->n, d->h_in, d->w_in, d->cin0, kc)) != LWB_OK) return fail(rc);
            if ((rc = encode_map(&L.p.w_hi, w_hi, 3, wd, ws, wb)) != LWB_OK) return fail(rc);
            if (split && (rc = encode_map(&L.p.w_lo, w_lo, 3, wd, ws, wb)) != LWB_OK) return fail(rc);
            const int ky_list[2][2] = {{1, -1}, {2, 0}}, d_list[2][2] = {{0, 0}, {0, 1}}, cnt[2] = {1, 2};
            int t = 0;
            for (int i = 0; i < cnt[a]; i++) for (int j = 0; j < cnt[b]; j++) {
                L.p.dy[t] = (signed char)d_list[a][i]; L.p.dx[t] = (signed char)d_list[b][j];
                L.p.tmap[t] = 0; L.p.wtap[t] = (short)(ky_list[a][i] * 3 + ky_list[b][j]);
                t++;
            }
            L.p.ntaps = t; L.p.chunks0 = d->cin0 / kc; L.p.chunks1 = 0;
            L.p.oy_mul = 2; L.p.ox_mul = 2; L.p.oy_add = a; L.p.ox_add = b;
            finish(L, d->h_in, d->w_in);
        }
        *plan_out = plan;
        return LWB_OK;
    }

    LWB_CHECK_ARG(d->stride == 1 || d->stride == 2, "stride must be 1 or 2");
    LWB_CHECK_ARG(d->stride == 1 || d->cin1 == 0, "concat input only with stride 1");
    Launch& L = plan->launches[plan->num++];
    memset(&L.p, 0, sizeof(L.p));
    int t = 0;
    for (int ky = 0; ky < d->kh; ky++) for (int kx = 0; kx < d->kw; kx++) {
        const int oy = ky * d->dil - d->pad, ox = kx * d->dil - (d->pad_w >= 0 ? d->pad_w : d->pad);   // input offset relative to stride*y
        if (oy < -127 || oy > 127 || ox < -127 || ox > 127) { lwb::set_error("conv_tc: tap offset out of range"); return fail(LWB_E_UNSUPPORTED); }
        if (d->stride == 1) {
            L.p.dy[t] = (signed char)oy; L.p.dx[t] = (signed char)ox; L.p.tmap[t] = 0;
        } else {
            // input coordinate 2y + oy = 2(y + floor(oy/2)) + (oy mod 2): parity view + index shift
            const int py = ((oy % 2) + 2) % 2, px = ((ox % 2) + 2) % 2;
            L.p.dy[t] = (signed char)((oy - py) / 2); L.p.dx[t] = (signed char)((ox - px) / 2);
            L.p.tmap[t] = (signed char)(py * 2 + px);
        }
        L.p.wtap[t] = (short)t;
        t++;
    }
    L.p.ntaps = t; L.p.chunks0 = d->cin0 / kc; L.p.chunks1 = d->cin1 / kc;
    L.p.oy_mul = 1; L.p.ox_mul = 1; L.p.oy_add = 0; L.p.ox_add = 0;
    // y-halo schedule (k_conv_tc2y): also opens the cta_group::2 path to N = 32 (folded heads, 16-channel gated layers)
    int rows = TILE_H;
    bool yh = false;
    {
        const bool pair_ok = !d->halo && (m_tiles0 % 2 == 0) && sms >= 2 && kc == KCHUNK && want_2sm;
        if ((two_sm || (pair_ok && n_tile == 32)) && kc == KCHUNK) {
            const int r = plan_yhalo(L.p, n_tile, split);
            if (r) { rows = r; yh = true; two_sm = true; cl = 2; }
        }
    }
    const uint32_t wb2[3] = {(uint32_t)kc, (uint32_t)(n_tile / cl), 1};
    if (d->stride == 1) {
        if ((rc = map_nhwc(&L.p.a_hi[0], x0_hi, d->n, d->h_in, d->w_in, d->cin0, kc, rows)) != LWB_OK) return fail(rc);
        if (split && (rc = map_nhwc(&L.p.a_lo[0], x0_lo, d->n, d->h_in, d->w_in, d->cin0, kc, rows)) != LWB_OK) return fail(rc);
        if (d->cin1) {
            if ((rc = map_nhwc(&L.p.a_hi[1], x1_hi, d->n, d->h_in, d->w_in, d->cin1, kc, rows)) != LWB_OK) return fail(rc);
            if (split && (rc = map_nhwc(&L.p.a_lo[1], x1_lo, d->n, d->h_in, d->w_in, d->cin1, kc, rows)) != LWB_OK) return fail(rc);
        }
    } else {
        for (int py = 0; py < 2; py++) for (int px = 0; px < 2; px++) {
            if ((rc = map_nhwc_parity(&L.p.a_hi[py * 2 + px], x0_hi, d->n, d->h_in, d->w_in, d->cin0, py, px, kc, rows)) != LWB_OK) return fail(rc);
            if (split && (rc = map_nhwc_parity(&L.p.a_lo[py * 2 + px], x0_lo, d->n, d->h_in, d->w_in, d->cin0, py, px, kc, rows)) != LWB_OK) return fail(rc);
        }
    }
    if ((rc = encode_map(&L.p.w_hi, w_hi, 3, wd, ws, wb2)) != LWB_OK) return fail(rc);
    if (split && (rc = encode_map(&L.p.w_lo, w_lo, 3, wd, ws, wb2)) != LWB_OK) return fail(rc);
    finish(L, d->h_out, d->w_out);
    L.yhalo = yh;
    *plan_out = plan;
    return LWB_OK;
}

extern "C" int lwb_conv_plan_fuse_norm(lwb_conv_plan* plan, const lwb_fused_norm* f)
{
    LWB_CHECK_ARG(plan && f, "null pointer");
    if (plan->num != 1) { lwb::set_error("fuse_norm: multi-launch plans (transposed convs) keep the separate norm pass"); return LWB_E_UNSUPPORTED; }
    Launch& L = plan->launches[0];
    ConvParams& p = L.p;
    if (L.halo || L.yhalo || !L.two_sm || !L.split || L.n_tile < 128 || !p.stats || p.oy_mul != 1 || p.ox_mul != 1) {
        lwb::set_error("fuse_norm: needs a split-mode cta_group::2 plan (not the y-halo variant) with N tile >= 128 and statistics");
        return LWB_E_UNSUPPORTED;
    }
    const int tiles_per_image = p.tiles_y * p.tiles_x;
    // every unit (image, N tile) must fit one round of the persistent grid (see epilogue_fused_loop)
    if (tiles_per_image > L.grid || (tiles_per_image & 1)) {
        lwb::set_error("fuse_norm: %d tiles per image exceed one round of the grid (%d CTAs)", tiles_per_image, L.grid);
        return LWB_E_UNSUPPORTED;
    }
    LWB_CHECK_ARG(f->counters && (f->y_hi || f->y_f32), "fuse_norm: counters and an output are required");
    LWB_CHECK_ARG(!f->warp_src || (f->T && f->th > 0 && f->tw > 0 && (f->src_batch == 1 || f->src_batch == p.n_img)), "bad warp arguments");
    LWB_CHECK_ARG(f->lo_format == 0 || (f->lo_format == 1 && p.cout % 64 == 0), "lo_format 1 needs channels in blocks of 64");
    FusedNorm& F = p.fn;
    F.gamma = f->gamma; F.beta = f->beta; F.eps = f->eps; F.relu = f->relu;
    F.residual = f->residual;
    F.warp_src = f->warp_src; F.src_batch = f->src_batch; F.T = f->T; F.th = f->th; F.tw = f->tw; F.align_corners = f->align_corners;
    F.y_f32 = f->y_f32; F.y_hi = (__half*)f->y_hi; F.y_lo = (uint8_t*)f->y_lo; F.lo_format = f->lo_format;
    F.range_flag = f->range_flag;
    F.counters = f->counters; F.tiles_per_image = tiles_per_image;
    F.inv_hw = 1.0 / ((double)p.dom_h * p.dom_w);
    L.fused = true;
    return LWB_OK;
}

extern "C" int lwb_conv_plan_run(const lwb_conv_plan* plan, lwb_stream_t stream)
{
    LWB_CHECK_ARG(plan, "null plan");
    for (int i = 0; i < plan->num; i++) {
        const int rc = launch(plan->launches[i], (cudaStream_t)stream);
        if (rc != LWB_OK) return rc;
    }
    return LWB_OK;
}

extern "C" void lwb_conv_plan_destroy(lwb_conv_plan* plan) { delete plan; }

extern "C" int lwb_conv_plan_num_launches(const lwb_conv_plan* plan) { return plan ? plan->num : 0; }

extern "C" int lwb_conv2d_nhwc(const lwb_conv_desc* d,
                               const uint16_t* x0_hi, const uint16_t* x0_lo,
                               const uint16_t* x1_hi, const uint16_t* x1_lo,
                               const uint16_t* w_hi, const uint16_t* w_lo,
                               float* out_raw, double* stats, lwb_stream_t stream)
{
    lwb_conv_plan* plan = nullptr;
    int rc = lwb_conv_plan_create(d, x0_hi, x0_lo, x1_hi, x1_lo, w_hi, w_lo, out_raw, stats, &plan);
    if (rc != LWB_OK) return rc;
    rc = lwb_conv_plan_run(plan, stream);
    lwb_conv_plan_destroy(plan);
    return rc;
}
