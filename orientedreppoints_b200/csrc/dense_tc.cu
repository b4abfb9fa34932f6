// dense_tc.cu - tensor-core (tcgen05 / TMEM / TMA) implicit-GEMM convolution for sm_100a, NHWC bf16.
//
// Replaces the cuDNN convolutions of the reference's backbone / FPN / head towers
// (mmdet/models/backbones/resnet.py:203-239,495-506; necks/fpn.py:138-178;
// anchor_heads/orientedreppoints_head.py:153-168) and, in its gathered-operand variant, the deformable
// im2col + SGEMM of mmdet/ops/dcn/src/deform_conv_cuda.cpp:152-260.
//
// GEMM view: D[M = 128 output pixels, N = BN output channels] += A[M, K] * B[N, K]^T with
// K = taps x Cin walked in 64-channel blocks.  One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer: the A tile of one (tap, channel block) is ONE 4-D box {64 ch, BW, BH, BI}
//               of the NHWC activation (tensor-map element strides = conv stride; out-of-bounds
//               coordinates are zero-filled by the TMA unit = the convolution's zero padding), landing in
//               shared memory as 128 rows x 128 B with the 128-byte swizzle - exactly the canonical
//               K-major operand layout of tcgen05.mma; the B tile is a 2-D box of the [Cout, K] weights.
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN,
//               K=16) x 4 per stage, accumulating fp32 in TMEM; tcgen05.commit releases the stage and,
//               after the last K block, publishes the accumulator.
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns), bias / residual / ReLU, bf16 or fp32 NHWC
//               stores.  TMEM holds two accumulators so the epilogue of tile i overlaps the main loop
//               of tile i+1.
//   warps 6-13  (deformable variant only) A-operand producers: per output pixel and tap the 4-corner
//               bilinear sample of the reference (deform_conv_cuda_kernel.cu:84-115) is computed in
//               fp32 from bf16 features and written to shared memory in the same swizzled layout.
// Several "problems" (the five FPN levels, which share the head weights) are served by ONE launch.
//
// Two operand modes share the kernel.  bf16: activations / weights rounded to bf16, one MMA per K step.
// f16x3 ("split"): every fp32 value is carried as an fp16 pair x = hi + lo (22 significand bits, activations
// stored [N,H,W,2,C]: hi channels then lo channels per pixel) and every product is evaluated as
// hi*hi + lo*hi + hi*lo with three MMAs into the same fp32 TMEM accumulator (the dropped lo*lo term is 2^-22
// relative) - fp32-faithful arithmetic at 1/3 of the tensor-pipe rate; this is the parity mode.  The K loop
// simply runs three "terms" per (tap, channel block); weights are stored [Cout][tap][channel block][2][64] = (hi, lo).
// A tcgen05.mma with M = 128 takes the same ~100 ns whatever N <= 256 (measured: 87 / 97 / 115 ns at N = 64 / 128 / 256), so
// layers with 64 or 128 output channels are issue-bound at a quarter / half of the tensor rate.  For those (BN <= 128) the
// terms are concatenated along N instead of K ("ncat"): per K step  x_hi * [w_hi | w_lo]  is ONE MMA of width 2 BN (main
// product into accumulator columns [0, BN), cross term into [BN, 2 BN)) and  x_lo * w_hi  a second one into [BN, 2 BN) -
// two instructions instead of three, the x_hi / x_lo tiles of a (tap, channel block) are loaded once, and the small cross
// terms own an accumulator (their sum never meets the large main sum before the epilogue adds the two in fp32).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace orp {
namespace {

constexpr int kMaxProb = 5;
constexpr int kStagesMax = 6;
constexpr int kBM = 128;
constexpr int kBK = 64;                       // bf16 elements per K block = 128 bytes = one swizzle row
constexpr int kABytes = kBM * kBK * 2;        // 16 KiB

// division by a runtime constant as multiply-high + shift (Granlund-Montgomery); exact for 0 <= n < 2^31
struct FastDiv {
    uint32_t mul, shr, d;
    __host__ void set(uint32_t div)
    {
        d = div;
        uint32_t l = 0;
        while ((1ull << l) < div) ++l;
        shr = l;
        mul = (uint32_t)(((1ull << 32) * ((1ull << l) - div)) / div + 1);
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return (__umulhi(n, mul) + n) >> shr; }   // n < 2^31: no overflow
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t &q, uint32_t &r) const { q = div(n); r = n - q * d; }
};

struct Problem {
    int N, H, W, Ho, Wo;
    FastDiv fd_tw, fd_th;                     // / tiles_w, / tiles_h
    int BW, BH, BI;                           // tile box: BW*BH*BI == 128 output pixels (powers of two)
    int lbw, lbh;                             // log2(BW), log2(BH)
    int tiles_w, tiles_h, tiles_i, tile_start;
    void *out;                                // bf16 / split-fp16 or fp32 NHWC [N,Ho,Wo,(2,)Cout]
    const __nv_bfloat16 *res;                 // optional residual, same layout as the 16-bit output
    const float *res32;                       // optional fp32 residual (head: refine += init)
    const __nv_bfloat16 *x;                   // activation base (deformable variant; fp16 pairs in split mode)
    const float *offset;                      // deformable: [N,Ho,Wo,2*taps] fp32
    const float *mask;                        // deformable, optional DCNv2 modulation: [N,Ho,Wo,taps] fp32
    double *gn_stats;                         // optional [N, 32, 2] (sum, sum of squares) of the output, GroupNorm(32)
};

struct alignas(64) TcParams {
    CUtensorMap tmA[kMaxProb];
    CUtensorMap tmB;
    CUtensorMap tmOut[kMaxProb];              // bf16 output tensors, box {64 ch, BW, BH, BI} (TMA-store epilogue)
    CUtensorMap tmRes[kMaxProb];              // bf16 residual tensors, same boxes
    CUtensorMap tmI;                          // 64 x 64 bf16 identity (residual add on the tensor core)
    Problem prob[kMaxProb];
    int tma_epi, epi_bufs;                    // TMA epilogue on/off; output staging buffers (1 or 2)
    int epi_merge;                            // split mode, memory-bound layers: hi and lo tiles of a 64-column group leave in ONE pass
    int ncat;                                 // split mode, BN <= 128: terms concatenated along N (see the kernel header)
    int ksplit;                               // split-K over the taps (launches with fewer tiles than SMs): partial sums meet in an fp32 buffer
    FastDiv fd_ks;                            // / ksplit
    long long ks_stride;                      // elements between the partial-sum slabs of consecutive K splits
    int b_resident;                           // short-K layers: the whole weight slab of this CTA's N tile stays in shared memory
    int res_mma;                              // residual added by the tensor core: extra K blocks  R[128x64] * I[64x64]
    int epi_split;                            // epilogue-bound layers: the two epilogue warpgroups work on alternate tiles (one per accumulator buffer)
    int gn_fused;                             // GroupNorm statistics accumulated in the TMA epilogue (Cout == 256)
    int dcat;                                 // deformable split mode: one stage = sampled x_hi | x_lo | w_hi | w_lo of a K block
    int stem;                                 // producers build conv1's 7x7/2 im2col rows from the NCHW fp32 image
    int s2d_stem;                             // conv1 in space-to-depth form (host bookkeeping: 147 useful K of 256)
    int split;                                // f16x3 mode: fp16 (hi, lo) operand pairs, three MMA terms per K block
    float oscale;                             // epilogue multiplier 2^-s undoing the power-of-two weight scale (split mode)
    unsigned int *ovf;                        // split mode: count of outputs beyond the fp16 range (saturated)
    int nprob, num_m_tiles, n_tiles_n, num_tiles;
    FastDiv fd_ntn;                           // / n_tiles_n
    int KH, KW, Cin, cin_blocks, stride, pad, Cout, relu;
    const float *bias;
};

// ----------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, %1;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_4d(void *smem, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *smem, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem, const CUtensorMap *tm, uint64_t *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, 128-byte swizzle: start address >> 4, LBO = 1 (ignored), SBO = 1024 B >> 4, version 1, layout 2
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int ID, int N>
__device__ __forceinline__ void named_bar()
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// split form: issue the load, do independent work, then tmem_ld_wait() before touching the registers
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32])
{
    // the registers are tied to the wait so that no use can be scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

// epilogue activation: 0 none, 1 ReLU, 2 exact (erf) GELU as nn.GELU (swin_transformer.py:24-29)
__device__ __forceinline__ float act_fn(float v, int act)
{
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    return v;
}

struct Ring {
    int stage = 0;
    uint32_t phase = 0;
    int n;
    __device__ explicit Ring(int n_) : n(n_) {}
    __device__ void next() { if (++stage == n) { stage = 0; phase ^= 1; } }
};

// Walk of the main-loop K blocks of one tile: (tap, term, channel block).
//   bf16:                     tap-major, one term.
//   split, TMA operands:      the cross terms (1: x_lo * w_hi, 2: x_hi * w_lo) of ALL taps first, then the main terms
//                             (0: x_hi * w_hi).  The tensor core truncates its fp32 accumulator on every K step (measured:
//                             error grows linearly with the step count, 6e-6 of max at 432 steps); while only the 2^-11
//                             times smaller cross terms have been added the accumulator's ulp - hence that loss - is 2^-11
//                             times smaller too, so the loss of a tile is that of K/16 steps instead of 3K/16.
//   split, deformable:        tap-major, one stage per (tap, channel block) carrying both halves of both operands (dcat): the
//                             producers sample once and hand the K block over once - with one stage per term they could
//                             only start the next gather after the THIRD stage of a block had been freed, which exposed
//                             two thirds of the gather time.
struct KIter {
    int tap, term, cb = 0, phase = 0;
    int tap0, taps, cbn, mode;                             // taps [tap0, taps); mode 0: one stage per K block, 1: split TMA (a stage per term)
    __device__ KIter(int taps_, int cbn_, int mode_, int tap0_ = 0) : tap0(tap0_), taps(taps_), cbn(cbn_), mode(mode_)
    {
        tap = tap0;
        term = (mode == 1) ? 1 : 0;
    }
    __device__ void next()
    {
        if (mode == 0) { if (++cb == cbn) { cb = 0; ++tap; } }
        else if (phase == 0) {
            if (++cb == cbn) { cb = 0; if (++term == 3) { term = 1; if (++tap == taps) { tap = tap0; term = 0; phase = 1; } } }
        } else { if (++cb == cbn) { cb = 0; ++tap; } }
    }
};

__device__ __forceinline__ int ksplit_of(const TcParams &P, int tile)
{
    if (P.ksplit <= 1) return 0;
    uint32_t q, r;
    P.fd_ks.divmod((uint32_t)tile, q, r);
    return (int)r;
}

__device__ __forceinline__ void decode_tile(const TcParams &P, int tile, int &pi, int &wb, int &hb, int &ib, int &nt)
{
    uint32_t mt_u, nt_u;
    if (P.ksplit > 1) tile = (int)P.fd_ks.div((uint32_t)tile);          // tile = (m tile, n tile, k split), k split fastest
    P.fd_ntn.divmod((uint32_t)tile, mt_u, nt_u);
    nt = (int)nt_u;
    const int mt = (int)mt_u;
    pi = 0;
#pragma unroll
    for (int k = 1; k < kMaxProb; ++k)
        if (k < P.nprob && mt >= P.prob[k].tile_start) pi = k;
    const Problem &pr = P.prob[pi];
    const int local = mt - pr.tile_start;
    uint32_t rest, wb_u, ib_u, hb_u;
    pr.fd_tw.divmod((uint32_t)local, rest, wb_u);
    pr.fd_th.divmod(rest, ib_u, hb_u);
    wb = (int)wb_u; hb = (int)hb_u; ib = (int)ib_u;
}

// 64 x 64 bf16 identity, the B operand that adds a residual tile into the accumulator
struct IdentBlock { unsigned short v[64 * 64]; };
constexpr IdentBlock make_ident()
{
    IdentBlock b{};
    for (int i = 0; i < 64; ++i) b.v[i * 64 + i] = 0x3F80;   // bf16 1.0
    return b;
}
__device__ IdentBlock g_ident = make_ident();
// split mode: 2^s * identity in fp16 for s = 0..15 (the weights of a layer carry a power-of-two scale 2^s that
// the epilogue removes, so the residual has to enter the accumulator scaled alike)
struct IdentBlocks16 { unsigned short v[16][64 * 64]; };
constexpr IdentBlocks16 make_ident16()
{
    IdentBlocks16 b{};
    for (int s = 0; s < 16; ++s)
        for (int i = 0; i < 64; ++i) b.v[s][i * 64 + i] = (unsigned short)((15 + s) << 10);   // fp16 2^s
    return b;
}
__device__ IdentBlocks16 g_ident16 = make_ident16();
__device__ unsigned int g_f16_overflow = 0;

// ----------------------------------------------------------------------------------------------- kernel
// BN: accumulator width (32..256).  OUT_F32: fp32 output (head predictions) instead of bf16.
constexpr int kStemPatchBytes = 24576;             // conv1's input patch in dynamic shared memory (5632 floats used)
constexpr int kDP = 256;                           // deformable A-operand producer threads (warps 6 .. 6 + kDP/32 - 1)
constexpr int kDItems = 1024 / kDP;                // (pixel row, 8-channel chunk) items of one 128 x 64 A block per thread
constexpr int kDRound = kDItems / 2;               // items gathered together (two rounds per K block)
// DEFORM: A operand produced by the warps from 6 on (bilinear gather) instead of TMA.
template <int BN, bool OUT_F32, bool DEFORM>
__global__ void __launch_bounds__(DEFORM ? 192 + kDP : 320, 1)
conv_tc_kernel(const __grid_constant__ TcParams P, int stages)
{
    // warp roles.  plain: 0 TMA | 1 MMA | 2-9 epilogue (two warps per TMEM lane quarter, splitting the columns).
    // deformable / stem: 0 TMA(B) | 1 MMA | 2-5 epilogue | 6-13 A-operand producers.
    constexpr int kEpiWarps = DEFORM ? 4 : 8;
    constexpr int kEpiThreads = kEpiWarps * 32;
    constexpr int kWG = kEpiWarps / 4;                 // epilogue warps per lane quarter
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic: [stages][A 16K | B BN*128]  then the epilogue staging tile [128 rows][HC*2 + 16 B]
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kBBytes = BN * kBK * 2;
    // b_resident: [B slab: kblocks x kBBytes] then A-only stages; otherwise every stage carries A | B
    const int kStageBytes = (P.ncat || P.dcat) ? (P.b_resident ? 2 * kABytes : 2 * kABytes + 2 * kBBytes)
                                   : (P.b_resident ? kABytes : kABytes + kBBytes);
    const int kblocks_all = P.KH * P.KW * P.cin_blocks * (P.split ? 2 : 1);   // weight K blocks (hi and lo halves in split mode)
    uint8_t *ident = smem;                             // [8 KiB] identity block when res_mma
    if (P.res_mma) smem += 8192;
    float *s_patch = reinterpret_cast<float *>(smem);  // [24 KiB] conv1's input patch (stem transform only)
    if (DEFORM && P.stem) smem += kStemPatchBytes;
    uint8_t *bres = smem;
    if (P.b_resident) smem += (size_t)kblocks_all * kBBytes;
    constexpr int HC = BN < 64 ? BN : 64;              // columns staged per epilogue pass
    constexpr int kPitch = HC * 2 + 16;                // bytes per staged row (+16: conflict-free 16-byte accesses)
    uint8_t *stage_out = smem + (size_t)stages * kStageBytes;
    __shared__ uint64_t bars[2 * kStagesMax + 4];
    __shared__ __align__(16) float s_bias_all[2][256];
    __shared__ uint64_t bres_bar;                // resident weight slab landed
    __shared__ uint32_t tmem_slot_s;
    uint64_t *full = bars;                       // [stages]  TMA bytes landed (+ producer arrivals when DEFORM)
    uint64_t *empty = bars + kStagesMax;         // [stages]  MMA finished reading the stage
    uint64_t *tfull = bars + 2 * kStagesMax;     // [2] accumulator ready
    uint64_t *tempty = tfull + 2;                // [2] accumulator drained
    uint32_t *tmem_slot = &tmem_slot_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // two accumulator buffers; BN <= 128 leaves room for the ncat layout (main | cross columns per buffer)
    constexpr uint32_t kAccCols = BN <= 128 ? 4 * BN : 2 * BN;
    constexpr uint32_t kTmemCols = (kAccCols <= 32) ? 32 : (kAccCols <= 64) ? 64 : (kAccCols <= 128) ? 128 : (kAccCols <= 256) ? 256 : 512;

    if (warp == 0 && elect_one()) {
#pragma unroll 1
        for (int p = 0; p < P.nprob; ++p)
            asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmA[p]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmB) : "memory");
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(&full[s], DEFORM ? 1 + (P.stem ? 8 : kDP / 32) : 1);          // TMA expect_tx arrival (+ one arrival per producer warp)
                mbar_init(&empty[s], 1);
            }
            for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], (kWG == 2 && P.epi_split) ? 4 : kEpiWarps); }
            mbar_init(&bres_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int taps_per = (P.KH * P.KW) / (P.ksplit > 1 ? P.ksplit : 1);              // taps of one K split
    const int kblocks = taps_per * P.cin_blocks * ((P.split && !P.ncat && !P.dcat) ? 3 : 1);     // main-loop K blocks (stages) per tile
    // Programmatic dependent launch: the next kernel in the stream may start its CTAs (barrier init, TMEM allocation,
    // descriptor prefetch - the code above) on SMs this grid has already left; nothing above touches global memory,
    // and everything below (loads AND stores) comes after the wait for the preceding grid to complete and flush.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            Ring r(stages);
            if ((P.b_resident || P.res_mma) && (int)blockIdx.x < P.num_tiles) {
                // gridDim.x is a multiple of n_tiles_n, so every tile of this CTA has the same N tile
                const int nt0 = blockIdx.x % P.n_tiles_n;
                mbar_expect_tx(&bres_bar, (uint32_t)((P.b_resident ? kblocks_all * kBBytes : 0) + (P.res_mma ? 8192 : 0)));   // kblocks_all counts weight blocks
                if (P.res_mma) tma_load_2d(ident, &P.tmI, &bres_bar, 0, 0);
                if (P.b_resident)
                    for (int kb = 0; kb < kblocks_all; ++kb)
                        tma_load_2d(bres + (size_t)kb * kBBytes, &P.tmB, &bres_bar, kb * kBK, nt0 * BN);
            }
            const int wterms = P.split ? 2 : 1, rterms = P.split ? 2 : 1;
            for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
                int pi, wb, hb, ib, nt;
                decode_tile(P, tile, pi, wb, hb, ib, nt);
                const Problem &pr = P.prob[pi];
                const int w0 = wb * pr.BW * P.stride - P.pad, h0 = hb * pr.BH * P.stride - P.pad, i0 = ib * pr.BI;
                const int tap_lo = ksplit_of(P, tile) * taps_per;
                KIter it(tap_lo + taps_per, P.cin_blocks, (P.split && !P.ncat && !P.dcat) ? 1 : 0, tap_lo);
                for (int j = 0; j < kblocks; ++j, it.next()) {
                    const int kh = it.tap / P.KW, kw = it.tap - kh * P.KW;
                    mbar_wait(&empty[r.stage], r.phase ^ 1);
                    uint8_t *sa = smem + (size_t)r.stage * kStageBytes;
                    const int wblk = ((it.tap * P.cin_blocks + it.cb) * wterms) * kBK;      // K offset of the (hi, lo) weight blocks
                    if (P.ncat || P.dcat) {
                        // one stage = x_hi tile | x_lo tile | w_hi block | w_lo block of this (tap, channel block); the
                        // deformable variant's x halves come from the producer warps
                        mbar_expect_tx(&full[r.stage], (DEFORM ? 0 : 2 * kABytes) + (P.b_resident ? 0 : 2 * kBBytes));
                        if (!DEFORM) {
                            tma_load_5d(sa, &P.tmA[pi], &full[r.stage], it.cb * kBK, 0, w0 + kw, h0 + kh, i0);
                            tma_load_5d(sa + kABytes, &P.tmA[pi], &full[r.stage], it.cb * kBK, 1, w0 + kw, h0 + kh, i0);
                        }
                        if (!P.b_resident) {
                            tma_load_2d(sa + 2 * kABytes, &P.tmB, &full[r.stage], wblk, nt * BN);
                            tma_load_2d(sa + 2 * kABytes + kBBytes, &P.tmB, &full[r.stage], wblk + kBK, nt * BN);
                        }
                    } else {
                        mbar_expect_tx(&full[r.stage], DEFORM ? kBBytes : (P.b_resident ? kABytes : kABytes + kBBytes));
                        if (!DEFORM) tma_load_5d(sa, &P.tmA[pi], &full[r.stage], it.cb * kBK, it.term == 1 ? 1 : 0, w0 + kw, h0 + kh, i0);
                        if (!P.b_resident)
                            tma_load_2d(sa + kABytes, &P.tmB, &full[r.stage], wblk + (it.term == 2 ? kBK : 0), nt * BN);
                    }
                    r.next();
                }
                if (!DEFORM && P.res_mma) {
                    // residual: one extra K block per 64 output channels (two in split mode: hi and lo), the A operand
                    // is the residual tile itself
                    for (int g = 0; g < BN / 64 && nt * BN + g * 64 < P.Cout; ++g)
                        for (int t = 0; t < rterms; ++t) {
                            mbar_wait(&empty[r.stage], r.phase ^ 1);
                            mbar_expect_tx(&full[r.stage], kABytes);
                            tma_load_5d(smem + (size_t)r.stage * kStageBytes, &P.tmRes[pi], &full[r.stage], nt * BN + g * 64, t,
                                        wb * pr.BW, hb * pr.BH, ib * pr.BI);
                            r.next();
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        // instruction descriptor: D fp32 (bit 4), A/B format (bits 7-9 / 10-12: 0 = fp16, 1 = bf16), N >> 3, M >> 4
        const uint32_t fmt = P.split ? 0u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        const uint32_t idesc64 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)((2 * BN > 256 ? 256 : 2 * BN) >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);   // N = 2 BN (ncat)
        Ring r(stages);
        int acc = 0;
        uint32_t acc_phase = 0;
        const int wterms = P.split ? 2 : 1, rterms = P.split ? 2 : 1;
        if ((P.b_resident || P.res_mma) && (int)blockIdx.x < P.num_tiles) mbar_wait(&bres_bar, 0);
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
            int res_groups = 0;
            if (!DEFORM && P.res_mma) {
                const int nt = tile - (int)P.fd_ntn.div((uint32_t)tile) * P.n_tiles_n;
                for (int g = 0; g < BN / 64 && nt * BN + g * 64 < P.Cout; ++g) ++res_groups;
            }
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t accw = P.ncat ? 2u * BN : (uint32_t)BN;          // TMEM columns of one accumulator buffer
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * accw;
            const int tap_lo = ksplit_of(P, tile) * taps_per;
            KIter it(tap_lo + taps_per, P.cin_blocks, (P.split && !P.ncat && !P.dcat) ? 1 : 0, tap_lo);   // same walk as the producer
            for (int kb = 0; kb < kblocks; ++kb, it.next()) {
                mbar_wait(&full[r.stage], r.phase);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + (size_t)r.stage * kStageBytes);
                    const uint64_t da = make_desc_sw128(sa);
                    const int slab = (it.tap * P.cin_blocks + it.cb) * wterms + (it.term == 2 ? 1 : 0);
                    if (P.ncat) {
                        // [w_hi | w_lo] are adjacent in the stage (and in the resident slab): one operand of 2 BN rows
                        const uint64_t dl = make_desc_sw128(sa + kABytes);
                        const uint64_t db = make_desc_sw128(P.b_resident ? smem_u32(bres + (size_t)slab * kBBytes) : sa + 2 * kABytes);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc2, (kb | k) ? 1u : 0u);        // x_hi * [w_hi | w_lo]
                            umma_bf16(d_tmem + (uint32_t)BN, dl + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);          // x_lo * w_hi -> cross columns
                        }
                    } else if (P.dcat) {
                        const uint64_t dl = make_desc_sw128(sa + kABytes);
                        const uint64_t dbh = make_desc_sw128(sa + 2 * kABytes), dbl = make_desc_sw128(sa + 2 * kABytes + kBBytes);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            umma_bf16(d_tmem, dl + (uint64_t)(k * 2), dbh + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);   // x_lo * w_hi
                            umma_bf16(d_tmem, da + (uint64_t)(k * 2), dbl + (uint64_t)(k * 2), idesc, 1u);                   // x_hi * w_lo
                            umma_bf16(d_tmem, da + (uint64_t)(k * 2), dbh + (uint64_t)(k * 2), idesc, 1u);                   // x_hi * w_hi
                        }
                    } else {
                        const uint64_t db = make_desc_sw128(P.b_resident ? smem_u32(bres + (size_t)slab * kBBytes) : sa + kABytes);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)
                            umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[r.stage]);
                    if (kb == kblocks - 1 && res_groups == 0) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                r.next();
            }
            for (int g = 0; g < res_groups * rterms; ++g) {
                // accumulator columns [64 g, 64 g + 64) += residual tile * (scaled) identity
                mbar_wait(&full[r.stage], r.phase);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_desc_sw128(smem_u32(smem + (size_t)r.stage * kStageBytes));
                    const uint64_t db = make_desc_sw128(smem_u32(ident));
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_bf16(d_tmem + (uint32_t)((g / rterms) * 64), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc64, 1u);
                    umma_commit(&empty[r.stage]);
                    if (g == res_groups * rterms - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                r.next();
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp < 2 + kEpiWarps) {
        // ===================================================== epilogue (TMEM lane quarter = warp % 4)
        const int q = warp & 3;
        const int wg = (warp - 2) >> 2;                    // epilogue warpgroup (0 or 1)
        // Two ways to use eight epilogue warps.  Default: both warpgroups work on the same tile, each warp owning one
        // 32-column half of a 64-column pass.  Split (epilogue-bound layers): the warpgroups are independent, group g
        // drains accumulator buffer g (= every second tile of this CTA) with its own staging tiles, barriers and bias
        // copy, so the fixed latencies of one group's pass (barriers, TMEM load, store issue) overlap the other's.
        const bool split = (kWG == 2) && P.epi_split;
        const int grp = split ? wg : 0;
        const int et = split ? ((threadIdx.x - 64) & 127) : (threadIdx.x - 64);   // thread index inside the group
        const int nthr = split ? 128 : kEpiThreads;
        const int bar_a = 1 + 3 * grp, bar_b = 3 + 2 * grp;                     // named barriers (1,3) / (4,5); 2 = producers
        const int ch0 = split ? 0 : wg, chs = split ? 1 : kWG;
        float *s_bias = s_bias_all[grp];
        auto bar_sync = [](int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); };
        int ob = 0;                                        // staging ring position (TMA epilogue)
        int bias_nt = -1;                                  // N tile whose bias slice is staged in s_bias
        int acc = split ? wg : 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x + (split ? wg * (int)gridDim.x : 0); tile < P.num_tiles;
             tile += (split ? 2 : 1) * (int)gridDim.x) {
            int pi, wb, hb, ib, nt;
            decode_tile(P, tile, pi, wb, hb, ib, nt);
            const Problem &pr = P.prob[pi];
            const int rrow = q * 32 + lane;
            auto row_pixel = [&](int r, size_t &pix) -> bool {
                const int iw = r & (pr.BW - 1), ih = (r >> pr.lbw) & (pr.BH - 1), ii = r >> (pr.lbw + pr.lbh);
                const int w = wb * pr.BW + iw, h = hb * pr.BH + ih, n = ib * pr.BI + ii;
                pix = ((size_t)n * pr.Ho + h) * pr.Wo + w;
                return (w < pr.Wo) && (h < pr.Ho) && (n < pr.N);
            };
            size_t pix;
            const bool valid = row_pixel(rrow, pix);
            if (OUT_F32) {
                // small fp32 outputs (head predictions, Cout <= 32): straight from registers
                mbar_wait(&tfull[acc], acc_phase);
                tcgen05_fence_after();
#pragma unroll 1
                for (int ch = wg; ch < BN / 32; ch += kWG) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + ch * 32), v);
                    const int c0 = nt * BN + ch * 32;
                    if (valid && c0 < P.Cout) {
                        float *op = reinterpret_cast<float *>(pr.out) + pix * P.Cout + c0;
                        const float *rp = pr.res32 ? pr.res32 + pix * P.Cout + c0 : nullptr;
                        if (P.ksplit > 1) {
                            // split-K: this tile holds the sum over its taps only and writes it to its own slab of the fp32
                            // workspace; the finishing kernel adds the slabs in a fixed order (bit-reproducible, unlike atomics)
                            // and applies bias / activation / the hi-lo split
                            float *sp = op + (long long)ksplit_of(P, tile) * P.ks_stride;
                            for (int j = 0; j < 32; ++j) {
                                if (c0 + j >= P.Cout) break;
                                sp[j] = __uint_as_float(v[j]) * P.oscale;
                            }
                        } else
                        for (int j = 0; j < 32; ++j) {
                            if (c0 + j >= P.Cout) break;
                            float f = __uint_as_float(v[j]) * P.oscale;
                            if (P.bias) f += P.bias[c0 + j];
                            if (pr.res && !P.split) f += __bfloat162float(pr.res[pix * P.Cout + c0 + j]);
                            if (rp) f += rp[j];
                            f = act_fn(f, P.relu);
                            op[j] = f;
                        }
                    }
                }
            } else if (P.tma_epi) {
                // bf16 outputs through the TMA unit, in 64-channel passes: results go to a 128-byte-swizzled staging
                // tile and leave with cp.async.bulk.tensor (coalescing and partial-tile clipping by hardware).  A
                // residual is already in the accumulator (added by the tensor core, see the MMA warp).
                // With eight epilogue warps each warp owns one 32-column half of the pass.
                const uint32_t slot = P.epi_merge ? 32768u : 16384u;     // a staging slot: one 16 KiB tile (hi | lo when merged)
                const uint32_t obuf_u = smem_u32(stage_out) + (uint32_t)(grp * P.epi_bufs) * slot;   // [epi_bufs][slot] per group
                const uint32_t bias_u = smem_u32(s_bias);
                const bool io = (et == 0);
                constexpr int kPasses = BN / 64;
                // split mode: a hi pass and a lo pass per 64 columns (compute-bound layers: one 16 KiB staging tile), or both
                // halves in one pass (memory-bound layers: half the TMEM reads, barriers and fp32 work per output value)
                const int oterms = (P.split && !P.epi_merge) ? 2 : 1;
                if (nt != bias_nt) {                                     // bias slice changes only with the N tile
                    bar_sync(bar_a, nthr);                               // previous tile's bias reads are done
                    for (int c = et; c < BN; c += nthr) s_bias[c] = (P.bias && nt * BN + c < P.Cout) ? P.bias[nt * BN + c] : 0.f;
                    bias_nt = nt;
                }
                float amax = 0.f;
#pragma unroll 1
                for (int pass = 0; pass < kPasses * oterms; ++pass) {
                    const int half = oterms == 2 ? (pass >> 1) : pass, oterm = oterms == 2 ? (pass & 1) : 0;
                    if (io) {
                        if (P.epi_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    bar_sync(bar_a, nthr);                               // staging buffer `ob` is free, bias staged
                    if (pass == 0) {
                        mbar_wait(&tfull[acc], acc_phase);
                        tcgen05_fence_after();
                    }
                    float gn_s = 0.f, gn_q = 0.f;
                    const uint32_t orow = obuf_u + (uint32_t)ob * slot + (uint32_t)rrow * 128u;
#pragma unroll 1
                    for (int ch = ch0; ch < 2; ch += chs) {
                        uint32_t v[32];
                        const uint32_t accw = P.ncat ? 2u * BN : (uint32_t)BN;
                        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * accw + (uint32_t)(half * 64 + ch * 32);
                        tmem_ld32_issue(tcol, v);
                        // bias slice of these 32 columns: eight back-to-back shared loads, in flight with the TMEM load
                        uint4 bu[8];
#pragma unroll
                        for (int j8 = 0; j8 < 8; ++j8) bu[j8] = lds128(bias_u + (uint32_t)(half * 64 + ch * 32 + j8 * 4) * 4u);
                        tmem_ld_wait(v);
                        if (P.ncat) {
                            // main + cross accumulators meet here, in fp32 with round-to-nearest
                            uint32_t vc[32];
                            tmem_ld32(tcol + (uint32_t)BN, vc);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(vc[j]));
                        }
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const int c16 = ch * 4 + j4;                  // 16-byte chunk inside the 128-byte row
                            const uint32_t sw = (uint32_t)(c16 ^ (rrow & 7)) << 4;
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j4 * 8 + j]);
                            const uint4 b0 = bu[2 * j4], b1 = bu[2 * j4 + 1];
                            if (P.split) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] *= P.oscale;      // exact: power of two
                            }
                            f[0] += __uint_as_float(b0.x); f[1] += __uint_as_float(b0.y);
                            f[2] += __uint_as_float(b0.z); f[3] += __uint_as_float(b0.w);
                            f[4] += __uint_as_float(b1.x); f[5] += __uint_as_float(b1.y);
                            f[6] += __uint_as_float(b1.z); f[7] += __uint_as_float(b1.w);
                            if (P.relu == 2) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] = act_fn(f[j], 2);
                            }
                            uint32_t pk[4];
                            if (P.split) {
                                // x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significand bits.  Values beyond the fp16
                                // range become inf / nan in the output AND are counted (P.ovf): never silent.
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    if (P.relu == 1) f[j] = fmaxf(f[j], 0.f);
                                    if (oterm == 0) amax = fmaxf(amax, fabsf(f[j]));
                                }
                                uint32_t pl[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const __half2 h2 = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
                                    pk[k] = *reinterpret_cast<const uint32_t *>(&h2);
                                    if (P.epi_merge || oterm) {
                                        const float2 hf = __half22float2(h2);
                                        const __half2 l2 = __floats2half2_rn(f[2 * k] - hf.x, f[2 * k + 1] - hf.y);
                                        pl[k] = *reinterpret_cast<const uint32_t *>(&l2);
                                    }
                                }
                                if (P.epi_merge) sts128(orow + 16384u + sw, make_uint4(pl[0], pl[1], pl[2], pl[3]));
                                else if (oterm) { pk[0] = pl[0]; pk[1] = pl[1]; pk[2] = pl[2]; pk[3] = pl[3]; }
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                                    // ReLU after rounding: rounding is monotone and keeps the sign, so the result is the same
                                    if (P.relu == 1) b2 = __hmax2(b2, __floats2bfloat162_rn(0.f, 0.f));
                                    pk[k] = *reinterpret_cast<uint32_t *>(&b2);
                                }
                            }
                            sts128(orow + sw, make_uint4(pk[0], pk[1], pk[2], pk[3]));
                            if (P.gn_fused && oterm == 0) {
                                // one 16-byte chunk = 8 channels = one GroupNorm group (Cout 256 / 32 groups)
                                float gs = 0.f, gq = 0.f;
                                if (valid) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) { gs += f[j]; gq = fmaf(f[j], f[j], gq); }
                                }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    gs += __shfl_xor_sync(0xffffffffu, gs, o);
                                    gq += __shfl_xor_sync(0xffffffffu, gq, o);
                                }
                                if (lane == c16) { gn_s = gs; gn_q = gq; }      // lane g keeps group g of this pass
                            }
                        }
                    }
                    if (P.gn_fused && oterm == 0 && lane < 8 && (chs == 1 || (lane >> 2) == wg)) {
                        // the 32 rows of a warp belong to one image (host guarantees BW*BH >= 32)
                        const int n_img = ib * pr.BI + ((q * 32) >> (pr.lbw + pr.lbh));
                        if (n_img < pr.N) {
                            double *st = pr.gn_stats + ((size_t)n_img * 32 + (size_t)(nt * BN + half * 64) / 8 + lane) * 2;
                            atomicAdd(st, (double)gn_s);
                            atomicAdd(st + 1, (double)gn_q);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    bar_sync(bar_b, nthr);                               // staging written by all
                    if (io) {
                        asm volatile(
                            "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                            ::"l"(&P.tmOut[pi]), "r"(obuf_u + (uint32_t)ob * slot), "r"(nt * BN + half * 64), "r"(oterm),
                              "r"(wb * pr.BW), "r"(hb * pr.BH), "r"(ib * pr.BI) : "memory");
                        if (P.epi_merge)
                            asm volatile(
                                "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                                ::"l"(&P.tmOut[pi]), "r"(obuf_u + (uint32_t)ob * slot + 16384u), "r"(nt * BN + half * 64), "r"(1),
                                  "r"(wb * pr.BW), "r"(hb * pr.BH), "r"(ib * pr.BI) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (++ob == P.epi_bufs) ob = 0;
                }
                if (P.split && valid && amax > 65504.f) atomicAdd(P.ovf, 1u);
            } else {
                // bf16 outputs: residual tile prefetched into shared memory with coalesced cp.async while the
                // MMA is still running, results written back to the same staging tile, then stored coalesced
                const bool vec_ok = (P.Cout & 7) == 0;
                named_bar<1, kEpiThreads>();
                for (int c = et; c < BN; c += kEpiThreads) s_bias[c] = (P.bias && nt * BN + c < P.Cout) ? P.bias[nt * BN + c] : 0.f;
                named_bar<1, kEpiThreads>();
#pragma unroll 1
                for (int half = 0; half < BN / HC; ++half) {
                    const int cbase = nt * BN + half * HC;           // first output channel of this pass
                    constexpr int kChunksPerRow = HC / 8;            // 16-byte chunks per staged row
                    if (pr.res && vec_ok) {
                        for (int c = et; c < 128 * kChunksPerRow; c += kEpiThreads) {
                            const int r = c / kChunksPerRow, k16 = c - r * kChunksPerRow;
                            size_t rp;
                            const bool ok = row_pixel(r, rp) && (cbase + k16 * 8 < P.Cout);
                            const uint32_t dst = smem_u32(stage_out + (size_t)r * kPitch + k16 * 16);
                            if (ok) {
                                const __nv_bfloat16 *src = pr.res + rp * P.Cout + cbase + k16 * 8;
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                            }
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    }
                    if (half == 0) {
                        mbar_wait(&tfull[acc], acc_phase);
                        tcgen05_fence_after();
                    }
                    if (pr.res && vec_ok) {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                        named_bar<1, kEpiThreads>();
                    }
#pragma unroll 1
                    for (int ch = wg; ch < HC / 32; ch += kWG) {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * HC + ch * 32), v);
                        const int c0 = cbase + ch * 32;
                        uint4 *sp = reinterpret_cast<uint4 *>(stage_out + (size_t)rrow * kPitch + ch * 64);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j4 * 8 + j]);
                            {
                                const float4 b0 = *reinterpret_cast<const float4 *>(&s_bias[half * HC + ch * 32 + j4 * 8]);
                                const float4 b1 = *reinterpret_cast<const float4 *>(&s_bias[half * HC + ch * 32 + j4 * 8 + 4]);
                                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                            }
                            if (pr.res && !vec_ok && valid) {
                                for (int j = 0; j < 8; ++j)
                                    if (c0 + j4 * 8 + j < P.Cout) f[j] += __bfloat162float(pr.res[pix * P.Cout + c0 + j4 * 8 + j]);
                            }
                            if (pr.res && vec_ok && valid) {
                                const uint4 u = sp[j4];
                                const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float2 r2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&uu[k]));
                                    f[2 * k] += r2.x;
                                    f[2 * k + 1] += r2.y;
                                }
                            }
                            if (P.relu) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] = act_fn(f[j], P.relu);
                            }
                            uint32_t pk[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                                pk[k] = *reinterpret_cast<uint32_t *>(&b2);
                            }
                            sp[j4] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                    named_bar<1, kEpiThreads>();
                    // coalesced copy-out: 16 threads cover one 256-byte row segment
                    for (int c = et; c < 128 * kChunksPerRow; c += kEpiThreads) {
                        const int r = c / kChunksPerRow, k16 = c - r * kChunksPerRow;
                        size_t rp;
                        if (row_pixel(r, rp) && (cbase + k16 * 8 < P.Cout)) {
                            const uint4 u = *reinterpret_cast<const uint4 *>(stage_out + (size_t)r * kPitch + k16 * 16);
                            __nv_bfloat16 *op = reinterpret_cast<__nv_bfloat16 *>(pr.out) + rp * P.Cout + cbase + k16 * 8;
                            if (vec_ok) {
                                *reinterpret_cast<uint4 *>(op) = u;
                            } else {
                                const __nv_bfloat16 *e8 = reinterpret_cast<const __nv_bfloat16 *>(&u);
                                for (int j = 0; j < 8; ++j) if (cbase + k16 * 8 + j < P.Cout) op[j] = e8[j];
                            }
                        }
                    }
                    named_bar<1, kEpiThreads>();
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (split) acc_phase ^= 1;                       // this group always drains the same accumulator buffer
            else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (P.tma_epi && et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores drained before exit
    } else if (DEFORM) {
        // ===================================================== deformable A-operand producers (warps 6-13)
        // Per tap: threads 0-127 compute the bilinear parameters of their output pixel (4 weights + 4 element
        // offsets, deform_conv_cuda_kernel.cu:84-115 + the validity test of :229) into a shared table; then all
        // 256 threads gather: 8 consecutive lanes fetch the 8 x 16 B of one pixel's 64-channel block for each of
        // the 4 corners (every load instruction covers whole 128-byte lines), blend in fp32, round to bf16 and
        // store to the stage in the 128-byte-swizzled K-major layout the MMA expects.
        __shared__ float4 s_w[2][128];
        __shared__ uint2 s_wh[2][128];                     // the same four weights as fp16 (split mode: blend of the lo halves)
        __shared__ int4 s_o[2][128];
        const int pt = threadIdx.x - 192;                  // 0..kDP-1 (the stem transform uses the first 256)
        Ring r(stages);
        int tb = 0;
        if (P.stem && pt >= 256) {
            // conv1's operand is built from a shared-memory patch: 8 warps are plenty
        } else if (P.stem) {
            // conv1 (resnet.py:495, 7x7 stride 2 pad 3, 3 input channels) as a GEMM with K = 192 (147 used),
            // k = (kh*7 + kw)*3 + c.  Per tile the input patch ((2*BH+5) x (2*BW+5) pixels x 3 channels per image
            // of the tile) is staged once in shared memory with coalesced loads of the NCHW fp32 image
            // (pr.offset); the three 64-wide K blocks of A rows are then built from shared memory.
            __shared__ __align__(16) int s_koff[192];                                      // k -> offset inside the patch (-1: k >= 147)
            {
                const Problem &p0 = P.prob[0];
                const int pw0 = 2 * p0.BW + 5, ph0 = 2 * p0.BH + 5;
                if (pt < 192) {
                    const int tap = pt / 3, c = pt - tap * 3, kh = tap / 7, kw = tap - kh * 7;
                    s_koff[pt] = pt < 147 ? (c * ph0 + kh) * pw0 + kw : -1;
                }
            }
            for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
                int pi, wb, hb, ib, nt;
                decode_tile(P, tile, pi, wb, hb, ib, nt);
                const Problem &pr = P.prob[pi];
                const float *img = pr.offset;
                const int pw = 2 * pr.BW + 5, ph = 2 * pr.BH + 5;            // patch extent per image
                const int x0 = wb * pr.BW * 2 - 3, y0 = hb * pr.BH * 2 - 3;
                const int per_img = 3 * ph * pw;
                asm volatile("bar.sync 2, 256;" ::: "memory");               // previous tile's reads of the patch are done
                {
                    // one patch row (image, channel, y) per iteration: the row decode is warp-uniform, the loads of
                    // different rows are independent (unrolled for memory-level parallelism) and coalesced along x
                    const int nrows = pr.BI * 3 * ph;
#pragma unroll 7
                    for (int rowi = 0; rowi < nrows; ++rowi) {
                        const int ii = rowi / (3 * ph), rem = rowi - ii * 3 * ph;
                        const int c = rem / ph, py = rem - c * ph;
                        const int n = ib * pr.BI + ii, yy = y0 + py;
                        const bool rok = (n < pr.N) && (yy >= 0) && (yy < pr.H);
                        const float *src = img + (((size_t)(rok ? n : 0) * 3 + c) * pr.H + (rok ? yy : 0)) * pr.W;
                        for (int px = pt; px < pw; px += 256) {
                            const int xx = x0 + px;
                            s_patch[rowi * pw + px] = (rok && xx >= 0 && xx < pr.W) ? __ldg(src + xx) : 0.f;
                        }
                    }
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                int pbase[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int row = (pt + it * 256) >> 3;
                    const int iw = row & (pr.BW - 1), ih = (row >> pr.lbw) & (pr.BH - 1), ii = row >> (pr.lbw + pr.lbh);
                    pbase[it] = ii * per_img + (ih * 2) * pw + iw * 2;
                }
                const int c16 = pt & 7;
                for (int kb = 0; kb < 3; ++kb) {
                    mbar_wait(&empty[r.stage], r.phase ^ 1);
                    uint8_t *sa = smem + (size_t)r.stage * kStageBytes;
                    const int4 o0 = *reinterpret_cast<const int4 *>(&s_koff[kb * 64 + c16 * 8]);
                    const int4 o1 = *reinterpret_cast<const int4 *>(&s_koff[kb * 64 + c16 * 8 + 4]);
                    const int off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int row = (pt + it * 256) >> 3;
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? s_patch[pbase[it] + off[j]] : 0.f;
                        uint32_t pk[4];
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) {
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * k2], v[2 * k2 + 1]);
                            pk[k2] = *reinterpret_cast<uint32_t *>(&b2);
                        }
                        *reinterpret_cast<uint4 *>(sa + (size_t)row * 128 + ((c16 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if ((pt & 31) == 0) mbar_arrive(&full[r.stage]);
                    r.next();
                }
            }
        } else
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
            int pi, wb, hb, ib, nt;
            decode_tile(P, tile, pi, wb, hb, ib, nt);
            const Problem &pr = P.prob[pi];
            bool valid = false;
            int w = 0, h = 0, n = 0;
            if (pt < 128) {
                const int iw = pt & (pr.BW - 1), ih = (pt >> pr.lbw) & (pr.BH - 1), ii = pt >> (pr.lbw + pr.lbh);
                w = wb * pr.BW + iw; h = hb * pr.BH + ih; n = ib * pr.BI + ii;
                valid = (w < pr.Wo) && (h < pr.Ho) && (n < pr.N);
            }
            const int taps = P.KH * P.KW;
            const size_t opix = ((size_t)(valid ? n : 0) * pr.Ho + (valid ? h : 0)) * pr.Wo + (valid ? w : 0);
            const float *offp = pr.offset + opix * (2 * taps);
            const float *mskp = pr.mask ? pr.mask + opix * taps : nullptr;
            const int cpp = P.Cin * (P.split ? 2 : 1);                  // 16-bit elements per pixel (hi and lo halves in split mode)
            const int img0 = (valid ? n : 0) * pr.H * pr.W * cpp;
            for (int tap = 0; tap < taps; ++tap) {
                if (pt < 128) {
                    const int kh = tap / P.KW, kw = tap - kh * P.KW;
                    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
                    int4 ov = make_int4(img0, img0, img0, img0);
                    if (valid) {
                        const float h_im = (float)(h * P.stride - P.pad + kh) + offp[2 * tap];
                        const float w_im = (float)(w * P.stride - P.pad + kw) + offp[2 * tap + 1];
                        if (h_im > -1.f && w_im > -1.f && h_im < (float)pr.H && w_im < (float)pr.W) {
                            const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
                            const int h_high = h_low + 1, w_high = w_low + 1;
                            const float lh = h_im - (float)h_low, lw = w_im - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
                            if (h_low >= 0 && w_low >= 0) { wv.x = hh * hw; ov.x = img0 + (h_low * pr.W + w_low) * cpp; }
                            if (h_low >= 0 && w_high <= pr.W - 1) { wv.y = hh * lw; ov.y = img0 + (h_low * pr.W + w_high) * cpp; }
                            if (h_high <= pr.H - 1 && w_low >= 0) { wv.z = lh * hw; ov.z = img0 + (h_high * pr.W + w_low) * cpp; }
                            if (h_high <= pr.H - 1 && w_high <= pr.W - 1) { wv.w = lh * lw; ov.w = img0 + (h_high * pr.W + w_high) * cpp; }
                            if (mskp) {
                                // DCNv2 (modulated_deformable_im2col_gpu_kernel, deform_conv_cuda_kernel.cu:570-633): the sample
                                // is multiplied by the mask value of (pixel, tap); folded into the four corner weights
                                const float m = mskp[tap];
                                wv.x *= m; wv.y *= m; wv.z *= m; wv.w *= m;
                            }
                        }
                    }
                    s_w[tb][pt] = wv;
                    s_o[tb][pt] = ov;
                    if (P.split) {
                        const __half2 w01 = __floats2half2_rn(wv.x, wv.y), w23 = __floats2half2_rn(wv.z, wv.w);
                        s_wh[tb][pt] = make_uint2(*reinterpret_cast<const uint32_t *>(&w01), *reinterpret_cast<const uint32_t *>(&w23));
                    }
                }
                asm volatile("bar.sync 2, %0;" ::"n"(kDP) : "memory");
                for (int cb = 0; cb < P.cin_blocks; ++cb) {
                    if (!P.split) {
                        mbar_wait(&empty[r.stage], r.phase ^ 1);
                        uint8_t *sa = smem + (size_t)r.stage * kStageBytes;
                        uint4 u[kDItems][4];
#pragma unroll
                        for (int it = 0; it < kDItems; ++it) {
                            const int item = pt + it * kDP, row = item >> 3, c16 = item & 7;
                            const int4 ov = s_o[tb][row];
                            const int co = cb * kBK + c16 * 8;
                            u[it][0] = *reinterpret_cast<const uint4 *>(pr.x + ov.x + co);
                            u[it][1] = *reinterpret_cast<const uint4 *>(pr.x + ov.y + co);
                            u[it][2] = *reinterpret_cast<const uint4 *>(pr.x + ov.z + co);
                            u[it][3] = *reinterpret_cast<const uint4 *>(pr.x + ov.w + co);
                        }
#pragma unroll
                        for (int it = 0; it < kDItems; ++it) {
                            const int item = pt + it * kDP, row = item >> 3, c16 = item & 7;
                            const float4 wv = s_w[tb][row];
                            const uint32_t *a1 = reinterpret_cast<const uint32_t *>(&u[it][0]);
                            const uint32_t *a2 = reinterpret_cast<const uint32_t *>(&u[it][1]);
                            const uint32_t *a3 = reinterpret_cast<const uint32_t *>(&u[it][2]);
                            const uint32_t *a4 = reinterpret_cast<const uint32_t *>(&u[it][3]);
                            uint32_t pk[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&a1[k]));
                                const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&a2[k]));
                                const float2 f3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&a3[k]));
                                const float2 f4 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&a4[k]));
                                const float vx = wv.x * f1.x + wv.y * f2.x + wv.z * f3.x + wv.w * f4.x;
                                const float vy = wv.x * f1.y + wv.y * f2.y + wv.z * f3.y + wv.w * f4.y;
                                __nv_bfloat162 b2 = __floats2bfloat162_rn(vx, vy);
                                pk[k] = *reinterpret_cast<uint32_t *>(&b2);
                            }
                            *reinterpret_cast<uint4 *>(sa + (size_t)row * 128 + ((c16 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic -> async proxy (UMMA reads smem)
                        __syncwarp();                                                   // every lane has fenced its own stores
                        if ((pt & 31) == 0) mbar_arrive(&full[r.stage]);
                        r.next();
                    } else {
                        // split mode: every corner is read as its (hi, lo) fp16 pair, the sample is formed in fp32 exactly as
                        // the reference does (deform_conv_cuda_kernel.cu:84-115), split again, and written as the x_hi and
                        // x_lo tiles of this K block's stage
                        const __half *xh = reinterpret_cast<const __half *>(pr.x);
                        uint32_t phi[kDItems][4], plo[kDItems][4];
#pragma unroll
                        for (int i2 = 0; i2 < 2; ++i2) {
                            uint4 u[kDRound][4][2];
#pragma unroll
                            for (int i1 = 0; i1 < kDRound; ++i1) {
                                const int item = pt + (i2 * kDRound + i1) * kDP, row = item >> 3, c16 = item & 7;
                                const int4 ov = s_o[tb][row];
                                const int co = cb * kBK + c16 * 8;
                                const int oo[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    u[i1][c][0] = *reinterpret_cast<const uint4 *>(xh + oo[c] + co);
                                    u[i1][c][1] = *reinterpret_cast<const uint4 *>(xh + oo[c] + P.Cin + co);
                                }
                            }
#pragma unroll
                            for (int i1 = 0; i1 < kDRound; ++i1) {
                                const int it = i2 * kDRound + i1, item = pt + it * kDP, row = item >> 3;
                                const float4 wv = s_w[tb][row];
                                const uint2 wh = s_wh[tb][row];
                                const float2 wc[4] = {make_float2(wv.x, wv.x), make_float2(wv.y, wv.y), make_float2(wv.z, wv.z), make_float2(wv.w, wv.w)};
                                const __half2 wl[4] = {__low2half2(*reinterpret_cast<const __half2 *>(&wh.x)), __high2half2(*reinterpret_cast<const __half2 *>(&wh.x)),
                                                       __low2half2(*reinterpret_cast<const __half2 *>(&wh.y)), __high2half2(*reinterpret_cast<const __half2 *>(&wh.y))};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    // hi halves: fp32, two channels per packed FMA; lo halves (<= 2^-11 of the value): blended in
                                    // half2 arithmetic with fp16 weights - an error of 2^-11 on a 2^-11 term
                                    float2 acc = make_float2(0.f, 0.f);
                                    __half2 accl = __float2half2_rn(0.f);
#pragma unroll
                                    for (int c = 0; c < 4; ++c) {
                                        const uint32_t uh = reinterpret_cast<const uint32_t *>(&u[i1][c][0])[k];
                                        const uint32_t ul = reinterpret_cast<const uint32_t *>(&u[i1][c][1])[k];
                                        acc = __ffma2_rn(wc[c], __half22float2(*reinterpret_cast<const __half2 *>(&uh)), acc);
                                        accl = __hfma2(wl[c], *reinterpret_cast<const __half2 *>(&ul), accl);
                                    }
                                    acc = __fadd2_rn(acc, __half22float2(accl));
                                    const __half2 h2 = __floats2half2_rn(acc.x, acc.y);
                                    const float2 hf = __half22float2(h2);
                                    const float2 rem = __fadd2_rn(acc, make_float2(-hf.x, -hf.y));
                                    const __half2 l2 = __floats2half2_rn(rem.x, rem.y);
                                    phi[it][k] = *reinterpret_cast<const uint32_t *>(&h2);
                                    plo[it][k] = *reinterpret_cast<const uint32_t *>(&l2);
                                }
                            }
                        }
                        {
                            mbar_wait(&empty[r.stage], r.phase ^ 1);
                            uint8_t *sa = smem + (size_t)r.stage * kStageBytes;
#pragma unroll
                            for (int it = 0; it < kDItems; ++it) {
                                const int item = pt + it * kDP, row = item >> 3, c16 = item & 7;
                                const size_t at = (size_t)row * 128 + ((c16 ^ (row & 7)) << 4);
                                *reinterpret_cast<uint4 *>(sa + at) = make_uint4(phi[it][0], phi[it][1], phi[it][2], phi[it][3]);
                                *reinterpret_cast<uint4 *>(sa + kABytes + at) = make_uint4(plo[it][0], plo[it][1], plo[it][2], plo[it][3]);
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if ((pt & 31) == 0) mbar_arrive(&full[r.stage]);
                            r.next();
                        }
                    }
                }
                tb ^= 1;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ----------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }

// event pool for orp_set_timing(1): every conv_tc launch is bracketed on its own stream
constexpr int kEvPool = 1024;
thread_local cudaEvent_t g_tc_ev[kEvPool][2];
thread_local int g_tc_ev_created = 0, g_tc_ev_used = 0;
thread_local double g_tc_flops = 0.0;
struct TcTrace { int nprob, N, H, W, Cin, Cout, K, stride, deform, BN, tiles, grid; double flops; };
thread_local TcTrace g_tc_trace[kEvPool];

template <int BN, bool OUT_F32, bool DEFORM>
int launch_tc(const TcParams &P, int stages, int grid, cudaStream_t st, int staging_bytes)
{
    const size_t stage_b = (P.ncat || P.dcat) ? (P.b_resident ? 2 * kABytes : 2 * kABytes + 2 * BN * kBK * 2)
                                  : (P.b_resident ? kABytes : kABytes + BN * kBK * 2);
    const size_t smem = 1024 + (size_t)stages * stage_b + (size_t)staging_bytes;
    auto kern = conv_tc_kernel<BN, OUT_F32, DEFORM>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncAttributes fa;
        ORP_CUDA(cudaFuncGetAttributes(&fa, kern));
        ORP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
        attr_set = true;
    }
    int slot = -1;
    if (g_timing && g_tc_ev_used < kEvPool) {
        slot = g_tc_ev_used++;
        if (slot >= g_tc_ev_created) {
            ORP_CUDA(cudaEventCreate(&g_tc_ev[slot][0]));
            ORP_CUDA(cudaEventCreate(&g_tc_ev[slot][1]));
            g_tc_ev_created = slot + 1;
        }
        ORP_CUDA(cudaEventRecord(g_tc_ev[slot][0], st));
        double fl = 0;
        for (int i = 0; i < P.nprob; ++i)
            fl += 2.0 * P.prob[i].N * P.prob[i].Ho * P.prob[i].Wo * (double)P.Cout *
                  (P.s2d_stem ? 147.0 : (double)P.KH * P.KW * (P.stem ? 147 : P.Cin));   // algorithmic K, not the padded one
        g_tc_flops += fl;
        g_tc_trace[slot] = TcTrace{P.nprob, P.prob[0].N, P.prob[0].H, P.prob[0].W, P.Cin, P.Cout, P.KH, P.stride, DEFORM ? 1 : 0,
                                   BN, P.num_tiles, grid, fl};
    }
    {
        static const bool pdl = getenv("ORP_TC_NO_PDL") == nullptr;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(DEFORM ? (unsigned)(192 + kDP) : 320u);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        ORP_CUDA(cudaLaunchKernelEx(&cfg, kern, P, stages));
    }
    ORP_LAUNCHED();
    if (slot >= 0) ORP_CUDA(cudaEventRecord(g_tc_ev[slot][1], st));
    return ORP_OK;
}

}  // namespace
}  // namespace orp

using namespace orp;

extern "C" int orp_tc_timing_collect(float *total_ms, int *launches, double *flops)
{
    if (!total_ms || !launches || !flops) return fail(ORP_EINVAL, "orp_tc_timing_collect: null");
    float sum = 0.f;
    for (int i = 0; i < g_tc_ev_used; ++i) {
        float ms = 0.f;
        ORP_CUDA(cudaEventSynchronize(g_tc_ev[i][1]));
        ORP_CUDA(cudaEventElapsedTime(&ms, g_tc_ev[i][0], g_tc_ev[i][1]));
        sum += ms;
        if (getenv("ORP_TC_TRACE")) {
            const TcTrace &t = g_tc_trace[i];
            fprintf(stderr, "tc[%3d] np=%d N=%d %4dx%-4d Cin=%4d Cout=%4d k=%d s=%d dcn=%d BN=%3d tiles=%5d grid=%3d  %8.1f us  %7.1f TFLOP/s\n",
                    i, t.nprob, t.N, t.H, t.W, t.Cin, t.Cout, t.K, t.stride, t.deform, t.BN, t.tiles, t.grid, ms * 1e3,
                    t.flops / (ms * 1e-3) / 1e12);
        }
    }
    *total_ms = sum; *launches = g_tc_ev_used; *flops = g_tc_flops;
    g_tc_ev_used = 0; g_tc_flops = 0.0;
    return ORP_OK;
}

static int conv2d_bf16_impl(int nprob, const orp_tc_problem *probs, const void *w, int Cout, int Cout_padded, int KH,
                            int KW, int Cin, int stride, int pad, const float *bias, int relu, int out_f32,
                            int deform, int stem, void *stream, int split = 0, int wscale_log2 = 0, int ksplit = 1);

/* see include/orp_b200.h */
extern "C" int orp_conv2d_bf16(int nprob, const orp_tc_problem *probs, const void *w, int Cout, int Cout_padded, int KH,
                               int KW, int Cin, int stride, int pad, const float *bias, int relu, int out_f32,
                               int deform, void *stream)
{
    return conv2d_bf16_impl(nprob, probs, w, Cout, Cout_padded, KH, KW, Cin, stride, pad, bias, relu, out_f32, deform, 0, stream);
}

/* see include/orp_b200.h */
extern "C" int orp_conv2d_f16x3(int nprob, const orp_tc_problem *probs, const void *w_split, int Cout, int Cout_padded, int KH,
                                int KW, int Cin, int stride, int pad, const float *bias, int wscale_log2, int relu,
                                int out_f32, int deform, void *stream)
{
    if (wscale_log2 < 0 || wscale_log2 > 15) return fail(ORP_EINVAL, "conv2d_f16x3: weight scale exponent must be in 0..15");
    return conv2d_bf16_impl(nprob, probs, w_split, Cout, Cout_padded, KH, KW, Cin, stride, pad, bias, relu, out_f32, deform, 0,
                            stream, 1, wscale_log2);
}

extern "C" int orp_stem_conv_s2d_f16x3(const void *x_s2d, int N, int H, int W, const void *w_split, const float *bias,
                                       int wscale_log2, int relu, void *out, void *stream)
{
    if (!x_s2d || !w_split || !out || N < 1 || H < 2 || W < 2 || (H & 1) || (W & 1))
        return fail(ORP_EINVAL, "stem_conv_s2d_f16x3: needs even H, W");
    if (wscale_log2 < 0 || wscale_log2 > 15) return fail(ORP_EINVAL, "stem_conv_s2d_f16x3: weight scale exponent must be in 0..15");
    orp_tc_problem q;
    memset(&q, 0, sizeof(q));
    q.x = x_s2d;
    q.N = N; q.H = H / 2 + 3; q.W = W / 2; q.out = out;
    return conv2d_bf16_impl(1, &q, w_split, 64, 64, 4, 1, 64, 1, 0, bias, relu, 0, 0, 2, stream, 1, wscale_log2);
}

extern "C" int orp_f16x3_overflow_count(unsigned int *count, int reset)
{
    if (!count) return fail(ORP_EINVAL, "f16x3_overflow_count: null");
    ORP_CUDA(cudaMemcpyFromSymbol(count, g_f16_overflow, sizeof(unsigned int)));
    if (reset) {
        const unsigned int z = 0;
        ORP_CUDA(cudaMemcpyToSymbol(g_f16_overflow, &z, sizeof(z)));
    }
    return ORP_OK;
}

/* see include/orp_b200.h */
extern "C" int orp_stem_conv_bf16(const float *img_nchw, int N, int H, int W, const void *w192, const float *bias, int relu,
                                  void *out, void *stream)
{
    if (!img_nchw || !w192 || !out || N < 1) return fail(ORP_EINVAL, "stem_conv_bf16: bad arguments");
    orp_tc_problem q;
    memset(&q, 0, sizeof(q));
    q.x = img_nchw;                       // never dereferenced as bf16: the producers read q.offset
    q.offset = img_nchw;
    q.N = N; q.H = H; q.W = W; q.out = out;
    return conv2d_bf16_impl(1, &q, w192, 64, 64, 1, 1, 192, 1, 0, bias, relu, 0, 1, 1, stream);
}

extern "C" int orp_stem_conv_s2d_bf16(const void *x_s2d, int N, int H, int W, const void *w256, const float *bias, int relu,
                                      void *out, void *stream)
{
    if (!x_s2d || !w256 || !out || N < 1 || H < 2 || W < 2 || (H & 1) || (W & 1))
        return fail(ORP_EINVAL, "stem_conv_s2d_bf16: needs even H, W");
    orp_tc_problem q;
    memset(&q, 0, sizeof(q));
    q.x = x_s2d;
    q.N = N; q.H = H / 2 + 3; q.W = W / 2; q.out = out;     // 4 x 1 taps over rows, 64 virtual channels, no padding
    return conv2d_bf16_impl(1, &q, w256, 64, 64, 4, 1, 64, 1, 0, bias, relu, 0, 0, 2, stream);
}

namespace orp {
namespace {
// split-K finish: fp32 sums [pixels, C] -> + bias -> ReLU -> bf16 [pixels, C] or split fp16 [pixels, 2, C]
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const float *__restrict__ ws, int ksplit, size_t pixels, int C, const float *__restrict__ bias, int relu, int split,
                     void *__restrict__ out, unsigned int *ovf)
{
    const int c8 = C / 8;
    const size_t total = pixels * c8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i / c8;
        const int c = (int)(i - pix * c8) * 8;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < ksplit; ++k) {                       // fixed order: the sum is reproducible bit for bit
            const float *p = ws + (size_t)k * pixels * C + pix * C + c;
            const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (bias) v[j] += bias[c + j];
            if (relu) v[j] = fmaxf(v[j], 0.f);
            amax = fmaxf(amax, fabsf(v[j]));
        }
        uint32_t hi[4], lo[4];
        if (split) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __half2 h2 = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                const float2 hf = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
                hi[k] = *reinterpret_cast<const uint32_t *>(&h2);
                lo[k] = *reinterpret_cast<const uint32_t *>(&l2);
            }
            __half *o = static_cast<__half *>(out) + pix * 2 * C + c;
            *reinterpret_cast<uint4 *>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(o + C) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            if (amax > 65504.f) atomicAdd(ovf, 1u);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                hi[k] = *reinterpret_cast<uint32_t *>(&b2);
            }
            *reinterpret_cast<uint4 *>(static_cast<__nv_bfloat16 *>(out) + pix * C + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        }
    }
}
}  // namespace
}  // namespace orp

/* see include/orp_b200.h */
extern "C" int orp_conv2d_tc_splitk(const orp_tc_problem *prob, const void *w, int Cout, int Cout_padded, int KH, int KW, int Cin,
                                    int stride, int pad, const float *bias, int f16x3, int wscale_log2, int relu, int ksplit,
                                    float *workspace, void *stream)
{
    if (!prob || !w || !workspace || ksplit < 2 || (Cout % 8)) return fail(ORP_EINVAL, "conv2d_tc_splitk: bad arguments");
    if (f16x3 && (wscale_log2 < 0 || wscale_log2 > 15)) return fail(ORP_EINVAL, "conv2d_tc_splitk: weight scale exponent must be in 0..15");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Ho = (prob->H + 2 * pad - (KH - 1) - 1) / stride + 1, Wo = (prob->W + 2 * pad - (KW - 1) - 1) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return fail(ORP_EINVAL, "conv2d_tc_splitk: bad problem");
    const size_t pixels = (size_t)prob->N * Ho * Wo;
    orp_tc_problem q = *prob;                                  // every (pixel, channel, split) element of the workspace is written
    q.out = workspace;
    q.gn_stats = nullptr;
    int rc = conv2d_bf16_impl(1, &q, w, Cout, Cout_padded, KH, KW, Cin, stride, pad, nullptr, 0, 1, 0, 0, stream, f16x3 ? 1 : 0,
                              wscale_log2, ksplit);
    if (rc) return rc;
    void *ovf = nullptr;
    ORP_CUDA(cudaGetSymbolAddress(&ovf, g_f16_overflow));
    const size_t items = pixels * (Cout / 8);
    size_t g = (items + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    splitk_finish_kernel<<<(unsigned)(g ? g : 1), 256, 0, st>>>(workspace, ksplit, pixels, Cout, bias, relu, f16x3 ? 1 : 0, prob->out,
                                                                static_cast<unsigned int *>(ovf));
    ORP_LAUNCHED();
    if (prob->gn_stats) {
        if (Cout != 256) return fail(ORP_EINVAL, "conv2d_tc_splitk: gn_stats needs 256 output channels");
        return f16x3 ? orp_gn_stats_f16x3(prob->out, prob->N, Ho * Wo, 256, 32, prob->gn_stats, stream)
                     : orp_gn_stats_bf16(prob->out, prob->N, Ho * Wo, 256, 32, prob->gn_stats, stream);
    }
    return ORP_OK;
}

static int conv2d_bf16_impl(int nprob, const orp_tc_problem *probs, const void *w, int Cout, int Cout_padded, int KH,
                            int KW, int Cin, int stride, int pad, const float *bias, int relu, int out_f32,
                            int deform, int stem, void *stream, int split, int wscale_log2, int ksplit)
{
    if (nprob < 1 || nprob > kMaxProb || !probs || !w) return fail(ORP_EINVAL, "conv2d_tc: bad arguments");
    if (Cin % 8) return fail(ORP_EINVAL, "conv2d_tc: Cin must be a multiple of 8 (16-byte channel rows)");
    if (deform && !stem && (Cin % kBK)) return fail(ORP_EINVAL, "conv2d_tc: deformable conv needs Cin % 64 == 0");
    if (Cout_padded % 32 || Cout_padded < Cout) return fail(ORP_EINVAL, "conv2d_tc: padded Cout must be a multiple of 32");
    if (split && stem == 1) return fail(ORP_EINVAL, "conv2d_tc: the direct stem has no f16x3 form (use the space-to-depth stem)");
    if (ksplit < 1) ksplit = 1;
    if (ksplit > 1 && (nprob != 1 || deform || stem || !out_f32 || (KH * KW) % ksplit || probs[0].residual_bf16 || probs[0].residual_f32 || bias || relu))
        return fail(ORP_EINVAL, "conv2d_tc: split-K serves one plain problem with an fp32 partial-sum output and taps % ksplit == 0");
    int rc = ensure_device();
    if (rc) return rc;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(ORP_ECUDA, "conv2d_tc: cuTensorMapEncodeTiled unavailable");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CUtensorMapDataType dt16 = split ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const int T = split ? 2 : 1;             // 16-bit planes per value (hi, lo)

    int BN = 256;
    if (Cout_padded % 256) BN = (Cout_padded % 128 == 0) ? 128 : (Cout_padded % 64 == 0) ? 64 : 32;
    {
        // narrower accumulators when the 128 x BN tiling would leave SMs idle
        long long mtiles = 0;
        for (int i = 0; i < nprob; ++i) {
            const int ho = (probs[i].H + 2 * pad - (KH - 1) - 1) / stride + 1, wo = (probs[i].W + 2 * pad - (KW - 1) - 1) / stride + 1;
            mtiles += ((long long)probs[i].N * ho * wo + 127) / 128;
        }
        while (BN > 64 && mtiles * (Cout_padded / BN) * ksplit < 120) BN /= 2;
    }
    TcParams P;
    memset(&P, 0, sizeof(P));
    // a partial last channel block is zero-filled by TMA (A operand) and by the weight layout (B operand)
    P.nprob = nprob; P.KH = KH; P.KW = KW; P.Cin = Cin; P.cin_blocks = (Cin + kBK - 1) / kBK;
    P.stride = stride; P.pad = pad;
    P.Cout = Cout; P.relu = relu; P.bias = bias; P.stem = (stem == 1) ? 1 : 0; P.s2d_stem = (stem == 2) ? 1 : 0;
    P.split = split ? 1 : 0;
    P.oscale = split ? ldexpf(1.f, -wscale_log2) : 1.f;
    {
        void *ovf = nullptr;
        ORP_CUDA(cudaGetSymbolAddress(&ovf, g_f16_overflow));
        P.ovf = static_cast<unsigned int *>(ovf);
    }
    P.n_tiles_n = Cout_padded / BN;
    P.fd_ntn.set((uint32_t)P.n_tiles_n);
    int mt = 0;
    for (int i = 0; i < nprob; ++i) {
        const orp_tc_problem &q = probs[i];
        Problem &pr = P.prob[i];
        pr.N = q.N; pr.H = q.H; pr.W = q.W;
        pr.Ho = (q.H + 2 * pad - (KH - 1) - 1) / stride + 1;
        pr.Wo = (q.W + 2 * pad - (KW - 1) - 1) / stride + 1;
        if (stem == 1) { pr.Ho = (q.H + 6 - 7) / 2 + 1; pr.Wo = (q.W + 6 - 7) / 2 + 1; }
        if (pr.Ho <= 0 || pr.Wo <= 0 || !q.x || !q.out) return fail(ORP_EINVAL, "conv2d_tc: bad problem");
        pr.BW = pow2_floor(pr.Wo < 128 ? pr.Wo : 128);
        if (stride * pr.BW > 256) pr.BW = 256 / stride;
        // (measured: 16 x 8 pixel tiles for the deformable variant change nothing - 1361 vs 1334 us per f16x3 launch, and neither
        // does the spread of the offsets: the gather runs at the ~47 GB/s per SM L2 -> SM ceiling whatever its locality)
        if (deform && !stem && pr.BW > 16 && getenv("ORP_TC_DCN_2DTILES")) pr.BW = 16;
        pr.BH = pow2_floor(pr.Ho < 128 / pr.BW ? pr.Ho : 128 / pr.BW);
        pr.BI = 128 / (pr.BW * pr.BH);
        pr.lbw = 0; while ((1 << pr.lbw) < pr.BW) ++pr.lbw;
        pr.lbh = 0; while ((1 << pr.lbh) < pr.BH) ++pr.lbh;
        pr.tiles_w = ceil_div(pr.Wo, pr.BW); pr.tiles_h = ceil_div(pr.Ho, pr.BH); pr.tiles_i = ceil_div(pr.N, pr.BI);
        pr.tile_start = mt;
        pr.fd_tw.set((uint32_t)pr.tiles_w); pr.fd_th.set((uint32_t)pr.tiles_h);
        mt += pr.tiles_w * pr.tiles_h * pr.tiles_i;
        pr.out = q.out; pr.res = static_cast<const __nv_bfloat16 *>(q.residual_bf16); pr.res32 = q.residual_f32;
        pr.x = static_cast<const __nv_bfloat16 *>(q.x); pr.offset = q.offset; pr.gn_stats = q.gn_stats; pr.mask = q.mask;
        if (deform && !q.offset) return fail(ORP_EINVAL, "conv2d_tc: deformable conv needs offsets");
        if (!deform) {
            // 5-D view {channel, plane (hi / lo), w, h, image}; bf16 tensors have a single plane.  Split activations are
            // [N,H,W,2,C]: the lo plane of a pixel follows its hi plane.
            cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)T, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)q.N};
            cuuint64_t gstr[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * 2 * T, (cuuint64_t)q.W * Cin * 2 * T,
                                  (cuuint64_t)q.H * q.W * Cin * 2 * T};
            if (stem == 2) {
                // space-to-depth stem: the tensor is [T][N, H, W + 3, 16] (planes outermost); a 64-element "pixel" row of
                // the GEMM is the 4 horizontally adjacent 16-channel pixels starting at w, so consecutive w overlap
                // (stride 32 bytes)
                gstr[0] = (cuuint64_t)q.N * q.H * (q.W + 3) * 32;
                gstr[1] = 32;
                gstr[2] = (cuuint64_t)(q.W + 3) * 32;
                gstr[3] = (cuuint64_t)q.H * (q.W + 3) * 32;
            }
            cuuint32_t box[5] = {(cuuint32_t)kBK, 1u, (cuuint32_t)(pr.BW * stride), (cuuint32_t)(pr.BH * stride), (cuuint32_t)pr.BI};
            cuuint32_t estr[5] = {1, 1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
            CUresult r = enc(&P.tmA[i], dt16, 5, const_cast<void *>(q.x), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(ORP_ECUDA, "conv2d_tc: cuTensorMapEncodeTiled(A) failed");
        }
    }
    {
        // weights [Cout_padded][tap][cin_blocks][T][64] (zero padded per tap; T = 2: the hi block, then the lo block)
        const cuuint64_t K = (cuuint64_t)KH * KW * T * P.cin_blocks * kBK;
        cuuint64_t gdim[2] = {K, (cuuint64_t)Cout_padded};
        cuuint64_t gstr[1] = {K * 2};
        cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&P.tmB, dt16, 2, const_cast<void *>(w), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(ORP_ECUDA, "conv2d_tc: cuTensorMapEncodeTiled(B) failed");
    }
    P.num_m_tiles = mt;
    P.num_tiles = mt * P.n_tiles_n * ksplit;
    P.ksplit = ksplit;
    P.fd_ks.set((uint32_t)ksplit);
    P.ks_stride = (long long)P.prob[0].N * P.prob[0].Ho * P.prob[0].Wo * Cout;
    // TMA epilogue: 16-bit outputs whose channel count is a multiple of 64
    bool any_res = false;
    for (int i = 0; i < nprob; ++i) any_res = any_res || (probs[i].residual_bf16 != nullptr);
    // (f16x3 also takes channel counts that are only a multiple of 8 - Swin's 96 / 288: the TMA unit clips the last box)
    P.tma_epi = (!out_f32 && ((Cout % 64 == 0) || (split && Cout % 8 == 0)) && BN >= 64) ? 1 : 0;
    if (getenv("ORP_TC_NO_TMA_EPI") && !split) P.tma_epi = 0;
    if (split && !out_f32 && !P.tma_epi) return fail(ORP_EINVAL, "conv2d_f16x3: 16-bit outputs need Cout % 8 == 0 and weights padded to a multiple of 64 rows");
    if (split && out_f32 && any_res) return fail(ORP_EINVAL, "conv2d_f16x3: fp32 outputs take an fp32 residual only");
    // GroupNorm statistics: fused into the TMA epilogue when every warp's 32 rows lie in one image
    bool want_gn = false, gn_ok = (P.tma_epi != 0) && Cout == 256 && !bias && !relu;
    for (int i = 0; i < nprob; ++i) {
        want_gn = want_gn || (probs[i].gn_stats != nullptr);
        if (P.prob[i].BW * P.prob[i].BH < 32) gn_ok = false;
    }
    P.gn_fused = (want_gn && gn_ok) ? 1 : 0;
    const bool mem_bound = any_res || (KH * KW * (Cin / kBK) <= 8);
    P.epi_bufs = mem_bound ? 2 : 1;          // a second staging tile costs compute-bound layers a main-loop stage
    if (const char *e = getenv("ORP_TC_EPI_BUFS")) P.epi_bufs = atoi(e) == 1 ? 1 : 2;
    P.epi_merge = (split && P.tma_epi && (mem_bound || stem == 2) && !getenv("ORP_TC_NO_MERGE")) ? 1 : 0;
    // terms concatenated along N for narrow layers (kernel header); the residual / deformable / fp32-output variants keep
    // the K-concatenated walk
    P.dcat = (split && deform && stem != 1) ? 1 : 0;
    P.ncat = (split && P.tma_epi && BN <= 128 && !deform && !any_res && stem != 1 && !getenv("ORP_TC_NO_NCAT")) ? 1 : 0;
    // epilogue-bound layers (at most 6 K blocks per tile incl. the residual's; measured: 7-15 lose a little to the
    // smaller staging/stage budget): independent epilogue warpgroups
    {
        const int kb_total = (KH * KW * P.cin_blocks) * ((split && !P.ncat) ? 3 : 1) + (any_res ? (BN / 64) * T : 0);
        P.epi_split = (P.tma_epi && !deform && !stem && kb_total <= 6 && !getenv("ORP_TC_NO_SPLIT")) ? 1 : 0;
        if (P.epi_split) P.epi_bufs = 2;
    }
    // residual through the tensor core (TMA epilogue only; the staged epilogue adds it itself)
    P.res_mma = (P.tma_epi && any_res) ? 1 : 0;
    if (P.res_mma) {
        if (deform) return fail(ORP_EINVAL, "conv2d_tc: residual is not supported on the deformable path");
        for (int i = 0; i < nprob; ++i)
            if (!probs[i].residual_bf16) return fail(ORP_EINVAL, "conv2d_tc: residual must be given for every problem or none");
        void *ident_ptr = nullptr;
        if (split) {
            ORP_CUDA(cudaGetSymbolAddress(&ident_ptr, g_ident16));
            ident_ptr = static_cast<char *>(ident_ptr) + (size_t)wscale_log2 * 8192;
        } else {
            ORP_CUDA(cudaGetSymbolAddress(&ident_ptr, g_ident));
        }
        cuuint64_t gdim[2] = {64, 64};
        cuuint64_t gstr[1] = {128};
        cuuint32_t box[2] = {64, 64};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&P.tmI, dt16, 2, ident_ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(ORP_ECUDA, "conv2d_tc: cuTensorMapEncodeTiled(identity) failed");
    }
    // layers whose whole weight slab for one N tile is <= 72 KiB keep it resident; stages then carry only the A tile
    P.b_resident = (!deform && ksplit == 1 && KH * KW * T * P.cin_blocks * BN * kBK * 2 <= 72 * 1024 && !getenv("ORP_TC_NO_BRES")) ? 1 : 0;
    if (P.tma_epi) {
        for (int i = 0; i < nprob; ++i) {
            const Problem &pr = P.prob[i];
            cuuint64_t gdim[5] = {(cuuint64_t)Cout, (cuuint64_t)T, (cuuint64_t)pr.Wo, (cuuint64_t)pr.Ho, (cuuint64_t)pr.N};
            cuuint64_t gstr[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)Cout * 2 * T, (cuuint64_t)pr.Wo * Cout * 2 * T,
                                  (cuuint64_t)pr.Ho * pr.Wo * Cout * 2 * T};
            cuuint32_t box[5] = {64u, 1u, (cuuint32_t)pr.BW, (cuuint32_t)pr.BH, (cuuint32_t)pr.BI};
            cuuint32_t estr[5] = {1, 1, 1, 1, 1};
            CUresult r = enc(&P.tmOut[i], dt16, 5, pr.out, gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(ORP_ECUDA, "conv2d_tc: cuTensorMapEncodeTiled(out) failed");
            if (pr.res) {
                r = enc(&P.tmRes[i], dt16, 5, const_cast<__nv_bfloat16 *>(pr.res), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail(ORP_ECUDA, "conv2d_tc: cuTensorMapEncodeTiled(residual) failed");
            }
        }
    }
    int sms = 148;
    {
        int dev = 0;
        ORP_CUDA(cudaGetDevice(&dev));
        ORP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int grid = P.num_tiles < sms ? P.num_tiles : sms;
    if (P.b_resident && grid >= P.n_tiles_n) grid -= grid % P.n_tiles_n;       // fixed N tile per CTA
    else if (P.b_resident) P.b_resident = 0;
    // shared-memory budget: output staging + resident weights + main-loop stages.  When the extras leave fewer than three
    // stages they are given up in order of least value: the second epilogue group, the second staging slot, the resident slab.
    int stage_bytes = 0, bres_bytes = 0, staging = 0, stages = 0;
    for (;;) {
        stage_bytes = (P.ncat || P.dcat) ? (P.b_resident ? 2 * kABytes : 2 * kABytes + 2 * BN * kBK * 2)
                                         : (P.b_resident ? kABytes : kABytes + BN * kBK * 2);
        bres_bytes = (P.b_resident ? KH * KW * T * P.cin_blocks * BN * kBK * 2 : 0) + (P.res_mma ? 8192 : 0) + (stem == 1 ? kStemPatchBytes : 0);
        const int hc = BN < 64 ? BN : 64;
        staging = out_f32 ? 0 : 128 * (hc * 2 + 16);
        if (P.tma_epi) staging = P.epi_bufs * (P.epi_merge ? 32768 : 16384) * (P.epi_split ? 2 : 1);
        stages = (int)((227 * 1024 - (deform || stem == 1 ? 14336 : 4096) - 1024 - staging - bres_bytes) / stage_bytes);   // static shared memory of the variant
        if (stages >= 3) break;
        if (P.epi_split) { P.epi_split = 0; P.epi_bufs = mem_bound ? 2 : 1; continue; }
        if (P.epi_bufs == 2) { P.epi_bufs = 1; continue; }
        if (P.b_resident) { P.b_resident = 0; continue; }
        if (stages >= 2) break;
        return fail(ORP_EINVAL, "conv2d_tc: shared-memory budget cannot hold two main-loop stages");
    }
    if (stages > kStagesMax) stages = kStagesMax;
    if (const char *e = getenv("ORP_TC_STAGES")) { const int v = atoi(e); if (v >= 2 && v < stages) stages = v; }   // experiments
    if (deform && stages > 3) stages = 3;     // leave L1 capacity for the bilinear gather (corner reuse between neighbouring pixels)
    int lrc = ORP_EINVAL;
    bool launched = false;
#define ORP_TC_DISPATCH(BNV)                                                                     \
    if (!launched && BN == BNV) {                                                                \
        launched = true;                                                                         \
        if (deform) lrc = out_f32 ? launch_tc<BNV, true, true>(P, stages, grid, st, staging + bres_bytes) : launch_tc<BNV, false, true>(P, stages, grid, st, staging + bres_bytes); \
        else lrc = out_f32 ? launch_tc<BNV, true, false>(P, stages, grid, st, staging + bres_bytes) : launch_tc<BNV, false, false>(P, stages, grid, st, staging + bres_bytes);       \
    }
    ORP_TC_DISPATCH(256)
    ORP_TC_DISPATCH(128)
    ORP_TC_DISPATCH(64)
    ORP_TC_DISPATCH(32)
#undef ORP_TC_DISPATCH
    if (!launched) return fail(ORP_EINVAL, "conv2d_tc: unsupported tile width");
    if (lrc) return lrc;
    if (want_gn && !P.gn_fused) {
        // statistics requested but not fusable for this shape: separate pass over the bf16 output
        if (out_f32 || Cout != 256) return fail(ORP_EINVAL, "conv2d_tc: gn_stats needs a 16-bit output with 256 channels");
        for (int i = 0; i < nprob; ++i)
            if (probs[i].gn_stats) {
                int r2 = split ? orp_gn_stats_f16x3(P.prob[i].out, P.prob[i].N, P.prob[i].Ho * P.prob[i].Wo, 256, 32, probs[i].gn_stats, stream)
                               : orp_gn_stats_bf16(P.prob[i].out, P.prob[i].N, P.prob[i].Ho * P.prob[i].Wo, 256, 32, probs[i].gn_stats, stream);
                if (r2) return r2;
            }
    }
    return ORP_OK;
}
