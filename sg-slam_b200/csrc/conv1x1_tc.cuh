// conv1x1_tc.cuh -- the detector's 1x1 convolutions (66 of the 70 convolutions of mobilenetv3_ssdlite_voc.param, 90 % of its MACs;
// Detector2D.cc:39-45 runs them through ncnn) as a Blackwell-native GEMM:
//     out[p][co] = tail(bias[co] + sum_ci X[p][ci] * W[co][ci])          X: NHWC activations [pixels of all frames][Cin], W: [Cout][Cin]
//   * operands reach shared memory by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle, out-of-range rows / channels zero-filled),
//   * the contraction is tcgen05.mma (kind::tf32, M = 128 pixels, N = a tile of <= 256 output channels, K = 8 per instruction) issued by one
//     thread, the FP32 accumulator lives in TMEM,
//   * FP32-grade accuracy: every operand is x = hi + lo with hi, lo exactly representable in TF32 (round-to-nearest split); each k-step issues
//     lo*hi + hi*lo + hi*hi (the dropped lo*lo term is < 2^-22 relative).  Weights are split once at load time; the activation tile is split
//     in shared memory by the CTA's 128 threads (element-wise, so the swizzled placement is irrelevant),
//   * epilogue: tcgen05.ld (one TMEM lane = one pixel per thread, 16 channels at a time) -> bias -> fused element-wise tail -> NHWC store.
// One CTA = one (128-pixel, N-tile) output tile, 128 threads; several CTAs are resident per SM (shared memory / TMEM permitting), so the
// load / split / MMA / epilogue phases of different tiles overlap without warp specialisation inside a CTA.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

namespace sgs {
namespace tc {

constexpr int kBM = 128;        // pixels per tile = UMMA M = TMEM lanes
constexpr int kBK = 32;         // floats per k-block = one 128-byte swizzle row (16 = one 64-byte row for layers with at most 16 input channels)
constexpr int kMaxStages = 8;
constexpr int kMaxNT = 256;      // output channels per tile (UMMA N <= 256; two accumulators = 512 TMEM columns)

// ---------------------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
#ifdef TC_DBG_SUSPEND
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#endif
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major tile whose rows are one swizzle row (ROWB = 128 or 64 bytes) and whose 8-row groups are 8 * ROWB
// bytes apart (cute::UMMA::SmemDescriptor: start >> 4 | LBO << 16 | SBO << 32 | version 1 << 46 | layout << 61; SWIZZLE_128B = 2, SWIZZLE_64B = 4).
template <int ROWB>
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr) {
    static_assert(ROWB == 128 || ROWB == 64, "one swizzle row per tile row");
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)((8 * ROWB) >> 4) << 32) | (1ull << 46) | ((uint64_t)(ROWB == 128 ? 2 : 4) << 61);
}
// cute::UMMA::InstrDescriptor for kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ inline uint32_t tf32_idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24); }

// x = hi + lo, both TF32 (round to nearest, ties away, by integer arithmetic on the bit pattern)
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
    hi = (x + 0x1000u) & 0xffffe000u;
    lo = (__float_as_uint(__fsub_rn(__uint_as_float(x), __uint_as_float(hi))) + 0x1000u) & 0xffffe000u;
}

#ifdef SGS_TC_TRACE
// pipeline trace of CTA (0, 0) (tools/umma_proto only): SM clock at the hand-over points of the four roles
__device__ long long g_tc_trace[4][1024];
#define TC_TRACE(role, slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (slot) < 1024) g_tc_trace[role][slot] = clock64(); } while (0)
#else
#define TC_TRACE(role, slot) do { } while (0)
#endif

struct TcGeom {
    int npix, Cin, Cout;          // rows of X, channels in / out
    int NT, KB, stages;           // output channels per tile (multiple of 16), k-blocks, pipeline depth of the operand ring
    int m_tiles;                  // 128-pixel tiles
    int b_resident;               // the CTA's weight tile (all k-blocks, hi + lo) is loaded once and stays in shared memory
    uint32_t tmem_cols;           // power of two >= 2 * NT: two accumulators (the epilogue of tile i overlaps the MMAs of tile i + 1)
    int HW;                       // pixels per frame (output addressing)
    int64_t frame_stride;         // floats between consecutive frames of the output
    int64_t base_off;             // float offset of frame 0 of the output inside `out`
    int pitch;                    // floats between consecutive pixels of the output
    int vec_ok;                   // 16-byte aligned rows: float4 stores
    int coalesce;                 // contiguous [pixel][channel] output (frame_stride == HW * pitch): a pixel's row address needs no division by HW
};

// warp 0: TMA producer, warp 1: MMA issuer (+ TMEM owner), warps 2 .. 2+EPW-1: epilogue (EPW = 4 or 8), the last four warps: operand split
constexpr int tc_threads(int epw) { return (2 + epw + 4) * 32; }

// Persistent, warp-specialised.  CTA (x, y) owns output-channel tile y and walks the pixel tiles x, x + gridDim.x, ...
//   producer  : TMA loads of the activation k-blocks (and of the weight k-blocks unless resident) into a ring of `stages` slots
//   split     : 128 threads turn each landed FP32 activation block into TF32 hi (in place) + lo
//   MMA       : one thread issues 3 tcgen05.mma per 8-wide k-step into one of two TMEM accumulators, commits to the ring / to the epilogue
//   epilogue  : 128 threads (one TMEM lane = one pixel each) read the accumulator, add bias, apply the tail, store NHWC
// mbarriers: full[s] (TMA -> split), ready[s] (split -> MMA), empty[s] (MMA -> producer), bfull (resident weights), acc_full[a] / acc_empty[a].
// EpiFn: struct with  template <int N> __device__ void run(float (&v)[N], int64_t idx0) const   applied to N consecutive channels of one pixel;
// idx0 = pixel * Cout + channel (the NHWC index of same-shape operand tensors).
template <class EpiFn, int BKT, int EPW>
__global__ void __launch_bounds__(tc_threads(EPW), EPW == 4 ? 2 : 1) conv1x1_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                                                                   const __grid_constant__ CUtensorMap mapBl, const float* __restrict__ bias,
                                                                   float* __restrict__ out, const TcGeom G, const EpiFn epi) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kMaxStages], s_ready[kMaxStages], s_empty[kMaxStages], s_bfull, s_acc_full[2], s_acc_empty[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * G.NT;
    constexpr int ROWB = BKT * 4;                                             // bytes per operand row = one swizzle row
    const uint32_t a_bytes = kBM * ROWB, b_bytes = (uint32_t)G.NT * ROWB;
    const uint32_t stage_bytes = 2 * a_bytes + (G.b_resident ? 0u : 2 * b_bytes);
    const uint32_t sbase = (smem_addr(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (sbase - smem_addr(smem_raw));
    const uint32_t bres = sbase + (uint32_t)G.stages * stage_bytes;          // resident weights: [kb][hi | lo]
    const uint32_t epi_off = (uint32_t)G.stages * stage_bytes + (G.b_resident ? (uint32_t)G.KB * 2 * b_bytes : 0u);      // per epilogue warp: 32 rows x 128 B
    const int my_tiles = (G.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < G.stages; ++s) { mbar_init(smem_addr(&s_full[s]), 1); mbar_init(smem_addr(&s_ready[s]), 4); mbar_init(smem_addr(&s_empty[s]), 1); }
        mbar_init(smem_addr(&s_bfull), 1);
        for (int a = 0; a < 2; ++a) { mbar_init(smem_addr(&s_acc_full[a]), 1); mbar_init(smem_addr(&s_acc_empty[a]), EPW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "r"(G.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            if (G.b_resident) {
                mbar_expect_tx(smem_addr(&s_bfull), (uint32_t)G.KB * 2 * b_bytes);
                for (int kb = 0; kb < G.KB; ++kb) {
                    tma_load_2d(bres + (uint32_t)kb * 2 * b_bytes, &mapBh, kb * BKT, n0, smem_addr(&s_bfull));
                    tma_load_2d(bres + (uint32_t)kb * 2 * b_bytes + b_bytes, &mapBl, kb * BKT, n0, smem_addr(&s_bfull));
                }
            }
            uint32_t g = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * kBM;
                for (int kb = 0; kb < G.KB; ++kb, ++g) {
                    const uint32_t s = g % (uint32_t)G.stages, ph = (g / (uint32_t)G.stages) & 1u;
                    mbar_wait(smem_addr(&s_empty[s]), ph ^ 1u);
                    const uint32_t st = sbase + s * stage_bytes, bar = smem_addr(&s_full[s]);
                    mbar_expect_tx(bar, a_bytes + (G.b_resident ? 0u : 2 * b_bytes));
                    tma_load_2d(st, &mapA, kb * BKT, m0, bar);
                    if (!G.b_resident) {
                        tma_load_2d(st + 2 * a_bytes, &mapBh, kb * BKT, n0, bar);
                        tma_load_2d(st + 2 * a_bytes + b_bytes, &mapBl, kb * BKT, n0, bar);
                    }
                    TC_TRACE(0, g);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = tf32_idesc(G.NT);
            if (G.b_resident) mbar_wait(smem_addr(&s_bfull), 0);
            uint32_t g = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const uint32_t ab = (uint32_t)it & 1u, aph = ((uint32_t)it >> 1) & 1u;
                mbar_wait(smem_addr(&s_acc_empty[ab]), aph ^ 1u);                   // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem + ab * (uint32_t)G.NT;
                for (int kb = 0; kb < G.KB; ++kb, ++g) {
                    const uint32_t s = g % (uint32_t)G.stages, ph = (g / (uint32_t)G.stages) & 1u;
                    mbar_wait(smem_addr(&s_ready[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    TC_TRACE(2, 2 * g);
                    const uint32_t st = sbase + s * stage_bytes;
                    const uint32_t bh0 = G.b_resident ? bres + (uint32_t)kb * 2 * b_bytes : st + 2 * a_bytes;
                    const int ksteps = min(BKT, G.Cin - kb * BKT + 7) >> 3;         // 8-wide k-steps that hold real channels (the rest is zero fill)
                    for (int j = 0; j < ksteps; ++j) {
                        const uint64_t ah = kmajor_desc<ROWB>(st + j * 32), al = kmajor_desc<ROWB>(st + a_bytes + j * 32);
                        const uint64_t bh = kmajor_desc<ROWB>(bh0 + j * 32), bl = kmajor_desc<ROWB>(bh0 + b_bytes + j * 32);
                        umma_tf32(acc, al, bh, idesc, (kb | j) != 0);
                        umma_tf32(acc, ah, bl, idesc, 1);
                        umma_tf32(acc, ah, bh, idesc, 1);
                    }
                    umma_commit(smem_addr(&s_empty[s]));                            // slot free once these MMAs have read it
                    TC_TRACE(2, 2 * g + 1);
                }
                umma_commit(smem_addr(&s_acc_full[ab]));                            // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 2 + EPW) {
        // ------------------------------------------------------------------------------------------------ operand split (128 threads)
        const int t = tid - (2 + EPW) * 32;
        uint32_t g = 0;
        for (int it = 0; it < my_tiles; ++it)
            for (int kb = 0; kb < G.KB; ++kb, ++g) {
                const uint32_t s = g % (uint32_t)G.stages, ph = (g / (uint32_t)G.stages) & 1u;
                mbar_wait(smem_addr(&s_full[s]), ph);
                if (t == 0) TC_TRACE(1, 2 * g);
                uint4* ahi = reinterpret_cast<uint4*>(gbase + (size_t)s * stage_bytes);
                uint4* alo = reinterpret_cast<uint4*>(gbase + (size_t)s * stage_bytes + a_bytes);
#pragma unroll
                for (int j = 0; j < (kBM * BKT / 4) / 128; ++j) {
                    const int i = t + j * 128;
                    const uint4 x = ahi[i];
                    uint4 h, l;
                    split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
                    ahi[i] = h; alo[i] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_addr(&s_ready[s]));
                if (t == 0) TC_TRACE(1, 2 * g + 1);
            }
    } else {
        // ------------------------------------------------------------------------------------------------ epilogue: warp w reads TMEM lane quadrant w % 4;
        // with eight epilogue warps the two warps of a quadrant take alternate 32-channel chunks
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
        for (int it = 0; it < my_tiles; ++it) {
            const uint32_t ab = (uint32_t)it & 1u, aph = ((uint32_t)it >> 1) & 1u;
            mbar_wait(smem_addr(&s_acc_full[ab]), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 64) TC_TRACE(3, 2 * it);
            const uint32_t tacc = trow + ab * (uint32_t)G.NT;
#ifdef SGS_TC_TRACE
            int tk = 0;
#define TC_TRACE_EPI() do { if (tid == 64 && tk < 16) { TC_TRACE(3, 64 + 16 * it + tk); ++tk; } } while (0)
#else
#define TC_TRACE_EPI() do { } while (0)
#endif
            const int ncol = min(G.NT, G.Cout - n0);                            // real output channels of this tile (uniform)
            const int cstep = 32 * (EPW / 4);
            for (int c0 = 32 * half; c0 < ncol; c0 += cstep) {
                const int nch = min(32, ncol - c0);                             // channels of this chunk (uniform)
                const bool last = c0 + cstep >= ncol;
                uint32_t r[2][16];
                tmem_ld16(tacc + (uint32_t)c0, r[0]);
                if (nch <= 16) {
                    // At most 16 channels (narrow layers, the left-over chunk of a tile): one pixel per lane, 64 contiguous bytes each -- the short
                    // dependency chain matters more here than the shape of the stores (tiles of narrow layers are small and come in quick succession).
                    tmem_ld_wait();
                    if (last) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_addr(&s_acc_empty[ab]));
                    }
                    const int p = ((int)blockIdx.x + it * (int)gridDim.x) * kBM + quad * 32 + lane;
                    if (G.coalesce && G.vec_ok && nch == 16 && p - lane + 32 <= G.npix) {          // the usual case as straight-line code
                        float* o = out + G.base_off + (int64_t)p * G.pitch + n0 + c0;
                        const int64_t idx = (int64_t)p * G.Cout + n0 + c0;
                        float4 bv[4];
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) bv[q4] = bias ? __ldg(reinterpret_cast<const float4*>(bias + n0 + c0) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            float v[4] = {__fadd_rn(__uint_as_float(r[0][4 * q4]), bv[q4].x), __fadd_rn(__uint_as_float(r[0][4 * q4 + 1]), bv[q4].y),
                                          __fadd_rn(__uint_as_float(r[0][4 * q4 + 2]), bv[q4].z), __fadd_rn(__uint_as_float(r[0][4 * q4 + 3]), bv[q4].w)};
                            epi.template run<4>(v, idx + 4 * q4);
                            *reinterpret_cast<float4*>(o + 4 * q4) = make_float4(v[0], v[1], v[2], v[3]);
                        }
                    } else if (p < G.npix) {
                        int64_t off = (int64_t)p * G.pitch;
                        if (!G.coalesce) { const int f = p / G.HW; off = (int64_t)f * G.frame_stride + (int64_t)(p - f * G.HW) * G.pitch; }
                        float* o = out + G.base_off + off + n0 + c0;
                        const int64_t idx = (int64_t)p * G.Cout + n0 + c0;
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            if (q < nch) {
                                float one[1] = {__fadd_rn(__uint_as_float(r[0][q]), bias ? __ldg(bias + n0 + c0 + q) : 0.f)};
                                epi.template run<1>(one, idx + q);
                                o[q] = one[0];
                            }
                        }
                    }
                    continue;
                }
                tmem_ld16(tacc + (uint32_t)c0 + 16u, r[1]);
                {
                    // The raw accumulators (one pixel per lane) are transposed through this warp's 4 KB of shared memory (16-byte pieces XOR-swizzled by the
                    // row: conflict-free both ways); on the row-major side a lane owns ONE channel quad of eight rows: its bias is fetched once, the tensor
                    // operands of the tail are read as coalesced rows, and every store instruction writes 4 rows x 128 contiguous bytes (4 memory wavefronts
                    // instead of the 32 of a one-pixel-per-lane store).
                    const int c = lane & 7, cq = 4 * c;
                    const int nq = min(4, nch - cq);                            // channels this lane owns on the row-major side: 4 = a whole quad, <= 0 = none
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias) {
                        if (nq == 4) bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + cq));
                        else if (nq > 0) { bv.x = __ldg(bias + n0 + c0 + cq); if (nq > 1) bv.y = __ldg(bias + n0 + c0 + cq + 1); if (nq > 2) bv.z = __ldg(bias + n0 + c0 + cq + 2); }
                    }
                    float4* stg = reinterpret_cast<float4*>(gbase + epi_off + (size_t)(warp - 2) * 4096);
                    tmem_ld_wait();
                    TC_TRACE_EPI();
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            stg[lane * 8 + ((4 * h + q) ^ (lane & 7))] = make_float4(__uint_as_float(r[h][4 * q]), __uint_as_float(r[h][4 * q + 1]), __uint_as_float(r[h][4 * q + 2]), __uint_as_float(r[h][4 * q + 3]));
                    if (last) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (last && lane == 0) mbar_arrive(smem_addr(&s_acc_empty[ab]));          // every TMEM read of this warp's quadrant has completed: the MMAs of tile it + 2 may start
                    TC_TRACE_EPI();
                    const int p0 = ((int)blockIdx.x + it * (int)gridDim.x) * kBM + quad * 32;
                    if (G.coalesce && G.vec_ok && (nch & 3) == 0 && p0 + 32 <= G.npix) {
                        // the usual case as straight-line code: whole quads, all 32 rows inside the layer, contiguous rows -- one pointer per lane, constant strides,
                        // the eight shared-memory reads in flight together
                        if (nq == 4) {
                            const int r0 = lane >> 3;
                            float* o = out + G.base_off + (int64_t)(p0 + r0) * G.pitch + n0 + c0 + cq;
                            const int64_t idx = (int64_t)(p0 + r0) * G.Cout + n0 + c0 + cq;
                            const int64_t ostep = 4 * (int64_t)G.pitch, istep = 4 * (int64_t)G.Cout;
                            float4 x[8];
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) x[i4] = stg[(i4 * 4 + r0) * 8 + (c ^ ((i4 * 4 + r0) & 7))];
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) {
                                float v[4] = {__fadd_rn(x[i4].x, bv.x), __fadd_rn(x[i4].y, bv.y), __fadd_rn(x[i4].z, bv.z), __fadd_rn(x[i4].w, bv.w)};
                                epi.template run<4>(v, idx + i4 * istep);
                                *reinterpret_cast<float4*>(o + i4 * ostep) = make_float4(v[0], v[1], v[2], v[3]);
                            }
                        }
                    } else if (nq > 0) {
                        // the last rows of a layer, slices of a concatenation buffer, unaligned rows, channel counts that are not multiples of 4
                        const bool vec = nq == 4 && G.vec_ok;
#pragma unroll 1
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const int row = i4 * 4 + (lane >> 3), p = p0 + row;
                            if (p < G.npix) {
                                const float4 x = stg[row * 8 + (c ^ (row & 7))];
                                float v[4] = {__fadd_rn(x.x, bv.x), __fadd_rn(x.y, bv.y), __fadd_rn(x.z, bv.z), __fadd_rn(x.w, bv.w)};
                                const int64_t idx = (int64_t)p * G.Cout + n0 + c0 + cq;
                                int64_t off = (int64_t)p * G.pitch;
                                if (!G.coalesce) { const int f = p / G.HW; off = (int64_t)f * G.frame_stride + (int64_t)(p - f * G.HW) * G.pitch; }
                                float* o = out + G.base_off + off + n0 + c0 + cq;
                                if (vec) {
                                    epi.template run<4>(v, idx);
                                    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                                } else {
#pragma unroll 1
                                    for (int q = 0; q < nq; ++q) {
                                        float one[1] = {q == 0 ? v[0] : q == 1 ? v[1] : q == 2 ? v[2] : v[3]};
                                        epi.template run<1>(one, idx + q);
                                        o[q] = one[0];
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    TC_TRACE_EPI();
                }
            }
            if (half == 1 && 32 >= ncol) {                                       // this warp had no chunk: nothing of the accumulator to wait for
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_addr(&s_acc_empty[ab]));
            }
            if (tid == 64) TC_TRACE(3, 2 * it + 1);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    __syncwarp();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(G.tmem_cols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_tmap_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tmap_encode tmap_encoder() {
    static PFN_tmap_encode fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_tmap_encode)p;
        cudaGetLastError();
    }
    return fn;
}
// 2-D FP32 tensor [rows][cols] with `pitch` floats between rows; box = bk floats (one 128- or 64-byte swizzle row) x box_rows, zero fill outside
inline bool encode_kmajor_map(CUtensorMap* m, const float* base, int64_t rows, int cols, int64_t pitch, int box_rows, int bk) {
    PFN_tmap_encode fn = tmap_encoder();
    if (!fn || ((uintptr_t)base & 15) || (pitch & 3) || box_rows < 1 || box_rows > 256 || (bk != 32 && bk != 16)) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t est[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct GemmPlan {
    int Cin = 0, Cout = 0, NT = 0, n_tiles = 0, KB = 0, BK = kBK, epw = 4, stages = 0, smem_bytes = 0, Kp = 0, Np = 0, b_resident = 0, ctas_per_sm = 1;
    uint32_t tmem_cols = 0;
    float* d_whi = nullptr; float* d_wlo = nullptr;
    CUtensorMap map_hi, map_lo;
};

inline void split_tf32_host(float x, float& hi, float& lo) {
    uint32_t b; memcpy(&b, &x, 4);
    uint32_t h = (b + 0x1000u) & 0xffffe000u;
    memcpy(&hi, &h, 4);
    const float r = x - hi;       // exact
    memcpy(&b, &r, 4);
    uint32_t l = (b + 0x1000u) & 0xffffe000u;
    memcpy(&lo, &l, 4);
}

// Tiling of one layer (no device access): output-channel tiles of at most kMaxNT channels, weights resident when all their k-blocks fit in 96 KB,
// ring depth from what is left of the shared memory; small weight tiles leave room for two CTAs per SM (more loads in flight, two MMA issuers).
// m_tiles (128-pixel tiles of the largest batch, 0 = unknown) picks the number of output-channel tiles: every tile of a pixel block re-reads the
// activations and pays the pipeline's fill, so fewer, wider tiles win as long as they still fill the machine -- measured (profiles/r02y_nt_sweep.txt):
// 112->672 at 19x19 takes 76 us with 6 x 112, 58 us with 3 x 224; 256->512 at 5x5 (25 pixel tiles) is fastest with 4 x 128.  The cost of a plan is
// modelled as  waves of the persistent grid x (NT + 100)  and the cheapest tile count between Cout / 256 and Cout / 128 is taken.
constexpr int kPlanSMs = 148;
inline void plan_tiling(int Cin, int Cout, GemmPlan* P, int force_nt = 0, int m_tiles = 0) {
    P->Cin = Cin; P->Cout = Cout;
    int nt = (Cout + 127) / 128;
    if (m_tiles > 0) {
        double best = 1e300;
        for (int n = (Cout + kMaxNT - 1) / kMaxNT; n <= (Cout + 127) / 128; ++n) {
            const int w = ((Cout + n - 1) / n + 15) & ~15;
            const long long waves = ((long long)m_tiles * n + kPlanSMs - 1) / kPlanSMs;
            const double cost = (double)waves * (w + 100);
            if (cost < best) { best = cost; nt = n; }
        }
    }
    P->NT = force_nt > 0 ? force_nt : ((Cout + nt - 1) / nt + 15) & ~15;
    P->n_tiles = (Cout + P->NT - 1) / P->NT;
    P->BK = Cin <= 16 ? 16 : kBK;                                        // at most 16 input channels: 64-byte operand rows, half the ring slot
    P->KB = (Cin + P->BK - 1) / P->BK;
    P->Kp = P->KB * P->BK; P->Np = P->n_tiles * P->NT;
    const int a_stage = 2 * kBM * P->BK * 4, b_block = 2 * P->NT * P->BK * 4;
    const int b_all = P->KB * b_block;
    P->b_resident = b_all <= 96 * 1024 ? 1 : 0;
    // resident weights + two ring slots + the four transpose buffers fit twice: two CTAs per SM
    P->ctas_per_sm = (P->b_resident && b_all + 2 * a_stage + 4 * 4096 + 1024 <= 110 * 1024) ? 2 : 1;
    P->epw = P->ctas_per_sm == 2 ? 4 : 8;                                // one CTA per SM: eight epilogue warps keep up with wide output tiles
    const int epi_stage = P->epw * 4096;                                 // the epilogue warps' transpose buffers
    const int budget = (P->ctas_per_sm == 2 ? 110 : 226) * 1024 - 1024 - epi_stage - (P->b_resident ? b_all : 0);
    const int stage_bytes = a_stage + (P->b_resident ? 0 : b_block);
    int st = budget / stage_bytes;
    if (st > kMaxStages) st = kMaxStages;
    if (st < 2) st = 2;
    P->stages = st;
    P->smem_bytes = st * stage_bytes + (P->b_resident ? b_all : 0) + epi_stage + 1024;
    uint32_t tc = 32;
    while ((int)tc < 2 * P->NT) tc <<= 1;
    P->tmem_cols = tc;
}

// Splits and uploads W [Cout][Cin], picks the tiling.  Returns false when TMA is unavailable or an allocation fails.
inline bool plan_weights(const float* W, int Cin, int Cout, GemmPlan* P, int force_nt = 0, int m_tiles = 0) {
    plan_tiling(Cin, Cout, P, force_nt, m_tiles);
    std::vector<float> hi((size_t)P->Np * P->Kp, 0.f), lo((size_t)P->Np * P->Kp, 0.f);
    for (int co = 0; co < Cout; ++co)
        for (int c = 0; c < Cin; ++c) split_tf32_host(W[(size_t)co * Cin + c], hi[(size_t)co * P->Kp + c], lo[(size_t)co * P->Kp + c]);
    if (cudaMalloc((void**)&P->d_whi, hi.size() * 4) != cudaSuccess) return false;
    if (cudaMalloc((void**)&P->d_wlo, lo.size() * 4) != cudaSuccess) return false;
    if (cudaMemcpy(P->d_whi, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    if (cudaMemcpy(P->d_wlo, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    return encode_kmajor_map(&P->map_hi, P->d_whi, P->Np, P->Kp, P->Kp, P->NT, P->BK) && encode_kmajor_map(&P->map_lo, P->d_wlo, P->Np, P->Kp, P->Kp, P->NT, P->BK);
}
inline void free_plan(GemmPlan* P) { cudaFree(P->d_whi); cudaFree(P->d_wlo); P->d_whi = P->d_wlo = nullptr; }

// ---- stand-alone tails of the unit harness (the detector passes its own functor built from the ncnn element-wise chain)
enum { TK_NONE = 0, TK_RELU, TK_CLIP, TK_HSWISH, TK_ADD_T, TK_SE_TAIL };
// x / c.  For c == 6 (the hard-swish / hard-sigmoid divisor of the model) the quotient comes from the reciprocal and one FMA correction:
// q = RN(x * r), q' = RN(q + RN(x - 6q) * r), which equals the correctly rounded RN(x / 6) for every float with |x| >= 2^-100 and for zeros up to the
// sign of zero (verified exhaustively on the host over all 2^32 bit patterns); below 2^-100 it may differ from the IEEE quotient in the last bit of a
// denormal.  Both the fused and the layer-by-layer execution use this function.  (The generic IEEE division takes its slow path whenever the dividend is
// zero -- half of all hard-swish outputs.)
__device__ __forceinline__ float div6(float x) {         // correctly rounded x / 6 without the division subroutine
    const float r = 0x1.555556p-3f;                     // RN(1/6)
    const float q = __fmul_rn(x, r);
    return __fmaf_rn(__fmaf_rn(-6.0f, q, x), r, q);
}
__device__ __forceinline__ float div_scalar(float x, float c) { return c == 6.0f ? div6(x) : __fdiv_rn(x, c); }

struct GemmTail {                  // the unit harness' description of a tail (tools/umma_proto); the detector has its own functor (detector.cu: EpiFnK)
    int kind;
    float a, lo, hi, b;
    const float* t1; const float* t2;
};
template <int KIND>
struct GemmTailK {                 // tail kind fixed at compile time: the epilogue carries the code of one tail only
    GemmTail t;
    template <int N>
    __device__ __forceinline__ void run(float (&v)[N], int64_t idx0) const {
#pragma unroll
        for (int q = 0; q < N; ++q) {
            if constexpr (KIND == TK_RELU) v[q] = fmaxf(v[q], 0.f);
            else if constexpr (KIND == TK_CLIP) v[q] = fminf(fmaxf(v[q], t.lo), t.hi);
            else if constexpr (KIND == TK_HSWISH) v[q] = div6(__fmul_rn(v[q], fminf(fmaxf(__fadd_rn(v[q], t.a), t.lo), t.hi)));
            else if constexpr (KIND == TK_ADD_T) v[q] = __fadd_rn(v[q], __ldg(t.t1 + idx0 + q));
            else if constexpr (KIND == TK_SE_TAIL) v[q] = __fadd_rn(__fmul_rn(__ldg(t.t1 + idx0 + q), div6(fminf(fmaxf(__fadd_rn(v[q], t.a), t.lo), t.hi))), __ldg(t.t2 + idx0 + q));
        }
    }
};

inline int sm_count() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148; }
    return n;
}

// mapA: encode_kmajor_map over the input activations (any number of rows >= npix; box kBM rows, P.BK columns)
template <class EpiFn>
inline bool launch_conv1x1_tc_map(const GemmPlan& P, const CUtensorMap& mapA, int npix, const float* bias, float* out, int HW, int64_t frame_stride, int64_t base_off,
                                  int pitch, const EpiFn& epi, cudaStream_t st) {
    TcGeom G;
    G.npix = npix; G.Cin = P.Cin; G.Cout = P.Cout; G.NT = P.NT; G.KB = P.KB; G.stages = P.stages; G.tmem_cols = P.tmem_cols;
    G.m_tiles = (npix + kBM - 1) / kBM; G.b_resident = P.b_resident;
    G.HW = HW; G.frame_stride = frame_stride; G.base_off = base_off; G.pitch = pitch;
    G.vec_ok = ((pitch & 3) == 0 && (frame_stride & 3) == 0 && (base_off & 3) == 0 && (((uintptr_t)out) & 15) == 0) ? 1 : 0;
    G.coalesce = (frame_stride == 0 || frame_stride == (int64_t)HW * pitch) ? 1 : 0;
    int gx = (sm_count() * P.ctas_per_sm) / P.n_tiles;                  // one wave of persistent CTAs
    if (gx < 1) gx = 1;
    if (gx > G.m_tiles) gx = G.m_tiles;
    const dim3 grid(gx, P.n_tiles);
    auto go = [&](auto kern, int threads) -> bool {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) return false;   // cheap; per device
        kern<<<grid, threads, P.smem_bytes, st>>>(mapA, P.map_hi, P.map_lo, bias, out, G, epi);
        return true;
    };
    bool ok;
    if (P.BK == 16) ok = P.epw == 4 ? go(conv1x1_tc_kernel<EpiFn, 16, 4>, tc_threads(4)) : go(conv1x1_tc_kernel<EpiFn, 16, 8>, tc_threads(8));
    else ok = P.epw == 4 ? go(conv1x1_tc_kernel<EpiFn, 32, 4>, tc_threads(4)) : go(conv1x1_tc_kernel<EpiFn, 32, 8>, tc_threads(8));
    if (!ok) return false;
    return cudaGetLastError() == cudaSuccess;
}

inline bool launch_conv1x1_tc(const GemmPlan& P, const float* x, int in_pitch, int npix, const float* bias, float* out, int out_pitch, const GemmTail& T, cudaStream_t st) {
    CUtensorMap mapA;
    if (!encode_kmajor_map(&mapA, x, npix, P.Cin, in_pitch, kBM, P.BK)) return false;
    const int HW = npix > 0 ? npix : 1;
    switch (T.kind) {
    case TK_RELU: return launch_conv1x1_tc_map(P, mapA, npix, bias, out, HW, 0, 0, out_pitch, GemmTailK<TK_RELU>{T}, st);
    case TK_CLIP: return launch_conv1x1_tc_map(P, mapA, npix, bias, out, HW, 0, 0, out_pitch, GemmTailK<TK_CLIP>{T}, st);
    case TK_HSWISH: return launch_conv1x1_tc_map(P, mapA, npix, bias, out, HW, 0, 0, out_pitch, GemmTailK<TK_HSWISH>{T}, st);
    case TK_ADD_T: return launch_conv1x1_tc_map(P, mapA, npix, bias, out, HW, 0, 0, out_pitch, GemmTailK<TK_ADD_T>{T}, st);
    case TK_SE_TAIL: return launch_conv1x1_tc_map(P, mapA, npix, bias, out, HW, 0, 0, out_pitch, GemmTailK<TK_SE_TAIL>{T}, st);
    default: return launch_conv1x1_tc_map(P, mapA, npix, bias, out, HW, 0, 0, out_pitch, GemmTailK<TK_NONE>{T}, st);
    }
}

}  // namespace tc
}  // namespace sgs
