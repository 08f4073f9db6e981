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
constexpr int kBK = 32;         // floats per k-block = one 128-byte swizzle row
constexpr int kMaxStages = 4;
constexpr int kMaxNT = 128;      // output channels per tile (TMEM columns per CTA)

// ---------------------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
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

// Shared-memory matrix descriptor of a K-major tile whose rows are 128 bytes (one swizzle row) and whose 8-row groups are 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start >> 4 | LBO << 16 | SBO << 32 | version 1 << 46 | layout SWIZZLE_128B (2) << 61).
__device__ __forceinline__ uint64_t kmajor_sw128_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ inline uint32_t tf32_idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24); }

// x = hi + lo, both TF32 (round to nearest, ties away, by integer arithmetic on the bit pattern)
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
    hi = (x + 0x1000u) & 0xffffe000u;
    lo = (__float_as_uint(__fsub_rn(__uint_as_float(x), __uint_as_float(hi))) + 0x1000u) & 0xffffe000u;
}

struct TcGeom {
    int npix, Cin, Cout;          // rows of X, channels in / out
    int NT, KB, stages;           // output channels per tile (multiple of 16), k-blocks, pipeline depth
    uint32_t tmem_cols;           // power of two >= max(32, NT)
    int HW;                       // pixels per frame (output addressing)
    int64_t frame_stride;         // floats between consecutive frames of the output
    int64_t base_off;             // float offset of frame 0 of the output inside `out`
    int pitch;                    // floats between consecutive pixels of the output
    int vec_ok;                   // 16-byte aligned rows: float4 stores
};

// EpiFn: struct with  template <int N> __device__ void run(float (&v)[N], int64_t idx0) const   applied to N consecutive channels of one pixel;
// idx0 = pixel * Cout + channel (the NHWC index of same-shape operand tensors).
template <class EpiFn>
__global__ void __launch_bounds__(kBM) conv1x1_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                                                         const __grid_constant__ CUtensorMap mapBl, const float* __restrict__ bias, float* __restrict__ out,
                                                         const TcGeom G, const EpiFn epi) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kMaxStages], s_empty[kMaxStages], s_accum;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * G.NT;
    // stage layout: A (hi in place of the landed FP32 tile) | A lo | B hi | B lo, every part a multiple of 1024 bytes
    const uint32_t a_bytes = kBM * kBK * 4, b_bytes = (uint32_t)G.NT * kBK * 4, stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const uint32_t sbase = (smem_addr(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (sbase - smem_addr(smem_raw));

    if (tid == 0) {
        for (int s = 0; s < G.stages; ++s) { mbar_init(smem_addr(&s_full[s]), 1); mbar_init(smem_addr(&s_empty[s]), 1); }
        mbar_init(smem_addr(&s_accum), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "r"(G.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    auto issue_loads = [&](int kb) {
        const int s = kb % G.stages;
        const uint32_t st = sbase + (uint32_t)s * stage_bytes, bar = smem_addr(&s_full[s]);
        mbar_expect_tx(bar, a_bytes + 2 * b_bytes);
        tma_load_2d(st, &mapA, kb * kBK, m0, bar);
        tma_load_2d(st + 2 * a_bytes, &mapBh, kb * kBK, n0, bar);
        tma_load_2d(st + 2 * a_bytes + b_bytes, &mapBl, kb * kBK, n0, bar);
    };
    if (tid == 0)
        for (int kb = 0; kb < G.KB && kb < G.stages; ++kb) issue_loads(kb);

    const uint32_t idesc = tf32_idesc(G.NT);
    for (int kb = 0; kb < G.KB; ++kb) {
        const int s = kb % G.stages;
        const uint32_t ph = (uint32_t)(kb / G.stages) & 1u;
        mbar_wait(smem_addr(&s_full[s]), ph);
        // split the activation tile: hi in place, lo next to it (same swizzled position)
        uint4* ahi = reinterpret_cast<uint4*>(gbase + (size_t)s * stage_bytes);
        uint4* alo = reinterpret_cast<uint4*>(gbase + (size_t)s * stage_bytes + a_bytes);
#pragma unroll
        for (int j = 0; j < (kBM * kBK / 4) / kBM; ++j) {
            const int i = tid + j * kBM;
            const uint4 x = ahi[i];
            uint4 h, l;
            split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
            ahi[i] = h; alo[i] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t st = sbase + (uint32_t)s * stage_bytes;
            const int ksteps = min(kBK, G.Cin - kb * kBK + 7) >> 3;          // 8-wide k-steps that hold real channels (the rest of the block is zero fill)
            for (int j = 0; j < ksteps; ++j) {
                const uint64_t ah = kmajor_sw128_desc(st + j * 32), al = kmajor_sw128_desc(st + a_bytes + j * 32);
                const uint64_t bh = kmajor_sw128_desc(st + 2 * a_bytes + j * 32), bl = kmajor_sw128_desc(st + 2 * a_bytes + b_bytes + j * 32);
                umma_tf32(tmem, al, bh, idesc, (kb | j) != 0);
                umma_tf32(tmem, ah, bl, idesc, 1);
                umma_tf32(tmem, ah, bh, idesc, 1);
            }
            umma_commit(smem_addr(&s_empty[s]));
            if (kb + G.stages < G.KB) {               // refill this stage once its MMAs have read it
                mbar_wait(smem_addr(&s_empty[s]), ph);
                issue_loads(kb + G.stages);
            } else if (kb == G.KB - 1) {
                umma_commit(smem_addr(&s_accum));
            }
        }
    }
    mbar_wait(smem_addr(&s_accum), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: this thread owns TMEM lane 32*warp + lane = pixel m0 + tid
    const int p = m0 + tid;
    const bool live = p < G.npix;
    float* orow = nullptr;
    int64_t idx_row = 0;
    if (live) {
        const int f = p / G.HW, pl = p - f * G.HW;
        orow = out + G.base_off + (int64_t)f * G.frame_stride + (int64_t)pl * G.pitch;
        idx_row = (int64_t)p * G.Cout;
    }
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < G.NT; c0 += 32) {
        if (n0 + c0 >= G.Cout) break;                                   // uniform: padding columns of the last tile
        uint32_t r[2][16];
        const bool two = c0 + 16 < G.NT && n0 + c0 + 16 < G.Cout;       // uniform
        tmem_ld16(trow + (uint32_t)c0, r[0]);
        if (two) tmem_ld16(trow + (uint32_t)c0 + 16u, r[1]);
        tmem_ld_wait();
        if (!live) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const int co = n0 + c0 + 16 * h;
            const int nv = min(16, G.Cout - co);
            float v[16];
            if (nv == 16) {
                if (bias) {
#pragma unroll
                    for (int q = 0; q < 16; q += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + co + q));
                        v[q] = __fadd_rn(__uint_as_float(r[h][q]), b.x); v[q + 1] = __fadd_rn(__uint_as_float(r[h][q + 1]), b.y);
                        v[q + 2] = __fadd_rn(__uint_as_float(r[h][q + 2]), b.z); v[q + 3] = __fadd_rn(__uint_as_float(r[h][q + 3]), b.w);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[h][q]);
                }
                epi.template run<16>(v, idx_row + co);
                if (G.vec_ok) {
#pragma unroll
                    for (int q = 0; q < 16; q += 4) *reinterpret_cast<float4*>(orow + co + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) orow[co + q] = v[q];
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    if (q < nv) {
                        float one[1] = {__fadd_rn(__uint_as_float(r[h][q]), bias ? __ldg(bias + co + q) : 0.f)};
                        epi.template run<1>(one, idx_row + co + q);
                        orow[co + q] = one[0];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(G.tmem_cols) : "memory");
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
// 2-D FP32 tensor [rows][cols] with `pitch` floats between rows; box = 32 floats x box_rows, 128-byte swizzle, zero fill outside
inline bool encode_kmajor_map(CUtensorMap* m, const float* base, int64_t rows, int cols, int64_t pitch, int box_rows) {
    PFN_tmap_encode fn = tmap_encoder();
    if (!fn || ((uintptr_t)base & 15) || (pitch & 3) || box_rows < 1 || box_rows > 256) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t est[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct GemmPlan {
    int Cin = 0, Cout = 0, NT = 0, n_tiles = 0, KB = 0, stages = 0, smem_bytes = 0, Kp = 0, Np = 0;
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

// Splits and uploads W [Cout][Cin], picks the tiling.  Returns false when TMA is unavailable or an allocation fails.
inline bool plan_weights(const float* W, int Cin, int Cout, GemmPlan* P) {
    P->Cin = Cin; P->Cout = Cout;
    const int nt = (Cout + kMaxNT - 1) / kMaxNT;                         // output-channel tiles of at most kMaxNT (several CTAs stay resident per SM)
    P->NT = ((Cout + nt - 1) / nt + 15) & ~15;
    P->n_tiles = (Cout + P->NT - 1) / P->NT;
    P->KB = (Cin + kBK - 1) / kBK;
    P->Kp = P->KB * kBK; P->Np = P->n_tiles * P->NT;
    const int stage_bytes = 2 * kBM * kBK * 4 + 2 * P->NT * kBK * 4;
    int st = P->KB >= 3 ? 2 : 1;                                         // short contractions: all loads are in flight at once anyway
    if (st > P->KB) st = P->KB;
    P->stages = st;
    P->smem_bytes = st * stage_bytes + 1024;
    uint32_t tc = 32;
    while ((int)tc < P->NT) tc <<= 1;
    P->tmem_cols = tc;
    std::vector<float> hi((size_t)P->Np * P->Kp, 0.f), lo((size_t)P->Np * P->Kp, 0.f);
    for (int co = 0; co < Cout; ++co)
        for (int c = 0; c < Cin; ++c) split_tf32_host(W[(size_t)co * Cin + c], hi[(size_t)co * P->Kp + c], lo[(size_t)co * P->Kp + c]);
    if (cudaMalloc((void**)&P->d_whi, hi.size() * 4) != cudaSuccess) return false;
    if (cudaMalloc((void**)&P->d_wlo, lo.size() * 4) != cudaSuccess) return false;
    if (cudaMemcpy(P->d_whi, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    if (cudaMemcpy(P->d_wlo, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    return encode_kmajor_map(&P->map_hi, P->d_whi, P->Np, P->Kp, P->Kp, P->NT) && encode_kmajor_map(&P->map_lo, P->d_wlo, P->Np, P->Kp, P->Kp, P->NT);
}
inline void free_plan(GemmPlan* P) { cudaFree(P->d_whi); cudaFree(P->d_wlo); P->d_whi = P->d_wlo = nullptr; }

// ---- stand-alone tails of the unit harness (the detector passes its own functor built from the ncnn element-wise chain)
enum { TK_NONE = 0, TK_RELU, TK_CLIP, TK_HSWISH, TK_ADD_T, TK_SE_TAIL };
struct GemmTail {
    int kind;
    float a, lo, hi, b;
    const float* t1; const float* t2;
    template <int N>
    __device__ __forceinline__ void run(float (&v)[N], int64_t idx0) const {
#pragma unroll
        for (int q = 0; q < N; ++q) {
            float x = v[q];
            switch (kind) {
            case TK_RELU: x = fmaxf(x, 0.f); break;
            case TK_CLIP: x = fminf(fmaxf(x, lo), hi); break;
            case TK_HSWISH: x = __fdiv_rn(__fmul_rn(x, fminf(fmaxf(__fadd_rn(x, a), lo), hi)), b); break;
            case TK_ADD_T: x = __fadd_rn(x, __ldg(t1 + idx0 + q)); break;
            case TK_SE_TAIL: x = __fadd_rn(__fmul_rn(__ldg(t1 + idx0 + q), __fdiv_rn(fminf(fmaxf(__fadd_rn(x, a), lo), hi), b)), __ldg(t2 + idx0 + q)); break;
            default: break;
            }
            v[q] = x;
        }
    }
};

template <class EpiFn>
inline bool launch_conv1x1_tc_geom(const GemmPlan& P, const float* x, int64_t in_pitch, int npix, const float* bias, float* out, int HW, int64_t frame_stride,
                                   int64_t base_off, int pitch, const EpiFn& epi, cudaStream_t st) {
    CUtensorMap mapA;
    if (!encode_kmajor_map(&mapA, x, npix, P.Cin, in_pitch, kBM)) return false;
    TcGeom G;
    G.npix = npix; G.Cin = P.Cin; G.Cout = P.Cout; G.NT = P.NT; G.KB = P.KB; G.stages = P.stages; G.tmem_cols = P.tmem_cols;
    G.HW = HW; G.frame_stride = frame_stride; G.base_off = base_off; G.pitch = pitch;
    G.vec_ok = ((pitch & 3) == 0 && (frame_stride & 3) == 0 && (base_off & 3) == 0 && (((uintptr_t)out) & 15) == 0) ? 1 : 0;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv1x1_tc_kernel<EpiFn>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return false;
        attr_set = true;
    }
    const dim3 grid((npix + kBM - 1) / kBM, P.n_tiles);
    conv1x1_tc_kernel<EpiFn><<<grid, kBM, P.smem_bytes, st>>>(mapA, P.map_hi, P.map_lo, bias, out, G, epi);
    return cudaGetLastError() == cudaSuccess;
}

inline bool launch_conv1x1_tc(const GemmPlan& P, const float* x, int in_pitch, int npix, const float* bias, float* out, int out_pitch, const GemmTail& T, cudaStream_t st) {
    return launch_conv1x1_tc_geom(P, x, in_pitch, npix, bias, out, npix > 0 ? npix : 1, 0, 0, out_pitch, T, st);
}

}  // namespace tc
}  // namespace sgs
