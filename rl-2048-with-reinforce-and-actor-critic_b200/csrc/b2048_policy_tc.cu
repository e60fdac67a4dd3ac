// b2048_policy_tc.cu — K3 policy_step on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands,
// fp32 accumulation.  Used for batches >= 4096 boards with the runner-default policy shape
// (16 -> 256 -> 256 -> 4, ReLU); everything else runs on the fp32 CUDA-core path (b2048_policy.cu).
//
// One CTA (128 threads = 4 warps = the 128 TMEM lanes) owns a tile of 128 boards and loops over tiles:
//
//   A1 [128 x 16] bf16  <- packed boards (each thread encodes its own board: 16 nibbles -> 16 bf16)
//   D1 = A1 . W1^T      one  tcgen05.mma  M128 N256 K16          -> TMEM columns   0..255
//   A2 = relu(D1 + b1)  tcgen05.ld 32x32b, bias + ReLU, bf16 pack, st.shared into the 128B-swizzled
//                       K-major operand layout                      (activations never leave the SM)
//   D2 = A2 . W2^T      sixteen tcgen05.mma M128 N256 K16         -> TMEM columns 256..511
//   logits = relu(D2 + b2) . W3 + b3   in the epilogue on CUDA cores (N = 4 is below the UMMA minimum),
//   masked softmax + inverse-CDF sample / greedy, one action byte per board.
//
// Weights arrive as a pre-arranged shared-memory IMAGE (bf16, already in the UMMA canonical layouts) that a
// small prep kernel builds from the reference-layout fp32 parameters; each CTA pulls the 142 KB image with
// bulk async copies (cp.async.bulk + mbarrier) once and keeps it resident for all of its tiles.
//
// Reference arithmetic: encode_observation / forward_logits / logits_to_probs / select_action
// (src/MLP.py:22-43, :139-196; src/reinforce_agent.py:126-192).  Parity bar for this path: 1e-2 relative.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "b2048_device.cuh"
#include "b2048_internal.h"

namespace b2 {

constexpr int TC_M = 128;          // boards per tile = TMEM lanes
constexpr int TC_H = 256;          // hidden width (both layers)
constexpr int TC_K1 = 16;          // input width

// ---- shared-memory image (byte offsets).  SW128 K-major slabs must be 1024-byte aligned.
constexpr int IMG_W2 = 0;                          // 4 slabs [256 rows x 128 B] = 131072 B (SWIZZLE_128B, K-major)
constexpr int IMG_W1 = 131072;                     // [32 row-groups][2 k-chunks][8 rows][16 B] = 8192 B (no swizzle)
constexpr int IMG_W3 = IMG_W1 + 8192;              // float [256][4] = 4096 B
constexpr int IMG_B1 = IMG_W3 + 4096;              // float [256]
constexpr int IMG_B2 = IMG_B1 + 1024;              // float [256]
constexpr int IMG_B3 = IMG_B2 + 1024;              // float [4] (+ pad to 16)
constexpr int IMG_BYTES = IMG_B3 + 16;             // 145424
// ---- per-CTA working buffers after the image
constexpr int SM_A2 = ((IMG_BYTES + 1023) / 1024) * 1024;   // 4 slabs [128 rows x 128 B] = 65536 B (SWIZZLE_128B)
constexpr int SM_A1 = SM_A2 + 65536;                        // [16 row-groups][2][8][16 B] = 4096 B (no swizzle)
constexpr int SM_BAR = SM_A1 + 4096;                        // mbarriers + tmem base
constexpr int SM_TOTAL = SM_BAR + 64;

// ------------------------------------------------------------------------------------------------ image prep
// W (reference layout [in][out] fp32) -> bf16 UMMA B operands stored [n][k] K-major.
__global__ void __launch_bounds__(256) policy_tc_prepare_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                                                 const float* __restrict__ W2, const float* __restrict__ b2,
                                                                 const float* __restrict__ W3, const float* __restrict__ b3,
                                                                 uint8_t* __restrict__ img) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    // W2: element (n, k) -> slab k/64, row n, 16-byte chunk ((k%64)/8) ^ (n%8), element k%8
    for (int idx = tid; idx < TC_H * TC_H; idx += nth) {
        int k = idx / TC_H, n = idx - k * TC_H;          // W2[k][n] is contiguous in n: coalesced reads
        int slab = k >> 6, kc = (k & 63) >> 3, ke = k & 7;
        size_t off = (size_t)IMG_W2 + (size_t)slab * 32768 + (size_t)n * 128 + (size_t)((kc ^ (n & 7)) * 16) + ke * 2;
        *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(W2[idx]);
    }
    // W1: element (n, k), k < 16 -> row-group n/8, k-chunk k/8, row n%8, element k%8 (no swizzle)
    for (int idx = tid; idx < TC_K1 * TC_H; idx += nth) {
        int k = idx / TC_H, n = idx - k * TC_H;
        size_t off = (size_t)IMG_W1 + (size_t)(n >> 3) * 256 + (size_t)(k >> 3) * 128 + (size_t)(n & 7) * 16 + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(W1[idx]);
    }
    for (int idx = tid; idx < TC_H * 4; idx += nth) reinterpret_cast<float*>(img + IMG_W3)[idx] = W3[idx];
    for (int idx = tid; idx < TC_H; idx += nth) {
        reinterpret_cast<float*>(img + IMG_B1)[idx] = b1[idx];
        reinterpret_cast<float*>(img + IMG_B2)[idx] = b2[idx];
    }
    if (tid < 4) reinterpret_cast<float*>(img + IMG_B3)[tid] = b3[tid];
}

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
// SWIZZLE_128B K-major operand: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// no-swizzle K-major operand with K = 16: core matrix = 8 rows x 16 B (128 B contiguous);
// the second 16-byte K chunk is LBO = 128 B away, the next 8-row group SBO = 256 B away
__device__ __forceinline__ uint64_t desc_nosw_k16(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_H >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

struct PolicyTcArgs {
    const uint8_t* img;
    const uint64_t* board;
    const uint8_t* mask_flags;
    uint8_t* action;
    float* probs;
    float* logits;
    int64_t n;
    PhiloxKeys keys;
    uint64_t gid0;
    uint32_t t;
    int greedy;
    int obs_mode;
    float obs_scale;
};

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(TC_M, 1) policy_tc_kernel(const __grid_constant__ PolicyTcArgs args) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    const uint32_t bar_img = s_u32(&bars[0]), bar_mma = s_u32(&bars[1]);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 32);

    if (tid == 0) {
        mbar_init(bar_img, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // one warp allocates all 512 TMEM columns (two fp32 accumulators of 256 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {    // weight image -> shared memory (bulk async copies, completion on bar_img)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_img), "r"((uint32_t)IMG_BYTES)
                     : "memory");
        constexpr uint32_t kChunk = 16384;
        for (uint32_t off = 0; off < (uint32_t)IMG_BYTES; off += kChunk) {
            uint32_t sz = (uint32_t)IMG_BYTES - off < kChunk ? (uint32_t)IMG_BYTES - off : kChunk;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             s_u32(smem + off)),
                         "l"(args.img + off), "r"(sz), "r"(bar_img)
                         : "memory");
        }
    }

    const float* sW3 = reinterpret_cast<const float*>(smem + IMG_W3);
    const float* sB1 = reinterpret_cast<const float*>(smem + IMG_B1);
    const float* sB2 = reinterpret_cast<const float*>(smem + IMG_B2);
    const float* sB3 = reinterpret_cast<const float*>(smem + IMG_B3);
    const uint32_t sA1 = s_u32(smem + SM_A1), sA2 = s_u32(smem + SM_A2);
    const uint32_t sW1 = s_u32(smem + IMG_W1), sW2 = s_u32(smem + IMG_W2);
    const uint32_t taddr_lane = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
    uint32_t mma_phase = 0;
    bool img_ready = false;

    const int64_t n_tiles = (args.n + TC_M - 1) / TC_M;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t s = tile * TC_M + tid;
        const bool valid = s < args.n;
        // ---- A1: this thread's board as 16 bf16 (row = tid), no-swizzle K-major core matrices
        uint64_t bd = valid ? args.board[s] : 0ull;
        uint32_t packed[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t e0 = (uint32_t)(bd >> (8 * j)) & 0xFu, e1 = (uint32_t)(bd >> (8 * j + 4)) & 0xFu;
            float v0, v1;
            if (args.obs_mode == B2048_OBS_RAW) { v0 = e0 ? (float)(1u << e0) : 0.0f; v1 = e1 ? (float)(1u << e1) : 0.0f; }
            else { v0 = (float)e0 * args.obs_scale; v1 = (float)e1 * args.obs_scale; }
            __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
            packed[j] = *reinterpret_cast<uint32_t*>(&p);
        }
        {
            uint8_t* a1 = smem + SM_A1 + (tid >> 3) * 256 + (tid & 7) * 16;
            *reinterpret_cast<uint4*>(a1) = make_uint4(packed[0], packed[1], packed[2], packed[3]);         // k 0..7
            *reinterpret_cast<uint4*>(a1 + 128) = make_uint4(packed[4], packed[5], packed[6], packed[7]);   // k 8..15
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (!img_ready) { mbar_wait(bar_img, 0); img_ready = true; }

        // ---- layer 1: D1 = A1 . W1^T
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_f16(tmem_base, desc_nosw_k16(sA1), desc_nosw_k16(sW1), kIdesc, 0u);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue 1: A2[row = tid][k] = bf16(relu(D1 + b1)), 128B-swizzled K-major slabs
#pragma unroll 1
        for (int c0 = 0; c0 < TC_H; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr_lane + (uint32_t)c0, r);
            uint8_t* rowp = smem + SM_A2 + (c0 >> 6) * 16384 + tid * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {                    // four 16-byte chunks of 8 columns
                uint32_t w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int c = c0 + q * 8 + u * 2;
                    float v0 = fmaxf(__uint_as_float(r[q * 8 + u * 2]) + sB1[c], 0.0f);
                    float v1 = fmaxf(__uint_as_float(r[q * 8 + u * 2 + 1]) + sB1[c + 1], 0.0f);
                    __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
                    w[u] = *reinterpret_cast<uint32_t*>(&p);
                }
                int chunk = ((c0 & 63) >> 3) + q;            // 16-byte chunk index inside the 128-byte row
                *reinterpret_cast<uint4*>(rowp + ((chunk ^ (tid & 7)) * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();

        // ---- layer 2: D2 = A2 . W2^T, K = 256 as 16 instructions of K = 16 (4 slabs x 4)
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                uint32_t a_addr = sA2 + (uint32_t)(kk >> 2) * 16384u + (uint32_t)(kk & 3) * 32u;
                uint32_t b_addr = sW2 + (uint32_t)(kk >> 2) * 32768u + (uint32_t)(kk & 3) * 32u;
                umma_f16(tmem_base + 256u, desc_sw128(a_addr), desc_sw128(b_addr), kIdesc, kk > 0 ? 1u : 0u);
            }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue 2: logits = relu(D2 + b2) . W3 + b3 on CUDA cores, then softmax / sampling
        float lg0 = sB3[0], lg1 = sB3[1], lg2 = sB3[2], lg3 = sB3[3];
#pragma unroll 1
        for (int c0 = 0; c0 < TC_H; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr_lane + 256u + (uint32_t)c0, r);
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                float h = fmaxf(__uint_as_float(r[u]) + sB2[c0 + u], 0.0f);
                float4 w = *reinterpret_cast<const float4*>(sW3 + (c0 + u) * 4);
                lg0 = fmaf(h, w.x, lg0); lg1 = fmaf(h, w.y, lg1); lg2 = fmaf(h, w.z, lg2); lg3 = fmaf(h, w.w, lg3);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // TMEM reads done before the next tile's MMAs

        if (valid) {
            uint32_t fl = 0xFu;
            const bool use_mask = args.mask_flags != nullptr;
            if (use_mask) fl = args.mask_flags[s];
            float l0 = (use_mask && !(fl & 1u)) ? -1e9f : lg0, l1 = (use_mask && !(fl & 2u)) ? -1e9f : lg1;
            float l2 = (use_mask && !(fl & 4u)) ? -1e9f : lg2, l3 = (use_mask && !(fl & 8u)) ? -1e9f : lg3;
            float mx = fmaxf(fmaxf(l0, l1), fmaxf(l2, l3));
            float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx), e3 = expf(l3 - mx);
            float sum = e0 + e1 + e2 + e3;
            float p0 = e0 / sum, p1 = e1 / sum, p2 = e2 / sum, p3 = e3 / sum;
            if (args.probs) *reinterpret_cast<float4*>(args.probs + s * 4) = make_float4(p0, p1, p2, p3);
            if (args.logits) *reinterpret_cast<float4*>(args.logits + s * 4) = make_float4(lg0, lg1, lg2, lg3);
            if (args.action) {
                uint32_t a;
                if (args.greedy) {
                    float q0 = (!use_mask || (fl & 1u)) ? p0 : 0.0f, q1 = (!use_mask || (fl & 2u)) ? p1 : 0.0f;
                    float q2 = (!use_mask || (fl & 4u)) ? p2 : 0.0f, q3 = (!use_mask || (fl & 8u)) ? p3 : 0.0f;
                    a = 0; float best = q0;
                    if (q1 > best) { best = q1; a = 1; }
                    if (q2 > best) { best = q2; a = 2; }
                    if (q3 > best) { best = q3; a = 3; }
                } else {
                    Rand4 rr = stream_keyed(args.keys, args.gid0 + (uint64_t)s, args.t, B2048_DOM_STEP);
                    float c0 = p0, c1 = c0 + p1, c2 = c1 + p2, c3 = c2 + p3;
                    float u = ((float)(rr.w3 >> 8) + 0.5f) * (1.0f / 16777216.0f) * c3;
                    a = (u >= c0 ? 1u : 0u) + (u >= c1 ? 1u : 0u) + (u >= c2 ? 1u : 0u);
                    float pa = a == 0 ? p0 : a == 1 ? p1 : a == 2 ? p2 : p3;
                    if (!(pa > 0.0f)) {
                        if (p3 > 0.0f) a = 3;
                        if (p2 > 0.0f) a = 2;
                        if (p1 > 0.0f) a = 1;
                        if (p0 > 0.0f) a = 0;
                    }
                }
                args.action[s] = (uint8_t)a;
            }
        }
        __syncthreads();   // A1 / A2 / TMEM are reused by the next tile
    }

    if (!img_ready) mbar_wait(bar_img, 0);   // never leave with bulk copies in flight
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Returns B2048_OK if the tensor-core path applies and was launched, B2048_ERR_UNSUPPORTED (without setting an
// error message) when the shape is outside what this kernel implements.
int launch_policy_tc(b2048_handle* h, const b2048_mlp_desc* mlp, const uint64_t* board, const uint8_t* mask_flags,
                     uint8_t* action, float* probs, float* logits, int64_t n, uint64_t seed, uint64_t gid0, uint32_t t,
                     int greedy, cudaStream_t stream) {
    if (mlp->n_layers != 3 || mlp->dims[0] != 16 || mlp->dims[1] != TC_H || mlp->dims[2] != TC_H || mlp->dims[3] != 4 ||
        mlp->activation != B2048_ACTV_RELU || (mlp->obs_mode != B2048_OBS_RAW && mlp->obs_mode != B2048_OBS_LOG2) ||
        h->smem_optin < SM_TOTAL)
        return B2048_ERR_UNSUPPORTED;
    if (!h->tc_image) {
        cudaError_t e = cudaMalloc(&h->tc_image, IMG_BYTES);
        if (e != cudaSuccess) return check_cuda(e, "cudaMalloc(tc_image)");
        e = cudaFuncSetAttribute(policy_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(policy_tc_kernel)");
    }
    // the image is rebuilt on every call (71 K parameters, ~2 us): the library never caches stale weights
    policy_tc_prepare_kernel<<<64, 256, 0, stream>>>(mlp->W[0], mlp->b[0], mlp->W[1], mlp->b[1], mlp->W[2], mlp->b[2],
                                                      h->tc_image);
    PolicyTcArgs a;
    a.img = h->tc_image; a.board = board; a.mask_flags = mask_flags; a.action = action; a.probs = probs; a.logits = logits;
    a.n = n; a.keys = make_keys(seed); a.gid0 = gid0; a.t = t; a.greedy = greedy; a.obs_mode = mlp->obs_mode;
    a.obs_scale = mlp->obs_log2_scale;
    int64_t tiles = (n + TC_M - 1) / TC_M;
    int grid = (int)(tiles < h->num_sms ? tiles : h->num_sms);
    policy_tc_kernel<<<grid, TC_M, SM_TOTAL, stream>>>(a);
    return check_cuda(cudaGetLastError(), "policy_tc_kernel launch");
}

}  // namespace b2
