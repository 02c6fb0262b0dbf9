// ngp_mlp_tf32.cu -- K3 wide layers on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// One hidden layer of the per-genome MLP (numpy_nn.NeuralNetwork.run, /root/reference/numpy_nn.py:126-129) for
// E environments of one genome is a small dense contraction  out[e][o] = sigmoid( sum_k W_g[o][k] * a[e][k] )
// with W_g = that genome's (n_out x (n_in+1)) slice of the genome vector (bias weight = last column).  For the
// [6,512,512,2] net and 64 environments per genome (BASELINE config 4) the 512x513 layer is 98 % of the FLOPs and
// its 1.05 MB of weights are used for one step only: the layer is bound by streaming weights from HBM, which FP32
// FFMA cannot keep up with (34 MFLOP per MB) but tensor cores can.
//
// Mapping (D^T form, so nothing is padded): CTA = 128 outputs x 64 environments of one genome.
//   A (UMMA "M" side) = weights  [128 outputs][K]  -- K-major exactly as the genome stores them
//   B (UMMA "N" side) = inputs   [ 64 envs   ][K]  -- K-major, bias column synthesised as 1.0, zero padding
//   D (TMEM)          = [128 lanes = outputs][64 columns = envs], FP32
// Precision: kind::tf32 keeps 10 mantissa bits, which cannot meet the reference tolerance (rtol 1e-5), so every
// operand is split x = hi + lo with hi = tf32(x) and the product is accumulated as hi*hi + lo*hi + hi*lo
// ("3xTF32", relative error ~2^-21).  Three MMAs per k-step still leave the layer HBM-bound (96 flop/B against a
// ridge of ~170 flop/B), so the split costs no time.
// Pipeline (details at the kernel): K is walked in chunks of 16 floats.  Eight producer warps copy raw FP32 rows with
// cp.async into private rings, split them into hi/lo tiles in the canonical no-swizzle K-major core-matrix layout (two
// shared-memory stages) and arrive on the stage's "filled" mbarrier; a ninth warp waits for the eight arrivals, issues the 6
// tcgen05.mma of the chunk and commits them to the stage's "free" mbarrier, which is what the stores of chunk i+2 wait on.
// There is no CTA-wide barrier in the loop.  The epilogue reads TMEM with tcgen05.ld (warp w: lanes 32 (w % 4).., columns
// 32 (w / 4)..), applies the sigmoid and writes out[e][o] coalesced along o.
#include "ngp_internal.h"

namespace tf32 {

constexpr int TM = 128;            // outputs per CTA  (UMMA M)
constexpr int TN = 64;             // environments per CTA (UMMA N)
constexpr int THREADS = 256;       // producer threads
// canonical K-major, no swizzle: 16-byte unit (row r, k-unit c) at  c*LBO + (r/8)*SBO + (r%8)*16
constexpr uint32_t SBO = 128;                              // next group of 8 rows
constexpr uint32_t LBO_A = TM * 16 + 16;                   // next 16-byte k unit (+16 B pad: conflict-free stores)
constexpr uint32_t LBO_B = TN * 16 + 16;
constexpr uint32_t TMEM_COLS = 64;
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, both K-major, N=64, M=128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo)
{
    // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout NONE
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------------
// Producer: cp.async ring.  Every producer warp copies its own rows of the next chunks with 4-byte cp.async (the rows are only
// 4-byte aligned: K = 513 floats) into a private 4-deep ring of raw FP32 tiles, three chunks ahead of its use; the same warp
// then converts its rows from the ring (16-byte shared loads) to the hi/lo UMMA tiles, so the ring needs no cross-warp
// synchronisation at all.  K chunks of 16 floats; two UMMA stages and the ring are 97 KB: two CTAs per SM overlap each
// other's prologue and epilogue.  (A register-staged producer -- loads two chunks ahead in registers, split, 4-byte stores
// -- measured the same 2.1-2.2 TB/s: profiles/README.md.)
namespace v2 {
constexpr int KC2 = 16;                                    // floats per chunk = 4 units of 16 bytes = 2 MMAs of K=8
constexpr int RING = 4;
constexpr uint32_t TILE_A2 = (KC2 / 4) * LBO_A, TILE_B2 = (KC2 / 4) * LBO_B;
constexpr uint32_t STAGE2 = 2 * TILE_A2 + 2 * TILE_B2;      // A_hi, A_lo, B_hi, B_lo
constexpr uint32_t RAW_A = (TM / 8) * KC2 * 4, RAW_B = (TN / 8) * KC2 * 4;    // per warp and ring slot: 16 / 8 rows of 64 bytes
constexpr uint32_t RAW_SLOT = RAW_A + RAW_B;
constexpr uint32_t RAW_BYTES = (THREADS / 32) * RING * RAW_SLOT;
constexpr uint32_t SMEM2 = 2 * STAGE2 + RAW_BYTES + 64;     // stages + rings + five mbarriers

__device__ __forceinline__ void cp_async4(uint32_t dst, const float *src, bool valid)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4u : 0u) : "memory");
}

constexpr int THREADS2 = THREADS + 32;                     // eight producer warps + one MMA warp

__global__ void __launch_bounds__(THREADS2, 2)
mlp_layer_tf32_v2_kernel(const float *__restrict__ genomes, size_t w_off, int G, const float *__restrict__ in, int envs, int ni, int no, int bias,
                         float *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.z, e0 = blockIdx.y * TN, o0 = blockIdx.x * TM;
    const int K = ni + bias;
    const int n_chunks = (K + KC2 - 1) / KC2;
    const float *W = genomes + (size_t)g * G + w_off;
    const float *A = in + (size_t)g * envs * ni;      // in[g][e][ni] (+ bias 1) x W_g[no][ni+bias] -> out[g][e][no] = sigmoid(.)
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar0 = smem_base + 2 * STAGE2 + RAW_BYTES;     // two "stage free", one "accumulator ready", two "stage filled"
    const uint32_t bar_done = bar0 + 16, bar_full = bar0 + 24;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(bar_done, 1);
        mbar_init(bar_full, THREADS / 32); mbar_init(bar_full + 8, THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == THREADS / 32) {
        // ---- MMA warp: one thread waits for a filled stage, issues its MMAs and hands the stage back ----
        if (lane == 0) {
            for (int c = 0; c < n_chunks; ++c) {
                const int st = c & 1;
                mbar_wait(bar_full + 8 * st, (c >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_base + st * STAGE2, a_lo = a_hi + TILE_A2, b_hi = a_hi + 2 * TILE_A2, b_lo = b_hi + TILE_B2;
#pragma unroll
                for (int j = 0; j < KC2 / 8; ++j) {
                    const uint32_t ka = 2 * j * LBO_A, kb = 2 * j * LBO_B;
                    umma_tf32(tmem_d, make_desc(a_lo + ka, LBO_A), make_desc(b_hi + kb, LBO_B), (c | j) ? 1u : 0u);   // small terms first
                    umma_tf32(tmem_d, make_desc(a_hi + ka, LBO_A), make_desc(b_lo + kb, LBO_B), 1u);
                    umma_tf32(tmem_d, make_desc(a_hi + ka, LBO_A), make_desc(b_hi + kb, LBO_B), 1u);
                }
                umma_commit(bar0 + 8 * st);
                if (c == n_chunks - 1) umma_commit(bar_done);
            }
        }
        __syncwarp();                                           // the other lanes park here instead of polling a barrier next to the issuing lane
    } else {
        // ---- producer warps.  Copy side: lane = (row parity, k): one cp.async instruction covers two rows x 16 floats ----
        const uint32_t raw_base = smem_base + 2 * STAGE2 + (uint32_t)warp * RING * RAW_SLOT;
        const int ck = lane & 15, cr = lane >> 4;
        uint32_t woff[TM / 16], xoff[TN / 16], wok = 0, xok = 0;     // byte offsets of this lane's rows, rows that exist
#pragma unroll
        for (int j = 0; j < TM / 16; ++j) {
            const int row = o0 + warp + 8 * (2 * j + cr);
            wok |= (row < no ? 1u : 0u) << j;
            woff[j] = (uint32_t)(min(row, no - 1) * K + ck) * 4u;
        }
#pragma unroll
        for (int j = 0; j < TN / 16; ++j) {
            const int row = e0 + warp + 8 * (2 * j + cr);
            xok |= (row < envs ? 1u : 0u) << j;
            xoff[j] = (uint32_t)(min(row, envs - 1) * ni + ck) * 4u;
        }
        const bool full_rows = wok == (1u << (TM / 16)) - 1u && xok == (1u << (TN / 16)) - 1u;
        auto issue_copy = [&](int c) {
            if (c < n_chunks) {
                const uint32_t slot = raw_base + (uint32_t)(c % RING) * RAW_SLOT + (uint32_t)(cr * KC2 + ck) * 4u;
                const char *wc = reinterpret_cast<const char *>(W) + (size_t)c * (KC2 * 4);
                const char *xc = reinterpret_cast<const char *>(A) + (size_t)c * (KC2 * 4);
                if (full_rows && (c + 1) * KC2 <= ni) {         // interior chunk of a full tile: no predicates
#pragma unroll
                    for (int j = 0; j < TM / 16; ++j) cp_async4(slot + (uint32_t)j * (2 * KC2 * 4), reinterpret_cast<const float *>(wc + woff[j]), true);
#pragma unroll
                    for (int j = 0; j < TN / 16; ++j) cp_async4(slot + RAW_A + (uint32_t)j * (2 * KC2 * 4), reinterpret_cast<const float *>(xc + xoff[j]), true);
                } else {
                    const int k = c * KC2 + ck;
#pragma unroll
                    for (int j = 0; j < TM / 16; ++j) {
                        const bool ok = ((wok >> j) & 1u) && k < K;
                        cp_async4(slot + (uint32_t)j * (2 * KC2 * 4), ok ? reinterpret_cast<const float *>(wc + woff[j]) : W, ok);
                    }
#pragma unroll
                    for (int j = 0; j < TN / 16; ++j) {
                        const bool ok = ((xok >> j) & 1u) && k < ni;
                        cp_async4(slot + RAW_A + (uint32_t)j * (2 * KC2 * 4), ok ? reinterpret_cast<const float *>(xc + xoff[j]) : A, ok);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // Convert side: lane owns 16-byte units; unit u of a tile = (row i = u / 4, k-unit c = u % 4)
        auto split_store = [&](uint8_t *hi_dst, uint32_t tile_bytes, float4 v) {
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
            *reinterpret_cast<float4 *>(hi_dst) = h;
            *reinterpret_cast<float4 *>(hi_dst + tile_bytes) = l;
        };
        const int bias_chunk = bias ? ni / KC2 : -1;            // the chunk that holds the synthesised bias input column
#pragma unroll
        for (int c = 0; c < RING - 1; ++c) issue_copy(c);
        for (int c = 0; c < n_chunks; ++c) {
            issue_copy(c + RING - 1);
            asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");
            __syncwarp();                                       // the warp's copies of chunk c have landed and are visible to all its lanes
            const int st = c & 1;
            if (c >= 2) mbar_wait(bar0 + 8 * st, ((c >> 1) - 1) & 1);  // the MMAs that read this stage two chunks ago are done
            uint8_t *stage = smem + st * STAGE2;
            const uint8_t *raw = smem + 2 * STAGE2 + (size_t)warp * RING * RAW_SLOT + (size_t)(c % RING) * RAW_SLOT;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int u = lane + 32 * j, i = u >> 2, cu = u & 3;
                const float4 v = *reinterpret_cast<const float4 *>(raw + i * (KC2 * 4) + cu * 16);
                split_store(stage + cu * LBO_A + i * SBO + warp * 16, TILE_A2, v);
            }
            {
                const int i = lane >> 2, cu = lane & 3;
                float4 v = *reinterpret_cast<const float4 *>(raw + RAW_A + i * (KC2 * 4) + cu * 16);
                if (c == bias_chunk) {                          // bias input column: 1.0 at k == ni for rows that exist
                    const int kb = c * KC2 + cu * 4;
                    if (kb <= ni && ni < kb + 4 && e0 + warp + 8 * i < envs) {
                        if (ni == kb) v.x = 1.0f; else if (ni == kb + 1) v.y = 1.0f; else if (ni == kb + 2) v.z = 1.0f; else v.w = 1.0f;
                    }
                }
                split_store(stage + 2 * TILE_A2 + cu * LBO_B + i * SBO + warp * 16, TILE_B2, v);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();                                       // also: every lane is done reading ring slot c before it is refilled
            if (lane == 0) mbar_arrive(bar_full + 8 * st);
        }
    }
    // ---- epilogue: TMEM -> registers -> sigmoid -> out[e][o] ----
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 8) {
        // warp w reads TMEM lanes 32 (w % 4) .. +31 (its outputs) and columns 32 (w / 4) .. +31 (its environments)
        const int o = o0 + (warp & 3) * 32 + lane, half = warp >> 2;
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (o < no) {
            // sigmoid with the fast exponential (ex2.approx, ~2 ulp) and an IEEE reciprocal: far inside the 1e-5 bar
            float *dst = out + ((size_t)g * envs + e0 + half * 32) * no + o;
            const int e_left = envs - (e0 + half * 32);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < e_left) *dst = __frcp_rn(1.0f + __expf(-__uint_as_float(v[j])));
                dst += no;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
}
}  // namespace v2

}  // namespace tf32

// Launch one wide hidden layer on the tensor-core path; returns NGP_ERR_UNSUPPORTED when the shape does not
// make it a real dense contraction (callers then use the FP32 FFMA kernel).
int ngp_mlp_layer_tf32(ngp_handle *h, const float *genomes, size_t w_off, const float *in, int n_genomes, int envs, int ni, int no, int bias,
                       float *out, cudaStream_t st)
{
    if (ni < 64 || no < 64 || envs < 16) return NGP_ERR_UNSUPPORTED;
    if (!h->tf32_attr_set) {
        NGP_CUDA(cudaFuncSetAttribute(tf32::v2::mlp_layer_tf32_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tf32::v2::SMEM2));
        h->tf32_attr_set = 1;
    }
    dim3 grid((no + tf32::TM - 1) / tf32::TM, (envs + tf32::TN - 1) / tf32::TN, n_genomes);
    tf32::v2::mlp_layer_tf32_v2_kernel<<<grid, tf32::v2::THREADS2, tf32::v2::SMEM2, st>>>(genomes, w_off, h->gene_size, in, envs, ni, no, bias, out);
    h->launches++;
    NGP_CUDA(cudaGetLastError());
    return NGP_OK;
}
