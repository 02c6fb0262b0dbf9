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
// Pipeline: K is walked in chunks of 32 floats through two shared-memory stages.  All 256 threads load a chunk
// (coalesced 128-byte rows), split it and store hi/lo tiles in the canonical no-swizzle K-major core-matrix
// layout, and every warp arrives on the stage's "filled" mbarrier; one thread waits for the eight arrivals, issues the 12
// tcgen05.mma of the chunk and commits them to the stage's "free" mbarrier, which is what the stores of chunk i+2 wait on.
// There is no CTA-wide barrier in the loop.  The epilogue reads TMEM with tcgen05.ld (warp w owns lanes 32w..32w+31),
// applies the sigmoid and writes out[e][o] coalesced along o.
#include "ngp_internal.h"

namespace tf32 {

constexpr int TM = 128;            // outputs per CTA  (UMMA M)
constexpr int TN = 64;             // environments per CTA (UMMA N)
constexpr int KC = 32;             // floats per K chunk = 8 chunks of 16 bytes = 4 MMAs of K=8
constexpr int THREADS = 256;
// canonical K-major, no swizzle: 16-byte unit (row r, k-unit c) at  c*LBO + (r/8)*SBO + (r%8)*16
constexpr uint32_t SBO = 128;                              // next group of 8 rows
constexpr uint32_t LBO_A = TM * 16 + 16;                   // next 16-byte k unit (+16 B pad: conflict-free stores)
constexpr uint32_t LBO_B = TN * 16 + 16;
constexpr uint32_t TILE_A = (KC / 4) * LBO_A;              // bytes of one A tile (hi or lo)
constexpr uint32_t TILE_B = (KC / 4) * LBO_B;
constexpr uint32_t STAGE = 2 * TILE_A + 2 * TILE_B;        // A_hi, A_lo, B_hi, B_lo
constexpr uint32_t SMEM_BYTES = 2 * STAGE + 64;          // stages + five mbarriers
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

// in[g][e][ni] (+ bias 1) x W_g[no][ni+bias] -> out[g][e][no] = sigmoid(.)
__global__ void __launch_bounds__(THREADS, 2)
mlp_layer_tf32_kernel(const float *__restrict__ genomes, size_t w_off, int G, const float *__restrict__ in, int envs, int ni, int no, int bias,
                      float *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.z, e0 = blockIdx.y * TN, o0 = blockIdx.x * TM;
    const int K = ni + bias;
    const int n_chunks = (K + KC - 1) / KC;
    const float *W = genomes + (size_t)g * G + w_off;
    const float *A = in + (size_t)g * envs * ni;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar0 = smem_base + 2 * STAGE;            // two "stage free" barriers, one "accumulator ready", two "stage filled"
    const uint32_t bar_done = bar0 + 16;
    const uint32_t bar_full = bar0 + 24;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(bar_done, 1);
        mbar_init(bar_full, THREADS / 32); mbar_init(bar_full + 8, THREADS / 32);      // one arrival per producer warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    // Register-staged producer.  lane = column k of the chunk (coalesced 128-byte row segments), warp w owns weight rows
    // w, w+8, .. and input rows w, w+8, ..; all pointers, predicates and shared-memory offsets are hoisted out of the
    // chunk loop, the 24 loads of a chunk are all in flight at once and one chunk ahead of the shared-memory stores.
    constexpr int WR = TM / 8, XR = TN / 8;                 // rows per thread: 16 weight rows, 8 input rows
    const uint32_t unit = (uint32_t)(lane >> 2), sub = (uint32_t)(lane & 3) * 4;
    const uint32_t a_off = unit * LBO_A + (uint32_t)warp * 16 + sub;               // row r = warp + 8 i  ->  + i * SBO
    const uint32_t b_off = 2 * TILE_A + unit * LBO_B + (uint32_t)warp * 16 + sub;
    const float *wp = W + (size_t)min(o0 + warp, no - 1) * K + lane;
    const float *xp = A + (size_t)min(e0 + warp, envs - 1) * ni + lane;
    uint32_t wmask = 0, xmask = 0;                           // rows of this thread that exist
#pragma unroll
    for (int i = 0; i < WR; ++i) wmask |= (o0 + warp + 8 * i < no ? 1u : 0u) << i;
#pragma unroll
    for (int i = 0; i < XR; ++i) xmask |= (e0 + warp + 8 * i < envs ? 1u : 0u) << i;
    // two chunks of loads (2 x 24 registers per thread = 48 KB per CTA) are kept in flight ahead of the stores: with ~1 us of
    // HBM latency under load it takes ~45 KB in flight per SM to stream at full bandwidth
    float wv0[WR], xv0[XR], wv1[WR], xv1[XR];
    // interior chunks of full tiles need no predicates at all (the common case: 512 outputs, 64 environments)
    const bool full_tile = (o0 + TM <= no) && (e0 + TN <= envs);
    const size_t wstride = (size_t)8 * K, xstride = (size_t)8 * ni;
    auto load_chunk = [&](int c, float (&wv)[WR], float (&xv)[XR]) {
        const int k = c * KC + lane;
        const float *w = wp + c * KC, *x = xp + c * KC;
        if (full_tile && (c + 1) * KC <= ni) {
#pragma unroll
            for (int i = 0; i < WR; ++i) { wv[i] = __ldg(w); w += wstride; }
#pragma unroll
            for (int i = 0; i < XR; ++i) { xv[i] = __ldg(x); x += xstride; }
            return;
        }
        const bool kw = k < K, kx = k < ni;
        const float fill = (k == ni && bias) ? 1.0f : 0.0f;  // bias input column / zero padding
#pragma unroll
        for (int i = 0; i < WR; ++i) wv[i] = (kw && ((wmask >> i) & 1)) ? __ldg(w + (size_t)i * wstride) : 0.f;
#pragma unroll
        for (int i = 0; i < XR; ++i) xv[i] = ((xmask >> i) & 1) ? (kx ? __ldg(x + (size_t)i * xstride) : fill) : 0.f;
    };
    auto consume_chunk = [&](int c, float (&wv)[WR], float (&xv)[XR]) {
        const int st = c & 1;
        uint8_t *stage = smem + st * STAGE;
        // the MMAs that read this stage two chunks ago must have completed
        if (c >= 2) mbar_wait(bar0 + 8 * st, ((c >> 1) - 1) & 1);
#pragma unroll
        for (int i = 0; i < WR; ++i) {
            const float v = wv[i], hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            *reinterpret_cast<float *>(stage + a_off + i * SBO) = hi;
            *reinterpret_cast<float *>(stage + a_off + i * SBO + TILE_A) = v - hi;
        }
#pragma unroll
        for (int i = 0; i < XR; ++i) {
            const float v = xv[i], hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            *reinterpret_cast<float *>(stage + b_off + i * SBO) = hi;
            *reinterpret_cast<float *>(stage + b_off + i * SBO + TILE_B) = v - hi;
        }
        if (c + 2 < n_chunks) load_chunk(c + 2, wv, xv);    // refill these registers two chunks ahead
        // generic-proxy stores -> visible to the tensor core's async proxy; then one arrival per warp on the stage's "filled"
        // barrier.  No CTA-wide barrier: only the MMA-issuing thread waits for all eight warps, the others run ahead into the
        // next chunk (bounded by the "stage free" barriers).
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * st);
        if (tid == 0) {
            mbar_wait(bar_full + 8 * st, (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_base + st * STAGE, a_lo = a_hi + TILE_A, b_hi = a_hi + 2 * TILE_A, b_lo = b_hi + TILE_B;
#pragma unroll
            for (int j = 0; j < KC / 8; ++j) {             // one MMA consumes two 16-byte k units
                const uint32_t ka = 2 * j * LBO_A, kb = 2 * j * LBO_B;
                umma_tf32(tmem_d, make_desc(a_lo + ka, LBO_A), make_desc(b_hi + kb, LBO_B), (c | j) ? 1u : 0u);   // small terms first
                umma_tf32(tmem_d, make_desc(a_hi + ka, LBO_A), make_desc(b_lo + kb, LBO_B), 1u);
                umma_tf32(tmem_d, make_desc(a_hi + ka, LBO_A), make_desc(b_hi + kb, LBO_B), 1u);
            }
            umma_commit(bar0 + 8 * st);                     // stage reusable when these MMAs are done
            if (c == n_chunks - 1) umma_commit(bar_done);   // accumulator complete
        }
    };
    load_chunk(0, wv0, xv0);
    if (n_chunks > 1) load_chunk(1, wv1, xv1);
    for (int c = 0; c < n_chunks; c += 2) {
        consume_chunk(c, wv0, xv0);
        if (c + 1 < n_chunks) consume_chunk(c + 1, wv1, xv1);
    }
    // ---- epilogue: TMEM -> registers -> sigmoid -> out[e][o] ----
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        const int o = o0 + warp * 32 + lane;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 32);
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
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
}

}  // namespace tf32

// Launch one wide hidden layer on the tensor-core path; returns NGP_ERR_UNSUPPORTED when the shape does not
// make it a real dense contraction (callers then use the FP32 FFMA kernel).
int ngp_mlp_layer_tf32(ngp_handle *h, const float *genomes, size_t w_off, const float *in, int n_genomes, int envs, int ni, int no, int bias,
                       float *out, cudaStream_t st)
{
    if (ni < 64 || no < 64 || envs < 16) return NGP_ERR_UNSUPPORTED;
    static bool attr_set[64];
    if (!attr_set[h->device]) {
        NGP_CUDA(cudaFuncSetAttribute(tf32::mlp_layer_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tf32::SMEM_BYTES));
        attr_set[h->device] = true;
    }
    dim3 grid((no + tf32::TM - 1) / tf32::TM, (envs + tf32::TN - 1) / tf32::TN, n_genomes);
    tf32::mlp_layer_tf32_kernel<<<grid, tf32::THREADS, tf32::SMEM_BYTES, st>>>(genomes, w_off, h->gene_size, in, envs, ni, no, bias, out);
    h->launches++;
    NGP_CUDA(cudaGetLastError());
    return NGP_OK;
}
