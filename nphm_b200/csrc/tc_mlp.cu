// tcgen05 / TMEM kernel for the forward-deformation backbone (DeepSDF MLP, hidden 512, 6 hidden layers, condition 232,
// 3 outputs) - the `compress` DeformationNetwork of scripts/configs/nphm_def.yaml.
//
// Reference semantics: DeepSDF.forward  src/NPHM/models/deepSDF.py:64-89   (called by DeformationNetwork.forward :237)
//
// Layer stack with the per-query-constant condition folded away (see simt.cu):
//   h0 = sp(W0x x + v0)                    K = 3                       -> CUDA cores
//   h1 = sp(W1 h0 + b1)                    512 x 512                   -> UMMA
//   h2 = sp(W2 h1 + b2)                    277 x 512  (N -> 288)
//   h3 = sp(W3a h2/r2 + W3x x/r2 + v3)     512 x 280  (K -> 288)
//   h4 = sp(W4 h3 + b4), h5 = sp(W5 h4 + b5)   512 x 512
//   y  = W6 h5 + b6                        3 x 512                     -> CUDA cores, fused into the h5 epilogue
// A 128-row tile does not fit: the K = 512 activations of one layer are 256 KB as fp16 hi/lo and the accumulator
// needs all 512 TMEM columns.  The kernel therefore works on M = 64 tiles: A (hi | lo, 128 KB) lives in shared
// memory in UMMA K-major no-swizzle order, chunk-major (for each 8-column K chunk: 64 rows x 16 B, so the epilogue's
// 16-byte stores are conflict free; LBO = 1024 B, SBO = 128 B); the accumulator of a layer is two N-halves
// D_a = TMEM[0,256) and D_b = TMEM[256,512), each filled by one long run of SS-form MMAs (an accumulator switch
// costs ~200 cycles, an M = 64 MMA costs max(70, N/2) cycles: profiles/r01_mma_microbench.txt).  Weights stream
// through a 5-slot ring of (N_half x 16) fp16 hi|lo slabs.  Same 3-pass fp16 split and log2-unit softplus as the
// ensemble kernel.  M = 64 uses TMEM lanes (r % 16) + 32 * (r / 16): the epilogue therefore reads the accumulator with the
// 16-lane shape tcgen05.ld.16x256b (thread = rows l/4 and l/4+8, column pairs 2*(l%4)), which keeps all 32 lanes of a warp
// busy and makes the fp16-pair stores of a warp one contiguous 128-byte line of the chunk-major A operand.
#include "engine.cuh"
#include "tc_common.cuh"

namespace nphm {
namespace tcm {
using namespace tc;

constexpr int kH = 512, kN2 = 277, kCond = 232, kOut = 3;
constexpr int kTL = 5;                                         // tensor layers (network layers 1..5)
__host__ __device__ constexpr int layer_ks(int t) { return t == 2 ? 18 : 32; }          // k-steps of 16
__host__ __device__ constexpr int layer_nh(int t) { return t == 1 ? 144 : 256; }        // rows per N half
__host__ __device__ constexpr int layer_np(int t) { return 2 * layer_nh(t); }           // padded N
__host__ __device__ constexpr size_t layer_bytes(int t) { return (size_t)2 * layer_ks(t) * layer_nh(t) * 64; }
constexpr size_t kWeightBytes = layer_bytes(0) + layer_bytes(1) + layer_bytes(2) + layer_bytes(3) + layer_bytes(4);
constexpr int kSlotBytes = 256 * 64, kSlots = 5;
constexpr int kABytes = 64 * kH * 2;                           // one plane (hi or lo)
// per-query record (floats): layer-0 rows, biases of the 5 tensor layers (padded), output layer
constexpr int kRecL0 = 0;                                      // 512 x float4
constexpr int kRecB = 2048;                                    // 512 | 288 | 512 | 512 | 512
__host__ __device__ constexpr int rec_bias_off(int t) { return kRecB + (t == 0 ? 0 : (t == 1 ? 512 : (t == 2 ? 800 : (t == 3 ? 1312 : 1824)))); }
constexpr int kRecW6 = kRecB + 2336;                           // 3 x 512
constexpr int kRecB6 = kRecW6 + 3 * kH;                        // 3 (+1)
constexpr int kRecFloats = kRecB6 + 4;
constexpr int kEpiWarps = 16, kParts = 4;
constexpr int kThreads = 32 * (kEpiWarps + 2);

struct __align__(128) Smem {
    uint8_t a_hi[kABytes], a_lo[kABytes];
    uint8_t slabs[kSlots][kSlotBytes];
    float partial[kParts - 1][64][4];
    uint64_t slab_full[kSlots], slab_empty[kSlots];
    uint64_t a_ready, d_ready;
    uint32_t tmem_base;
};

struct Params {
    const uint8_t *weights;
    const float *recs;          // [n_queries][kRecFloats]
    const float *xyz;           // [n_queries][n_points][3]
    long long n_points;
    int n_queries;
    float *out;                 // [n_queries][n_points][3]
    const int *live;            // optional device counter: the launch does nothing when *live == 0 (Broyden: nobody is active)
};

// TMEM column of output n of tensor layer t
__device__ __forceinline__ int d_col(int t, int n)
{
    const int nh = layer_nh(t);
    return n < nh ? n : 256 + (n - nh);
}
// 16 lanes x 16 accumulator columns: r[4g + 0/1] = (row l/4, cols 8g + 2(l%4) + 0/1), r[4g + 2/3] = same columns, row l/4 + 8
__device__ __forceinline__ void tc_ld16x16(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// one fp16 pair (hi and lo planes) of row `row`, K chunk `chunk`, pair index j (K columns 8*chunk + 2j, +1)
__device__ __forceinline__ void store_a_pair(Smem &sm, int chunk, int row, int j, float v0, float v1)
{
    uint32_t hi, lo;
    split2(v0, v1, hi, lo);
    *reinterpret_cast<uint32_t *>(sm.a_hi + (size_t)chunk * 1024 + row * 16 + j * 4) = hi;
    *reinterpret_cast<uint32_t *>(sm.a_lo + (size_t)chunk * 1024 + row * 16 + j * 4) = lo;
}

__global__ void __launch_bounds__(kThreads, 1) mlp_tc_kernel(const Params p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tiles_per_query = (p.n_points + 63) / 64;
    const long long n_tiles = tiles_per_query * p.n_queries;
    if (p.live && *p.live == 0) return;                    // uniform over the grid: every sample of the search is frozen

    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(&sm.slab_full[i], 1); mbar_init(&sm.slab_empty[i], 1); }
        mbar_init(&sm.a_ready, kEpiWarps);
        mbar_init(&sm.d_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kEpiWarps + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&sm.tmem_base)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == kEpiWarps) {
        // =========================================================================== producer: weight slabs
        if (lane == 0) {
            int slot = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const uint8_t *w = p.weights;
#pragma unroll 1
                for (int t = 0; t < kTL; ++t) {
                    const uint32_t bytes = layer_nh(t) * 64;
#pragma unroll 1
                    for (int j = 0; j < 2 * layer_ks(t); ++j) {
                        mbar_wait(&sm.slab_empty[slot], ph ^ 1);
                        mbar_expect_tx(&sm.slab_full[slot], bytes);
                        bulk_g2s(sm.slabs[slot], w, bytes, &sm.slab_full[slot]);
                        w += bytes;
                        if (++slot == kSlots) { slot = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // =========================================================================== MMA issuer (SS form, M = 64)
        if (lane == 0) {
            int slot = 0;
            uint32_t ph = 0, a_ph = 0;
            const uint32_t a_hi = smem_u32(sm.a_hi), a_lo = smem_u32(sm.a_lo);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
                for (int t = 0; t < kTL; ++t) {
                    const int nh = layer_nh(t), ks = layer_ks(t);
                    const uint32_t idesc = make_idesc_m(64, nh);
                    mbar_wait(&sm.a_ready, a_ph);
                    a_ph ^= 1;
                    tc_fence_after();
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t d = tmem + half * 256;
#pragma unroll 1
                        for (int j = 0; j < ks; ++j) {
                            mbar_wait(&sm.slab_full[slot], ph);
                            tc_fence_after();
                            const uint32_t b = smem_u32(sm.slabs[slot]);
                            const uint64_t b_hi = make_desc(b, 128, 256), b_lo = make_desc(b + nh * 32, 128, 256);
                            const uint64_t ah = make_desc(a_hi + j * 2048, 1024, 128), al = make_desc(a_lo + j * 2048, 1024, 128);
                            tc_mma_ss(d, ah, b_hi, idesc, j == 0 ? 0 : 1);      // first k-step overwrites: the bias is added in the epilogue
                            tc_mma_ss(d, ah, b_lo, idesc, 1);
                            tc_mma_ss(d, al, b_hi, idesc, 1);
                            tc_commit(&sm.slab_empty[slot]);
                            if (++slot == kSlots) { slot = 0; ph ^= 1; }
                        }
                    }
                    tc_commit(&sm.d_ready);
                }
            }
        }
    } else {
        // =========================================================================== compute / epilogue warps
        // warp = (TMEM lane quarter q, column group part).  Rows 16q..16q+15 of the tile sit in lanes 0-15 of the quarter; a
        // thread owns rows rA = 16q + l/4 and rB = rA + 8 and the column pairs 2*(l%4) of every 8-column group it loads.
        // 16-column block b of a layer belongs to part b & 3.
        const int q = warp & 3, part = warp >> 2;
        const int j = lane & 3;
        const int rowA = q * 16 + (lane >> 2), rowB = rowA + 8;
        const int row0 = q * 16 + (lane & 15), half0 = lane >> 4;      // layer 0: (row, 4-element half of a chunk)
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t d_ph = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int qi = (int)(tile / tiles_per_query);
            const long long base = (tile - (long long)qi * tiles_per_query) * 64;
            auto coords = [&](int row, float &x, float &y, float &z) {
                const long long idx = base + row;
                const float *pp = p.xyz + ((size_t)qi * p.n_points + (idx < p.n_points ? idx : 0)) * 3;
                x = kS * pp[0]; y = kS * pp[1]; z = kS * pp[2];         // coordinates in log2 units
            };
            const float *rec = p.recs + (size_t)qi * kRecFloats;

            // ---------------- layer 0 on CUDA cores -> A of tensor layer 0 (all 32 lanes: 16 rows x 2 chunk halves)
            {
                float cx, cy, cz;
                coords(row0, cx, cy, cz);
                const float4 *l0 = reinterpret_cast<const float4 *>(rec + kRecL0);
#pragma unroll 1
                for (int c = part; c < 64; c += kParts) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float4 w = __ldg(l0 + c * 8 + half0 * 4 + e);
                        const float t = fmaf(w.x, cx, fmaf(w.y, cy, fmaf(w.z, cz, w.w)));
                        v[e] = (e & 1) ? sp_t_poly(t) : sp_t(t);
                    }
                    store_a_pair(sm, c, row0, half0 * 2, v[0], v[1]);
                    store_a_pair(sm, c, row0, half0 * 2 + 1, v[2], v[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.a_ready);

            float axA, ayA, azA, axB, ayB, azB;
            coords(rowA, axA, ayA, azA);
            coords(rowB, axB, ayB, azB);
            float accA[3] = {0.f, 0.f, 0.f}, accB[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
            for (int t = 0; t < kTL; ++t) {
                const int n_blocks = layer_np(t) / 16;
                const float *bias = rec + rec_bias_off(t);
                mbar_wait(&sm.d_ready, d_ph);
                d_ph ^= 1;
                tc_fence_after();
#pragma unroll 1
                for (int b = part; b < n_blocks; b += kParts) {
                    const int n0 = b * 16;
                    uint32_t r[8];
                    tc_ld16x16(tl + d_col(t, n0), r);
                    tc_wait_ld();
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const int n = n0 + 8 * g + 2 * j;                         // this thread's column pair of group g
                        const float2 bb = __ldg(reinterpret_cast<const float2 *>(bias + n));
                        float vA0 = sp_t(__uint_as_float(r[4 * g + 0]) + bb.x), vA1 = sp_t_poly(__uint_as_float(r[4 * g + 1]) + bb.y);
                        float vB0 = sp_t(__uint_as_float(r[4 * g + 2]) + bb.x), vB1 = sp_t(__uint_as_float(r[4 * g + 3]) + bb.y);
                        if (t == 1) {
                            // 277 outputs; K of the skip layer = [h2 (277), x (3), zero padding to 288]
                            if (n == 276) { vA1 = axA; vB1 = axB; }
                            else if (n == 278) { vA0 = ayA; vA1 = azA; vB0 = ayB; vB1 = azB; }
                            else if (n >= 280) { vA0 = vA1 = vB0 = vB1 = 0.f; }
                        }
                        if (t < kTL - 1) {
                            store_a_pair(sm, n >> 3, rowA, j, vA0, vA1);
                            store_a_pair(sm, n >> 3, rowB, j, vB0, vB1);
                        } else {
                            const float *w6 = rec + kRecW6 + n;
#pragma unroll
                            for (int o = 0; o < 3; ++o) {
                                const float2 w = __ldg(reinterpret_cast<const float2 *>(w6 + o * kH));
                                accA[o] = fmaf(vA0, w.x, fmaf(vA1, w.y, accA[o]));
                                accB[o] = fmaf(vB0, w.x, fmaf(vB1, w.y, accB[o]));
                            }
                        }
                    }
                }
                if (t < kTL - 1) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.a_ready);
                }
            }
            // ---------------- output layer: reduce the partial dot products over the column pairs of a row (4 lanes) and
            // over the column groups (4 warps of the quarter)
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                accA[o] += __shfl_xor_sync(0xffffffffu, accA[o], 1); accA[o] += __shfl_xor_sync(0xffffffffu, accA[o], 2);
                accB[o] += __shfl_xor_sync(0xffffffffu, accB[o], 1); accB[o] += __shfl_xor_sync(0xffffffffu, accB[o], 2);
            }
            if (part != 0 && j == 0) {
#pragma unroll
                for (int o = 0; o < 3; ++o) { sm.partial[part - 1][rowA][o] = accA[o]; sm.partial[part - 1][rowB][o] = accB[o]; }
            }
            tc_fence_before();
            asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * kParts) : "memory");
            tc_fence_after();
            if (part == 0 && j == 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = h ? rowB : rowA;
                    const long long idx = base + row;
                    if (idx < p.n_points) {
                        float *o = p.out + ((size_t)qi * p.n_points + idx) * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float sum = (h ? accB[c] : accA[c]) + __ldg(rec + kRecB6 + c);
#pragma unroll
                            for (int i = 0; i < kParts - 1; ++i) sum += sm.partial[i][row][c];
                            o[c] = sum;
                        }
                    }
                }
            }
            // the partial buffer is reused by the next tile: a second bar.sync keeps its readers ahead of the next writers
            asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * kParts) : "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kEpiWarps + 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
}

// weight slabs in consumption order: layer t, half, k-step; each slab N_half x 16 fp16 hi then lo in core-matrix order
__global__ void pack_mlp_slabs_kernel(const float *__restrict__ W1, const float *__restrict__ W2, const float *__restrict__ W3,
                                      const float *__restrict__ W4, const float *__restrict__ W5, uint8_t *__restrict__ out)
{
    const float inv_sqrt2 = 0.70710678118654752440f;
    const size_t total = kWeightBytes / 4;                   // (hi, lo) element pairs
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        size_t r = e;
        int t = 0;
        size_t off = 0;
        for (; t < kTL; ++t) {
            const size_t le = layer_bytes(t) / 4;
            if (r < le) break;
            r -= le; off += layer_bytes(t);
        }
        const int nh = layer_nh(t), ks = layer_ks(t);
        const int slab = (int)(r / ((size_t)nh * 16));       // = half * ks + j
        r -= (size_t)slab * nh * 16;
        const int half = slab / ks, j = slab % ks;
        const int nl = (int)(r / 16), kk = (int)(r % 16);
        const int n = half * nh + nl, k = j * 16 + kk;
        float v = 0.f;
        if (t == 0) { if (n < kH && k < kH) v = W1[(size_t)n * kH + k]; }
        else if (t == 1) { if (n < kN2 && k < kH) v = W2[(size_t)n * kH + k]; }
        else if (t == 2) { if (n < kH && k < kN2 + 3) v = W3[(size_t)n * kH + k] * inv_sqrt2; }   // [h2 | xyz | cond]: cond folded
        else if (t == 3) { if (n < kH && k < kH) v = W4[(size_t)n * kH + k]; }
        else { if (n < kH && k < kH) v = W5[(size_t)n * kH + k]; }
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t o = (size_t)(nl >> 3) * 256 + (size_t)(kk >> 3) * 128 + (size_t)(nl & 7) * 16 + (size_t)(kk & 7) * 2;
        uint8_t *base = out + off + (size_t)slab * nh * 64;
        *reinterpret_cast<__half *>(base + o) = hi;
        *reinterpret_cast<__half *>(base + (size_t)nh * 32 + o) = lo;
    }
}

// per-query record from the folded constants (cvec: bias + latent part of every layer, simt.cu:cvec_kernel)
__global__ void mlp_records_kernel(const float *__restrict__ cvec, int cvec_stride, const int *__restrict__ coff,
                                   const float *__restrict__ W0, const float *__restrict__ W6, float *__restrict__ recs)
{
    const int qi = blockIdx.x;
    const float *cv = cvec + (size_t)qi * cvec_stride;
    float *rec = recs + (size_t)qi * kRecFloats;
    const int d_in = 3 + kCond;
    for (int n = threadIdx.x; n < kH; n += blockDim.x) {
        const float *w = W0 + (size_t)n * d_in;
        rec[kRecL0 + 4 * n + 0] = w[0]; rec[kRecL0 + 4 * n + 1] = w[1]; rec[kRecL0 + 4 * n + 2] = w[2];
        rec[kRecL0 + 4 * n + 3] = kS * cv[coff[0] + n];
    }
    for (int t = 0; t < kTL; ++t) {
        const int n_real = t == 1 ? kN2 : kH;
        for (int n = threadIdx.x; n < layer_np(t); n += blockDim.x)
            rec[rec_bias_off(t) + n] = n < n_real ? kS * cv[coff[t + 1] + n] : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * kH; i += blockDim.x) rec[kRecW6 + i] = W6[i] / kS;
    if (threadIdx.x < 4) rec[kRecB6 + threadIdx.x] = threadIdx.x < 3 ? cv[coff[6] + threadIdx.x] : 0.f;
}

}  // namespace tcm

bool tc_mlp_supported(const nphm_mlp *h)
{
    return h->cfg.hidden_dim == tcm::kH && h->cfg.n_layers == 6 && h->cfg.lat_dim == tcm::kCond && h->cfg.out_dim == tcm::kOut &&
           h->dims.N[2] == tcm::kN2 && h->dims.skip == 3;
}

int tc_mlp_pack(nphm_mlp *h, cudaStream_t stream)
{
    h->tc_ready = false;
    if (!tc_mlp_supported(h)) return NPHM_OK;
    int rc;
    if ((rc = h->tc_weights.reserve(tcm::kWeightBytes))) return rc;
    tcm::pack_mlp_slabs_kernel<<<512, 256, 0, stream>>>(h->weights.W[1].as<float>(), h->weights.W[2].as<float>(),
                                                        h->weights.W[3].as<float>(), h->weights.W[4].as<float>(),
                                                        h->weights.W[5].as<float>(), h->tc_weights.as<uint8_t>());
    NPHM_CUDA_CHECK(cudaGetLastError());
    if ((rc = h->tc_coff.reserve(kMaxLayers * sizeof(int)))) return rc;
    NPHM_CUDA_CHECK(cudaMemcpyAsync(h->tc_coff.ptr, h->dims.coff, kMaxLayers * sizeof(int), cudaMemcpyHostToDevice, stream));
    NPHM_CUDA_CHECK(cudaStreamSynchronize(stream));
    h->tc_ready = true;
    return NPHM_OK;
}

int tc_mlp_launch(nphm_mlp *h, const float *xyz, const float *cvec, int n_queries, long long n_points, float *out,
                  cudaStream_t stream)
{
    NPHM_REQUIRE(h->tc_ready, "tcgen05 MLP kernel: weights not packed");
    int rc;
    if ((rc = h->tc_consts.reserve((size_t)n_queries * tcm::kRecFloats * sizeof(float)))) return rc;
    if (!h->tc_records_fresh) {
        tcm::mlp_records_kernel<<<n_queries, 256, 0, stream>>>(cvec, h->dims.cvec_stride, h->tc_coff.as<int>(),
                                                              h->weights.W[0].as<float>(), h->weights.W[6].as<float>(),
                                                              h->tc_consts.as<float>());
        NPHM_CUDA_CHECK(cudaGetLastError());
    }
    tcm::Params p{h->tc_weights.as<uint8_t>(), h->tc_consts.as<float>(), xyz, n_points, n_queries, out, h->tc_live};
    const long long n_tiles = ceil_div(n_points, 64) * n_queries;
    const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    const int smem = (int)sizeof(tcm::Smem);
    NPHM_CUDA_CHECK(cudaFuncSetAttribute(tcm::mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tcm::mlp_tc_kernel<<<grid, tcm::kThreads, smem, stream>>>(p);
    NPHM_CUDA_CHECK(cudaGetLastError());
    return NPHM_OK;
}

}  // namespace nphm
