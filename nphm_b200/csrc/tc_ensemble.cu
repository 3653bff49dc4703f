// tcgen05 / TMEM kernel for the identity-SDF ensemble (NPHM configuration: 40 members, hidden 200, 4 hidden
// layers, condition 64+32) - the dominant kernel of the hot path.
//
// Reference semantics: FastEnsembleDeepSDFMirrored.forward  src/NPHM/models/EnsembledDeepSDF.py:203-267
//
// Math (SURVEY.md 8a'): per member, per point
//   h0 = sp(W0x c + v0)            K=3   -> CUDA cores (v0 = latent part + bias, per query/member)
//   h1 = sp(W1 h0 + b1)            128x112x208 UMMA   (N 101 -> 112, K 200 -> 208)
//   h2 = sp(W2a h1/r2 + W2x c/r2 + v2)   128x208x112 UMMA   (K = 101 + 3 -> 112)
//   h3 = sp(W3 h2 + b3)            128x208x208 UMMA
//   s  = w4 . h3 + b4              CUDA cores, fused into the h3 epilogue; Gaussian anchor blend in registers.
// Precision: tensor cores run kind::f16 with fp32 accumulation; every operand is split in two fp16 terms
// (x = hi + lo, 22 significant bits) and each product is evaluated as hi*hi + hi*lo + lo*hi (3 MMAs), which
// keeps the result at fp32 round-off level (measured ~1e-7 abs against the fp32 reference, tolerance 1e-5).
// Activations are kept in "log2 units": t = a * 100*log2(e), sp'(t) = max(t,0) + lg2(1 + 2^-|t|) = 100*log2(e) *
// softplus_100(a), so the softplus costs 2 MUFU + 3 ALU and the unit change is folded into biases / w4.
//
// Data flow per CTA (persistent, one CTA per SM, 18 warps):
//   warp 16 : bulk-async-copy (TMA engine, cp.async.bulk) producer: streams pre-split fp16 weight slabs
//             (N x 16 K-columns, hi|lo, UMMA no-swizzle K-major core-matrix order) from L2 into a 14-slot ring,
//             and the per-(query,member) constant record (layer-0 weights, biases, w4, anchor) into a 2-slot ring.
//   warp 17 : allocates TMEM, single-thread tcgen05.mma issuer: A (activations) from TMEM, B (weights) from smem,
//             D (fp32) in TMEM; tcgen05.commit releases ring slots and signals the epilogue.
//   warps 0-15: thread = point (TMEM lane) x column group: read D with tcgen05.ld, softplus, split to fp16 hi/lo,
//             write the next layer's A operand back to TMEM with tcgen05.st, pre-load D with the next bias.
// TMEM map (columns): D [0,208)  A_hi [208,312)  A_lo [312,416)   (fp16 pairs, 2 K-values per column).
#include "tc_ensemble.cuh"
#include <cstdlib>

namespace nphm {
namespace tc {

// store 8 consecutive activations as fp16 hi/lo pairs: col_hi / col_lo = TMEM column of the first pair (2 K values per column)
__device__ __forceinline__ void store_a8(uint32_t col_hi, uint32_t col_lo, const float (&v)[8])
{
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
    tc_st4(col_hi, hi);
    tc_st4(col_lo, lo);
}
__device__ __forceinline__ void store_a4(uint32_t col_hi, uint32_t col_lo, const float (&v)[4])
{
    uint32_t h0, l0, h1, l1;
    split2(v[0], v[1], h0, l0); split2(v[2], v[3], h1, l1);
    tc_st2(col_hi, h0, h1);
    tc_st2(col_lo, l0, l1);
}

// ------------------------------------------------------------------------------------------------ packing
// Weight slabs: for weight set s, tensor layer L (1..3), k-step j: N x 16 fp16 hi then N x 16 fp16 lo, each in UMMA
// no-swizzle K-major core-matrix order: byte offset of (n, kk) = (n/8)*256 + (kk/8)*128 + (n%8)*16 + (kk%8)*2.
// Bias rows: the K padding of layers 1 and 3 (k = 200) carries S * b_l (per weight set, no latent part); the A operand has
// the constant 1.0 at that k, so the MMAs add the bias and the epilogues do not (layer 2's constant depends on the latent:
// its k-step-6 slab is re-built per (query, member), see l2_slab_kernel).
__global__ void pack_slabs_kernel(const float *__restrict__ W1, const float *__restrict__ W2, const float *__restrict__ W3,
                                  const float *__restrict__ b1, const float *__restrict__ b3,
                                  int n_sets, uint8_t *__restrict__ out)
{
    const int total_per_set = (kKS1 * kNP1 + (kKS2 + kKS3) * kNP2) * 16;       // (slab, n, kk) triples
    const float inv_sqrt2 = 0.70710678118654752440f;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < (size_t)n_sets * total_per_set;
         t += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(t / total_per_set);
        int r = (int)(t % total_per_set);
        int layer, j, n, kk, np;
        size_t slab_off;
        if (r < kKS1 * kNP1 * 16) {
            layer = 1; np = kNP1; j = r / (np * 16); r -= j * np * 16; slab_off = (size_t)j * kSlab1Bytes;
        } else if ((r -= kKS1 * kNP1 * 16) < kKS2 * kNP2 * 16) {
            layer = 2; np = kNP2; j = r / (np * 16); r -= j * np * 16; slab_off = kL1Bytes + (size_t)j * kSlabBytes;
        } else {
            r -= kKS2 * kNP2 * 16;
            layer = 3; np = kNP3; j = r / (np * 16); r -= j * np * 16; slab_off = kL1Bytes + kL2Bytes + (size_t)j * kSlabBytes;
        }
        n = r / 16; kk = r % 16;
        const int k = j * 16 + kk;
        float v = 0.f;
        if (layer == 1) {
            if (n < kN1 && k < kH) v = W1[((size_t)s * kN1 + n) * kH + k];
            else if (n < kN1 && k == kH) v = kS * b1[(size_t)s * kN1 + n];
        } else if (layer == 2) {
            // reference input order of the skip layer: [h1 (101), xyz (3), cond (96)] / sqrt(2); the cond part is folded
            if (n < kH && k < kN1 + 3) v = W2[((size_t)s * kH + n) * kH + k] * inv_sqrt2;
        } else {
            if (n < kH && k < kH) v = W3[((size_t)s * kH + n) * kH + k];
            else if (n < kH && k == kH) v = kS * b3[(size_t)s * kH + n];
        }
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t off = (size_t)(n >> 3) * 256 + (size_t)(kk >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(kk & 7) * 2;
        uint8_t *base = out + (size_t)s * kSetBytes + slab_off;
        *reinterpret_cast<__half *>(base + off) = hi;
        *reinterpret_cast<__half *>(base + (size_t)np * 32 + off) = lo;
    }
}

// record of (query, member): see kRec* offsets.  cvec holds b_l + latent-folded parts (simt.cu:cvec_kernel).
__global__ void records_kernel(const float *__restrict__ cvec, int cvec_stride, const int *__restrict__ coff,
                               const float *__restrict__ W0, const float *__restrict__ W4,
                               const float *__restrict__ anchors, int n_members, int n_symm, float *__restrict__ recs)
{
    const int m = blockIdx.x, qi = blockIdx.y;
    const int set = m < 2 * n_symm ? (m >> 1) : m - n_symm;
    const float *cv = cvec + ((size_t)qi * n_members + m) * cvec_stride;
    float *rec = recs + ((size_t)qi * n_members + m) * kRecFloats;
    const int d_in = 3 + kCond;
    for (int n = threadIdx.x; n < kNP2; n += blockDim.x) {
        const float *w = W0 + ((size_t)set * kH + (n < kH ? n : 0)) * d_in;
        const bool real = n < kH;
        rec[kRecL0 + 4 * n + 0] = real ? w[0] : 0.f;
        rec[kRecL0 + 4 * n + 1] = real ? w[1] : 0.f;
        rec[kRecL0 + 4 * n + 2] = real ? w[2] : 0.f;
        rec[kRecL0 + 4 * n + 3] = real ? kS * cv[coff[0] + n] : 0.f;
    }
    for (int n = threadIdx.x; n < kNP1; n += blockDim.x) rec[kRecB1 + n] = n < kN1 ? kS * cv[coff[1] + n] : 0.f;
    for (int n = threadIdx.x; n < kNP2; n += blockDim.x) {
        rec[kRecB2 + n] = n < kH ? kS * cv[coff[2] + n] : 0.f;
        rec[kRecB3 + n] = n < kH ? kS * cv[coff[3] + n] : 0.f;
        rec[kRecW4 + n] = n < kH ? W4[(size_t)set * kH + n] / kS : 0.f;
    }
    if (threadIdx.x == 0) {
        const bool has_anchor = m < n_members - 1;
        rec[kRecMisc + 0] = cv[coff[4]];
        rec[kRecMisc + 1] = has_anchor ? anchors[((size_t)qi * (n_members - 1) + m) * 3 + 0] : 0.f;
        rec[kRecMisc + 2] = has_anchor ? anchors[((size_t)qi * (n_members - 1) + m) * 3 + 1] : 0.f;
        rec[kRecMisc + 3] = has_anchor ? anchors[((size_t)qi * (n_members - 1) + m) * 3 + 2] : 0.f;
        rec[kRecMisc + 4] = has_anchor ? 1.f : 0.f;
        rec[kRecMisc + 5] = ((m & 1) && m < 2 * n_symm) ? 1.f : 0.f;
        rec[kRecMisc + 6] = 0.f;
        rec[kRecMisc + 7] = 0.f;
    }
}

// Layer 2's per-column constant v2 = S * (b2 + W2u u / sqrt(2)) depends on the latent, i.e. on (query, member): the last
// k-step slab of layer 2 (k 96..111, of which 96..103 are real inputs) is copied per (query, member) and its K-padding row
// k = 104 filled with v2 (fp16 hi | lo); the A operand carries 1.0 there.  13 KB per (query, member).
__global__ void l2_slab_kernel(const uint8_t *__restrict__ weights, const float *__restrict__ cvec, int cvec_stride,
                               const int *__restrict__ coff, int n_members, int n_symm, uint8_t *__restrict__ out)
{
    const int m = blockIdx.x, qi = blockIdx.y;
    const int set = m < 2 * n_symm ? (m >> 1) : m - n_symm;
    const uint4 *src = reinterpret_cast<const uint4 *>(weights + (size_t)set * kSetBytes + kL1Bytes + (size_t)(kKS2 - 1) * kSlabBytes);
    uint8_t *dst = out + ((size_t)qi * n_members + m) * kSlabBytes;
    for (int i = threadIdx.x; i < kSlabBytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(dst)[i] = src[i];
    __syncthreads();
    const float *cv = cvec + ((size_t)qi * n_members + m) * cvec_stride;
    const int kk = (kN1 + 3) - (kKS2 - 1) * 16;          // 104 - 96 = 8
    for (int n = threadIdx.x; n < kNP2; n += blockDim.x) {
        const float v = n < kH ? kS * cv[coff[2] + n] : 0.f;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t off = (size_t)(n >> 3) * 256 + (size_t)(kk >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(kk & 7) * 2;
        *reinterpret_cast<__half *>(dst + off) = hi;
        *reinterpret_cast<__half *>(dst + (size_t)kNP2 * 32 + off) = lo;
    }
}

// ------------------------------------------------------------------------------------------------ MMA self test
// One CTA: D[128 x n] = A[128 x 16*ks] * B[n x 16*ks]^T with the exact operand plumbing of the main kernel
// (fp16 hi/lo split, A in TMEM, B slabs in smem).  `variant` bit0: swap LBO/SBO, bit1: swap the fp16 pair order.
__global__ void __launch_bounds__(160, 1) mma_selftest_kernel(const float *__restrict__ A, const uint8_t *__restrict__ slabs,
                                                              int n, int ks, int variant, float *__restrict__ D)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slab_bytes = n * 64;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ks * slab_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(base)[i] = reinterpret_cast<const uint4 *>(slabs)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (warp < 4) {
        const int row = warp * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        for (int k0 = 0; k0 < ks * 16; k0 += 8) {
            float v[8];
            for (int e = 0; e < 8; ++e) v[e] = A[(size_t)row * ks * 16 + k0 + e];
            if (variant & 2) for (int e = 0; e < 8; e += 2) { const float t = v[e]; v[e] = v[e + 1]; v[e + 1] = t; }
            store_a8(tl + kColAhi + (k0 >> 1), tl + kColAlo + (k0 >> 1), v);
        }
        const uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < n; c += 8) tc_st8(tl + kColD + c, zero);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 128) {
        const uint32_t idesc = make_idesc(n);
        const uint32_t lbo = (variant & 1) ? 256 : 128, sbo = (variant & 1) ? 128 : 256;
        for (int j = 0; j < ks; ++j) {
            const uint32_t b = smem_u32(base + (size_t)j * slab_bytes);
            const uint64_t b_hi = make_desc(b, lbo, sbo), b_lo = make_desc(b + n * 32, lbo, sbo);
            tc_mma_ts(tmem + kColD, tmem + kColAhi + j * 8, b_hi, idesc, 1);
            tc_mma_ts(tmem + kColD, tmem + kColAhi + j * 8, b_lo, idesc, 1);
            tc_mma_ts(tmem + kColD, tmem + kColAlo + j * 8, b_hi, idesc, 1);
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (warp < 4) {
        const int row = warp * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < n; c += 8) {
            uint32_t r[8];
            tc_ld8(tl + kColD + c, r);
            tc_wait_ld();
            for (int e = 0; e < 8; ++e) D[(size_t)row * n + c + e] = __uint_as_float(r[e]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
}

__global__ void pack_test_slabs_kernel(const float *__restrict__ B, int n, int ks, uint8_t *__restrict__ out)
{
    // B: [n][16*ks] fp32 row-major -> ks slabs (hi | lo) in core-matrix order
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * ks * 16; t += gridDim.x * blockDim.x) {
        const int row = t / (ks * 16), k = t % (ks * 16), j = k / 16, kk = k % 16;
        const float v = B[t];
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t off = (size_t)(row >> 3) * 256 + (size_t)(kk >> 3) * 128 + (size_t)(row & 7) * 16 + (size_t)(kk & 7) * 2;
        uint8_t *base = out + (size_t)j * n * 64;
        *reinterpret_cast<__half *>(base + off) = hi;
        *reinterpret_cast<__half *>(base + (size_t)n * 32 + off) = lo;
    }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------ host side
bool tc_ensemble_supported(const nphm_ensemble *h)
{
    return h->cfg.hidden_dim == tc::kH && h->cfg.n_layers == 4 && h->cfg.lat_dim_glob + h->cfg.lat_dim_loc == tc::kCond &&
           h->dims.N[1] == tc::kN1;
}

int tc_ensemble_pack(nphm_ensemble *h, cudaStream_t stream)
{
    h->tc_ready = false;
    if (!tc_ensemble_supported(h)) return NPHM_OK;
    int rc;
    if ((rc = h->tc_weights.reserve((size_t)h->n_sets * tc::kSetBytes))) return rc;
    tc::pack_slabs_kernel<<<512, 256, 0, stream>>>(h->weights.W[1].as<float>(), h->weights.W[2].as<float>(),
                                                   h->weights.W[3].as<float>(), h->weights.b[1].as<float>(),
                                                   h->weights.b[3].as<float>(), h->n_sets, h->tc_weights.as<uint8_t>());
    NPHM_CUDA_CHECK(cudaGetLastError());
    if ((rc = h->tc_coff.reserve(kMaxLayers * sizeof(int)))) return rc;
    NPHM_CUDA_CHECK(cudaMemcpyAsync(h->tc_coff.ptr, h->dims.coff, kMaxLayers * sizeof(int), cudaMemcpyHostToDevice, stream));
    NPHM_CUDA_CHECK(cudaStreamSynchronize(stream));      // dims.coff is host memory of the handle; keep it simple
    h->tc_ready = true;
    return NPHM_OK;
}

int tc_ensemble_launch(nphm_ensemble *h, const SimtQuery &q, cudaStream_t stream)
{
    NPHM_REQUIRE(h->tc_ready, "tcgen05 ensemble kernel: weights not packed");
    int rc;
    if ((rc = h->tc_consts.reserve((size_t)q.n_queries * h->n_members * tc::kRecFloats * sizeof(float)))) return rc;
    dim3 grid(h->n_members, q.n_queries);
    tc::records_kernel<<<grid, 256, 0, stream>>>(q.cvec, h->dims.cvec_stride, h->tc_coff.as<int>(), h->weights.W[0].as<float>(),
                                                 h->weights.W[4].as<float>(), q.anchors, h->n_members, h->cfg.n_symm_pairs,
                                                 h->tc_consts.as<float>());
    NPHM_CUDA_CHECK(cudaGetLastError());
    if ((rc = h->tc_l2slabs.reserve((size_t)q.n_queries * h->n_members * tc::kSlabBytes))) return rc;
    tc::l2_slab_kernel<<<grid, 256, 0, stream>>>(h->tc_weights.as<uint8_t>(), q.cvec, h->dims.cvec_stride, h->tc_coff.as<int>(),
                                                 h->n_members, h->cfg.n_symm_pairs, h->tc_l2slabs.as<uint8_t>());
    NPHM_CUDA_CHECK(cudaGetLastError());
    tc::Params p{};
    p.weights = h->tc_weights.as<uint8_t>();
    p.l2_slabs = h->tc_l2slabs.as<uint8_t>();
    p.recs = h->tc_consts.as<float>();
    p.xyz = q.xyz; p.axes = q.axes; p.res = q.res; p.first = q.first; p.total = q.total; p.n_points = q.n_points;
    p.n_queries = q.n_queries; p.quirk_period = q.quirk_period; p.out = q.out; p.members_out = q.members_out; p.acts_out = q.acts_out; p.acts_packed_out = q.acts_packed_out; p.acts_packed_tile_steps = q.acts_packed_tile_steps;
    p.n_members = h->n_members; p.n_symm = h->cfg.n_symm_pairs;
    const bool prune = h->tc_prune && !q.exact;
    NPHM_REQUIRE(!q.acts_out || (q.n_queries == 1 && q.exact && q.acts_packed_out && q.acts_packed_tile_steps >= tc::kActPackedSteps), "activation dump needs a single exact query");
    p.anchors = q.anchors; p.prune_tau = h->tc_prune_tau;
    p.blocked = 0; p.px0 = p.px1 = 0; p.by = p.bz = 1;
    long long n_tiles = ceil_div(q.n_points, 128) * q.n_queries;
    if (prune && !q.xyz && q.n_points > 0) {
        // compact blocks over the x-plane range that contains [first, first + n_points)
        const long long rr = (long long)q.res * q.res;
        p.px0 = (int)(q.first / rr);
        p.px1 = (int)((q.first + q.n_points - 1) / rr);
        p.by = (q.res + 3) / 4; p.bz = (q.res + 3) / 4;
        n_tiles = (long long)((p.px1 - p.px0 + 8) / 8) * p.by * p.bz;
        p.blocked = 1;
    }
    p.n_tiles = n_tiles;
    p.member_groups = 1;
    if (q.acts_out) {
        // fitting: a few dozen tiles - split the members of a tile over several CTAs so that every SM has work.  Cost model:
        // waves * (members per group + pipeline fill of a work item); groups must all be non-empty.
        double best = 1e30;
        for (int g = 1; g <= h->n_members; ++g) {
            const int per = (h->n_members + g - 1) / g;
            if (per * (g - 1) >= h->n_members) continue;
            const double cost = (double)ceil_div(n_tiles * g, (long long)sm_count()) * (per + 0.7);
            if (cost < best - 1e-9) { best = cost; p.member_groups = g; }
        }
    }
    const long long n_items = n_tiles * p.member_groups;
    const int grid_x = (int)(n_items < sm_count() ? n_items : sm_count());
    if ((rc = tc::launch_ensemble_v8(p, prune, q.acts_out != nullptr, grid_x, stream))) return rc;
    return NPHM_OK;
}

}  // namespace nphm

// Debug entry (not part of the public ABI): D = A * B^T through the tensor-core operand path.
extern "C" int nphm_debug_tc_mma(const float *a_dev, const float *b_dev, int n, int ks, int variant, float *d_dev,
                                 void *stream_)
{
    using namespace nphm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NPHM_REQUIRE(n % 16 == 0 && n >= 16 && n <= 208 && ks >= 1 && ks <= 13, "nphm_debug_tc_mma: bad shape");
    uint8_t *slabs = nullptr;
    NPHM_CUDA_CHECK(cudaMalloc(&slabs, (size_t)ks * n * 64));
    tc::pack_test_slabs_kernel<<<64, 256, 0, stream>>>(b_dev, n, ks, slabs);
    const int smem = ks * n * 64 + 1024;
    NPHM_CUDA_CHECK(cudaFuncSetAttribute(tc::mma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc::mma_selftest_kernel<<<1, 160, smem, stream>>>(a_dev, slabs, n, ks, variant, d_dev);
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(slabs);
    if (e != cudaSuccess) { set_error("nphm_debug_tc_mma: %s", cudaGetErrorString(e)); return NPHM_ERR_CUDA; }
    return NPHM_OK;
}

// Micro-benchmark: `iters` MMAs (M=128, K=16, A from TMEM, B from smem) accumulating into one D (alternate=0) or
// round-robin into `alternate` disjoint D ranges; returns SM cycles from first issue to commit completion.
namespace nphm { namespace tc {
__global__ void __launch_bounds__(160, 1) mma_bench_kernel(int n, int iters, int alternate, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 208 * 64 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 128) {
        const uint32_t idesc = make_idesc(n);
        const uint64_t b = make_desc(smem_u32(smem_raw), 128, 256);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int d = alternate > 1 ? (i % alternate) * n : 0;
            tc_mma_ts(tmem + d, tmem + 448, b, idesc, 1);
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        cycles[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
}
}}
extern "C" int nphm_debug_tc_mma_bench(int n, int iters, int alternate, long long *cycles_host)
{
    using namespace nphm;
    long long *d = nullptr;
    NPHM_CUDA_CHECK(cudaMalloc(&d, 8));
    const int smem = 208 * 64 + 1024;
    NPHM_CUDA_CHECK(cudaFuncSetAttribute(tc::mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc::mma_bench_kernel<<<1, 160, smem>>>(n, iters, alternate, d);
    cudaError_t e = cudaMemcpy(cycles_host, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("nphm_debug_tc_mma_bench: %s", cudaGetErrorString(e)); return NPHM_ERR_CUDA; }
    return NPHM_OK;
}

// Layout probe for M = 64, A and B both from shared memory (SS form): D[64 x n] = A[64 x 16*ks] * B[n x 16*ks]^T.
// A is stored K-chunk-major (for each 8-column K chunk: 64 rows x 16 B contiguous; LBO = 1024 B, SBO = 128 B), hi plane
// then lo plane.  The whole TMEM accumulator region (128 lanes x n columns) is dumped so the host can recover the
// row -> lane mapping.
namespace nphm { namespace tc {
__global__ void __launch_bounds__(160, 1) mma_m64_probe_kernel(const float *__restrict__ A, const uint8_t *__restrict__ slabs,
                                                               int n, int ks, int variant, float *__restrict__ dump)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slab_bytes = n * 64;
    uint8_t *a_hi = smem_raw, *a_lo = smem_raw + ks * 4096, *b_base = smem_raw + ks * 8192;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ks * slab_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(b_base)[i] = reinterpret_cast<const uint4 *>(slabs)[i];
    const int m_rows = (variant & 8) ? 128 : 64;
    const int vlay = variant & 3;
    if (threadIdx.x < m_rows) {
        const int row = threadIdx.x;
        for (int kc = 0; kc < ks * 2; ++kc) {            // 8-column K chunks
            uint32_t hi[4], lo[4];
            for (int i = 0; i < 4; ++i) split2(A[(size_t)row * ks * 16 + kc * 8 + 2 * i], A[(size_t)row * ks * 16 + kc * 8 + 2 * i + 1], hi[i], lo[i]);
            // variant 0/1: K-chunk-major (chunk kc: 64 rows x 16 B); variant 2: like the B slabs (row-group-major per k-step)
            const size_t off = vlay == 2 ? (size_t)(kc >> 1) * (m_rows * 32) + (size_t)(row >> 3) * 256 + (size_t)(kc & 1) * 128 + (size_t)(row & 7) * 16
                                         : (size_t)kc * (m_rows * 16) + (size_t)row * 16;
            *reinterpret_cast<uint4 *>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (warp < 4) {
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t fill[8];
        for (int e = 0; e < 8; ++e) fill[e] = __float_as_uint(0.f);
        for (int c = 0; c < n; c += 8) tc_st8(tl + c, fill);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 128) {
        const int m_rows2 = (variant & 8) ? 128 : 64;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m_rows2 >> 4) << 24);
        for (int j = 0; j < ks; ++j) {
            const uint32_t albo = (variant & 3) == 2 ? 128 : m_rows2 * 16, asbo = (variant & 3) == 2 ? 256 : 128;
            const uint32_t astep = m_rows2 * 32;
            uint64_t ah = make_desc(smem_u32(a_hi + j * astep), albo, asbo), al = make_desc(smem_u32(a_lo + j * astep), albo, asbo);
            const uint32_t b = smem_u32(b_base + (size_t)j * slab_bytes);
            uint64_t bh = make_desc(b, 128, 256), bl = make_desc(b + n * 32, 128, 256);
            if (variant & 4) { uint64_t t = ah; ah = bh; bh = t; t = al; al = bl; bl = t; }      // swap operand order
            tc_mma_ss(tmem, ah, bh, idesc, 1);
            tc_mma_ss(tmem, ah, bl, idesc, 1);
            tc_mma_ss(tmem, al, bh, idesc, 1);
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (warp < 4) {
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < n; c += 8) {
            uint32_t r[8];
            tc_ld8(tl + c, r);
            tc_wait_ld();
            for (int e = 0; e < 8; ++e) dump[(size_t)(warp * 32 + lane) * n + c + e] = __uint_as_float(r[e]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
}
}}
extern "C" int nphm_debug_tc_mma_m64(const float *a_dev, const float *b_dev, int n, int ks, int variant, float *dump_dev, void *stream_)
{
    using namespace nphm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NPHM_REQUIRE(n % 16 == 0 && n >= 16 && n <= 256 && ks >= 1 && ks <= 8, "nphm_debug_tc_mma_m64: bad shape");
    uint8_t *slabs = nullptr;
    NPHM_CUDA_CHECK(cudaMalloc(&slabs, (size_t)ks * n * 64));
    tc::pack_test_slabs_kernel<<<64, 256, 0, stream>>>(b_dev, n, ks, slabs);
    const int smem = ks * 8192 + ks * n * 64 + 1024;
    NPHM_CUDA_CHECK(cudaFuncSetAttribute(tc::mma_m64_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc::mma_m64_probe_kernel<<<1, 160, smem, stream>>>(a_dev, slabs, n, ks, variant, dump_dev);
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(slabs);
    if (e != cudaSuccess) { set_error("nphm_debug_tc_mma_m64: %s", cudaGetErrorString(e)); return NPHM_ERR_CUDA; }
    return NPHM_OK;
}
