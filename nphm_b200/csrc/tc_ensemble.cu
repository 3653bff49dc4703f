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

struct __align__(128) Smem {
    uint8_t wbuf[2][kGroupBytes];            // double-buffered weight groups (one bulk copy + one barrier each)
    float rec[kRecSlots][kRecFloats];
    float partial[kParts - 1][128];
    uint64_t w_full[2], w_empty[2];
    uint64_t rec_full[kRecSlots], rec_empty[kRecSlots];
    uint64_t a0a_ready, a0b_ready, a1_ready[3], a2_ready, d_ready, mask_ready;
    unsigned long long maskq[2][4];
    uint32_t tmem_base;
};


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
// accumulator columns + per-column constant (bias / folded latent part) from the shared-memory record
__device__ __forceinline__ void add_bias8(uint32_t (&r)[8], const float *bias)
{
    const float4 b0 = *reinterpret_cast<const float4 *>(bias);
    const float4 b1 = *reinterpret_cast<const float4 *>(bias + 4);
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + b[e]);
}
__device__ __forceinline__ void add_bias4(uint32_t (&r)[4], const float *bias)
{
    const float4 b0 = *reinterpret_cast<const float4 *>(bias);
    const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + b[e]);
}

// ACTS: also write the hidden activations for the fitting backward (separate instantiation: no cost on the query path)
template <bool PRUNE, bool ACTS>
__global__ void __launch_bounds__(kThreads, 1) ensemble_tc_kernel(const Params p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tiles_per_query = p.blocked ? p.n_tiles : (p.n_points + 127) / 128;
    const long long n_tiles = p.n_tiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
        for (int i = 0; i < kRecSlots; ++i) { mbar_init(&sm.rec_full[i], 1); mbar_init(&sm.rec_empty[i], kEpiWarps); }
        mbar_init(&sm.a0a_ready, kEpiWarps);
        mbar_init(&sm.a0b_ready, kEpiWarps);
        for (int i = 0; i < 3; ++i) mbar_init(&sm.a1_ready[i], kEpiWarps);
        mbar_init(&sm.a2_ready, kEpiWarps);
        mbar_init(&sm.d_ready, 1);
        mbar_init(&sm.mask_ready, 4);
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
        // =========================================================================== producer (bulk async copies)
        if (lane == 0) {
            int wb = 0;
            uint32_t wph = 0, tcount = 0, rcount = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                const int qi = (int)(tile / tiles_per_query);
                unsigned long long mask = (1ull << p.n_members) - 1;
                if (PRUNE) {
                    mbar_wait(&sm.mask_ready, tcount & 1);
                    const unsigned long long *mq = sm.maskq[tcount & 1];
                    mask = mq[0] | mq[1] | mq[2] | mq[3];
                }
                for (int m = 0; m < p.n_members; ++m) {
                    if (PRUNE && !((mask >> m) & 1)) continue;
                    {
                        const int rslot = rcount % kRecSlots;
                        mbar_wait(&sm.rec_empty[rslot], ((rcount / kRecSlots) & 1) ^ 1);
                        mbar_expect_tx(&sm.rec_full[rslot], kRecFloats * 4);
                        bulk_g2s(sm.rec[rslot], p.recs + ((size_t)qi * p.n_members + m) * kRecFloats, kRecFloats * 4,
                                 &sm.rec_full[rslot]);
                        ++rcount;
                    }
                    const int set = m < 2 * p.n_symm ? (m >> 1) : m - p.n_symm;
                    const uint8_t *w = p.weights + (size_t)set * kSetBytes;
#pragma unroll 1
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t bytes = g == 0 ? kL1Bytes : (g == 1 ? kL2Bytes : (g == 2 ? 7 * kSlabBytes : 6 * kSlabBytes));
                        mbar_wait(&sm.w_empty[wb], wph ^ 1);
                        mbar_expect_tx(&sm.w_full[wb], bytes);
                        bulk_g2s(sm.wbuf[wb], w, bytes, &sm.w_full[wb]);
                        w += bytes;
                        if (++wb == 2) { wb = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // =========================================================================== MMA issuer
        if (lane == 0) {
            int wb = 0;
            uint32_t wph = 0, mph = 0, tcount = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                unsigned long long mask = (1ull << p.n_members) - 1;
                if (PRUNE) {
                    mbar_wait(&sm.mask_ready, tcount & 1);
                    const unsigned long long *mq = sm.maskq[tcount & 1];
                    mask = mq[0] | mq[1] | mq[2] | mq[3];
                }
                for (int m = 0; m < p.n_members; ++m) {
                    if (PRUNE && !((mask >> m) & 1)) continue;
                    // one k-step = 3 MMAs (hi*hi + hi*lo + lo*hi); `fresh` overwrites the accumulator (no bias preload)
                    auto kstep = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t slab, int n, uint32_t idesc, bool fresh) {
                        const uint64_t b_hi = make_desc(slab, 128, 256);
                        const uint64_t b_lo = make_desc(slab + n * 32, 128, 256);
                        tc_mma_ts(d, a_hi, b_hi, idesc, fresh ? 0 : 1);
                        tc_mma_ts(d, a_hi, b_lo, idesc, 1);
                        tc_mma_ts(d, a_lo, b_hi, idesc, 1);
                    };
                    auto next_buf = [&]() { if (++wb == 2) { wb = 0; wph ^= 1; } };
                    // ---- layer 1 (N 112): k-steps 0-5 read the operand written one member ahead, 6-12 the rest of layer 0
                    {
                        const uint32_t idesc = make_idesc(kNP1);
                        mbar_wait(&sm.a0a_ready, mph);
                        mbar_wait(&sm.w_full[wb], wph);
                        tc_fence_after();
                        const uint32_t base = smem_u32(sm.wbuf[wb]);
#pragma unroll 1
                        for (int j = 0; j < kNA / 16; ++j)
                            kstep(tmem + kColD1, tmem + kColSpareHi + j * 8, tmem + kColSpareLo + j * 8, base + j * kSlab1Bytes,
                                  kNP1, idesc, j == 0);
                        mbar_wait(&sm.a0b_ready, mph);
                        tc_fence_after();
#pragma unroll 1
                        for (int j = kNA / 16; j < kKS1; ++j)
                            kstep(tmem + kColD1, tmem + kColA0bHi + (j - kNA / 16) * 8, tmem + kColA0bLo + (j - kNA / 16) * 8,
                                  base + j * kSlab1Bytes, kNP1, idesc, false);
                        tc_commit(&sm.w_empty[wb]);
                        tc_commit(&sm.d_ready);
                        next_buf();
                    }
                    // ---- layer 2 (N 208, K 112): issued group by group while the layer-1 epilogue produces its A operand
                    {
                        const uint32_t idesc = make_idesc(kNP2);
                        mbar_wait(&sm.w_full[wb], wph);
                        const uint32_t base = smem_u32(sm.wbuf[wb]);
                        bool fresh = true;
#pragma unroll 1
                        for (int grp = 0; grp < 3; ++grp) {
                            const int j0 = grp == 0 ? 4 : (grp == 1 ? 2 : 0), j1 = grp == 0 ? 7 : (grp == 1 ? 4 : 2);
                            mbar_wait(&sm.a1_ready[grp], mph);
                            tc_fence_after();
                            for (int j = j0; j < j1; ++j) {
                                kstep(tmem + kColD2, tmem + kColA1Hi + j * 8, tmem + kColA1Lo + j * 8, base + j * kSlabBytes, kNP2,
                                      idesc, fresh);
                                fresh = false;
                            }
                        }
                        tc_commit(&sm.w_empty[wb]);
                        tc_commit(&sm.d_ready);
                        next_buf();
                    }
                    // ---- layer 3 (N 208, K 208) in two weight groups
                    {
                        const uint32_t idesc = make_idesc(kNP3);
                        mbar_wait(&sm.a2_ready, mph);
                        mbar_wait(&sm.w_full[wb], wph);
                        tc_fence_after();
                        uint32_t base = smem_u32(sm.wbuf[wb]);
#pragma unroll 1
                        for (int j = 0; j < 7; ++j)
                            kstep(tmem + kColD3, tmem + kColA2Hi + j * 8, tmem + kColA2Lo + j * 8, base + j * kSlabBytes, kNP3, idesc,
                                  j == 0);
                        tc_commit(&sm.w_empty[wb]);
                        next_buf();
                        mbar_wait(&sm.w_full[wb], wph);
                        tc_fence_after();
                        base = smem_u32(sm.wbuf[wb]);
#pragma unroll 1
                        for (int j = 7; j < kKS3; ++j)
                            kstep(tmem + kColD3, tmem + kColA2Hi + j * 8, tmem + kColA2Lo + j * 8, base + (j - 7) * kSlabBytes, kNP3,
                                  idesc, false);
                        tc_commit(&sm.w_empty[wb]);
                        tc_commit(&sm.d_ready);
                        next_buf();
                    }
                    mph ^= 1;
                }
            }
        }
    } else {
        // =========================================================================== compute / epilogue warps
        // thread = (point row, column group): warp w serves TMEM lanes 32*(w&3).. and column group part = w>>2.
        const int q = warp & 3, part = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);       // this warp's TMEM lane quarter
        uint32_t d_ph = 0, m_ph = 0, tcount = 0, rcount = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            int qi;
            long long idx, g;
            bool valid;
            float x, y, z;
            if (p.blocked) {
                // compact 8 x 4 x 4 block of grid points (z fastest inside the block)
                qi = 0;
                const long long tz = tile % p.bz, txy = tile / p.bz;
                const int ty = (int)(txy % p.by), tx = (int)(txy / p.by);
                const int ix = p.px0 + tx * 8 + (row >> 4), iy = ty * 4 + ((row >> 2) & 3), iz = (int)tz * 4 + (row & 3);
                g = ((long long)ix * p.res + iy) * p.res + iz;
                valid = ix <= p.px1 && iy < p.res && iz < p.res && g >= p.first && g < p.first + p.n_points;
                idx = g - p.first;
                const int cx_ = min(ix, p.res - 1), cy_ = min(iy, p.res - 1), cz_ = min(iz, p.res - 1);
                x = __ldg(p.axes + cx_); y = __ldg(p.axes + p.res + cy_); z = __ldg(p.axes + 2 * p.res + cz_);
                if (!valid) g = p.first;
            } else {
                qi = (int)(tile / tiles_per_query);
                idx = (tile - (long long)qi * tiles_per_query) * 128 + row;
                valid = idx < p.n_points;
                g = p.first + (valid ? idx : 0);
                if (p.xyz) {
                    const float *pp = p.xyz + ((size_t)qi * p.n_points + (valid ? idx : 0)) * 3;
                    x = pp[0]; y = pp[1]; z = pp[2];
                } else {
                    const long long rr = (long long)p.res * p.res;
                    const int ix = (int)(g / rr), iy = (int)((g - ix * rr) / p.res), iz = (int)(g % p.res);
                    x = __ldg(p.axes + ix); y = __ldg(p.axes + p.res + iy); z = __ldg(p.axes + 2 * p.res + iz);
                }
            }
            const bool quirk = p.quirk_period > 0 && ((g % p.quirk_period) == p.quirk_period - 1 || g == p.total - 1);
            float num = 0.f, den = 0.f;
            unsigned long long mask = (1ull << p.n_members) - 1;
            if (PRUNE) {
                // blend weights of all members for this thread's point: S = sum_k w_k; a member is needed by the tile if
                // w_k >= tau * (S + 1e-6) for at least one of its points (dropped mass per point < n_members * tau).
                if (part == 0) {
                    const float *anc = p.anchors + (size_t)qi * (p.n_members - 1) * 3;
                    float S = 0.f;
                    for (int k = 0; k < p.n_members; ++k) {
                        float d = -0.2f;
                        if (k < p.n_members - 1) {
                            const float dx = __ldg(anc + 3 * k) - x, dy = __ldg(anc + 3 * k + 1) - y, dz = __ldg(anc + 3 * k + 2) - z;
                            const float nrm = sqrtf(dx * dx + dy * dy + dz * dz) + 10e-6f;
                            d = -(nrm * nrm);
                        }
                        S += expf(__fdiv_rn(d, 0.01f));
                    }
                    den = S;
                    const float thr = p.prune_tau * (S + 1e-6f);
                    unsigned long long wm = 0;
                    for (int k = 0; k < p.n_members; ++k) {
                        float d = -0.2f;
                        if (k < p.n_members - 1) {
                            const float dx = __ldg(anc + 3 * k) - x, dy = __ldg(anc + 3 * k + 1) - y, dz = __ldg(anc + 3 * k + 2) - z;
                            const float nrm = sqrtf(dx * dx + dy * dy + dz * dz) + 10e-6f;
                            d = -(nrm * nrm);
                        }
                        const bool need = valid && expf(__fdiv_rn(d, 0.01f)) >= thr;
                        if (__any_sync(0xffffffffu, need)) wm |= 1ull << k;
                    }
                    wm |= 1ull << (p.n_members - 1);      // every tile evaluates >= 1 member: keeps all warps in lock step
                    if (lane == 0) {
                        sm.maskq[tcount & 1][q] = wm;
                        mbar_arrive(&sm.mask_ready);
                    }
                }
                mbar_wait(&sm.mask_ready, tcount & 1);
                const unsigned long long *mq = sm.maskq[tcount & 1];
                mask = mq[0] | mq[1] | mq[2] | mq[3];
            }

            // first 96 layer-0 outputs (chunks part, part+4, part+8) -> spare TMEM columns.  Normally computed for the NEXT
            // member while the current member's layer-3 MMAs run; at the start of a tile it is computed in place.
            // hidden activations for the fitting backward (h = v / S), feature-major so that a warp store is one 128-byte line
            auto save8 = [&](float *ab, int f0, int n_real, const float (&v)[8]) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (e < n_real) ab[(size_t)(f0 + e) * 128 + row] = v[e] * (1.0f / kS);
            };
            auto acts_of = [&](int member) -> float * {
                return ACTS ? p.acts_out + ((size_t)member * tiles_per_query + (tile % tiles_per_query)) * kActFeat * 128 : nullptr;
            };
            auto layer0_a = [&](const float *r, float ccx, float ccy, float ccz, float *ab) {
                const float4 *l0a = reinterpret_cast<const float4 *>(r + kRecL0);
#pragma unroll 1
                for (int c = part; c < 12; c += kParts) {
                    const int n0 = c * 8;
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float4 w = l0a[n0 + e];
                        const float t = fmaf(w.x, ccx, fmaf(w.y, ccy, fmaf(w.z, ccz, w.w)));
                        v[e] = (e & 1) ? sp_t_poly(t) : sp_t(t);
                    }
                    if (ACTS) save8(ab, n0, 8, v);
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
                    tc_st4(tl + kColSpareHi + (n0 >> 1), hi);
                    tc_st4(tl + kColSpareLo + (n0 >> 1), lo);
                }
            };
            auto member_coords = [&](const float *r, float &ccx, float &ccy, float &ccz) {
                ccx = x - r[kRecMisc + 1]; ccy = y - r[kRecMisc + 2]; ccz = z - r[kRecMisc + 3];
                if (r[kRecMisc + 5] != 0.f) ccx = -ccx;          // mirrored member
                ccx *= kS; ccy *= kS; ccz *= kS;                // coordinates in log2 units
            };
            bool a_done = false;

            for (int m = 0; m < p.n_members; ++m) {
                if (!((mask >> m) & 1)) continue;
                const int rslot = rcount % kRecSlots;
                mbar_wait(&sm.rec_full[rslot], (rcount / kRecSlots) & 1);
                const float *rec = sm.rec[rslot];
                const float ax = rec[kRecMisc + 1], ay = rec[kRecMisc + 2], az = rec[kRecMisc + 3];
                float cx, cy, cz;
                member_coords(rec, cx, cy, cz);

                // Column ownership (balanced: 6.5 chunks per warp and layer): 8-column chunks c = part + 4i (i < 6)
                // cover columns 0..191, the 4-column piece 192 + 4*part covers 192..207; for the 112-column layer 1:
                // chunks part + 4i (i < 3) and the piece 96 + 4*part.
                // Hazard notes (TMEM regions are reused, see the column map): a region is only overwritten after an
                // mbarrier has proven that its previous readers are done -
                //   A0b/A1 [0,112)  <- previous readers are MMAs that completed before the d_ready this warp waited on;
                //   D1 [112,224)    <- overlaps D3: the issuer waits for a0a_ready, which a warp signals after its last D3 read;
                //   D2 [304,512)    <- overlaps D3 and the A0a columns: both dead once layer 1 of this member was issued;
                //   A2 [0,208)      <- D1/A1 readers finished before d_ready (layer 2) fired;
                //   A0a [416,512)   <- overlaps D2: written only after a2_ready COMPLETED (every warp finished reading D2).
                const float4 *l0 = reinterpret_cast<const float4 *>(rec + kRecL0);
                auto publish = [&](uint64_t *bar) {          // my TMEM stores are visible to the MMA issuer after this
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar);
                };

                // ---------------- layer 0 on CUDA cores -> A operand of layer 1
                float *const ab = acts_of(m);
                if (!a_done) {
                    layer0_a(rec, cx, cy, cz, ab);
                    publish(&sm.a0a_ready);
                }
#pragma unroll 1
                for (int c = 12 + part; c < 24; c += kParts) {
                    const int n0 = c * 8;
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float4 w = l0[n0 + e];
                        const float t = fmaf(w.x, cx, fmaf(w.y, cy, fmaf(w.z, cz, w.w)));
                        v[e] = (e & 1) ? sp_t_poly(t) : sp_t(t);
                    }
                    store_a8(tl + kColA0bHi + ((n0 - kNA) >> 1), tl + kColA0bLo + ((n0 - kNA) >> 1), v);
                    if (ACTS) save8(ab, n0, 8, v);
                }
                {
                    const int n0 = 192 + 4 * part;               // rows >= 200 are zero: sp(0) meets zero weights
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float4 w = l0[n0 + e];
                        const float t = fmaf(w.x, cx, fmaf(w.y, cy, fmaf(w.z, cz, w.w)));
                        v[e] = (e & 1) ? sp_t_poly(t) : sp_t(t);
                    }
                    if (ACTS && n0 < kH) {
                        const float v8[8] = {v[0], v[1], v[2], v[3], 0.f, 0.f, 0.f, 0.f};
                        save8(ab, n0, 4, v8);
                    }
                    store_a4(tl + kColA0bHi + ((n0 - kNA) >> 1), tl + kColA0bLo + ((n0 - kNA) >> 1), v);
                }
                publish(&sm.a0b_ready);

                // ---------------- epilogue of layer 1 (N = 101 -> K of layer 2 = [h1, c, 0...]), published in three groups
                // (k-steps 4-6 | 2-3 | 0-1) so that the layer-2 MMAs start while the rest is still being converted
                mbar_wait(&sm.d_ready, d_ph);
                d_ph ^= 1;
                tc_fence_after();
                {
                    const int n0 = 96 + 4 * part;                // 96..99 | 100, c | padding | padding
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (part < 2) {
                        uint32_t r[4];
                        tc_ld4(tl + kColD1 + n0, r);
                        tc_wait_ld();
                        add_bias4(r, rec + kRecB1 + n0);
                        if (part == 0) sp4(r, v);
                        else { v[0] = sp_t(__uint_as_float(r[0])); v[1] = cx; v[2] = cy; v[3] = cz; }
                        if (ACTS) {
                            const float v8[8] = {v[0], v[1], v[2], v[3], 0.f, 0.f, 0.f, 0.f};
                            save8(ab, kH + n0, part == 0 ? 4 : 1, v8);          // h1 features 96..99 | 100
                        }
                    }
                    store_a4(tl + kColA1Hi + (n0 >> 1), tl + kColA1Lo + (n0 >> 1), v);
                }
#pragma unroll 1
                for (int grp = 0; grp < 3; ++grp) {
                    const int n0 = (8 - 4 * grp + part) * 8;     // chunk 8+part, 4+part, part
                    uint32_t r[8];
                    tc_ld8(tl + kColD1 + n0, r);
                    tc_wait_ld();
                    add_bias8(r, rec + kRecB1 + n0);
                    float v[8];
                    sp8(r, v);
                    store_a8(tl + kColA1Hi + (n0 >> 1), tl + kColA1Lo + (n0 >> 1), v);
                    if (ACTS) save8(ab, kH + n0, 8, v);
                    publish(&sm.a1_ready[grp]);
                }

                // ---------------- epilogue of layer 2
                mbar_wait(&sm.d_ready, d_ph);
                d_ph ^= 1;
                tc_fence_after();
#pragma unroll 1
                for (int c = part; c < 24; c += kParts) {
                    const int n0 = c * 8;
                    uint32_t r[8];
                    tc_ld8(tl + kColD2 + n0, r);
                    tc_wait_ld();
                    add_bias8(r, rec + kRecB2 + n0);
                    float v[8];
                    sp8(r, v);
                    store_a8(tl + kColA2Hi + (n0 >> 1), tl + kColA2Lo + (n0 >> 1), v);
                    if (ACTS) save8(ab, kH + kN1 + n0, 8, v);
                }
                {
                    const int n0 = 192 + 4 * part;
                    uint32_t r[4];
                    tc_ld4(tl + kColD2 + n0, r);
                    tc_wait_ld();
                    add_bias4(r, rec + kRecB2 + n0);
                    float v[4];
                    sp4(r, v);
                    store_a4(tl + kColA2Hi + (n0 >> 1), tl + kColA2Lo + (n0 >> 1), v);
                    if (ACTS && n0 < kH) {
                        const float v8[8] = {v[0], v[1], v[2], v[3], 0.f, 0.f, 0.f, 0.f};
                        save8(ab, kH + kN1 + n0, 4, v8);
                    }
                }
                publish(&sm.a2_ready);

                // ---------------- in the shadow of the layer-3 MMAs: first part of layer 0 of the next member of this tile
                {
                    const unsigned long long rest = (m + 1 < 64) ? (mask >> (m + 1)) : 0ull;
                    a_done = rest != 0;
                    if (a_done) {
                        const uint32_t nslot = (rcount + 1) % kRecSlots;
                        mbar_wait(&sm.rec_full[nslot], ((rcount + 1) / kRecSlots) & 1);
                        const float *nrec = sm.rec[nslot];
                        float nx, ny, nz;
                        member_coords(nrec, nx, ny, nz);
                        mbar_wait(&sm.a2_ready, m_ph);           // the A0a columns overlap D2: every warp must be done reading it
                        layer0_a(nrec, nx, ny, nz, acts_of(m + 1 + (__ffsll((long long)rest) - 1)));
                    }
                }

                // ---------------- epilogue of layer 3 fused with the output layer (dot with w4) and the blend
                mbar_wait(&sm.d_ready, d_ph);
                d_ph ^= 1;
                tc_fence_after();
                float acc = 0.f;
#pragma unroll 1
                for (int c = part; c < 24; c += kParts) {
                    const int n0 = c * 8;
                    uint32_t r[8];
                    tc_ld8(tl + kColD3 + n0, r);
                    tc_wait_ld();
                    add_bias8(r, rec + kRecB3 + n0);
                    const float4 w0 = *reinterpret_cast<const float4 *>(rec + kRecW4 + n0);
                    const float4 w1 = *reinterpret_cast<const float4 *>(rec + kRecW4 + n0 + 4);
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    float v[8];
                    sp8(r, v);
                    if (ACTS) save8(ab, 2 * kH + kN1 + n0, 8, v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc = fmaf(v[e], w[e], acc);                  // w4 pad = 0
                }
                {
                    const int n0 = 192 + 4 * part;
                    uint32_t r[4];
                    tc_ld4(tl + kColD3 + n0, r);
                    tc_wait_ld();
                    add_bias4(r, rec + kRecB3 + n0);
                    const float4 w0 = *reinterpret_cast<const float4 *>(rec + kRecW4 + n0);
                    const float w[4] = {w0.x, w0.y, w0.z, w0.w};
                    float v[4];
                    sp4(r, v);
                    if (ACTS && n0 < kH) {
                        const float v8[8] = {v[0], v[1], v[2], v[3], 0.f, 0.f, 0.f, 0.f};
                        save8(ab, 2 * kH + kN1 + n0, 4, v8);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc = fmaf(v[e], w[e], acc);
                }
                // next member's first layer-1 k-steps may go: its A0a is written and this warp no longer reads D3 (= D1's columns)
                if (a_done) publish(&sm.a0a_ready);
                m_ph ^= 1;
                if (part != 0) sm.partial[part - 1][row] = acc;
                tc_fence_before();
                asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * kParts) : "memory");   // the warps of this lane quarter
                tc_fence_after();
                if (part == 0) {
                    float s = acc + rec[kRecMisc + 0];
#pragma unroll
                    for (int i = 0; i < kParts - 1; ++i) s += sm.partial[i][row];
                    if (p.members_out && valid) p.members_out[((size_t)qi * p.n_points + idx) * p.n_members + m] = s;
                    float d;
                    if (rec[kRecMisc + 4] != 0.f) {
                        const float dx = ax - x, dy = ay - y, dz = az - z;
                        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz) + 10e-6f;
                        d = -(nrm * nrm);
                    } else {
                        d = -0.2f;
                    }
                    const float w = expf(__fdiv_rn(d, 0.01f));
                    num = fmaf(w, quirk ? 1.0f : s, num);
                    if (!PRUNE) den += w;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.rec_empty[rslot]);
                ++rcount;
            }
            if (part == 0 && valid) p.out[(size_t)qi * p.n_points + idx] = __fdiv_rn(num, den + 1e-6f);
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kEpiWarps + 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ packing
// Weight slabs: for weight set s, tensor layer L (1..3), k-step j: N x 16 fp16 hi then N x 16 fp16 lo, each in UMMA
// no-swizzle K-major core-matrix order: byte offset of (n, kk) = (n/8)*256 + (kk/8)*128 + (n%8)*16 + (kk%8)*2.
__global__ void pack_slabs_kernel(const float *__restrict__ W1, const float *__restrict__ W2, const float *__restrict__ W3,
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
        } else if (layer == 2) {
            // reference input order of the skip layer: [h1 (101), xyz (3), cond (96)] / sqrt(2); the cond part is folded
            if (n < kH && k < kN1 + 3) v = W2[((size_t)s * kH + n) * kH + k] * inv_sqrt2;
        } else {
            if (n < kH && k < kH) v = W3[((size_t)s * kH + n) * kH + k];
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
                                                   h->weights.W[3].as<float>(), h->n_sets, h->tc_weights.as<uint8_t>());
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
    tc::Params p{};
    p.weights = h->tc_weights.as<uint8_t>();
    p.recs = h->tc_consts.as<float>();
    p.xyz = q.xyz; p.axes = q.axes; p.res = q.res; p.first = q.first; p.total = q.total; p.n_points = q.n_points;
    p.n_queries = q.n_queries; p.quirk_period = q.quirk_period; p.out = q.out; p.members_out = q.members_out; p.acts_out = q.acts_out;
    p.n_members = h->n_members; p.n_symm = h->cfg.n_symm_pairs;
    const bool prune = h->tc_prune && !q.exact;
    NPHM_REQUIRE(!q.acts_out || (q.n_queries == 1 && q.exact), "activation dump needs a single exact query");
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
    const int grid_x = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    // kernel generation: v8 (in-place operand conversion, tc_ensemble_v8.cu) unless NPHM_TC_KERNEL=v6 asks for the older
    // column-rotation kernel of this file (kept for A/B runs and as a cross-check of the new one)
    static const bool use_v6 = []() { const char *e = getenv("NPHM_TC_KERNEL"); return e && e[0] == 'v' && e[1] == '6'; }();
    if (!use_v6) return tc::launch_ensemble_v8(p, prune, q.acts_out != nullptr, grid_x, stream);
    const int smem = (int)sizeof(tc::Smem);
    auto kern = q.acts_out ? tc::ensemble_tc_kernel<false, true>
                           : (prune ? tc::ensemble_tc_kernel<true, false> : tc::ensemble_tc_kernel<false, false>);
    NPHM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid_x, tc::kThreads, smem, stream>>>(p);
    NPHM_CUDA_CHECK(cudaGetLastError());
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
