// tcgen05 / TMEM kernel for the identity-SDF ensemble, generation "v8": IN-PLACE operand conversion.
//
// Reference semantics: FastEnsembleDeepSDFMirrored.forward  src/NPHM/models/EnsembledDeepSDF.py:203-267
// Same math, same weight slabs, same per-(query, member) records and same warp roles as tc_ensemble.cu (see the header
// of that file); what changes is where the operands live in TMEM and, with it, how much of the chain overlaps.
//
// An fp32 accumulator column holds one value; the fp16 (hi, lo) split of that value needs 2 x 16 bit = one column too.
// So the epilogue of layer L converts its accumulator D_L into the next layer's A operand IN PLACE: a thread reads the
// 16 accumulator columns of one k-step (tcgen05.ld.x16), applies bias + softplus, splits, and writes 8 columns of packed
// hi pairs + 8 columns of packed lo pairs back to the SAME 16 columns.  One 208-column region therefore serves a layer
// as D and the next one as A, and two regions P, Q (+ a 96-column side buffer S) carry the whole chain:
//
//     S [416,512)  A0 k-steps 0-5 of the NEXT member (layer 0 on CUDA cores, computed in the shadow of layer 3's MMAs)
//     P [0,208)    A0 k-steps 6-12 -> (layer 1 reads S + P) ... D2 -> A2 in place
//     Q [208,416)  D1 (112 cols) -> A1 in place ... D3
//
//     layer 1: A0 (S, P)   -> D1 in Q[0,112)      layer 2: A1 (Q[0,112)) -> D2 in P      layer 3: A2 (P) -> D3 in Q
//
// Because D_{L+1} never shares columns with the region that is being converted, the MMAs of EVERY layer start k-step
// round by k-step round while the previous layer's epilogue is still running (v6 could do that for layer 2 only: D2 + A2 +
// D3 = 624 > 512 columns forced layer 3 to wait for the complete layer-2 epilogue), and the next member's layer 1 runs
// while the output layer (dot with w4) of the current member is still being evaluated:
//
//   compute warps, member m:   [E3a(m-1): D3 cols 0..111] -> q_lo_free
//                              E0b(m): A0 k-steps 6-12 -> P            | tensor pipe: layer 1 of m (D1 -> Q[0,112))
//                              E3b(m-1): D3 cols 112..207, blend       | layer 1 tail
//                              E1(m): Q[0,112) in place, 2 rounds      | layer 2 (D2 -> P), round by round
//                              E2(m): P in place, 4 rounds             | layer 3 (D3 -> Q), round by round
//                              E0a(m+1): A0 k-steps 0-5 -> S           | layer 3 tail
//
// Hazards (every reuse is separated by an mbarrier or by the in-order execution of the tensor pipe):
//   S      written after d_ready(layer 1 of m) proved its last readers (layer-1 k-steps 0-5 of m) complete
//   P      E0b(m) writes after d_ready(layer 3 of m-1): A2(m-1) is dead.  D2(m) overwrites A0(m): issued after layer 1 of m
//   Q lo   D1(m) may only be written once every warp has finished E3a(m-1)  -> q_lo_free
//   Q      D3(m) overwrites A1(m) (issued after layer 2 of m) and Q[112,208), whose last reader E3b(m-1) precedes every
//          warp's a2_ready arrival
#include "tc_ensemble.cuh"

namespace nphm {
namespace tc {
namespace v8 {

constexpr int kParts = 4;                       // column-unit groups per TMEM lane quarter
constexpr int kEpiWarps = 4 * kParts;
constexpr int kThreads = 32 * (kEpiWarps + 2);
constexpr int kColP = 0, kColQ = 208, kColS = 416;
constexpr int kU0a = 6;                         // A0 units (k-steps of 16) that live in S
constexpr int kUnits208 = 13, kUnits112 = 7;
constexpr int rounds(int units) { return (units + kParts - 1) / kParts; }
constexpr int kR0b = rounds(kUnits208 - kU0a), kR1 = rounds(kUnits112), kR2 = rounds(kUnits208);
constexpr int kE3aUnits = 7;                    // D3 columns [0,112) = the columns the next member's D1 needs
constexpr int kM2aUnits = 7;                    // layer-2 column split: units [0,7) = columns [0,112), units [7,13) = [112,208)

// Optional timeline trace (build with -DNPHM_TC_TRACE, tools/build_variant.sh): CTA 0 records clock64() at the phase
// boundaries of its second tile - [member][event], events 0-15 compute warp 0, 16-31 MMA issuer, 32-47 compute warp 13.
#ifdef NPHM_TC_TRACE
__device__ long long g_trace[64][48];
#define TRACE_EVT(cond, m, id) do { if ((cond) && blockIdx.x == 0 && tcount == 1 && (threadIdx.x & 31) == 0) g_trace[m][id] = clock64(); } while (0)
#else
#define TRACE_EVT(cond, m, id) do { } while (0)
#endif

struct __align__(128) Smem {
    uint8_t wbuf[2][kGroupBytes];            // double-buffered weight groups (one bulk copy + one barrier each)
    float rec[kRecSlots][kRecFloats];
    float partial[kParts - 1][128];
    uint64_t w_full[2], w_empty[2];
    uint64_t rec_full[kRecSlots], rec_empty[kRecSlots];
    uint64_t a0a_ready, a0b_ready[kR0b], a1_ready[kR1], a2_ready[kR2], q_lo_free, d_ready, d2b_ready, mask_ready;
    unsigned long long maskq[2][4];
    uint32_t tmem_base;
};

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}

// warp-uniform election of one lane (the lane that issues tcgen05.mma / tcgen05.commit)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// half a k-step of A operand (8 activations, half h of the unit at column `col`): packed fp16 hi pairs go to columns
// [col + 4h, +4), lo pairs to [col + 8 + 4h, +4)  (unit layout: 8 columns of hi pairs, then 8 columns of lo pairs)
__device__ __forceinline__ void store_half(uint32_t col, int h, const float (&v)[8])
{
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
    tc_st4(col + 4 * h, hi);
    tc_st4(col + 8 + 4 * h, lo);
}

// softplus in log2 units with lg2(1 + e) on the FMA pipe: e * P5(e), |error| < 2.2e-6 log2 units = 1.5e-8 in SDF units
// (fp32 evaluation included), i.e. at the round-off level of the activations it produces.  One MUFU (ex2) instead of two.
__device__ __forceinline__ float sp_t_poly5(float t)
{
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-fabsf(t)));
    float pl = fmaf(e, -0.02645743577f, 0.1234514468f);
    pl = fmaf(pl, e, -0.2795380944f);
    pl = fmaf(pl, e, 0.4582707841f);
    pl = fmaf(pl, e, -0.7182819141f);
    pl = fmaf(pl, e, 1.442553145f);
    return fmaf(pl, e, fmaxf(t, 0.0f));
}
// which of 8 consecutive elements take the polynomial (bit set) and which the second MUFU (lg2): balances the MUFU pipe
// (16 lanes/clk/SM) against the issue slots - measured, see DESIGN.md
#ifndef NPHM_POLY_MASK_V8
#define NPHM_POLY_MASK_V8 0xAA
#endif
__device__ __forceinline__ float sp_sel(float t, int e)
{
#if NPHM_V8_POLY5
    return ((NPHM_POLY_MASK_V8 >> (e & 7)) & 1) ? sp_t_poly5(t) : sp_t(t);
#else
    return ((NPHM_POLY_MASK >> (e & 7)) & 1) ? sp_t_poly(t) : sp_t(t);
#endif
}

// half an accumulator unit (already in registers) -> softplus.  The per-column constants (bias, latent-folded input part)
// are already inside the accumulator: the MMAs add them through the K-padding row of B (A carries 1.0 there).
__device__ __forceinline__ void half_sp(const uint32_t (&r)[16], int h, float (&v)[8])
{
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = sp_sel(__uint_as_float(r[8 * h + e]), e);
}

// unit -> compute-warp group assignment.  A phase covers N consecutive units starting at U0 in rounds of kParts units;
// in a round, group `part` takes the unit with in-round index (part - OFF) mod kParts.  The offsets are chosen so that the
// phases a warp runs back to back WITHOUT waiting for the tensor pipe in between add up to the same number of units for
// every group (a round costs ~900 cycles; with OFF = 0 everywhere a member took 14 rounds, now 12):
//   E3a (7 units) + E0b (7) + E3b (6) = 5 + 5 + 5 + 5     E1 (7) = 2 rounds     E2 (13) + E0a (6) = 5 + 5 + 5 + 4
static_assert(kParts == 4, "phase offsets below are worked out for four groups");
template <int V> struct IC { static constexpr int value = V; };
#ifndef NPHM_V8_OFFSETS
#define NPHM_V8_OFFSETS 1
#endif
#ifndef NPHM_V8_PREFETCH
#define NPHM_V8_PREFETCH 0        // measured: keeping the next unit's tcgen05.ld in flight made the kernel 45 % SLOWER
#endif
#ifndef NPHM_V8_POLY5
#define NPHM_V8_POLY5 1
#endif
#ifndef NPHM_V8_E0A_EARLY
#define NPHM_V8_E0A_EARLY 0
#endif
// layer 2 as two column halves (D2 columns [0,112) first, then [112,208)): the layer-2 epilogue starts on the first half while
// the tensor pipe still produces the second - the MMAs of the second half would otherwise be pure tail behind the layer-1 epilogue
#ifndef NPHM_V8_M2_SPLIT
#define NPHM_V8_M2_SPLIT 1
#endif
// layer 0 of the NEXT member's first units in two halves: one in the wait for the first layer-2 columns, one behind the layer-2
// epilogue.  Measured SLOWER (A/B on one GPU: 360 vs 354 ms): kept as an option, off.
#ifndef NPHM_V8_E0A_SPLIT
#define NPHM_V8_E0A_SPLIT 0
#endif
#if NPHM_V8_OFFSETS
constexpr int kOffE3a = 0, kOffE0b = 1, kOffE3b = 3, kOffE1 = 0, kOffE2 = 0, kOffE0a = 1;
constexpr int kOffE2b = 3;          // split layer-2 epilogue: E2a (7) + E2b (6) + E0a (6) = 5 + 5 + 5 + 4 over the four groups
constexpr int kOffE0a1 = 1, kOffE0a2 = 0;       // E0a in two halves of 3 units: groups 1-3, then groups 0-2
#else
constexpr int kOffE3a = 0, kOffE0b = 0, kOffE3b = 0, kOffE1 = 0, kOffE2 = 0, kOffE0a = 0, kOffE2b = 0, kOffE0a1 = 0, kOffE0a2 = 0;
#endif

template <bool PRUNE, bool ACTS>
__global__ void __launch_bounds__(kThreads, 1) ensemble_tc_kernel_v8(const Params p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tiles_per_query = p.blocked ? p.n_tiles : (p.n_points + 127) / 128;
    const long long n_tiles = p.n_tiles;
    // ACTS (fitting: few points, so few tiles): the members of a tile are split over `member_groups` CTAs - a work item is
    // (tile, group of consecutive members), the in-kernel blend is meaningless then and `out` is not written
    const int n_groups = ACTS ? p.member_groups : 1;
    const long long n_items = n_tiles * n_groups;
    auto group_mask = [&](long long item) -> unsigned long long {
        const unsigned long long all = (1ull << p.n_members) - 1;
        if (!ACTS || n_groups == 1) return all;
        const int per = (p.n_members + n_groups - 1) / n_groups, lo = (int)(item % n_groups) * per;
        const int hi = min(p.n_members, lo + per);
        return ((1ull << hi) - 1) & ~((1ull << lo) - 1);
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
        for (int i = 0; i < kRecSlots; ++i) { mbar_init(&sm.rec_full[i], 1); mbar_init(&sm.rec_empty[i], kEpiWarps); }
        mbar_init(&sm.a0a_ready, kEpiWarps);
        for (int i = 0; i < kR0b; ++i) mbar_init(&sm.a0b_ready[i], kEpiWarps);
        for (int i = 0; i < kR1; ++i) mbar_init(&sm.a1_ready[i], kEpiWarps);
        for (int i = 0; i < kR2; ++i) mbar_init(&sm.a2_ready[i], kEpiWarps);
        mbar_init(&sm.q_lo_free, kEpiWarps);
        mbar_init(&sm.d_ready, 1);
        mbar_init(&sm.d2b_ready, 1);
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
            for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
                const long long tile = ACTS ? item / n_groups : item;
                const int qi = (int)(tile / tiles_per_query);
                unsigned long long mask = group_mask(item);
                if (PRUNE) {
                    mbar_wait(&sm.mask_ready, tcount & 1);
                    const unsigned long long *mq = sm.maskq[tcount & 1];
                    mask = mq[0] | mq[1] | mq[2] | mq[3];
                }
                // records run ONE member ahead of the weights (the compute warps start the next member's layer 0 while the
                // current member's layer-2 MMAs are still running); the first record of a tile is loaded at its start
                auto load_rec = [&](int member) {
                    const int rslot = rcount % kRecSlots;
                    mbar_wait(&sm.rec_empty[rslot], ((rcount / kRecSlots) & 1) ^ 1);
                    mbar_expect_tx(&sm.rec_full[rslot], kRecFloats * 4);
                    bulk_g2s(sm.rec[rslot], p.recs + ((size_t)qi * p.n_members + member) * kRecFloats, kRecFloats * 4,
                             &sm.rec_full[rslot]);
                    ++rcount;
                };
                load_rec(__ffsll((long long)mask) - 1);
                for (int m = 0; m < p.n_members; ++m) {
                    if ((PRUNE || ACTS) && !((mask >> m) & 1)) continue;
                    {
                        const unsigned long long rest = (m + 1 < 64) ? (mask >> (m + 1)) : 0ull;
                        if (rest) load_rec(m + 1 + (__ffsll((long long)rest) - 1));
                    }
                    const int set = m < 2 * p.n_symm ? (m >> 1) : m - p.n_symm;
                    const uint8_t *w = p.weights + (size_t)set * kSetBytes;
#pragma unroll 1
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t bytes = g == 0 ? kL1Bytes : (g == 1 ? kL2Bytes : (g == 2 ? 7 * kSlabBytes : 6 * kSlabBytes));
                        mbar_wait(&sm.w_empty[wb], wph ^ 1);
                        mbar_expect_tx(&sm.w_full[wb], bytes);
                        if (g == 1) {
                            // layer 2: the last k-step slab carries this (query, member)'s bias row (l2_slab_kernel)
                            bulk_g2s(sm.wbuf[wb], w, bytes - kSlabBytes, &sm.w_full[wb]);
                            bulk_g2s(sm.wbuf[wb] + (bytes - kSlabBytes), p.l2_slabs + ((size_t)qi * p.n_members + m) * kSlabBytes,
                                     kSlabBytes, &sm.w_full[wb]);
                        } else {
                            bulk_g2s(sm.wbuf[wb], w, bytes, &sm.w_full[wb]);
                        }
                        w += bytes;
                        if (++wb == 2) { wb = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // =========================================================================== MMA issuer
        // The WHOLE warp runs the control flow (waits, loops: warp-uniform, so descriptors live in uniform registers and
        // the k-step loops unroll to straight-line code); one elected lane issues the tcgen05 instructions.  A `lane == 0`
        // branch around everything made the compiler wrap every MMA in an ELECT/BRA.U.ANY waterfall and rebuild the
        // descriptors per k-step (~40 instructions per k-step on a warp that competes with four busy epilogue warps for
        // issue slots: the issue rate, not the tensor pipe, set the pace of the N = 112 layer).
        {
            const bool leader = elect_one();
            int wb = 0;
            uint32_t wph = 0, mph = 0, tcount = 0;
            const uint32_t idesc1 = make_idesc(kNP1), idesc2 = make_idesc(kNP2);
            const uint32_t idesc2a = make_idesc(16 * kM2aUnits), idesc2b = make_idesc(kNP2 - 16 * kM2aUnits);
            // A2 units per round of the layer-2 epilogue (= k-steps of layer 3 per a2_ready barrier)
            constexpr int kA2Round[kR2 + 1] = {0, 4, NPHM_V8_M2_SPLIT ? kM2aUnits : 8, NPHM_V8_M2_SPLIT ? kM2aUnits + 4 : 12, kUnits208};
            // one k-step = 3 MMAs (hi*hi + hi*lo + lo*hi) on the in-place operand unit at column `a`; `b` = descriptor of
            // the slab's hi half, its lo half lies n * 32 bytes (n * 2 descriptor units) further
            auto kstep = [&](uint32_t d, uint32_t a, uint64_t b, int n, uint32_t idesc, bool fresh) {
                if (leader) {
                    tc_mma_ts(d, a, b, idesc, fresh ? 0 : 1);
                    tc_mma_ts(d, a, b + (uint64_t)(n * 2), idesc, 1);
                    tc_mma_ts(d, a + 8, b, idesc, 1);
                }
            };
            auto commit = [&](uint64_t *bar) { if (leader) tc_commit(bar); };
            for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
                const long long tile = ACTS ? item / n_groups : item;
                unsigned long long mask = group_mask(item);
                if (PRUNE) {
                    mbar_wait(&sm.mask_ready, tcount & 1);
                    const unsigned long long *mq = sm.maskq[tcount & 1];
                    mask = mq[0] | mq[1] | mq[2] | mq[3];
                }
                for (int m = 0; m < p.n_members; ++m) {
                    if ((PRUNE || ACTS) && !((mask >> m) & 1)) continue;
                    auto next_buf = [&]() { if (++wb == 2) { wb = 0; wph ^= 1; } };
                    // ---- layer 1 (N 112): A0 units 0-5 from S (written one member ahead), 6-12 from P; D1 -> Q[0,112)
                    {
                        mbar_wait(&sm.q_lo_free, mph);
                        mbar_wait(&sm.a0a_ready, mph);
                        mbar_wait(&sm.w_full[wb], wph);
                        tc_fence_after();
                        TRACE_EVT(true, m, 16);
                        const uint64_t b0 = make_desc(smem_u32(sm.wbuf[wb]), 128, 256);
#pragma unroll
                        for (int j = 0; j < kU0a; ++j)
                            kstep(tmem + kColQ, tmem + kColS + 16 * j, b0 + (uint64_t)(j * (kSlab1Bytes >> 4)), kNP1, idesc1, j == 0);
                        TRACE_EVT(true, m, 17);
#pragma unroll
                        for (int r = 0; r < kR0b; ++r) {
                            mbar_wait(&sm.a0b_ready[r], mph);
                            tc_fence_after();
#pragma unroll
                            for (int j = kU0a + r * kParts; j < kU0a + (r + 1) * kParts; ++j)
                                if (j < kUnits208)
                                    kstep(tmem + kColQ, tmem + kColP + 16 * j, b0 + (uint64_t)(j * (kSlab1Bytes >> 4)), kNP1, idesc1, false);
                        }
                        commit(&sm.w_empty[wb]);
                        commit(&sm.d_ready);
                        TRACE_EVT(true, m, 18);
                        next_buf();
                    }
                    // ---- layer 2 (N 208, K 112): A1 in Q[0,112), D2 -> P, round by round behind the layer-1 epilogue
                    {
                        mbar_wait(&sm.w_full[wb], wph);
                        const uint64_t b0 = make_desc(smem_u32(sm.wbuf[wb]), 128, 256);
#pragma unroll
                        for (int r = 0; r < kR1; ++r) {
                            mbar_wait(&sm.a1_ready[r], mph);
                            tc_fence_after();
                            TRACE_EVT(true, m, 19 + r);
#pragma unroll
                            for (int j = r * kParts; j < (r + 1) * kParts; ++j)
                                if (j < kUnits112)
                                    kstep(tmem + kColP, tmem + kColQ + 16 * j, b0 + (uint64_t)(j * (kSlabBytes >> 4)), kNP2,
                                          NPHM_V8_M2_SPLIT ? idesc2a : idesc2, j == 0);
                        }
#if NPHM_V8_M2_SPLIT
                        commit(&sm.d_ready);                          // D2 columns [0,112) complete
                        // second column half: rows [112,208) of every slab (row block 14: 14 * 256 B further), D2 columns [112,208)
#pragma unroll
                        for (int j = 0; j < kUnits112; ++j)
                            kstep(tmem + kColP + 16 * kM2aUnits, tmem + kColQ + 16 * j,
                                  b0 + (uint64_t)(j * (kSlabBytes >> 4)) + (uint64_t)(16 * kM2aUnits / 8 * 256 / 16), kNP2, idesc2b, j == 0);
                        commit(&sm.w_empty[wb]);
                        commit(&sm.d2b_ready);
#else
                        commit(&sm.w_empty[wb]);
                        commit(&sm.d_ready);
#endif
                        TRACE_EVT(true, m, 22);
                        next_buf();
                    }
                    // ---- layer 3 (N 208, K 208): A2 in P, D3 -> Q, round by round behind the layer-2 epilogue; two weight groups
                    {
                        mbar_wait(&sm.w_full[wb], wph);
                        uint64_t b0 = make_desc(smem_u32(sm.wbuf[wb]), 128, 256);
#pragma unroll
                        for (int r = 0; r < kR2; ++r) {
                            mbar_wait(&sm.a2_ready[r], mph);
                            tc_fence_after();
                            TRACE_EVT(true, m, 23 + r);
#pragma unroll
                            for (int j = kA2Round[r]; j < kA2Round[r + 1]; ++j) {
                                if (j == 7) {                       // second weight group (k-steps 7-12)
                                    commit(&sm.w_empty[wb]);
                                    next_buf();
                                    mbar_wait(&sm.w_full[wb], wph);
                                    tc_fence_after();
                                    b0 = make_desc(smem_u32(sm.wbuf[wb]), 128, 256);
                                }
                                kstep(tmem + kColQ, tmem + kColP + 16 * j, b0 + (uint64_t)((j < 7 ? j : j - 7) * (kSlabBytes >> 4)), kNP3, idesc2, j == 0);
                            }
                        }
                        commit(&sm.w_empty[wb]);
                        commit(&sm.d_ready);
                        TRACE_EVT(true, m, 28);
                        next_buf();
                    }
                    mph ^= 1;
                }
            }
        }
    } else {
        // =========================================================================== compute / epilogue warps
        // thread = (point row, column-unit group): warp w serves TMEM lanes 32*(w&3).. and units u with u % kParts == w>>2.
        const int q = warp & 3, part = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);       // this warp's TMEM lane quarter
        uint32_t d_ph = 0, d2_ph = 0, tcount = 0, rcount = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
                const long long tile = ACTS ? item / n_groups : item;
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
            unsigned long long mask = group_mask(item);
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

            // for the fitting backward: the derivative of every hidden activation, sigma'(pre) = 1 - 2^(-softplus) (v is the
            // softplus in log2 units).  Per (member, 128-point tile) a block of kActLd features x 128 points, feature-major - a
            // warp store is one 128-byte line, and the layer-wise backward GEMMs (fit.cu, tc_linear.cu `blocked`) read it with
            // thread = point.  Columns beyond a layer's width receive don't-care values.
            auto save_half = [&](float *ab, int col0, const float (&v)[8]) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float ex;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-v[e]));
                    ab[(size_t)(col0 + e) * 128 + row] = 1.0f - ex;
                }
            };
            auto acts_of = [&](int member) -> float * {
                return ACTS ? p.acts_out + ((size_t)member * tiles_per_query + (tile % tiles_per_query)) * kActLd * 128 : nullptr;
            };
            // sigma'3 is the A operand of the first backward GEMM: saved operand-ready (tc_linear.cuh "packed": per k-step = unit
            // of 16 features [128 x 16 fp16 hi | 128 x 16 fp16 lo], core-matrix order), zeros in the K padding
            auto save_packed3 = [&](const float *ab, int u, int h, const float (&v)[8], bool zero) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float e0, e1;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-v[2 * i]));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-v[2 * i + 1]));
                    split2(zero ? 0.f : 1.0f - e0, zero ? 0.f : 1.0f - e1, hi[i], lo[i]);
                }
                const size_t blk = (size_t)(ab - p.acts_out) / ((size_t)kActLd * 128);
                uint8_t *dst = p.acts_packed_out + (blk * p.acts_packed_tile_steps + u) * 8192 + (size_t)(row >> 3) * 256 + (size_t)h * 128 +
                               (size_t)(row & 7) * 16;
                *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(dst + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            };
            auto publish = [&](uint64_t *bar) {          // my TMEM stores are visible to the MMA issuer after this
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar);
            };
            auto member_coords = [&](const float *r, float &ccx, float &ccy, float &ccz) {
                ccx = x - r[kRecMisc + 1]; ccy = y - r[kRecMisc + 2]; ccz = z - r[kRecMisc + 3];
                if (r[kRecMisc + 5] != 0.f) ccx = -ccx;          // mirrored member
                ccx *= kS; ccy *= kS; ccz *= kS;                // coordinates in log2 units
            };
            // layer 0 on CUDA cores: one unit (16 outputs) of h0 -> A operand unit at TMEM column `col`
            auto layer0_unit = [&](const float *r, int u, float ccx, float ccy, float ccz, uint32_t col, float *ab) {
                const float4 *l0 = reinterpret_cast<const float4 *>(r + kRecL0) + 16 * u;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (h == 1 && u == kUnits208 - 1) v[0] = 1.0f;      // k = 200: bias row of layer 1; 201..207: zero weight rows
                    else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 w = l0[8 * h + e];
                            const float t = fmaf(w.x, ccx, fmaf(w.y, ccy, fmaf(w.z, ccz, w.w)));
                            v[e] = sp_sel(t, e);
                        }
                    }
                    if (ACTS) save_half(ab, kActOff0 + 16 * u + 8 * h, v);
                    store_half(col, h, v);
                }
            };
            // a phase of layer 0: units [U0, U0 + N) -> region (S for units 0-5, P for the rest), one publish per round
            auto layer0_phase = [&](auto tU0, auto tN, auto tOff, const float *r, float ccx, float ccy, float ccz, uint32_t region,
                                    float *ab, uint64_t *bars) {
                constexpr int U0 = decltype(tU0)::value, N = decltype(tN)::value, OFF = decltype(tOff)::value;
                constexpr int R = (N + kParts - 1) / kParts;
                const int i = (part + kParts - OFF) % kParts;
#pragma unroll 1
                for (int rr = 0; rr < R; ++rr) {
                    const int k = rr * kParts + i;
                    if (k < N) layer0_unit(r, U0 + k, ccx, ccy, ccz, tl + region + 16 * (U0 + k), ab);
                    if (bars != nullptr) publish(&bars[rr]);
                }
            };
            // a phase over accumulator units [U0, U0 + N) of `region`: the tcgen05.ld of the NEXT unit is in flight while the
            // current one is converted (two register buffers, the round loop is unrolled); body(u, v) gets the activations
            // (tPad = the unit whose second half is zero padding: features 200..207 / the c + padding tail of layer 1 - no MUFU
            //  work is spent there, body() gets zeros)
            auto tmem_phase = [&](auto tU0, auto tN, auto tOff, auto tPad, uint32_t region, uint64_t *bars, auto &&body) {
                constexpr int U0 = decltype(tU0)::value, N = decltype(tN)::value, OFF = decltype(tOff)::value, PAD = decltype(tPad)::value;
                constexpr int R = (N + kParts - 1) / kParts;
                const int i = (part + kParts - OFF) % kParts;
#if NPHM_V8_PREFETCH
                uint32_t buf[2][16];
                if (i < N) tc_ld16(tl + region + 16 * (U0 + i), buf[0]);
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const int k = rr * kParts + i;
                    if (k < N) {
                        tc_wait_ld();
                        if (rr + 1 < R && k + kParts < N) tc_ld16(tl + region + 16 * (U0 + k + kParts), buf[(rr + 1) & 1]);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            if (!(h == 1 && U0 + k == PAD)) half_sp(buf[rr & 1], h, v);
                            body(U0 + k, h, v);
                        }
                    }
                    if (bars != nullptr) publish(&bars[rr]);
                }
#else
#pragma unroll 1
                for (int rr = 0; rr < R; ++rr) {
                    const int k = rr * kParts + i;
                    if (k < N) {
                        uint32_t buf[16];
                        tc_ld16(tl + region + 16 * (U0 + k), buf);
                        tc_wait_ld();
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            if (!(h == 1 && U0 + k == PAD)) half_sp(buf, h, v);
                            body(U0 + k, h, v);
                        }
                    }
                    if (bars != nullptr) publish(&bars[rr]);
                }
#endif
            };
            // output layer on CUDA cores: a phase of D3 units -> partial dot with w4
            auto layer3_dot = [&](auto tU0, auto tN, auto tOff, const float *r, float *ab, float acc) -> float {
                tmem_phase(tU0, tN, tOff, IC<kUnits208 - 1>(), kColQ, nullptr, [&](int u, int h, const float (&v)[8]) {
                    if (u == kUnits208 - 1 && h == 1) {                 // padding: w4 is zero there
                        if (ACTS) save_packed3(ab, u, h, v, true);
                        return;
                    }
                    if (ACTS) save_packed3(ab, u, h, v, false);
#pragma unroll
                    for (int i4 = 0; i4 < 2; ++i4) {
                        const float4 w = *reinterpret_cast<const float4 *>(r + kRecW4 + 16 * u + 8 * h + 4 * i4);     // w4 pad = 0
                        acc = fmaf(v[4 * i4], w.x, acc); acc = fmaf(v[4 * i4 + 1], w.y, acc);
                        acc = fmaf(v[4 * i4 + 2], w.z, acc); acc = fmaf(v[4 * i4 + 3], w.w, acc);
                    }
                });
                return acc;
            };
            // member output s = w4 . h3 + b4 (reduced over the column-unit groups of this lane quarter) and the anchor blend
            auto finalize = [&](const float *r, int member, uint32_t rslot, float acc) {
                if (part != 0) sm.partial[part - 1][row] = acc;
                tc_fence_before();
                asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * kParts) : "memory");   // the warps of this lane quarter
                tc_fence_after();
                if (part == 0) {
                    float s = acc + r[kRecMisc + 0];
#pragma unroll
                    for (int i = 0; i < kParts - 1; ++i) s += sm.partial[i][row];
                    if (p.members_out && valid) p.members_out[((size_t)qi * p.n_points + idx) * p.n_members + member] = s;
                    float d;
                    if (r[kRecMisc + 4] != 0.f) {
                        const float dx = r[kRecMisc + 1] - x, dy = r[kRecMisc + 2] - y, dz = r[kRecMisc + 3] - z;
                        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz) + 10e-6f;
                        d = -(nrm * nrm);
                    } else {
                        d = -0.2f;
                    }
                    const float w = expf(__fdiv_rn(d, 0.01f));
                    num = fmaf(w, quirk ? 1.0f : s, num);
                    if (!PRUNE) den += w;
                }
                // (sm.partial is written again one member later, behind barriers that need every warp of this quarter)
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.rec_empty[rslot]);
            };

            bool a_done = false, have_prev = false;
            float acc_prev = 0.f;
            const float *rec_prev = nullptr;
            float *ab_prev = nullptr;
            int m_prev = 0;
            uint32_t rslot_prev = 0;

            for (int m = 0; m < p.n_members; ++m) {
                if (!((mask >> m) & 1)) continue;
                const uint32_t rslot = rcount % kRecSlots;
                const int tr0 = warp == 0 ? 0 : 32;
                const bool trw = warp == 0 || warp == 13;
                TRACE_EVT(trw, m, tr0 + 11);
                mbar_wait(&sm.rec_full[rslot], (rcount / kRecSlots) & 1);
                const float *rec = sm.rec[rslot];
                float cx, cy, cz;
                member_coords(rec, cx, cy, cz);
                float *const ab = acts_of(m);
                TRACE_EVT(trw, m, tr0 + 0);

                // ---------------- layer 0 on CUDA cores -> A operand of layer 1 (units 0-5 normally exist already)
                if (!a_done) {
                    layer0_phase(IC<0>(), IC<kU0a>(), IC<kOffE0a>(), rec, cx, cy, cz, kColS, ab, nullptr);
                    publish(&sm.a0a_ready);
                }
                if (!have_prev) {                        // first member of the tile: nobody is reading Q[0,112)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.q_lo_free);
                }
                layer0_phase(IC<kU0a>(), IC<kUnits208 - kU0a>(), IC<kOffE0b>(), rec, cx, cy, cz, kColP, ab, sm.a0b_ready);
                TRACE_EVT(trw, m, tr0 + 1);
                // ---------------- rest of the previous member's output layer, in the shadow of this member's layer-1 MMAs
                if (have_prev) {
                    acc_prev = layer3_dot(IC<kE3aUnits>(), IC<kUnits208 - kE3aUnits>(), IC<kOffE3b>(), rec_prev, ab_prev, acc_prev);
                    finalize(rec_prev, m_prev, rslot_prev, acc_prev);
                }
                TRACE_EVT(trw, m, tr0 + 2);

                // ---------------- epilogue of layer 1: Q[0,112) in place (K of layer 2 = [h1 (101), c (3), 0...])
                mbar_wait(&sm.d_ready, d_ph);
                d_ph ^= 1;
                tc_fence_after();
                TRACE_EVT(trw, m, tr0 + 3);
                tmem_phase(IC<0>(), IC<kUnits112>(), IC<kOffE1>(), IC<kUnits112 - 1>(), kColQ, sm.a1_ready, [&](int u, int h, float (&v)[8]) {
                    if (u == kUnits112 - 1) {                // features 96..100 | c || 1.0 (bias row of layer 2) | zeros
                        if (h == 0) { v[5] = cx; v[6] = cy; v[7] = cz; }
                        else v[0] = 1.0f;
                    }
                    if (ACTS) save_half(ab, kActOff1 + 16 * u + 8 * h, v);
                    store_half(tl + kColQ + 16 * u, h, v);
                });
                TRACE_EVT(trw, m, tr0 + 4);

                // ---------------- A0 units 0-5 of the next member of this tile -> S (last read by this member's layer 1, which is
                // complete), placed in the shadow of the layer-2 MMAs (NPHM_V8_E0A_EARLY, their last round cannot start
                // before the layer-1 epilogue is complete) or of the layer-3 MMAs
                const unsigned long long rest = (m + 1 < 64) ? (mask >> (m + 1)) : 0ull;
                a_done = rest != 0;
                auto next_layer0_units = [&](auto tU0, auto tN, auto tOff, bool last) {
                    if (a_done) {
                        const uint32_t nslot = (rcount + 1) % kRecSlots;
                        mbar_wait(&sm.rec_full[nslot], ((rcount + 1) / kRecSlots) & 1);
                        const float *nrec = sm.rec[nslot];
                        float nx, ny, nz;
                        member_coords(nrec, nx, ny, nz);
                        layer0_phase(tU0, tN, tOff, nrec, nx, ny, nz, kColS, acts_of(m + 1 + (__ffsll((long long)rest) - 1)), nullptr);
                        if (last) publish(&sm.a0a_ready);
                    }
                };
                auto next_layer0_a = [&]() { next_layer0_units(IC<0>(), IC<kU0a>(), IC<kOffE0a>(), true); };
#if NPHM_V8_E0A_SPLIT
                // first half here, in the wait for the first layer-2 columns (S is free: layer 1 is complete); second half behind
                // the layer-2 epilogue, in the shadow of the layer-3 MMAs - all of it there made the output layer wait
                next_layer0_units(IC<0>(), IC<kU0a / 2>(), IC<kOffE0a1>(), false);
#else
                if (NPHM_V8_E0A_EARLY) next_layer0_a();
#endif
                TRACE_EVT(trw, m, tr0 + 10);

                // ---------------- epilogue of layer 2: P in place
                mbar_wait(&sm.d_ready, d_ph);
                d_ph ^= 1;
                tc_fence_after();
                TRACE_EVT(trw, m, tr0 + 5);
                auto e2_body = [&](int u, int h, float (&v)[8]) {
                    if (u == kUnits208 - 1 && h == 1) v[0] = 1.0f;          // k = 200: bias row of layer 3
                    if (ACTS) save_half(ab, kActOff2 + 16 * u + 8 * h, v);
                    store_half(tl + kColP + 16 * u, h, v);
                };
#if NPHM_V8_M2_SPLIT
                tmem_phase(IC<0>(), IC<kM2aUnits>(), IC<kOffE2>(), IC<-1>(), kColP, sm.a2_ready, e2_body);
                mbar_wait(&sm.d2b_ready, d2_ph);                              // D2 columns [112,208)
                d2_ph ^= 1;
                tc_fence_after();
                tmem_phase(IC<kM2aUnits>(), IC<kUnits208 - kM2aUnits>(), IC<kOffE2b>(), IC<kUnits208 - 1>(), kColP, sm.a2_ready + 2,
                           e2_body);
#else
                tmem_phase(IC<0>(), IC<kUnits208>(), IC<kOffE2>(), IC<kUnits208 - 1>(), kColP, sm.a2_ready, e2_body);
#endif
                TRACE_EVT(trw, m, tr0 + 6);
#if NPHM_V8_E0A_SPLIT
                next_layer0_units(IC<kU0a / 2>(), IC<kU0a - kU0a / 2>(), IC<kOffE0a2>(), true);
#else
                if (!NPHM_V8_E0A_EARLY) next_layer0_a();
#endif
                TRACE_EVT(trw, m, tr0 + 7);
                // ---------------- output layer, first part: D3 columns [0,112) (the columns the next member's D1 lands in)
                mbar_wait(&sm.d_ready, d_ph);
                d_ph ^= 1;
                tc_fence_after();
                TRACE_EVT(trw, m, tr0 + 8);
                acc_prev = layer3_dot(IC<0>(), IC<kE3aUnits>(), IC<kOffE3a>(), rec, ab, 0.f);
                if (a_done) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.q_lo_free);
                }
                TRACE_EVT(trw, m, tr0 + 9);
                have_prev = true;
                rec_prev = rec; ab_prev = ab; m_prev = m; rslot_prev = rslot;
                ++rcount;
            }
            if (have_prev) {                             // last member of the tile
                acc_prev = layer3_dot(IC<kE3aUnits>(), IC<kUnits208 - kE3aUnits>(), IC<kOffE3b>(), rec_prev, ab_prev, acc_prev);
                finalize(rec_prev, m_prev, rslot_prev, acc_prev);
            }
            if (part == 0 && valid && n_groups == 1) p.out[(size_t)qi * p.n_points + idx] = __fdiv_rn(num, den + 1e-6f);
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kEpiWarps + 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }
}

}  // namespace v8

#ifdef NPHM_TC_TRACE
}}
extern "C" int nphm_debug_tc_trace(long long *host, int n_ll)
{
    using namespace nphm;
    NPHM_CUDA_CHECK(cudaDeviceSynchronize());
    NPHM_CUDA_CHECK(cudaMemcpyFromSymbol(host, tc::v8::g_trace, (size_t)n_ll * 8));
    return NPHM_OK;
}
namespace nphm { namespace tc {
#endif

int launch_ensemble_v8(const Params &p, bool prune, bool acts, int grid_x, cudaStream_t stream)
{
    const int smem = (int)sizeof(v8::Smem);
    auto kern = acts ? v8::ensemble_tc_kernel_v8<false, true>
                     : (prune ? v8::ensemble_tc_kernel_v8<true, false> : v8::ensemble_tc_kernel_v8<false, false>);
    NPHM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid_x, v8::kThreads, smem, stream>>>(p);
    NPHM_CUDA_CHECK(cudaGetLastError());
    return NPHM_OK;
}

}  // namespace tc
}  // namespace nphm
