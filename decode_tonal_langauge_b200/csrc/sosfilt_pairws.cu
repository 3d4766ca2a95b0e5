// K3 cascade pair, warp-specialised: the notch half (four float64 sections) and the band-pass half (four
// float32 delta-form sections, sos_common.cuh: Bp32) of ONE chunk run in TWO threads of the same CTA, one
// pipeline stage apart, so that an SM holds four warps per scheduler -- two that feed the FP64 pipe and two
// that run on the FP32 pipe beside them -- WITHOUT more chunks (every extra chunk repeats the warm-up).
// Replaces frequency_filter.py:218-229 twice (58-62 Hz band-stop, 70-150 Hz band-pass), like sosfilt_pair.cu.
//
// Data path = sosfilt_tma.cu: the (C, T) array is one 2-D tensor [C * nChunks][L]; a stage is a box of
// P chunks x 32 samples (SWIZZLE_128B; P = a run-time multiple of 32, 224 at C2) brought in by one
// cp.async.bulk.tensor and completed on an mbarrier.
// Iteration i:  threads 0..P-1   ("A") filter stage i   of their chunk through the notch, float32 result in place;
//               threads P..2P-1  ("B") filter stage i-1 (the tile A finished one iteration earlier) through the
//               band-pass, in place; after the CTA barrier one thread stores tile i-1 with cp.async.bulk.tensor
//               and prefetches stage i+2 into the slot whose store (tile i-2) has been read.  Four slots of P x 128 bytes.
// Row ends: every chunk starts from a ZERO state `tail` samples early (stages that would reach in front of the
// row are skipped); there is no filtfilt start-up here -- the caller (ops.sosfilt_pair) overwrites the first /
// last `tail` samples of every row with the exact sequential result, as it does for the other pair kernels.
// The product of the two cascade gains is applied in float32 between the halves (one FMUL instead of a DMUL).
#include <cuda.h>

#include "sos_common.cuh"
#include "tma_common.cuh"

namespace ecog {

constexpr int kWsMaxChunks = 256;     // chunks per CTA = box rows: a run-time multiple of 32 (ecog_sos_plan.threads), chosen by
                                      // the host so that the CTAs fill the SMs (125 CTAs of 256 chunks leave 23 of 148 SMs idle at C2)
constexpr int kWsSub = 32;            // samples per stage

// packed float32 pairs (sm_100 fma.rn.f32x2 / add.rn.f32x2): lane x = an even section, lane y = the next one
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; "
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

// Band-pass half with the sections packed two by two: on diagonal d section 2k works on sample d - 2k and section
// 2k+1 on sample d - 2k - 1 -- independent, so the pair (2k, 2k+1) is ONE packed recursion step (5 packed operations
// for two sections instead of 10 scalar ones; same roundings as bp32_section).
struct Bp32x2 { float2 c1[2], ne2[2], w1[2], d[2]; };
__device__ __forceinline__ void bp32x2_init(Bp32x2& b, const Bp32Coef& k) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        b.c1[p] = make_float2(k.c1[2 * p], k.c1[2 * p + 1]);
        b.ne2[p] = make_float2(-k.e2[2 * p], -k.e2[2 * p + 1]);
        b.w1[p] = make_float2(0.f, 0.f);
        b.d[p] = make_float2(0.f, 0.f);
    }
}
template <int P>
__device__ __forceinline__ float2 bp32x2_step(float2 v, Bp32x2& b) {
    const float2 dn = fadd2(ffma2(b.c1[P], b.w1[P], ffma2(b.ne2[P], b.d[P], v)), b.d[P]);
    const float2 y = fadd2(dn, b.d[P]);
    b.w1[P] = fadd2(b.w1[P], dn);
    b.d[P] = dn;
    return y;
}
// 32 samples; the input is scaled by `gain`.  The two halves of a pair are offset by one sample, the pairs by two:
// a block runs over 32 + 3 diagonals.  Lanes that are outside the block on a ramp diagonal work on a zero input
// from a saved state (restored afterwards), so every diagonal is the same packed code.
__device__ __forceinline__ void bp32_block32(const float (&x)[kWsSub], float (&y)[kWsSub], float gain, Bp32x2& b) {
    float2 y0 = make_float2(0.f, 0.f), y1 = make_float2(0.f, 0.f);       // outputs of the previous diagonal
#pragma unroll
    for (int d = 0; d < kWsSub + 3; ++d) {
        // pair 1 (sections 2, 3): section 2 takes section 1's output of the previous diagonal (sample d - 2)
        if (d >= 2) {
            const bool lo = d - 2 < kWsSub, hi = d - 3 >= 0 && d - 3 < kWsSub;      // compile time
            const float w1x = b.w1[1].x, dx = b.d[1].x, w1y = b.w1[1].y, dy = b.d[1].y;
            const float2 o = bp32x2_step<1>(make_float2(y0.y, y1.x), b);
            if (!lo) { b.w1[1].x = w1x; b.d[1].x = dx; }
            if (!hi) { b.w1[1].y = w1y; b.d[1].y = dy; }
            y1 = o;
            if (hi) y[d - 3] = o.y;
        }
        // pair 0 (sections 0, 1): section 0 takes sample d, section 1 section 0's output of the previous diagonal
        if (d < kWsSub + 1) {
            const bool lo = d < kWsSub, hi = d - 1 >= 0;
            const float w1x = b.w1[0].x, dx = b.d[0].x, w1y = b.w1[0].y, dy = b.d[0].y;
            const float2 o = bp32x2_step<0>(make_float2(lo ? gain * x[lo ? d : 0] : 0.f, y0.x), b);
            if (!lo) { b.w1[0].x = w1x; b.d[0].x = dx; }
            if (!hi) { b.w1[0].y = w1y; b.d[0].y = dy; }
            y0 = o;
        }
    }
}

// N samples through the float64 sections [J0, J1) in wavefront order (sos_block for a sub-range of the notch)
template <int J0, int J1, int NUM, int N, bool STORE>
__device__ __forceinline__ void notch_block(const float (&x)[N], float (&y)[N], const double (&c)[4][5], double (&s)[4][2]) {
    constexpr int NS = J1 - J0;
    double pipe[NS + 1];
#pragma unroll
    for (int d = 0; d < N + NS - 1; ++d) {
#pragma unroll
        for (int k = NS - 1; k >= 0; --k) {
            const int n = d - k;
            if (n >= 0 && n < N) {
                const double u = k == 0 ? (double)x[n] : pipe[k];
                switch (J0 + k) {     // compile-time after unrolling
                    case 0: pipe[k + 1] = sos_range<0, 1, NUM, 4>(u, c, s); break;
                    case 1: pipe[k + 1] = sos_range<1, 2, NUM, 4>(u, c, s); break;
                    case 2: pipe[k + 1] = sos_range<2, 3, NUM, 4>(u, c, s); break;
                    default: pipe[k + 1] = sos_range<3, 4, NUM, 4>(u, c, s); break;
                }
            }
        }
        if (STORE && d >= NS - 1) y[d - (NS - 1)] = (float)pipe[NS];
    }
}

// NA = 1: threads 0..P-1 run the whole notch (A), P..2P-1 the band-pass (B).
// NA = 2 (measured slower, instantiated only with ECOG_PAIR_THREE_ROLES): threads 0..P-1 run notch sections 0-1 (A1),
//         P..2P-1 sections 2-3 (A2, float32 hand-over in the tile), 2P..3P-1 the band-pass.
template <bool REV, int NUM, int NA>
__global__ void __launch_bounds__(kWsMaxChunks * (NA + 1), 1)
sos_pair_ws_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                   int64_t nq, int L, int tail, int tail_b, int nChunks, SosCoef coef, float gain, Bp32Coef bpc, int P) {
    const int tileFloats = P * kWsSub;             // one tile: P chunks x 32 samples
    const uint32_t tileBytes = (uint32_t)tileFloats * 4u;
    constexpr int NR = NA + 1;                 // roles = pipeline depth in stages
    constexpr int SLOTS = NR + 2;
    extern __shared__ __align__(1024) unsigned char ws_smem[];
    unsigned char* base = ws_smem + ((1024u - (smem_u32(ws_smem) & 1023u)) & 1023u);    // SWIZZLE_128B repeats every 1024 B
    float* tiles = reinterpret_cast<float*>(base);                                      // [SLOTS][P][32]
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)SLOTS * tileBytes);     // [SLOTS]
    const int tid = threadIdx.x;
    const int role = tid / P;                  // warp-uniform (P % 32 == 0)
    const int ch = tid - role * P;
    const int64_t q0 = (int64_t)blockIdx.x * P;
    const int64_t q = q0 + ch;
    const bool valid = q < nq;
    const int j = valid ? (int)(q % nChunks) : 0;
    const int jj = REV ? nChunks - 1 - j : j;                    // chunk number in sweep order
    const int64_t before = (int64_t)jj * L;                      // samples between the row edge and the chunk
    const int nStages = L / kWsSub;
    const int first = -(tail / kWsSub);
    const int s_lo = valid ? (before < tail ? -(int)(before / kWsSub) : first) : 0;      // first stage that is filtered
    const int s_b = s_lo > -(tail_b / kWsSub) ? s_lo : -(tail_b / kWsSub);               // first stage of the band-pass half

    // box of stage st: forward  st >= 0: (32 st, q0)         st < 0: (L + 32 st, q0 - 1)
    //                  backward st >= 0: (L - 32 (st+1), q0)  st < 0: (-32 (st+1), q0 + 1)
    auto coords = [&](int st, int& c0, int& c1) {
        if (!REV) { c0 = st >= 0 ? kWsSub * st : L + kWsSub * st; c1 = (int)q0 - (st < 0 ? 1 : 0); }
        else      { c0 = st >= 0 ? L - kWsSub * (st + 1) : -kWsSub * (st + 1); c1 = (int)q0 + (st < 0 ? 1 : 0); }
    };
    auto load = [&](int st) {
        if (st < nStages) {
            const int slot = (st - first) % SLOTS;
            int c0, c1;
            coords(st, c0, c1);
            mbar_expect_tx(&full[slot], tileBytes);
            tma_load_2d(tiles + (size_t)slot * tileFloats, &in_map, c0, c1, &full[slot]);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        load(first);
        load(first + 1);
    }

    const int swz = ch & 7;
    // this thread's 128-byte row of a tile, 16-byte piece p (in sweep order) <-> float4
    auto rd = [&](const float* mine, float (&xs)[kWsSub]) {
#pragma unroll
        for (int p = 0; p < kWsSub / 4; ++p) {
            float4 v;
            if (!REV) {
                v = *reinterpret_cast<const float4*>(mine + 4 * (p ^ swz));
            } else {
                const float4 t4 = *reinterpret_cast<const float4*>(mine + 4 * ((kWsSub / 4 - 1 - p) ^ swz));
                v = make_float4(t4.w, t4.z, t4.y, t4.x);
            }
            xs[4 * p] = v.x; xs[4 * p + 1] = v.y; xs[4 * p + 2] = v.z; xs[4 * p + 3] = v.w;
        }
    };
    auto wr = [&](float* mine, const float (&ys)[kWsSub]) {
#pragma unroll
        for (int p = 0; p < kWsSub / 4; ++p) {
            if (!REV) *reinterpret_cast<float4*>(mine + 4 * (p ^ swz)) = make_float4(ys[4 * p], ys[4 * p + 1], ys[4 * p + 2], ys[4 * p + 3]);
            else *reinterpret_cast<float4*>(mine + 4 * ((kWsSub / 4 - 1 - p) ^ swz)) = make_float4(ys[4 * p + 3], ys[4 * p + 2], ys[4 * p + 1], ys[4 * p]);
        }
    };
    // after the CTA barrier of iteration `it` (stage st = first + it entered the pipeline): store the tile the last
    // role has finished and prefetch two stages ahead, into the slot whose store has been read
    auto advance = [&](int it) {
        const int st = first + it;
        const int sdone = st - NA;
        if (sdone >= 0 && sdone < nStages) {
            int c0, c1;
            coords(sdone, c0, c1);
            tma_store_2d(&out_map, c0, c1, tiles + (size_t)((it - NA) % SLOTS) * tileFloats);
            tma_commit();
            tma_wait_read<1>();         // the store of tile sdone - 1 has read its slot: refill it
        }
        load(st + 2);                   // lands in the slot of stage st + 2 - SLOTS = sdone - 1
    };

    const int nIter = nStages - first + NA;
    if (role < NA) {
        // ------------------------------------------------------------ A: notch sections, float64, role r is r stages behind
        double c[4][5], s[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int k = 0; k < 5; ++k) c[i][k] = coef.c[i][k];      // kernel parameters: uniform-register operands of the DFMAs
            s[i][0] = 0.0; s[i][1] = 0.0;
        }
        for (int it = 0; it < nIter; ++it) {
            const int st = first + it - role;
            if (it >= role && st < nStages) {
                const int slot = (it - role) % SLOTS;
                float* mine = tiles + (size_t)slot * tileFloats + ch * kWsSub;
                if (role == 0) mbar_wait(&full[slot], (uint32_t)(((it - role) / SLOTS) & 1));
                if (st >= s_lo) {
                    float xs[kWsSub], ys[kWsSub];
                    rd(mine, xs);
                    if (NA == 1) {
                        if (st >= s_b) {
                            notch_block<0, 4, NUM, kWsSub, true>(xs, ys, c, s);
                            wr(mine, ys);
                            fence_async_smem();     // this slot is refilled by the TMA unit later (generic -> async proxy)
                        } else {
                            notch_block<0, 4, NUM, kWsSub, false>(xs, ys, c, s);     // early warm-up: nothing handed on
                        }
                    } else if (role == 0) {
                        notch_block<0, 2, NUM, kWsSub, true>(xs, ys, c, s);
                        wr(mine, ys);
                        fence_async_smem();
                    } else {
                        if (st >= s_b) {
                            notch_block<2, 4, NUM, kWsSub, true>(xs, ys, c, s);
                            wr(mine, ys);
                            fence_async_smem();
                        } else {
                            notch_block<2, 4, NUM, kWsSub, false>(xs, ys, c, s);
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) advance(it);
        }
        if (tid == 0) tma_wait_all<0>();
    } else {
        // ------------------------------------------------------------ B: band-pass half, float32, NA stages behind
        Bp32x2 bp;
        bp32x2_init(bp, bpc);
        for (int it = 0; it < nIter; ++it) {
            const int st = first + it - NA;
            if (it >= NA && st >= s_b) {
                float* mine = tiles + (size_t)((it - NA) % SLOTS) * tileFloats + ch * kWsSub;
                float xs[kWsSub], ys[kWsSub];
                rd(mine, xs);
                bp32_block32(xs, ys, gain, bp);
                if (st >= 0) {
                    wr(mine, ys);
                    fence_async_smem();         // generic-proxy writes of this tile -> visible to the TMA store
                }
            }
            __syncthreads();
        }
    }
}

template <bool REV, int NUM, int NA>
static int launch_ws(const float* in, float* out, int64_t C, const ecog_sos_plan& p, int nChunks, const SosCoef& coef,
                     float gain, const Bp32Coef& bpc, cudaStream_t st) {
    const int P = p.threads;            // chunks per CTA (validated by the caller: 32 .. 256, multiple of 32)
    CUtensorMap in_map, out_map;
    ECOG_TRY(make_chunk_map(&in_map, in, C * nChunks, p.chunk, kWsSub, P));
    ECOG_TRY(make_chunk_map(&out_map, out, C * nChunks, p.chunk, kWsSub, P));
    constexpr int SLOTS = NA + 3;
    const size_t smem = (size_t)SLOTS * P * kWsSub * 4 + SLOTS * sizeof(uint64_t) + 1024;     // + alignment slack
    ECOG_TRY((smem_attr<sos_pair_ws_kernel<REV, NUM, NA>>(smem)));
    const unsigned grid = (unsigned)ceil_div(C * nChunks, (int64_t)P);
    sos_pair_ws_kernel<REV, NUM, NA><<<grid, P * (NA + 1), smem, st>>>(in_map, out_map, C * nChunks, p.chunk, p.tail, p.tail_b, nChunks,
                                                                      coef, gain, bpc, P);
    return check_launch(REV ? "sos_pair_ws_bwd" : "sos_pair_ws_fwd");
}

// Zero-phase sweep pair, notch half of form 2 or 8 (sos_common.cuh::unit_form), band-pass half of form 5.
// Same requirements as run_sos_warm_tma.  `coef_in`: sections 0-3 = notch, 4-7 = band-pass, the product of the
// gains in section 0's b0 (design.py::pair_design).
int run_sos_pair_ws(const float* x, float* y, int64_t C, int64_t T, const ecog_sos_plan& p, const SosCoef& coef_in,
                    float* tmp, cudaStream_t st) {
    const int nChunks = (int)(T / p.chunk);
    if (p.threads < 32 || p.threads > kWsMaxChunks || p.threads % 32)
        return fail(ECOG_E_VALUE, "ecog_sosfilt (float32 band-pass half): plan.threads = chunks per CTA must be a multiple of 32 in 32..256 (got %d)", p.threads);
    SosCoef coef = coef_in;
    const int na = unit_form(coef, 0, 4, 0), nb = unit_form(coef, 4, 8, -1);
    if (!((na == 2 || na == 8) && nb == 5))
        return fail(ECOG_E_UNSUPPORTED, "ecog_sosfilt (float32 band-pass half): cascade pair forms (%d, %d) are not instantiated", na, nb);
    const double gain = prepare_form(coef, 0, 4, 0, na);
    (void)prepare_form(coef, 4, 8, -1, nb);
    Bp32Coef bpc;       // c1 = -(1 + a1 + a2), e2 = 1 - a2, formed in float64, rounded once
    for (int j = 0; j < 4; ++j) {
        bpc.c1[j] = (float)(-(1.0 + coef.c[4 + j][3] + coef.c[4 + j][4]));
        bpc.e2[j] = (float)(1.0 - coef.c[4 + j][4]);
    }
#define ECOG_WS(NUMV, NAV)                                                                            \
    do {                                                                                              \
        ECOG_TRY((launch_ws<false, NUMV, NAV>(x, tmp, C, p, nChunks, coef, (float)gain, bpc, st)));   \
        return launch_ws<true, NUMV, NAV>(tmp, y, C, p, nChunks, coef, (float)gain, bpc, st);         \
    } while (0)
#ifdef ECOG_PAIR_THREE_ROLES      // measured at C2: 10.0 ms against 8.9 ms for two roles (more conversions, a float32 hand-over, 80 registers)
    if (na == 8) ECOG_WS(8, 2);
#endif
    if (na == 2) ECOG_WS(2, 1);
    ECOG_WS(8, 1);
#undef ECOG_WS
}

}  // namespace ecog
