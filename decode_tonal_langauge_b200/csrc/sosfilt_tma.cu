// K3 warm-up sweep with TMA-staged time tiles (the north star's "TMA-staged time tiles in shared
// memory" for frequency_filter.py:218-229).
//
// Same algorithm as sos_warm_kernel (sos_common.cuh): one thread = one chunk of L samples, its start
// state re-created by a zero-state warm-up over the `tail` samples in front of it, exact filtfilt
// start-up at the row ends.  What changes is how the samples travel:
//   * rows are contiguous (ld == T) and T == nChunks * L, so the whole (C, T) array is ONE 2-D tensor
//     [C * nChunks][L]: chunk q of the flat list starts at element q * L;
//   * a CTA owns 256 consecutive chunks.  One stage = 32 samples of all 256 chunks = a 256 x 32 box:
//     ONE cp.async.bulk.tensor.2d issued by one elected thread brings it into shared memory
//     (32 KB, SWIZZLE_128B: 16-byte piece p of box row r lands at piece p ^ (r & 7), so every thread
//     reads its own 128-byte row with conflict-free LDS.128), completion is counted on an mbarrier
//     (complete_tx) -- no per-thread copy instructions, no address arithmetic, no zero-fill predicates;
//   * results overwrite the tile row in place and leave with ONE cp.async.bulk.tensor store;
//   * warm-up stages address (sample L - 32 k, chunk q - 1): the previous chunk's tail in memory
//     order; boxes that stick out of the tensor are zero-filled by the TMA unit.
// Three tile slots per CTA (96 KB, two CTAs per SM): at iteration st the elected thread stores
// slot(st), waits until the store of slot(st-1) has been READ (cp.async.bulk.wait_group.read 1) and
// refills that slot with stage st+2.
#include <cuda.h>

#include "sos_common.cuh"
#include "tma_common.cuh"

namespace ecog {

constexpr int kTNT = 256;            // threads = chunks per CTA = box rows
constexpr int kTSub = 32;            // samples per stage (128-byte box rows)
constexpr int kTSlots = 3;
constexpr int kTileBytes = kTNT * kTSub * 4;

// Thread q = blockIdx.x * 256 + tid owns MEMORY chunk q (row q / nChunks, chunk j = q % nChunks) in both
// sweep directions; the backward sweep walks its chunk from the end and warms up on chunk q + 1.
template <int NSEC, bool REV, int NUM, int NUMB>
__global__ void __launch_bounds__(kTNT, (NSEC >= 8 ? 1 : 2))
sos_warm_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                    const float* __restrict__ x, int64_t C, int64_t T, int L, int tail, int nChunks, int padlen,
                    int zero_phase, SosCoef coef, double* __restrict__ padbuf, double gain, int tail_b) {
    extern __shared__ __align__(1024) unsigned char tma_smem[];
    // SWIZZLE_128B patterns repeat every 1024 bytes of SHARED address: align the tiles there
    unsigned char* base = tma_smem + ((1024u - (smem_u32(tma_smem) & 1023u)) & 1023u);
    float* tiles = reinterpret_cast<float*>(base);                                  // [kTSlots][256][32]
    uint64_t* full = reinterpret_cast<uint64_t*>(base + kTSlots * kTileBytes);      // [kTSlots]
    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.x * kTNT;
    const int64_t q = q0 + tid;
    const bool valid = q < C * nChunks;
    const int64_t row = valid ? q / nChunks : 0;
    const int j = valid ? (int)(q - row * nChunks) : 0;
    const int jj = REV ? nChunks - 1 - j : j;                    // chunk number in sweep order
    const int64_t before = (int64_t)jj * L;                      // samples between the row edge and the chunk
    const int ulo = valid ? (before < tail ? -(int)before : -tail) : 0;       // first logical offset that is filtered
    const int s_inject = valid && before <= tail ? -(int)(before / kTSub) : (1 << 30);
    const int s_full = NUMB < 0 ? -(1 << 30) : (s_inject < -(tail_b / kTSub) ? s_inject : -(tail_b / kTSub));
    const int nStages = L / kTSub;
    const int first = -(tail / kTSub);

    double c[NSEC][5], s[NSEC][2];
#define IN(v) ((NUM >> 1) ? gain * (v) : (v))
#pragma unroll
    for (int i = 0; i < NSEC; ++i) {
#pragma unroll
        for (int k = 0; k < 5; ++k) c[i][k] = coef.c[i][k];
        s[i][0] = 0.0; s[i][1] = 0.0;
    }
    auto step = [&](double u) -> double { return sos_step<NSEC, NUM, NUMB>(u, c, s); };

    // box of stage st: forward  st >= 0: (32 st, q0)         st < 0: (L + 32 st, q0 - 1)
    //                  backward st >= 0: (L - 32 (st+1), q0)  st < 0: (-32 (st+1), q0 + 1)
    auto coords = [&](int st, int& c0, int& c1) {
        if (!REV) { c0 = st >= 0 ? kTSub * st : L + kTSub * st; c1 = (int)q0 - (st < 0 ? 1 : 0); }
        else      { c0 = st >= 0 ? L - kTSub * (st + 1) : -kTSub * (st + 1); c1 = (int)q0 + (st < 0 ? 1 : 0); }
    };
    auto load = [&](int st) {
        if (st < nStages) {
            const int slot = (st - first) % kTSlots;
            int c0, c1;
            coords(st, c0, c1);
            mbar_expect_tx(&full[slot], kTileBytes);
            tma_load_2d(tiles + (size_t)slot * kTNT * kTSub, &in_map, c0, c1, &full[slot]);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kTSlots; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        load(first);
        load(first + 1);
    }

    const int swz = tid & 7;
    for (int st = first; st < nStages; ++st) {
        const int it = st - first;
        const int slot = it % kTSlots;
        float* mine = tiles + (size_t)slot * kTNT * kTSub + tid * kTSub;
        mbar_wait(&full[slot], (uint32_t)((it / kTSlots) & 1));
        if (st == s_inject && zero_phase) {
            // filtfilt start-up: zi * ext[0], then the odd-extension pad
            if (!REV) {
                const float* xr = x + row * T;
                const float x0 = xr[0];
                const float e0 = 2.0f * x0 - xr[padlen];
#pragma unroll
                for (int i = 0; i < NSEC; ++i) { s[i][0] = coef.zi[i][0] * (double)e0; s[i][1] = coef.zi[i][1] * (double)e0; }
                for (int i = 0; i < padlen; ++i) {
                    const float e = 2.0f * x0 - xr[padlen - i];
                    (void)step(IN((double)e));
                }
            } else {
                const double* pb = padbuf + row * padlen;
                const double y0 = pb[padlen - 1];
#pragma unroll
                for (int i = 0; i < NSEC; ++i) { s[i][0] = coef.zi[i][0] * y0; s[i][1] = coef.zi[i][1] * y0; }
                for (int i = padlen - 1; i >= 0; --i) (void)step(IN(pb[i]));
            }
        }
        const bool write = st >= 0;
        const int sbase = st * kTSub;
        if (sbase >= ulo) {                     // stages are whole: inside the filtered range or not at all
            const bool early = NUMB >= 0 && st < s_full;     // pair, early warm-up: first cascade only, nothing stored
            // HALF samples per wavefront block: the ramp-up / ramp-down of a block costs NSEC - 1 diagonals of
            // reduced parallelism, so the 8-section pair runs the whole 32-sample stage as one block (one CTA per
            // SM, registers to spare); the 4-section cascades keep two blocks of 16 (two CTAs per SM, 128 registers)
            constexpr int HALF = NSEC >= 8 ? kTSub : kTSub / 2;
#pragma unroll
            for (int h = 0; h < kTSub / HALF; ++h) {
                float xs[HALF], ys[HALF];
#pragma unroll
                for (int v = 0; v < HALF / 4; ++v) {
                    const int p = (HALF / 4) * h + v;
                    float4 q;
                    if (!REV) {
                        q = *reinterpret_cast<const float4*>(mine + 4 * (p ^ swz));
                    } else {
                        const float4 t4 = *reinterpret_cast<const float4*>(mine + 4 * ((kTSub / 4 - 1 - p) ^ swz));
                        q = make_float4(t4.w, t4.z, t4.y, t4.x);
                    }
                    xs[4 * v] = q.x; xs[4 * v + 1] = q.y; xs[4 * v + 2] = q.z; xs[4 * v + 3] = q.w;
                }
                if (early) {
                    sos_block<NSEC, (NSEC / 2 > 0 ? NSEC / 2 : 1), NUM, NUMB, HALF, (NUM >> 1) != 0, false>(xs, ys, gain, c, s);
                } else {
                    sos_block<NSEC, NSEC, NUM, NUMB, HALF, (NUM >> 1) != 0, true>(xs, ys, gain, c, s);
                    if (write) {
#pragma unroll
                        for (int v = 0; v < HALF / 4; ++v) {
                            const int p = (HALF / 4) * h + v;
                            if (!REV) *reinterpret_cast<float4*>(mine + 4 * (p ^ swz)) = make_float4(ys[4 * v], ys[4 * v + 1], ys[4 * v + 2], ys[4 * v + 3]);
                            else *reinterpret_cast<float4*>(mine + 4 * ((kTSub / 4 - 1 - p) ^ swz)) = make_float4(ys[4 * v + 3], ys[4 * v + 2], ys[4 * v + 1], ys[4 * v]);
                        }
                    }
                }
            }
        }
        if (write) fence_async_smem();          // generic-proxy writes of this tile -> visible to the TMA store
        __syncthreads();                        // everyone is done with slot(st) (and with slot(st-1) long ago)
        if (tid == 0) {
            if (write) {
                int c0, c1;
                coords(st, c0, c1);
                tma_store_2d(&out_map, c0, c1, tiles + (size_t)slot * kTNT * kTSub);
                tma_commit();
                tma_wait_read<1>();             // the store of slot(st-1) has read its tile: refill it
            }
            load(st + kTSlots - 1);             // lands in slot(st-1)
        }
    }
    if (tid == 0) tma_wait_all<0>();

    // forward sweep: the thread that owns a row's last chunk runs on through the right odd-extension
    // pad (float32 like scipy's odd_ext) and keeps the filtered pad in float64 for the backward start-up
    if (!REV && zero_phase && valid && j == nChunks - 1) {
        const float* xr = x + row * T;
        const float xe = xr[T - 1];
        double* pb = padbuf + row * padlen;
        for (int i = 0; i < padlen; ++i) {
            const float e = 2.0f * xe - xr[T - 2 - i];
            pb[i] = step(IN((double)e));
        }
    }
#undef IN
}

static int make_map(CUtensorMap* map, const float* base, int64_t nq, int L) {
    return make_chunk_map(map, base, nq, L, kTSub, kTNT);
}

template <int NSEC, bool REV, int NUM, int NUMB>
static int launch_tma(const float* in, float* out, int64_t C, int64_t T, const ecog_sos_plan& p, int nChunks,
                      const SosCoef& coef, double gain, double* padbuf, cudaStream_t st) {
    CUtensorMap in_map, out_map;
    ECOG_TRY(make_map(&in_map, in, C * nChunks, p.chunk));
    ECOG_TRY(make_map(&out_map, out, C * nChunks, p.chunk));
    const size_t smem = (size_t)kTSlots * kTileBytes + kTSlots * sizeof(uint64_t) + 1024;     // + alignment slack
    ECOG_TRY((smem_attr<sos_warm_tma_kernel<NSEC, REV, NUM, NUMB>>(smem)));
    const unsigned grid = (unsigned)ceil_div(C * nChunks, kTNT);
    sos_warm_tma_kernel<NSEC, REV, NUM, NUMB><<<grid, kTNT, smem, st>>>(in_map, out_map, in, C, T, p.chunk, p.tail, nChunks,
                                                                        p.padlen, p.zero_phase, coef, padbuf, gain, p.tail_b);
    return check_launch(NUMB >= 0 ? (REV ? "sos_warm_tma_pair_bwd" : "sos_warm_tma_pair_fwd")
                                  : (REV ? "sos_warm_tma_bwd" : "sos_warm_tma_fwd"));
}

// Zero-phase sweeps through TMA tiles.  Requirements (checked by the caller): contiguous rows
// (ld == T everywhere), T == nChunks * chunk, chunk and tail multiples of 32, tail <= chunk.
// Instantiated for the 4-section single cascades (forms 0, 2, 5, 8) and the cascade pairs (2,5), (8,5).
int run_sos_warm_tma(const float* x, float* y, int64_t C, int64_t T, const ecog_sos_plan& p, const SosCoef& coef_in,
                     float* tmp, double* padbuf, cudaStream_t st) {
    const int nChunks = (int)(T / p.chunk);
    SosCoef coef = coef_in;
    double gain = 1.0;
#define ECOG_TMA2(NSECV, NUMV, NUMBV)                                                                                  \
    do {                                                                                                               \
        ECOG_TRY((launch_tma<NSECV, false, NUMV, NUMBV>(x, tmp, C, T, p, nChunks, coef, gain, padbuf, st)));           \
        return launch_tma<NSECV, true, NUMV, NUMBV>(tmp, y, C, T, p, nChunks, coef, gain, padbuf, st);                 \
    } while (0)
    if (p.split) {
        const int na = unit_form(coef, 0, 4, 0), nb = unit_form(coef, 4, 8, -1);
        if (!((na == 2 || na == 8) && nb == 5))
            return fail(ECOG_E_UNSUPPORTED, "ecog_sosfilt (TMA): cascade pair forms (%d, %d) are not instantiated", na, nb);
        gain = prepare_form(coef, 0, 4, 0, na);
        (void)prepare_form(coef, 4, 8, -1, nb);
        if (p.split & ECOG_SOS_SPLIT_F32B)
            return run_sos_pair_ws(x, y, C, T, p, coef_in, tmp, st);       // notch threads (float64) + band-pass threads (float32)
        if (na == 2) ECOG_TMA2(8, 2, 5);
        ECOG_TMA2(8, 8, 5);
    }
    if (p.nsec != 4) return fail(ECOG_E_UNSUPPORTED, "ecog_sosfilt (TMA): single cascades of 4 sections only");
    const int num = unit_form(coef, 0, 4, 0);
    gain = prepare_form(coef, 0, 4, 0, num);
    if (num == 2) ECOG_TMA2(4, 2, -1);
    if (num == 5) ECOG_TMA2(4, 5, -1);
    if (num == 8) ECOG_TMA2(4, 8, -1);
    ECOG_TMA2(4, 0, -1);
#undef ECOG_TMA2
}

}  // namespace ecog
