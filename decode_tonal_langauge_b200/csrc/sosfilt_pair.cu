// K3 cascade pair: two 4-section zero-phase cascades (two consecutive filtfilt steps of the
// reference chain, frequency_filter.py:218-229 twice) as ONE forward and ONE backward sweep of
// the warm-up kernel (sos_common.cuh).  See include/ecog_sm100.h (ecog_sos_plan.split).
#include "sos_common.cuh"

namespace ecog {

template <bool REV, int NUMA, int NUMB, int NT>
static int launch_warm_pair_nt(const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                               const ecog_sos_plan& p, int nChunks, const SosCoef& coef, double gain, double* padbuf,
                               cudaStream_t st) {
    const size_t smem = ((size_t)kWarmRing * NT * kWPitch) * sizeof(float) + (size_t)NT * (2 * sizeof(int64_t) + sizeof(int2));
    const unsigned grid = (unsigned)ceil_div(C * nChunks, NT);
    auto k = sos_warm_kernel<8, REV, true, NT, NUMA, NUMB>;
    ECOG_TRY((smem_attr<sos_warm_kernel<8, REV, true, NT, NUMA, NUMB>>(smem)));
    k<<<grid, NT, smem, st>>>(x, y, C, T, ldx, ldy, p.chunk, p.tail, nChunks, p.padlen, p.zero_phase, coef, padbuf, gain,
                              p.tail_b);
    return check_launch(REV ? "sos_warm_pair_bwd" : "sos_warm_pair_fwd");
}

template <bool REV, int NUMA, int NUMB>
static int launch_warm_pair(const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                            const ecog_sos_plan& p, int nChunks, const SosCoef& coef, double gain, double* padbuf,
                            cudaStream_t st) {
    if (p.threads >= 512)
        return launch_warm_pair_nt<REV, NUMA, NUMB, 512>(x, y, C, T, ldx, ldy, p, nChunks, coef, gain, padbuf, st);
    if (p.threads == 384)       // one CTA per SM, three warps per scheduler
        return launch_warm_pair_nt<REV, NUMA, NUMB, 384>(x, y, C, T, ldx, ldy, p, nChunks, coef, gain, padbuf, st);
    return launch_warm_pair_nt<REV, NUMA, NUMB, 256>(x, y, C, T, ldx, ldy, p, nChunks, coef, gain, padbuf, st);
}

// Numerator forms of the two halves (sos_common.cuh::unit_form): 2 / 5 = unit forms, 8 = monic general in
// direct form II (the exact factors of the reference's rounded numerator); the gains of both halves ride on
// the input (the host folds the second half's gain into section 0, design.py::pair_design).
// Supported: (2,5) (8,5) (5,2) (5,8) (5,5) (2,2) (8,8).
int run_sos_warm_pair(const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                      const ecog_sos_plan& p, const SosCoef& coef_in, float* tmp, int64_t ldt, double* padbuf,
                      cudaStream_t st) {
    const int nChunks = (int)ceil_div(T, p.chunk);
    const bool vec = aligned16(x) && aligned16(y) && aligned16(tmp) && T % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ldt % 4 == 0;
    if (!vec) return fail(ECOG_E_UNSUPPORTED, "ecog_sosfilt: the cascade pair needs 16-byte aligned rows (T, ld multiples of 4)");
    if (!p.zero_phase) return fail(ECOG_E_UNSUPPORTED, "ecog_sosfilt: the cascade pair is a zero-phase path");
    SosCoef coef = coef_in;
    const int na = unit_form(coef, 0, 4, 0), nb = unit_form(coef, 4, 8, -1);
    const int key = 10 * na + nb;
    if (key != 25 && key != 85 && key != 52 && key != 58 && key != 55 && key != 22 && key != 88)
        return fail(ECOG_E_UNSUPPORTED, "ecog_sosfilt: unsupported numerator forms (%d, %d) for the cascade pair", na, nb);
    const double gain = prepare_form(coef, 0, 4, 0, na);
    (void)prepare_form(coef, 4, 8, -1, nb);
#define ECOG_PAIR(REVV, Y_IN, Y_OUT, LD_IN, LD_OUT)                                                                          \
    (key == 25 ? launch_warm_pair<REVV, 2, 5>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, st)           \
     : key == 85 ? launch_warm_pair<REVV, 8, 5>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, st)         \
     : key == 52 ? launch_warm_pair<REVV, 5, 2>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, st)         \
     : key == 58 ? launch_warm_pair<REVV, 5, 8>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, st)         \
     : key == 55 ? launch_warm_pair<REVV, 5, 5>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, st)         \
     : key == 22 ? launch_warm_pair<REVV, 2, 2>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, st)         \
                 : launch_warm_pair<REVV, 8, 8>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, st))
    ECOG_TRY(ECOG_PAIR(false, x, tmp, ldx, ldt));
    return ECOG_PAIR(true, tmp, y, ldt, ldy);
#undef ECOG_PAIR
}

}  // namespace ecog
